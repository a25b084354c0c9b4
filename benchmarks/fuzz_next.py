#!/usr/bin/env python
"""Fuzz of the kernels behind the batched ground segment on one GPU:

  K7  ksl_newton_elements     random catalog objects at random offsets from their epoch: the solved mean elements must
                              reproduce the target state through the ORACLE's SGP4 (1e-6 km, 1e-9 km/s)
  K8  ksl_kepler_partials     elements and the 6x6 Jacobian against the host mirror of util.py:187-302 (torch autograd)
      ksl_states_at_fine      rows of propagate + upsample recomputed by the oracle (SGP4 at the coarse epochs, torch-style lerp)
      ksl_pc_mvn_batched      counts against the oracle's all-pairs count of the very points the kernel drew; keyed by CDM id

Not part of the test suite:   python benchmarks/fuzz_next.py --seconds 120 [--seed0 0]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import common  # noqa: E402
from kessler_b200 import engine, util  # noqa: E402
from oracle import c_oracle as C  # noqa: E402

MU_M = 398600.5e9


def fail(msg):
    raise SystemExit("MISMATCH " + msg)


def physical(n, seed, e_max=0.1):
    """catalog objects on which a round trip is meaningful: perigee above 200 km, moderate drag, not deep space"""
    el = common.synthetic_catalog(4 * n + 64, seed=seed, include_edge=False)
    _, e0 = C.sgp4_init(el)
    a = (398600.5 / (el[0] / 60.0) ** 2) ** (1.0 / 3.0)
    keep = (e0 == 0) & (a * (1.0 - el[1]) > 6578.0) & (np.abs(el[6]) < 1e-3) & (el[1] < e_max) & (el[1] > 1e-4) & \
           (el[2] > 0.05) & (el[2] < np.pi - 0.05)
    return np.ascontiguousarray(el[:, keep][:, :n])


def newton(rng, seed, stats):
    el = physical(int(rng.choice([1, 63, 64, 65, 1000])), seed)
    n = el.shape[1]
    if n == 0:
        return
    dt = float(rng.uniform(-4320, 4320))
    st, e = C.tca_states(el, dt)                     # metres
    ok = (e == 0) & np.isfinite(st).all(axis=1) & (np.linalg.norm(st[:, :3], axis=1) > 6.5e6)
    targets, bstar = st[ok] / 1e3, el[6, ok]
    if len(targets) == 0:
        return
    y, resid, iters = engine.newton_elements(targets, bstar)
    y, resid, iters = y.cpu().numpy(), resid.cpu().numpy(), iters.cpu().numpy()
    if not np.isfinite(y).all():
        fail(f"newton non-finite seed {seed}")
    back, _ = C.tca_states(np.concatenate([y.T, bstar[None, :]], axis=0), 0.0)
    dp, dv = np.abs(back[:, :3] / 1e3 - targets[:, :3]).max(), np.abs(back[:, 3:] / 1e3 - targets[:, 3:]).max()
    if dp > 1e-6 or dv > 1e-9:
        w = int(np.argmax(np.abs(back[:, :3] / 1e3 - targets[:, :3]).max(axis=1)))
        fail(f"newton round trip seed {seed}: {dp} km {dv} km/s, iters {iters[w]}, resid {resid[w]}, el {el[:, ok][:, w]}")
    stats["newton_states"] = stats.get("newton_states", 0) + len(targets)
    stats["newton_max_iters"] = max(stats.get("newton_max_iters", 0), int(iters.max()))


def partials(rng, seed, stats):
    el = physical(40, seed, e_max=0.3)
    st, e = C.tca_states(el, float(rng.uniform(-1440, 1440)))
    st = st[(e == 0) & np.isfinite(st).all(axis=1)]
    if len(st) == 0:
        return
    els, jac = engine.kepler_partials(st, MU_M)
    els, jac = els.cpu().numpy(), jac.cpu().numpy()
    for k in range(len(st)):
        s2 = torch.from_numpy(st[k].reshape(2, 3))
        want_el = np.array([float(v) for v in util.from_cartesian_to_keplerian(s2[0], s2[1], MU_M)])
        if not np.allclose(els[k], want_el, rtol=1e-10, atol=1e-10):
            fail(f"kepler elements seed {seed} state {st[k]}: {els[k]} vs {want_el}")
        want = util.keplerian_cartesian_partials(st[k].reshape(2, 3), MU_M)
        scale = np.abs(want).max(axis=1, keepdims=True)
        # (acos-based angles: next to 0 and pi the derivative 1 / sqrt(1 - c^2) amplifies the last bits of both evaluations)
        rel = float((np.abs(jac[k] - want) / scale).max())
        if rel > 1e-6:
            fail(f"kepler partials seed {seed} state {st[k]}: rel {rel}")
        stats["partials_max_rel"] = max(stats.get("partials_max_rel", 0.0), rel)
    stats["partials_states"] = stats.get("partials_states", 0) + len(st)


def at_fine(rng, seed, stats):
    el = physical(int(rng.choice([1, 5, 40])), seed)
    n_rec = el.shape[1]
    if n_rec == 0:
        return
    _, ep0, _, _, _ = common.golden_pair()
    n_coarse = int(rng.choice([2, 33, 100, 601]))
    n_fine = int(rng.choice([(n_coarse - 1) * 100 + 1, (n_coarse - 1) * 7 + 3, n_coarse, max(1, n_coarse // 2)]))
    tc = (ep0 - 1.0) + np.arange(n_coarse) * (2.0 / max(n_coarse - 1, 1))
    n_items = int(rng.choice([1, 127, 128, 129, 3000]))
    rec = rng.integers(0, n_rec, n_items)
    epoch = ep0 + rng.uniform(-2, 2, n_rec)
    idx = rng.integers(0, n_fine, n_items)
    idx[:2] = (0, n_fine - 1)[:min(2, n_items)]
    c_ref, _ = C.sgp4_init(el)
    got, err = engine.states_at_fine(c_ref, rec, epoch[rec], idx, tc, n_fine)
    got, err = got.cpu().numpy(), err.cpu().numpy()
    for r in np.unique(rec):
        rv, e = C.sgp4_prop(np.ascontiguousarray(c_ref[:, [r]]), (tc - epoch[r]) * 1440.0)
        fine = C.upsample(rv[0] * 1e3, n_fine)
        sel = rec == r
        if e[0] == 0 and (np.abs(got[sel][:, :3] - fine[idx[sel]][:, :3]).max() > 1e-3 or np.abs(got[sel][:, 3:] - fine[idx[sel]][:, 3:]).max() > 1e-6):
            fail(f"states_at_fine seed {seed} record {r}: {np.abs(got[sel] - fine[idx[sel]]).max(axis=0)}")
        if e[0] == 0 and (err[sel] != 0).any():      # (an item only sees its two coarse epochs: no flag without a flagged record)
            fail(f"states_at_fine flags seed {seed} record {r}")
    stats["at_fine_items"] = stats.get("at_fine_items", 0) + n_items


def mvn(rng, seed, stats):
    n_cdm, n_pts = int(rng.choice([1, 2, 17, 300])), int(rng.choice([1, 31, 257, 1000]))
    mean = rng.standard_normal((n_cdm, 2, 3)) * 40.0
    A = rng.standard_normal((n_cdm, 2, 3, 3)) * float(rng.choice([5.0, 25.0, 200.0]))
    L = np.linalg.cholesky(A @ np.transpose(A, (0, 1, 3, 2)) + 10.0 * np.eye(3))
    id0 = int(rng.integers(0, 1 << 30))
    counts, pts = engine.pc_mvn_batched(mean, L, n_pts, 70.0, seed=seed, cdm_id0=id0, return_points=True)
    counts, pts = counts.cpu().numpy(), pts.cpu().numpy()
    for k in rng.choice(n_cdm, min(n_cdm, 8), replace=False):
        if counts[k] != C.pc_count(np.ascontiguousarray(pts[k, 0]), np.ascontiguousarray(pts[k, 1]), 70.0, 0):
            fail(f"pc_mvn count seed {seed} cdm {k} of {n_cdm} x {n_pts}")
    if n_cdm > 2:
        h = int(rng.integers(1, n_cdm))
        tail = engine.pc_mvn_batched(mean[h:], L[h:], n_pts, 70.0, seed=seed, cdm_id0=id0 + h).cpu().numpy()
        if not np.array_equal(tail, counts[h:]):
            fail(f"pc_mvn keyed by CDM id seed {seed}")
    stats["mvn_pairs"] = stats.get("mvn_pairs", 0) + n_cdm * n_pts * n_pts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120.0)
    ap.add_argument("--seed0", type=int, default=0)
    a = ap.parse_args()
    torch.cuda.set_device(0)
    stats = dict(cases=0)
    t0, seed = time.time(), a.seed0
    parts = (newton, partials, at_fine, mvn)
    while time.time() - t0 < a.seconds:
        parts[seed % len(parts)](np.random.default_rng(seed), seed, stats)
        stats["cases"] += 1
        seed += 1
    stats.update(seconds=round(time.time() - t0, 1), seed0=a.seed0, seed_end=seed, mismatches=0)
    print(json.dumps(stats))


if __name__ == "__main__":
    main()
