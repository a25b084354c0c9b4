#!/usr/bin/env python
"""Shape fuzz of the remaining entry points on one GPU against the C oracle: random sizes around every tiling constant.

  K2  ksl_sgp4_prop(_counted)   records x epochs, shared and per-sample time grids (persistent tiles of 256 epochs, two per thread)
  K4  ksl_tca_moments           samples from a catalog with every awkward branch; states, counts and first moments
  K5  ksl_pc_count              all-pairs and element-wise counts, ragged sizes, accumulated over row shards
      ksl_upsample / ksl_find_conjunction   bit-equal to the oracle's restatement of the reference functions
      ksl_screen_host           host buffers, batches around the chunking threshold (2^18), against ksl_screen_elems

Not part of the test suite:   python benchmarks/fuzz_shapes.py --seconds 240 [--seed0 0]
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
from kessler_b200 import engine  # noqa: E402
from oracle import c_oracle as C  # noqa: E402
from oracle.sgp4_ref import CONST_NAMES as S_NAMES  # noqa: E402

POS_TOL_KM = 1e-6


def _np(t):
    return t.detach().cpu().numpy()


def fail(msg):
    raise SystemExit("MISMATCH " + msg)


def sane_orbit(el):
    """perigee above 120 km and a physical drag term: outside, SGP4 returns unflagged numbers at 1e6 km that only compare relatively"""
    a = (398600.5 / (el[0] / 60.0) ** 2) ** (1.0 / 3.0)
    return (a * (1.0 - el[1]) > 6378.0 + 120.0) & (np.abs(el[6]) < 5e-3) & (el[1] < 0.5)


def k2(rng, seed, stats):
    n = int(rng.choice([1, 2, 3, 31, 257, 1000, 3001]))
    m = int(rng.choice([1, 2, 127, 128, 129, 255, 256, 257, 511, 512, 513, 1000, 6000, 6001]))
    if n * m > 4_000_000:
        m = max(1, 4_000_000 // n)
    el = common.synthetic_catalog(n, seed=seed, include_edge=bool(rng.random() < 0.5))
    per_sample = bool(rng.random() < 0.4)
    span = float(rng.choice([10.0, 1440.0, 1.0e4, 5.0e4]))
    ts = rng.uniform(-span, span, (n, m) if per_sample else (m,))
    if not per_sample and rng.random() < 0.5:
        ts = np.sort(ts)
    c_ref, e0 = C.sgp4_init(el)
    rv_ref, e_ref = C.sgp4_prop(c_ref, ts, per_sample_t=per_sample)
    c_own, e0g = engine.sgp4_init(el)
    if not np.array_equal(_np(e0g), e0):
        fail(f"k2 init flags seed {seed}")
    if rng.random() < 0.5:
        rv, e = engine.sgp4_prop(c_own, ts, per_sample_t=per_sample)
    else:
        rv, e, n_exact = engine.sgp4_prop_counted(c_own, ts, per_sample_t=per_sample)
        if not 0 <= n_exact <= n * m:
            fail(f"k2 exact-route counter seed {seed}: {n_exact}")
    rv, e = _np(rv), _np(e)
    ok = e0 == 0
    if not np.array_equal(e[ok], e_ref[ok]):
        fail(f"k2 error bits seed {seed} n {n} m {m}")
    # ... and a drag series that has not blown up: |cc1 t| < 0.05 keeps the semi-major axis within 10 % of its epoch value
    # (perigee 147 km with bstar = -2.6e-3 is 18 000 km out after 20 days, where the last bits of the constants are worth 2e-6 km)
    tmax = np.abs(ts).max(axis=-1) if per_sample else np.abs(ts).max()
    good = ok & (e_ref == 0) & sane_orbit(el) & (np.abs(c_ref[list(S_NAMES).index("cc1")]) * tmax < 0.05)
    if good.any():
        dp = np.abs(rv[good][..., :3] - rv_ref[good][..., :3]).max()
        dv = np.abs(rv[good][..., 3:] - rv_ref[good][..., 3:]).max()
        if not (dp < POS_TOL_KM and dv < 1e-9):
            fail(f"k2 states seed {seed} n {n} m {m} per_sample {per_sample}: {dp} km {dv} km/s")
        stats["k2_max_dpos_km"] = max(stats.get("k2_max_dpos_km", 0.0), float(dp))
    bad = ok & (e_ref != 0)
    if not np.array_equal(np.isnan(rv[bad]), np.isnan(rv_ref[bad])):
        fail(f"k2 NaN pattern seed {seed}")
    stats["k2_evals"] = stats.get("k2_evals", 0) + n * m


def k4(rng, seed, stats):
    n = int(rng.choice([1, 15, 16, 17, 255, 4097, 50000]))
    el = common.synthetic_catalog(n, seed=seed, include_edge=bool(rng.random() < 0.5))
    _, e0 = C.sgp4_init(el)
    el = np.ascontiguousarray(el[:, e0 == 0])
    n = el.shape[1]
    if n == 0:
        return
    ts = float(rng.uniform(-7200, 7200))
    st_ref, e_ref = C.tca_states(el, ts)
    fin = np.isfinite(st_ref).all(axis=1)
    ref = st_ref[fin][0] if fin.any() else np.zeros(6)
    mom, errc, st = engine.tca_moments(el, ts, ref, want_states=True)
    mom, st = _np(mom), _np(st)
    if not np.array_equal(np.isfinite(st).all(axis=1), fin):
        fail(f"k4 finiteness seed {seed}")
    good = fin & (e_ref == 0) & sane_orbit(el)
    if good.any() and np.abs(st[good, :3] - st_ref[good, :3]).max() > 1e-3:
        fail(f"k4 states seed {seed}: {np.abs(st[good, :3] - st_ref[good, :3]).max()} m")
    if int(errc[0]) != int((e_ref != 0).sum()) or mom[0] != fin.sum():
        fail(f"k4 counts seed {seed}: {int(errc[0])} {int((e_ref != 0).sum())} {mom[0]} {fin.sum()}")
    d = st[fin] - ref
    if fin.any() and not np.allclose(mom[1:7], d.sum(axis=0), rtol=1e-9, atol=1e-6 * np.abs(d).sum(axis=0).max() + 1e-9):
        fail(f"k4 first moments seed {seed}")
    stats["k4_samples"] = stats.get("k4_samples", 0) + n


def k5(rng, seed, stats):
    nt, nc = int(rng.choice([1, 2, 31, 32, 33, 255, 257, 1023, 1025, 4096, 10000])), int(rng.choice([1, 3, 64, 127, 129, 1000, 5000]))
    scale = float(rng.choice([20.0, 60.0, 500.0]))
    t = rng.standard_normal((3, nt)) * scale
    c = rng.standard_normal((3, nc)) * scale
    if rng.random() < 0.2:
        c[:, 0] = np.nan
    got = int(engine.pc_count(t, c, 70.0)[0])
    if got != C.pc_count(t, c, 70.0, 0):
        fail(f"k5 all pairs seed {seed} {nt}x{nc}")
    m = min(nt, nc)
    if int(engine.pc_count(np.ascontiguousarray(t[:, :m]), np.ascontiguousarray(c[:, :m]), 70.0, pairing=1)[0]) != \
            C.pc_count(np.ascontiguousarray(t[:, :m]), np.ascontiguousarray(c[:, :m]), 70.0, 1):
        fail(f"k5 element-wise seed {seed} {m}")
    cnt, step = None, max(1, nt // int(rng.integers(1, 5)))
    for lo in range(0, nt, step):
        cnt = engine.pc_count(np.ascontiguousarray(t[:, lo:lo + step]), c, 70.0, count=cnt)
    if int(cnt[0]) != got:
        fail(f"k5 row shards seed {seed}")
    stats["k5_pairs"] = stats.get("k5_pairs", 0) + nt * nc


def small_ops(rng, seed, stats):
    n_in, ch = int(rng.choice([1, 2, 3, 7, 50, 257, 6000])), int(rng.choice([1, 3, 6]))
    n_out = int(rng.choice([1, 2, 5, 13, 257, 6001, 60000]))
    x = rng.standard_normal((n_in, ch)) * 1e4
    if not np.array_equal(_np(engine.upsample(x, n_out)), C.upsample(x, n_out)):
        fail(f"upsample seed {seed} {n_in}->{n_out} x{ch}")
    n = int(rng.choice([1, 2, 255, 256, 257, 70000]))
    a = np.cumsum(rng.standard_normal((n, 3)) * 30.0, axis=0)
    b = a + rng.standard_normal((n, 3)) * float(rng.choice([1e2, 4e3, 1e5]))
    if rng.random() < 0.2:
        b[int(rng.integers(0, n)), 1] = np.nan
    oi, od = engine.find_conjunction(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), 5e3)
    oi, od = _np(oi), _np(od)
    i_min, d_min, i_conj, d_conj = C.find_conjunction(a, b, 5e3)
    if oi[0] != i_min or oi[1] != i_conj or not (od[0] == d_min or (np.isnan(od[0]) and np.isnan(d_min))) or \
            (i_conj >= 0 and od[1] != d_conj):
        fail(f"find_conjunction seed {seed} n {n}: {oi} {od} vs {(i_min, d_min, i_conj, d_conj)}")
    stats["small_ops"] = stats.get("small_ops", 0) + 1


def host_path(rng, seed, stats):
    n = int(rng.choice([1, 1000, (1 << 18) - 1, 1 << 18, (1 << 18) + 3, 300001]))
    shared = bool(rng.random() < 0.4)
    gt, ept, gc, epc, _ = common.golden_pair()
    elc = common.perturbed_samples(gc, n, seed)
    elt = gt[:, None].copy() if shared else common.perturbed_samples(gt, n, seed + 1)
    ep_t = np.array([ept]) if shared else np.full(n, ept)
    ep_c = np.full(n, epc)
    n_coarse = int(rng.choice([33, 100, 257]))
    n_fine = (n_coarse - 1) * int(rng.integers(2, 60)) + 1
    tc = (ept - 0.25) + np.arange(n_coarse) * (0.5 / (n_coarse - 1))
    h = engine.screen_host(elt, ep_t, elc, ep_c, tc, n_fine, 5e3)
    d = {k: _np(v) for k, v in engine.screen_elems(elt, ep_t, elc, ep_c, tc, n_fine, 5e3).items()}
    for k in ("i_min", "i_conj", "err", "d_min", "d_conj"):
        if not np.array_equal(np.asarray(h[k]), d[k], equal_nan=True):
            fail(f"screen_host.{k} seed {seed} n {n} shared {shared}")
    stats["host_traces"] = stats.get("host_traces", 0) + n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=240.0)
    ap.add_argument("--seed0", type=int, default=0)
    a = ap.parse_args()
    torch.cuda.set_device(0)
    stats = dict(cases=0)
    t0, seed = time.time(), a.seed0
    parts = (k2, k4, k5, small_ops, host_path)
    while time.time() - t0 < a.seconds:
        part = parts[seed % len(parts)]
        if part is host_path and (seed // len(parts)) % 8:      # the big host batches: one round in eight
            part = small_ops
        part(np.random.default_rng(seed), seed, stats)
        stats["cases"] += 1
        seed += 1
    stats.update(seconds=round(time.time() - t0, 1), seed0=a.seed0, seed_end=seed, mismatches=0)
    print(json.dumps(stats))


if __name__ == "__main__":
    main()
