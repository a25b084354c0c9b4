#!/usr/bin/env python
"""Differential fuzz of the Monte Carlo kernel K6 (`ksl_mc_pc`) on one GPU: random element clouds (near-circular, general,
straddling the fast path's bounds, negative eccentricity draws clamped to zero, high drag, days from the epoch), every
sample dumped and redone by the C oracle on the CPU from the SAME drawn elements: states to the tolerance of record,
collision count and error count exactly, first and second moments against the dumped states, and independence of the
split of the sample range.  Not part of the test suite:

    python benchmarks/fuzz_mc.py --seconds 240 [--seed0 0]
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

MU = 398600.5e9


def cloud(rng, el):
    """mean (a [m], e, i, raan, argp, M) and a random lower-triangular factor around one catalog object"""
    n_rad_s = el[0] / 60.0
    mean = np.array([(MU / n_rad_s ** 2) ** (1.0 / 3.0), el[1], el[2], el[3], el[4], el[5]])
    scale = np.array([10.0 ** rng.uniform(-1, 3), 10.0 ** rng.uniform(-7, -3.3), 10.0 ** rng.uniform(-7, -3), 10.0 ** rng.uniform(-7, -3),
                      10.0 ** rng.uniform(-6, -2), 10.0 ** rng.uniform(-6, -2)])
    L = np.tril(rng.normal(size=(6, 6))) * 0.3
    L[np.diag_indices(6)] = 1.0
    return mean, L * scale[:, None]


def case(seed, stats):
    rng = np.random.default_rng(seed)
    cat = common.synthetic_catalog(64, seed=seed, include_edge=False)
    _, e0 = C.sgp4_init(cat)
    ok = np.flatnonzero((e0 == 0) & (cat[1] < 0.2))
    it, ic = rng.choice(ok, 2, replace=False)
    el_t, el_c = cat[:, it].copy(), cat[:, ic].copy()
    kind = rng.choice(["near_circular", "tiny_e", "general", "edge_small", "edge_fast"])
    for el in (el_t, el_c):
        el[1] = dict(near_circular=rng.uniform(5e-4, 6e-3), tiny_e=rng.uniform(0.0, 3e-4), general=rng.uniform(0.01, 0.1),
                     edge_small=rng.uniform(7e-3, 1.1e-2), edge_fast=rng.uniform(0.115, 0.125))[str(kind)]
        a_min = 6.6e6 / (1.0 - el[1] - 0.015)       # keep the perigee of the whole cloud above ~200 km: below the surface
        if (MU / (el[0] / 60.0) ** 2) ** (1.0 / 3.0) < a_min:   # SGP4 returns unflagged garbage (1e9 m) nobody can compare
            el[0] = np.sqrt(MU / a_min ** 3) * 60.0
    mean_t, L_t = cloud(rng, el_t)
    mean_c, L_c = cloud(rng, el_c)
    if kind == "edge_fast":
        L_t[1, :], L_c[1, :] = 0.0, 0.0
        L_t[1, 1] = L_c[1, 1] = 2e-3
    ts_t, ts_c = float(rng.uniform(-7200, 7200)), float(rng.uniform(-7200, 7200))
    n = int(rng.choice([1, 31, 1000, 20000]))
    thr = float(rng.choice([70.0, 1e5, 5e6, 3e7]))
    ref_t, ref_c = rng.normal(size=6) * 1e6, rng.normal(size=6) * 1e6
    off = int(rng.integers(0, 1 << 40))
    args = (mean_t, L_t, el_t[6], ts_t, ref_t, mean_c, L_c, el_c[6], ts_c, ref_c, seed, off)
    r = engine.mc_pc(*args, n, thr, n_dump=n)
    d = r["dump"].cpu().numpy()
    s_t, s_c, x_t, x_c = d[:, :7].T.copy(), d[:, 7:14].T.copy(), d[:, 14:20], d[:, 20:26]
    tag = dict(seed=seed, kind=str(kind), n=n, ts=(ts_t, ts_c), thr=thr)
    o_t, et = C.tca_states(s_t, ts_t)
    o_c, ec = C.tca_states(s_c, ts_c)
    fin = np.isfinite(o_t).all(axis=1) & np.isfinite(o_c).all(axis=1)
    clean = fin & (et == 0) & (ec == 0)
    if not np.array_equal(np.isfinite(x_t).all(axis=1) & np.isfinite(x_c).all(axis=1), fin):
        raise SystemExit(f"MISMATCH finiteness {tag}")
    for name, x, o in (("target", x_t, o_t), ("chaser", x_c, o_c)):
        sane = clean & (np.abs(o[:, :3]).max(axis=1) < 1e8)      # (an unflagged object with bstar = -0.02 is 5e9 m away after a day)
        if sane.any():
            dp, dv = np.abs(x[sane, :3] - o[sane, :3]).max(), np.abs(x[sane, 3:] - o[sane, 3:]).max()
            if dp > 1e-3 or dv > 1e-6:
                raise SystemExit(f"MISMATCH {name} states {tag}: {dp} m, {dv} m/s")
            stats["max_dpos_m"] = max(stats["max_dpos_m"], float(dp))
        far = clean & ~sane
        if far.any() and (np.abs(x[far] - o[far]) > 1e-9 * np.abs(o[far]).max(axis=1, keepdims=True)).any():
            raise SystemExit(f"MISMATCH {name} far states {tag}")
    want = C.pc_count(np.ascontiguousarray(x_t[fin, :3].T), np.ascontiguousarray(x_c[fin, :3].T), thr, 1) if fin.any() else 0
    if int(r["count"][0]) != want:
        raise SystemExit(f"MISMATCH count {tag}: {int(r['count'][0])} vs {want}")
    if int(r["err_count"][0]) != int((((et | ec) != 0) & fin).sum()) and int(r["err_count"][0]) != int(((et | ec) != 0).sum()):
        raise SystemExit(f"MISMATCH err_count {tag}: {int(r['err_count'][0])} vs {int(((et | ec) != 0).sum())}")
    mt, mc = r["moments_t"].cpu().numpy(), r["moments_c"].cpu().numpy()
    for m, x, ref in ((mt, x_t, ref_t), (mc, x_c, ref_c)):
        f = np.isfinite(x).all(axis=1)
        if m[0] != f.sum():
            raise SystemExit(f"MISMATCH moment count {tag}: {m[0]} vs {f.sum()}")
        dx = x[f] - ref
        if f.any() and not np.allclose(m[1:7], dx.sum(axis=0), rtol=1e-9, atol=1e-6 * np.abs(dx).sum(axis=0).max()):
            raise SystemExit(f"MISMATCH first moments {tag}")
    if n >= 31:                                     # the same draws and count whatever the split of the sample range
        h = int(rng.integers(1, n))
        a = engine.mc_pc(*args, h, thr, n_dump=h)
        b = engine.mc_pc(*args[:-1], off + h, n - h, thr, n_dump=min(n - h, 64), count=a["count"], err_count=a["err_count"])
        if int(b["count"][0]) != want or not np.array_equal(a["dump"].cpu().numpy(), d[:h], equal_nan=True) or \
                not np.array_equal(b["dump"].cpu().numpy(), d[h:h + min(n - h, 64)], equal_nan=True):
            raise SystemExit(f"MISMATCH split {tag} at {h}")
    stats["cases"] += 1
    stats["samples"] += n
    stats["hits"] += want
    stats["flagged"] += int(((et | ec) != 0).sum())
    stats[f"kind_{kind}"] = stats.get(f"kind_{kind}", 0) + 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=240.0)
    ap.add_argument("--seed0", type=int, default=0)
    a = ap.parse_args()
    torch.cuda.set_device(0)
    stats = dict(cases=0, samples=0, hits=0, flagged=0, max_dpos_m=0.0)
    t0, seed = time.time(), a.seed0
    while time.time() - t0 < a.seconds:
        case(seed, stats)
        seed += 1
    stats.update(seconds=round(time.time() - t0, 1), seed0=a.seed0, seed_end=seed, mismatches=0)
    print(json.dumps(stats))


if __name__ == "__main__":
    main()
