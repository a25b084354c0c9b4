#!/usr/bin/env python
"""Secondary BASELINE.json configs (the headline bench line is bench.py = configs[1]).

    python benchmarks/run_configs.py [--config 3 4] [--n3 1e8] [--n4 30000]
    torchrun --nproc-per-node N benchmarks/run_configs.py --config 3      # configs[2] sharded over N GPUs

configs[2]  1e8-sample Monte Carlo collision-probability estimate of one pair (device-side draws,
            sgp4init + sgp4 at TCA for both objects, elementwise count + moments), NCCL count all-reduce.
configs[3]  one-vs-catalog screening: 30k synthetic LEO TLEs vs one target over 7 days at the default
            resolution (shared-target path of the fused kernel; the altitude pre-filter of model.py:567
            reported next to the unfiltered rate).
Each prints one JSON line (rank 0).  bench.py's `secondary` carries the driver-run form of all three (device-timed, configs[2] as
strong scaling); this script is the stand-alone, wall-clock form."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from kessler_b200 import distributed as D  # noqa: E402
from kessler_b200 import engine, montecarlo as MC, workloads as W  # noqa: E402
from kessler_b200.tle import MU_EARTH, R_EARTH  # noqa: E402


def synthetic_leo_catalog(n, seed=3):
    """n element sets drawn from default_prior() (model.py:54-153) with seed 3, rejecting period >= 225 min
    and perigee < 120 km -- the same catalog bench.py's secondary.configs3_catalog screens."""
    import bench
    return bench.default_prior_catalog(int(n), seed=seed)


def config3(n_total):
    rank, local, world = D.init()
    torch.cuda.set_device(local)
    t_tle, c_tle = W.pair_a()
    st_t, st_c = MC.nominal_state_m(t_tle, 0.0).reshape(2, 3), MC.nominal_state_m(c_tle, 0.0).reshape(2, 3)
    tca = t_tle.date_mjd + 1.0
    args = (t_tle, st_t, W.T_COV_RTN, c_tle, st_c, W.C_COV_RTN, tca)
    MC.collision_probability(*args, n_total=max(world * 100000, 100000), seed=2)      # warm-up
    torch.cuda.synchronize()
    D.barrier()
    t0 = time.perf_counter()
    r = MC.collision_probability(*args, n_total=int(n_total), collision_threshold=70.0, seed=2)
    torch.cuda.synchronize()
    dt = D.max_over_ranks(time.perf_counter() - t0)
    if rank == 0:
        print(json.dumps({"config": "BASELINE configs[2]: Monte Carlo Pc of one pair, device-side draws", "n_samples": int(n_total),
                          "n_gpus": world, "seconds": dt, "samples_per_s": n_total / dt,
                          "sgp4_init_plus_eval_per_s": 2 * n_total / dt, "count": r["count"], "pc": r["pc"],
                          "err_count": r["err_count"], "sigma_rtn_t_m": np.sqrt(np.diag(r["cov_rtn_t"])[:3]).tolist(),
                          "sigma_rtn_c_m": np.sqrt(np.diag(r["cov_rtn_c"])[:3]).tolist()}), flush=True)
    D.barrier()


def config4(n_cat):
    torch.cuda.set_device(0)
    t_tle, _ = W.pair_a()
    cat = synthetic_leo_catalog(int(n_cat))
    time0 = t_tle.date_mjd - 3.5
    times = W.time_grid(time0)
    tc = np.ascontiguousarray(times[::100])
    n = cat.shape[1]
    ep_c = np.full(n, t_tle.date_mjd)
    el_t = t_tle.elements().reshape(7, 1)
    # altitude pre-filter of model.py:567 (vectorised)
    a_c = (MU_EARTH / (cat[0] / 60.0) ** 2) ** (1.0 / 3.0)
    apo_c, per_c = a_c * (1 + cat[1]) - R_EARTH, a_c * (1 - cat[1]) - R_EARTH
    margin = 5e3 + 1000
    pre = ((t_tle.apogee_alt() + margin) < per_c) | ((apo_c + margin) < t_tle.perigee_alt())
    res = {}
    for tag, sel in (("unfiltered", np.ones(n, bool)), ("prefiltered", ~pre)):
        el = np.ascontiguousarray(cat[:, sel])
        engine.screen_host(el_t, [t_tle.date_mjd], el, ep_c[sel], tc, len(times), 5e3)         # warm-up
        t0 = time.perf_counter()
        out = engine.screen_host(el_t, [t_tle.date_mjd], el, ep_c[sel], tc, len(times), 5e3)
        dt = time.perf_counter() - t0
        res[tag] = {"objects": int(sel.sum()), "seconds": dt, "objects_per_s": float(sel.sum() / dt),
                    "sgp4_evals_per_s": float((sel.sum() + 1) * len(tc) / dt), "conjunctions_below_5km": int((out["i_conj"] > 0).sum()),
                    "min_d_min_m": float(np.nanmin(out["d_min"]))}
    print(json.dumps({"config": "BASELINE configs[3]: one-vs-catalog screening, 7 days, default resolution", "n_catalog": n,
                      "rejected_by_altitude_filter": int(pre.sum()), **res}), flush=True)


def config5(n_events):
    """BASELINE configs[4] at reduced size: synthetic CDM-sequence events by batched rejection sampling
    (forward_batch on the GPU) + the traced forward() of each accepted draw (Newton solve, Monte Carlo
    covariance and 1e4 x 1e4 collision count per CDM)."""
    from kessler_b200 import tle as ktle
    from kessler_b200.model import ConjunctionSimplified
    torch.cuda.set_device(0)
    pop = json.load(open(os.path.join(ROOT, "tests", "golden", "population.json")))
    tles = ktle.load_from_lines([x for p in pop for x in ("0 " + p["name"], p["line1"], p["line2"])])
    torch.manual_seed(0)
    np.random.seed(0)
    model = ConjunctionSimplified(tles=tles, time0=60727.13899462018)
    import contextlib, io
    from kessler_b200 import events
    # (a) the reference's own shape: one traced forward() per accepted draw (get_conjunction)
    n_seq = min(int(n_events), 20)
    t0 = time.perf_counter()
    n_cdms, iters = 0, 0
    with contextlib.redirect_stdout(io.StringIO()):
        for _ in range(n_seq):
            trace, it = model.get_conjunction(batch_size=4096)
            n_cdms += len(trace.nodes["cdms"]["infer"]["cdms"])
            iters += it
    dt_seq = time.perf_counter() - t0
    # (b) batched ground segment: every accepted draw of a screened batch at once (kessler_b200.events)
    events.generate_events(model, 8, batch_size=16384, seed=99)            # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out, screened = events.generate_events(model, int(n_events), batch_size=65536, seed=1)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    n_ev = sum(len(r["hits"]) for r in out)
    n_cdm_b = int(sum(r["issued"].sum() for r in out))
    print(json.dumps({"config": "BASELINE configs[4]: CDM-sequence events from the 20-TLE sample population (ConjunctionSimplified)",
                      "batched": {"events": n_ev, "cdms": n_cdm_b, "candidates_screened": screened, "seconds": dt,
                                  "events_per_s": n_ev / dt, "cdms_per_s": n_cdm_b / dt},
                      "per_event_forward": {"events": n_seq, "cdms": n_cdms, "candidates_screened": iters, "seconds": dt_seq,
                                            "events_per_s": n_seq / dt_seq, "cdms_per_s": n_cdms / dt_seq}}), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, nargs="+", default=[3, 4])
    ap.add_argument("--n3", type=float, default=1e8)
    ap.add_argument("--n4", type=float, default=30000)
    ap.add_argument("--n5", type=float, default=2000)
    a = ap.parse_args()
    if 3 in a.config:
        config3(a.n3)
    if 4 in a.config and int(os.environ.get("RANK", 0)) == 0:
        config4(a.n4)
    if 5 in a.config and int(os.environ.get("RANK", 0)) == 0:
        config5(a.n5)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()
