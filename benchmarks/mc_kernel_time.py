#!/usr/bin/env python
"""Device time of the Monte Carlo Pc kernel alone (CUDA events around ksl_mc_pc; configs[2] parameters).
KSL_LIB_PATH selects a library variant for tuning experiments."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from kessler_b200 import engine, montecarlo as MC, workloads as W  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100000000
torch.cuda.set_device(0)
t_tle, c_tle = W.pair_a()
st_t, st_c = MC.nominal_state_m(t_tle, 0.0).reshape(2, 3), MC.nominal_state_m(c_tle, 0.0).reshape(2, 3)
tca = t_tle.date_mjd + 1.0
mean_t, L_t = MC.element_distribution(t_tle, st_t, W.T_COV_RTN)
mean_c, L_c = MC.element_distribution(c_tle, st_c, W.C_COV_RTN)
ts_t, ts_c = (tca - t_tle.date_mjd) * 1440.0, (tca - c_tle.date_mjd) * 1440.0
ref_t, ref_c = MC.nominal_state_m(t_tle, ts_t), MC.nominal_state_m(c_tle, ts_c)
run = lambda: engine.mc_pc(mean_t, L_t, t_tle._bstar, ts_t, ref_t, mean_c, L_c, c_tle._bstar, ts_c, ref_c, 2, 0, n, 70.0)
for _ in range(2):
    r = run()
torch.cuda.synchronize()
ms = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    r = run()
    b.record()
    b.synchronize()
    ms.append(a.elapsed_time(b))
print(json.dumps({"lib": os.environ.get("KSL_LIB_PATH", "default"), "n": n, "ms": sorted(ms), "samples_per_s": n / (min(ms) * 1e-3),
                  "count": int(r["count"][0]), "err": int(r["err_count"][0]), "n_ok_t": float(r["moments_t"][0])}))
