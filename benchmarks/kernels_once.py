#!/usr/bin/env python
"""One launch of every kernel other than the fused screen at a representative size -- the command behind the ncu
captures of profiles/<tag>_kernels_*.json (ncu --set full -k regex:<name> ...)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from kessler_b200 import engine, montecarlo as MC, workloads as W  # noqa: E402

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
t_tle, c_tle = W.pair_a()
st_t, st_c = MC.nominal_state_m(t_tle, 0.0).reshape(2, 3), MC.nominal_state_m(c_tle, 0.0).reshape(2, 3)
wl = W.config2_batch(1000000, seed=5, rank=0, state_t=st_t, state_c=st_c)
el = wl["elems_t"].to(dev)
# K1 sgp4init, 1e6 records
consts, _ = engine.sgp4_init(el)
# K2 propagate, 8192 records x 6000 epochs
ts = torch.from_numpy((wl["times_coarse"] - wl["epoch_t"]) * 1440.0).to(dev)
rv, _ = engine.sgp4_prop(consts[:, :8192].contiguous(), ts)
del rv
# K4 Monte Carlo states + moments, 1e6 samples at one epoch
mom, errc, _ = engine.tca_moments(el, 1440.0, np.zeros(6))
# K5 all-pairs count at the reference size
rng = np.random.default_rng(4)
cnt = engine.pc_count(rng.standard_normal((3, 10000)) * 60.0, rng.standard_normal((3, 10000)) * 60.0, 70.0)
# K6 Monte Carlo Pc, 2e7 samples
pp = MC.prepare_pair(t_tle, st_t, W.T_COV_RTN, c_tle, st_c, W.C_COV_RTN, t_tle.date_mjd + 1.0)
MC.run_prepared(pp, 20000000, seed=2, to_host=False)
# K7 Newton solve + K8 Kepler partials, 1e5 states
rv0, _ = engine.sgp4_prop(consts[:, :100000].contiguous(), torch.zeros(1, dtype=torch.float64, device=dev))
y, resid, it = engine.newton_elements(rv0[:, 0, :], el[6, :100000])
elk, jac = engine.kepler_partials(rv0[:, 0, :] * 1e3, 398600.5e9)
# FP64 peak probe (the roofline denominator)
tf = engine.fp64_peak_tflops(iters=20000, reps=1)
torch.cuda.synchronize()
print("ok", float(mom[0]), int(cnt[0]), float(resid.max()), tf)
