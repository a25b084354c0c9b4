#!/usr/bin/env python
"""ksl_sgp4_prop (K2) device time at several record counts: is it bound by the 48 B per evaluation it writes?"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from kessler_b200 import _native as NV, engine, montecarlo as MC, workloads as W  # noqa: E402
from kessler_b200._native import ptr, stream_ptr  # noqa: E402

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
t_tle, c_tle = W.pair_a()
st_t, st_c = MC.nominal_state_m(t_tle, 0.0).reshape(2, 3), MC.nominal_state_m(c_tle, 0.0).reshape(2, 3)
wl = W.config2_batch(16384, seed=5, rank=0, state_t=st_t, state_c=st_c)
consts_all, _ = engine.sgp4_init(wl["elems_t"].to(dev))
ts = torch.from_numpy((wl["times_coarse"] - wl["epoch_t"]) * 1440.0).to(dev)
for n_rec in (128, 256, 1024, 8192, 16384):
    consts = consts_all[:, :n_rec].contiguous()
    m = ts.numel()
    rv = torch.empty((n_rec, m, 6), dtype=torch.float64, device=dev)
    err = torch.zeros(n_rec, dtype=torch.int32, device=dev)
    run = lambda: NV.check(NV.lib().ksl_sgp4_prop(ptr(consts), n_rec, ptr(ts), m, 0, ptr(rv), ptr(err), stream_ptr()), "k2")
    t = bench.device_ms(run, reps=7, warm=2)
    print(json.dumps({"records": n_rec, "epochs": m, "out_MB": n_rec * m * 48 / 1e6, "ms": t, "evals_per_s": n_rec * m / (t * 1e-3),
                      "write_GBs": n_rec * m * 48 / (t * 1e-3) / 1e9}))
    del rv
