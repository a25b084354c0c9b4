#!/usr/bin/env python
"""Device time of the one-vs-catalog screen (BASELINE configs[3]; bench.py's secondary.configs3_catalog workload) and
what the catalog is made of (share of near-circular / fast-path / exact-path objects)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from kessler_b200 import engine, workloads as W  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 30000
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
t_tle, _ = W.pair_a()
cat = bench.default_prior_catalog(n, seed=3)
times = W.time_grid(t_tle.date_mjd - 3.5)
tc = np.ascontiguousarray(times[::100])
d_el, d_ep = torch.from_numpy(cat).to(dev), torch.full((n,), t_tle.date_mjd, dtype=torch.float64, device=dev)
d_t, d_te = torch.from_numpy(t_tle.elements().reshape(7, 1)).to(dev), torch.tensor([t_tle.date_mjd], dtype=torch.float64, device=dev)
d_tc = torch.from_numpy(tc).to(dev)
run = lambda: engine.screen_elems(d_t, d_te, d_el, d_ep, d_tc, len(times), 5e3)
ms = bench.device_ms(run, reps=7, warm=2)
out = {k: v.cpu().numpy() for k, v in run().items()}
print(json.dumps({"objects": n, "ms": ms, "sgp4_evals_per_s": (n + 1) * len(tc) / (ms * 1e-3), "lib": os.environ.get("KSL_LIB_PATH", "default"),
                  "e_lt_0.008": float((cat[1] < 0.008).mean()), "e_lt_0.12": float((cat[1] < 0.12).mean()),
                  "err_free": float((out["err"] == 0).mean()), "conj": int((out["i_conj"] > 0).sum())}))
