#!/usr/bin/env python
"""cProfile of the batched CDM-sequence generation (BASELINE configs[4]) -- where the host time goes."""
import cProfile
import io
import json
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from kessler_b200 import events, tle as ktle  # noqa: E402
from kessler_b200.model import ConjunctionSimplified  # noqa: E402

n_events = int(float(sys.argv[1])) if len(sys.argv) > 1 else 5000
torch.cuda.set_device(0)
pop = json.load(open(os.path.join(ROOT, "tests", "golden", "population.json")))
tles = ktle.load_from_lines([x for p in pop for x in ("0 " + p["name"], p["line1"], p["line2"])])
torch.manual_seed(0)
np.random.seed(0)
model = ConjunctionSimplified(tles=tles, time0=60727.13899462018)
events.generate_events(model, 8, batch_size=16384, seed=99)
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
out, screened = events.generate_events(model, n_events, batch_size=65536, seed=1)
torch.cuda.synchronize()
pr.disable()
dt = time.perf_counter() - t0
n_ev = sum(len(r["hits"]) for r in out)
print(json.dumps({"events": n_ev, "screened": screened, "seconds": dt, "events_per_s": n_ev / dt}))
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
