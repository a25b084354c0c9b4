#!/usr/bin/env python
"""DFMA throughput of the B200 FP64 pipe against resident warps per SM and independent chains per thread
(ksl_fp64_chain_kernel): the context for the screen kernel's 75 % pipe activity at 16 warps x 2 chains."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from kessler_b200 import _native as N  # noqa: E402
from kessler_b200._native import ptr, stream_ptr  # noqa: E402

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
sms = torch.cuda.get_device_properties(dev).multi_processor_count
sink = torch.zeros(1, dtype=torch.float64, device=dev)
out = {}
MODES = {0: "fma_const", 1: "fma_3reg", 2: "fma_2reg", 3: "dmul", 4: "dadd", 5: "fma_R_UR_R", 6: "fma_R_R_UR"}
for warps in (4, 8, 16, 32, 64):
    row = {}
    for mode, chain_set in ((0, (1, 2, 4, 8)), (1, (1, 2, 4, 8)), (2, (2, 8)), (3, (2, 8)), (4, (2, 8)), (5, (2, 8)), (6, (2, 8))):
        for chains in chain_set:
            threads = 32 * warps if warps <= 32 else 1024
            blocks = sms * (1 if warps <= 32 else warps // 32)
            iters = 200000 // chains
            run = lambda: N.check(N.lib().ksl_fp64_chain_kernel(blocks, threads, iters, chains + 16 * mode, ptr(sink), stream_ptr()), "chain")
            run()
            torch.cuda.synchronize()
            best = 0.0
            for _ in range(3):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); run(); b.record(); b.synchronize()
                best = max(best, 1.0 * chains * iters * blocks * threads / (a.elapsed_time(b) * 1e-3) / 1e12)   # tera-instructions/s
            row[f"{MODES[mode]}_x{chains}"] = round(best, 2)
    out[f"warps_per_sm_{warps}"] = row
    print(f"warps/SM {warps:3d}: " + "  ".join(f"{k}={v:5.2f}" for k, v in row.items()), flush=True)
json.dump({"unit": "1e12 thread-instructions/s (x2 for the flop of a DFMA); the pipe can issue 148 SM x 64 lanes x 1.965 GHz = 18.6", "sms": sms, "table": out}, open(os.path.join(ROOT, "gpurun_out", "fp64_ilp.json"), "w"), indent=1)
