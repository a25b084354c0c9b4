#!/usr/bin/env python
"""bench.py -- Kessler Monte Carlo conjunction hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--samples S]

Workload (BASELINE.json configs[1]): S = 1e6 perturbed-TLE samples of one conjunction pair
(the default LEO pair, /root/reference/tests/test_model.py:20-26) PER GPU; one step = one
pass of the fused path over that batch: sgp4init (2 x S records) -> fused trace kernel
(2 x 6000 SGP4 evaluations + 600 000-point closest-approach search per sample) -> summary
(counts / moments) -> [N > 1] NCCL all-reduce of the summary.  Metric = conjunction
traces/s (whole job); SGP4 sample-epoch evaluations/s = 12 000 x that.

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every key.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "conjunction traces/s (1 trace = sgp4init x2 + 2x6000 SGP4 sample-epoch evals + 600000-point TCA search)"
UNIT = "traces/s"
N_COARSE, N_FINE, UPSAMPLE = 6000, 600000, 100
EVALS_PER_TRACE = 2 * N_COARSE
FLOP_PER_EVAL = 940.0          # canonical algorithmic FP64 flop per SGP4 evaluation
EXECUTED_FLOP_PER_EVAL = 354.0  # ncu, profiles/r01i_k_screen_warp_sass_opcode_mix.csv: 2 x 125.2 DFMA + 77.3 DMUL + 26.4 DADD
FLOP_PER_INIT = 600.0
BYTES_PER_TRACE = 2 * 7 * 8 + 32  # algorithmic bytes: 2x7 doubles in, 4 values out
THR_M, COLLISION_M = 5e3, 70.0


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle's NumPy/torch port, structured like the reference (one trace at a time)
# ------------------------------------------------------------------------------------------
def _cpu_worker(args):
    import torch
    torch.set_num_threads(1)
    from oracle import kessler_path as K
    el_t, ept, el_c, epc, times, idx = args
    out = []
    for i in idx:
        r = K.screen_trace(el_t[:, i], ept, el_c[:, i], epc, times, UPSAMPLE, THR_M, use_torch=True)
        out.append((r["i_min"], r["d_min"], r["i_conj"]))
    return out


def cpu_port_rate(wl, n_traces, cores):
    """traces/s of the oracle port on `cores` processes (fork; 1 thread each)."""
    import multiprocessing as mp
    el_t, el_c = wl["elems_t"].numpy(), wl["elems_c"].numpy()
    idx = np.arange(n_traces)
    chunks = [c for c in np.array_split(idx, cores) if len(c)]
    args = [(el_t, wl["epoch_t"], el_c, wl["epoch_c"], wl["times"], c) for c in chunks]
    ctx = mp.get_context("fork")
    with ctx.Pool(len(chunks)) as pool:
        pool.map(_cpu_worker, [(el_t, wl["epoch_t"], el_c, wl["epoch_c"], wl["times"], c[:1]) for c in chunks])  # warm
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, args)
        dt = time.perf_counter() - t0
    return n_traces / dt, dt, [r for chunk in res for r in chunk]


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_only_workload(n, seed=1):
    """Same pair / grid as the GPU workload; element samples around the TLE values (no GPU needed)."""
    from kessler_b200 import workloads as W
    from oracle import sgp4_ref as S
    t_tle, c_tle = W.pair_a()
    states = []
    for t in (t_tle, c_tle):
        c, _ = S.sgp4_init(*t.elements())
        r, v, _ = S.sgp4_propagate(c, np.array([0.0]))
        states.append(np.stack([r[0], v[0]]) * 1e3)
    return W.config2_batch(n, seed=seed, rank=0, state_t=states[0], state_c=states[1])


# ------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s in sm if s >= 0.5 * max(sm)] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The literal reference
    cannot run (dsgp4 / pyro are not installable offline), so this times the
    oracle port that restates it, on all host cores, on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cores = host_cores()
    per_step = max(16 * cores, 64)
    wl = cpu_only_workload(per_step, seed=1)
    rates, times_s = [], []
    for step in range(args.warmup + args.steps):
        rate, dt, _ = cpu_port_rate(wl, per_step, cores)
        if step >= args.warmup:
            rates.append(rate); times_s.append(dt)
    value = per_step * len(times_s) / sum(times_s)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(times_s) / len(times_s), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "sgp4_evals_per_s": value * EVALS_PER_TRACE,
            "config": {"workload": "BASELINE configs[1]: perturbed-TLE samples of the default LEO pair, 7-day window, "
                                   "6000 coarse SGP4 epochs -> 600000 fine points, 5 km threshold",
                       "samples_per_step": per_step, "n_coarse": N_COARSE, "n_fine": N_FINE},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{per_step} traces per step of the same workload, oracle NumPy/torch port "
                                       "(oracle/kessler_path.py), one process per core, 1 thread each"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "literal reference unavailable offline (dsgp4/pyro missing); this is the pinned oracle port"}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    from kessler_b200 import _native as NV
    from kessler_b200 import distributed as D
    from kessler_b200 import engine, workloads as W

    NV.require_cuda()
    rank, local_rank, world = D.init()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    n = int(args.samples)
    mode = NV.SCREEN_BRUTE if args.mode == "brute" else NV.SCREEN_ANALYTIC

    # ---- synthetic workload (host, pinned) ------------------------------------------------
    t_tle, c_tle = W.pair_a()
    states = []
    for t in (t_tle, c_tle):
        c, _ = engine.sgp4_init(t.elements().reshape(7, 1))
        rv, _ = engine.sgp4_prop(c, [0.0])
        states.append(rv[0, 0].reshape(2, 3).cpu().numpy() * 1e3)
    wl = W.config2_batch(n, seed=1, rank=rank, state_t=states[0], state_c=states[1])
    h_el_t, h_el_c = wl["elems_t"].pin_memory(), wl["elems_c"].pin_memory()
    h_ep_t = torch.full((n,), wl["epoch_t"], dtype=torch.float64).pin_memory()
    h_ep_c = torch.full((n,), wl["epoch_c"], dtype=torch.float64).pin_memory()
    tc = wl["times_coarse"]
    assert len(tc) == N_COARSE and wl["n_fine"] == N_FINE

    # ---- device-resident arm ---------------------------------------------------------------
    d_el_t, d_el_c = h_el_t.to(dev), h_el_c.to(dev)
    d_ep_t, d_ep_c = h_ep_t.to(dev), h_ep_c.to(dev)
    d_tc = torch.from_numpy(tc).to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    screen_ms = []

    def step(timed):
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        res = engine.screen_elems(d_el_t, d_ep_t, d_el_c, d_ep_c, d_tc, N_FINE, THR_M, mode)
        if timed:
            e1.record()
            screen_ms.append((e0, e1))
        counts, moments = engine.screen_summary(res, COLLISION_M)
        D.allreduce_summary(counts, moments)
        return res, counts, moments

    for _ in range(args.warmup):
        step(False)
        flush.zero_()
    torch.cuda.synchronize()
    D.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = engine.launch_count()
    torch.cuda.synchronize()
    evs = []
    for _ in range(args.steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        res, counts, moments = step(True)
        b.record()
        evs.append((a, b))
        flush.zero_()                       # L2 flush between timed iterations (outside the event pair)
    torch.cuda.synchronize()
    launches = engine.launch_count() - launches0
    D.barrier()
    clocks = sampler.stop() if rank == 0 else None
    dev_s = sum(a.elapsed_time(b) for a, b in evs) * 1e-3
    dev_s = D.max_over_ranks(dev_s)
    scr_s = sum(a.elapsed_time(b) for a, b in screen_ms) * 1e-3 / max(len(screen_ms), 1)
    value = world * n * args.steps / dev_s
    counts_h = counts.cpu().numpy().tolist()
    moments_h = moments.cpu().numpy().tolist()

    # ---- end-to-end arm: host buffers through the C ABI (H2D + kernels + D2H every step) -----
    out = dict(i_min=torch.empty(n, dtype=torch.int64).pin_memory().numpy(), d_min=torch.empty(n, dtype=torch.float64).pin_memory().numpy(),
               i_conj=torch.empty(n, dtype=torch.int64).pin_memory().numpy(), d_conj=torch.empty(n, dtype=torch.float64).pin_memory().numpy(),
               err=torch.empty(n, dtype=torch.int32).pin_memory().numpy())
    np_in = (h_el_t.numpy(), h_ep_t.numpy(), h_el_c.numpy(), h_ep_c.numpy())

    def e2e_step():
        r = engine.screen_host(np_in[0], np_in[1], np_in[2], np_in[3], tc, N_FINE, THR_M, mode, COLLISION_M, device=local_rank, out=out)
        if world > 1:
            c = torch.from_numpy(r["counts"].astype(np.int64)).to(dev)
            m = torch.from_numpy(r["moments"]).to(dev)
            D.allreduce_summary(c, m)
            c.cpu()
        return r

    e2e_step()
    torch.cuda.synchronize()
    D.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r_host = e2e_step()
    torch.cuda.synchronize()
    e2e_s = D.max_over_ranks(time.perf_counter() - t0)
    D.barrier()
    e2e_value = world * n * args.steps / e2e_s
    h2d = (2 * 7 * 8 + 2 * 8) * n + 8 * N_COARSE
    d2h = (4 * 8 + 4) * n + 8 * (NV.SUM_NCOUNT + NV.SUM_NMOMENT)
    same = bool(np.array_equal(r_host["i_min"], res["i_min"].cpu().numpy()))

    if rank != 0:
        D.barrier()
        return
    # ---- roofline of the dominant kernel (k_screen), FP64 vector pipe ------------------------
    peak_tf = engine.fp64_peak_tflops()
    flop_per_launch = n * EVALS_PER_TRACE * FLOP_PER_EVAL
    achieved_tf = flop_per_launch / scr_s / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    traffic, pipe_pct = None, None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum of k_screen_warp from the committed ncu --set full capture
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01i_k_screen_warp_traffic.json")))
        traffic = tr["dram_bytes_total"] * (n / tr["traces_per_launch"])
        pipe_pct = tr.get("sm__pipe_fp64_cycles_active_pct")
    except Exception:
        pass
    roofline = {"bound": "fp64", "kernel": "k_screen_warp (fused SGP4 x2 + closed-form TCA search, one warp per trace)", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved_tf / peak_tf, "traffic": traffic,
                "traffic_source": "profiles/r01i_k_screen_warp_traffic.json (ncu --set full at 1e6 traces/launch; scaled by n if --samples differs)",
                "peak_source": "measured in this run: dependent-free DFMA probe (ksl_fp64_peak_kernel); MEASURED_PEAKS.json has no FP64 entry",
                "flop_per_eval": FLOP_PER_EVAL, "evals_per_launch": n * EVALS_PER_TRACE, "launch_ms": scr_s * 1e3,
                # frac is ALGORITHMIC flop (the reference algorithm's 940 flop / evaluation) over the
                # measured peak, so it exceeds 1 once the kernel needs fewer operations than that count; what the
                # hardware actually ran is in the keys below (committed ncu capture, profiles/r01i_*)
                "executed_flop_per_eval": EXECUTED_FLOP_PER_EVAL,
                "executed_frac": n * EVALS_PER_TRACE * EXECUTED_FLOP_PER_EVAL / scr_s / 1e12 / peak_tf,
                "fp64_pipe_active_pct_ncu": pipe_pct,
                "hbm": {"achieved_gbs": n * BYTES_PER_TRACE / scr_s / 1e9, "peak_gbs": hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)",
                        "note": "algorithmic 144 B/trace; the path is FP64-bound, not HBM-bound"}}
    # ---- CPU baseline beside it (bounded sample, N = 1 only) ---------------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        cores = host_cores()
        n_cpu = max(16 * cores, 64)
        rate, dt, res_cpu = cpu_port_rate(wl, n_cpu, cores)
        agree = all(int(res["i_min"][i]) == res_cpu[i][0] and int(res["i_conj"][i]) == res_cpu[i][2] for i in range(n_cpu))
        # second CPU figure: the scalar C oracle (brute-force fine grid, OpenMP over traces) -- a far tighter CPU
        # implementation than the reference's Python/torch structure; reported so the GPU/CPU ratio is not flattering
        c_rate, c_agree = None, None
        try:
            from oracle import c_oracle as CO
            n_c = max(8 * cores, 64)
            elt_c = np.ascontiguousarray(wl["elems_t"].numpy()[:, :n_c])
            elc_c = np.ascontiguousarray(wl["elems_c"].numpy()[:, :n_c])
            t0 = time.perf_counter()
            ref_c = CO.screen(elt_c, np.full(n_c, wl["epoch_t"]), elc_c, np.full(n_c, wl["epoch_c"]), tc, N_FINE, THR_M)
            c_rate = n_c / (time.perf_counter() - t0)
            # the C oracle enumerates all 600 000 fine points: its indices are the checker for the same traces
            c_agree = bool(np.array_equal(ref_c["i_min"], res["i_min"][:n_c].cpu().numpy()) and
                           np.array_equal(ref_c["i_conj"], res["i_conj"][:n_c].cpu().numpy()) and
                           np.abs(ref_c["d_min"] - res["d_min"][:n_c].cpu().numpy()).max() < 1e-3)
        except Exception:
            pass
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "c_oracle_openmp_traces_per_s": c_rate,
               "c_oracle_indices_and_dmin_agree_with_gpu": c_agree,
               "sample": f"first {n_cpu} traces of the same batch, oracle NumPy/torch port structured like the reference "
                         f"(one trace at a time), {cores} processes x 1 thread, {dt:.1f} s",
               "indices_agree_with_gpu": bool(agree)}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_s * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "sgp4_evals_per_s": value * EVALS_PER_TRACE,
            "config": {"workload": "BASELINE configs[1]: 1e6 perturbed-TLE samples (per GPU) of the default LEO pair "
                                   "(44227/44229), 7-day window, 6000 coarse SGP4 epochs -> 600000 fine points, 5 km threshold",
                       "samples_per_gpu": n, "n_coarse": N_COARSE, "n_fine": N_FINE, "search": args.mode,
                       "l2": "256 MiB flush between timed steps; per-step working set (constants 576 MB) > L2",
                       "parallelism": f"dp{world} (samples sharded, no data-path collective; summary all-reduce)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s * 1e3 / args.steps, "matches_device_arm": same},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "summary": {"counts": counts_h, "moments": moments_h}}
    print(json.dumps(line), flush=True)
    D.barrier()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--samples", type=float, default=1e6, help="samples per GPU per step")
    ap.add_argument("--mode", default="analytic", choices=["analytic", "brute"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        try:
            run_ours(args)
        finally:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                dist.destroy_process_group()


if __name__ == "__main__":
    main()
