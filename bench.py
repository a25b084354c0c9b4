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
# What the committed ncu captures say about the dominant kernel (profiles/current.json names them).  bench.py DERIVES the
# executed-flop figure from the opcode-mix CSV at run time instead of carrying a literal, so the roofline record cannot
# drift from the evidence; tests/test_bench_contract.py checks that manifest, CSV and metrics JSON agree.
def kernel_sources_sha():
    import hashlib
    h = hashlib.sha256()
    for name in ("screen_kernels.cuh", "sgp4_device.cuh", "screen_device.cuh", "sincos_grid_table.cuh"):
        h.update(open(os.path.join(ROOT, "kessler_b200", "csrc", name), "rb").read())
    return h.hexdigest()[:16]


def load_profile(kernel="k_screen_warp"):
    """-> dict(executed_flop_per_unit, fp64_instr_per_unit, pipe_active_pct, issue_active_pct, dram_bytes_per_trace,
    files, matches_source) from profiles/current.json, or None."""
    try:
        man = json.load(open(os.path.join(ROOT, "profiles", "current.json")))[kernel]
        per = {}
        for ln in open(os.path.join(ROOT, "profiles", man["opcode_mix"])).read().strip().splitlines()[1:]:
            op, _, per_unit = ln.split(",")
            per[op] = float(per_unit)
        met = json.load(open(os.path.join(ROOT, "profiles", man["metrics"])))
        return {"executed_flop_per_unit": 2.0 * per.get("DFMA", 0.0) + per.get("DMUL", 0.0) + per.get("DADD", 0.0),
                "fp64_instr_per_unit": sum(per.get(k, 0.0) for k in ("DFMA", "DMUL", "DADD", "DSETP")),
                "instr_per_unit": sum(per.values()),
                "pipe_active_pct": met.get("sm__pipe_fp64_cycles_active_pct"), "issue_active_pct": met.get("smsp__issue_active_pct"),
                "dram_bytes_per_trace": met.get("bytes_per_trace"), "traces_per_launch": met.get("traces_per_launch"),
                "operand_limited_cycles_per_warp_eval": met.get("operand_limited_cycles_per_warp_eval"),
                "dfma_three_register_operands_per_eval": met.get("dfma_three_register_operands_per_eval"),
                "files": [man["opcode_mix"], man["metrics"]],
                "matches_source": man.get("kernel_sources_sha16") == kernel_sources_sha()}
    except Exception:
        return None


def device_ms(fn, reps=3, warm=1):
    """median CUDA-event time of fn() on the current stream"""
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms))


def default_prior_catalog(n, seed=3):
    """n TLE element sets drawn from default_prior() (model.py:54-153) with seed `seed`, rejecting
    period >= 225 min and perigee < 120 km.  Host sampler (the reference's own distributions), TLE-text quantised."""
    import torch
    from kessler_b200.model import Conjunction
    from kessler_b200.tle import MU_EARTH, R_EARTH
    import contextlib
    import io
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):      # (the constructor announces its default instruments on stdout)
        m = Conjunction()
    out = np.zeros((7, 0))
    while out.shape[1] < n:
        raw, el, _ = m.sample_elements('c', 2 * n)
        a = (MU_EARTH / np.maximum(raw['mean_motion'], 1e-12) ** 2) ** (1.0 / 3.0)
        keep = (raw['mean_motion'] > 2 * np.pi / (225.0 * 60.0)) & (a * (1 - raw['eccentricity']) - R_EARTH > 120e3)
        out = np.concatenate([out, el[:, keep]], axis=1)
    return np.ascontiguousarray(out[:, :n])


def secondary_workloads(engine, NV, W, D, rank, world, dev, peak_tf, hbm_peak, n_events=100000):
    """BASELINE configs[2] (strong scaling over the ranks), and -- single GPU only -- configs[3], K2 and K5.
    Every figure is device time (CUDA events) unless it says e2e."""
    import torch
    from kessler_b200 import montecarlo as MC
    from kessler_b200.tle import MU_EARTH, R_EARTH
    sec = {}
    t_tle, c_tle = W.pair_a()
    # ---- configs[2]: 1e8-sample Monte Carlo Pc of one pair, TOTAL fixed, sharded over the ranks -----------------------
    n_total = 100000000
    st_t, st_c = MC.nominal_state_m(t_tle, 0.0).reshape(2, 3), MC.nominal_state_m(c_tle, 0.0).reshape(2, 3)
    pp = MC.prepare_pair(t_tle, st_t, W.T_COV_RTN, c_tle, st_c, W.C_COV_RTN, t_tle.date_mjd + 1.0)
    MC.run_prepared(pp, n_total, seed=2, to_host=False)
    torch.cuda.synchronize()
    D.barrier()
    ms = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        D.barrier()
        a.record()
        packed, mom = MC.run_prepared(pp, n_total, seed=2, to_host=False)     # kernel(s) + the one collective
        b.record()
        b.synchronize()
        ms.append(D.max_over_ranks(a.elapsed_time(b)))
    t = float(np.median(ms))
    sec["configs2_mc_pc"] = {"workload": "BASELINE configs[2]: 1e8-sample Monte Carlo Pc of the default pair, device-side draws, "
                                         "sgp4init + sgp4 at TCA for both objects, count + 2 x 28 moments, one packed all-gather",
                             "n_total": n_total, "scaling": "strong", "n_gpus": world, "ms": t, "samples_per_s": n_total / (t * 1e-3),
                             "sgp4_init_plus_eval_per_s": 2 * n_total / (t * 1e-3), "count": int(packed[0]), "err_count": int(packed[1]),
                             "n_ok": float(mom[0])}
    if world > 1 or rank != 0:
        return sec
    # ---- configs[3]: 30 000 objects drawn from default_prior() (seed 3) against one target, 7 days ---------------------
    cat = default_prior_catalog(30000, seed=3)
    time0 = t_tle.date_mjd - 3.5
    times = W.time_grid(time0)
    tc = np.ascontiguousarray(times[::UPSAMPLE])
    ep_c = np.full(cat.shape[1], t_tle.date_mjd)
    el_t = t_tle.elements().reshape(7, 1)
    a_c = (MU_EARTH / (cat[0] / 60.0) ** 2) ** (1.0 / 3.0)
    apo_c, per_c = a_c * (1 + cat[1]) - R_EARTH, a_c * (1 - cat[1]) - R_EARTH
    pre = ((t_tle.apogee_alt() + THR_M + 1000) < per_c) | ((apo_c + THR_M + 1000) < t_tle.perigee_alt())     # model.py:567
    c3 = {"workload": "BASELINE configs[3]: 30000 objects from default_prior() (seed 3; period < 225 min, perigee > 120 km) vs the "
                      "pair-A target, 7 days, 6000 coarse epochs -> 600000 fine points", "objects": int(cat.shape[1]),
          "rejected_by_altitude_filter": int(pre.sum())}
    for tag, sel in (("unfiltered", np.ones(cat.shape[1], bool)), ("prefiltered", ~pre)):
        el = np.ascontiguousarray(cat[:, sel])
        d_el, d_ep = torch.from_numpy(el).to(dev), torch.from_numpy(ep_c[sel]).to(dev)
        d_t, d_te, d_tc = torch.from_numpy(el_t).to(dev), torch.tensor([t_tle.date_mjd], dtype=torch.float64, device=dev), torch.from_numpy(tc).to(dev)
        t = device_ms(lambda: engine.screen_elems(d_t, d_te, d_el, d_ep, d_tc, len(times), THR_M), reps=5)
        t0 = time.perf_counter()
        out = engine.screen_host(el_t, [t_tle.date_mjd], el, ep_c[sel], tc, len(times), THR_M)
        e2e = (time.perf_counter() - t0) * 1e3
        nobj = int(sel.sum())
        fast = int(((out["err"] >> 8) == 0).sum())
        c3[tag] = {"objects": nobj, "ms": t, "objects_per_s": nobj / (t * 1e-3), "sgp4_evals_per_s": (nobj + 1) * len(tc) / (t * 1e-3),
                   "e2e_ms_pageable": e2e, "conjunctions_below_5km": int((out["i_conj"] > 0).sum()),
                   "min_d_min_m": float(np.nanmin(out["d_min"])), "objects_without_error_bits": fast}
    sec["configs3_catalog"] = c3
    # ---- K2: dsgp4.propagate / propagate_batch standalone (states written: 48 B per evaluation) ------------------------
    n_rec, m_ep = 8192, 6000
    wl = W.config2_batch(n_rec, seed=5, rank=0, state_t=st_t, state_c=st_c)
    consts, _ = engine.sgp4_init(wl["elems_t"].to(dev))
    ts = torch.from_numpy((tc - wl["epoch_t"]) * 1440.0).to(dev)
    rv = torch.empty((n_rec, m_ep, 6), dtype=torch.float64, device=dev)
    err = torch.zeros(n_rec, dtype=torch.int32, device=dev)
    from kessler_b200._native import ptr, stream_ptr
    run_k2 = lambda: NV.check(NV.lib().ksl_sgp4_prop(ptr(consts), n_rec, ptr(ts), m_ep, 0, ptr(rv), ptr(err), stream_ptr()), "ksl_sgp4_prop")
    t = device_ms(run_k2, reps=5)
    evals = n_rec * m_ep
    sec["k2_sgp4_prop"] = {"workload": f"ksl_sgp4_prop: {n_rec} records x {m_ep} epochs, full state [n][m][6] written",
                           "evals": evals, "ms": t, "evals_per_s": evals / (t * 1e-3),
                           "roofline_hbm": {"bytes_per_eval": 48, "achieved_gbs": 48 * evals / (t * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                                            "frac": 48 * evals / (t * 1e-3) / 1e9 / hbm_peak},
                           "roofline_fp64_algorithmic": {"flop_per_eval": FLOP_PER_EVAL, "achieved_tflops": FLOP_PER_EVAL * evals / (t * 1e-3) / 1e12,
                                                         "peak_tflops": peak_tf}}
    del rv
    # ---- K5: model.py:432-437 at the reference size, 1e4 x 1e4 pairs (8 non-fused FP64 operations + compare per pair) ------
    rng = np.random.default_rng(4)
    tp = torch.from_numpy(rng.standard_normal((3, 10000)) * 60.0).to(dev)
    cp = torch.from_numpy(rng.standard_normal((3, 10000)) * 60.0).to(dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    t = device_ms(lambda: engine.pc_count(tp, cp, COLLISION_M, count=cnt), reps=5, warm=2)
    pairs = 1e8
    sec["k5_pc_all_pairs"] = {"workload": "ksl_pc_count: 1e4 x 1e4 position samples, all pairs, bit-exact integer count",
                              "pairs": pairs, "ms": t, "pairs_per_s": pairs / (t * 1e-3),
                              "roofline_fp64": {"fp64_instr_per_pair": 9, "note": "3 DADD + 3 DMUL + 2 DADD + 1 DSETP, none fused (the count must "
                                                "match the reference's roundings): the pipe issues 32 lanes / 2 cycles / SMSP",
                                                "achieved_ginstr_per_s": 9 * pairs / (t * 1e-3) / 1e9,
                                                "peak_ginstr_per_s": peak_tf * 1e12 / 2 / 1e9,
                                                "frac": 9 * pairs / (t * 1e-3) / (peak_tf * 1e12 / 2)}}
    # ---- configs[4]: 1e5 synthetic CDM-sequence events (rejection sampling + the batched ground segment) -----------------
    if n_events > 0:
        import contextlib
        import io
        from kessler_b200 import events, tle as ktle
        from kessler_b200.model import ConjunctionSimplified
        pop = json.load(open(os.path.join(ROOT, "tests", "golden", "population.json")))
        tles = ktle.load_from_lines([x for p in pop for x in ("0 " + p["name"], p["line1"], p["line2"])])
        torch.manual_seed(0)
        with contextlib.redirect_stdout(io.StringIO()):
            model = ConjunctionSimplified(tles=tles, time0=60727.13899462018)
        events.generate_events(model, 64, batch_size=16384, seed=99)        # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out, screened = events.generate_events(model, int(n_events), batch_size=65536, seed=1)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        n_ev = sum(len(r["hits"]) for r in out)
        n_cdm = int(sum(r["issued"].sum() for r in out))
        feats = sum(events.feature_tensor(r)[0].shape[0] for r in out[:1])
        sec["configs4_events"] = {"workload": "BASELINE configs[4]: CDM-sequence events from the 20-TLE sample population (ConjunctionSimplified; "
                                              "model.py:554-732 batched: candidates drawn / pre-filtered / screened / compacted on the GPU, then Newton "
                                              "solve, Kepler partials, 100-sample covariance clouds and the 1e4 x 1e4 Pc count for every CDM)",
                                  "events": n_ev, "cdms": n_cdm, "candidates_screened": int(screened), "seconds": dt, "events_per_s": n_ev / dt,
                                  "cdms_per_s": n_cdm / dt, "timing": "wall clock, host included (this config is a host + device pipeline)",
                                  "lstm_feature_rows_first_pass": int(feats)}
    return sec


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
        # one launch does sgp4init (in the kernel's prologue) + 2 x 6000 SGP4 evaluations + the search per trace
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
    def make_out(pinned):
        mk = (lambda dt: torch.empty(n, dtype=dt).pin_memory().numpy()) if pinned else (lambda dt: torch.empty(n, dtype=dt).numpy())
        return dict(i_min=mk(torch.int64), d_min=mk(torch.float64), i_conj=mk(torch.int64), d_conj=mk(torch.float64), err=mk(torch.int32))

    def e2e_run(np_in, out):
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
        s = D.max_over_ranks(time.perf_counter() - t0)
        D.barrier()
        return s, r_host

    e2e_s, r_host = e2e_run((h_el_t.numpy(), h_ep_t.numpy(), h_el_c.numpy(), h_ep_c.numpy()), make_out(True))
    e2e_value = world * n * args.steps / e2e_s
    same = bool(np.array_equal(r_host["i_min"], res["i_min"].cpu().numpy()) and
                np.array_equal(r_host["i_conj"], res["i_conj"].cpu().numpy()))
    # ... and with plain (pageable) NumPy arrays, what a caller that never heard of pinned memory hands over
    pg_in = tuple(np.array(a.numpy(), copy=True) for a in (h_el_t, h_ep_t, h_el_c, h_ep_c))
    e2e_pg_s, _ = e2e_run(pg_in, make_out(False))
    h2d = (2 * 7 * 8 + 2 * 8) * n + 8 * N_COARSE
    d2h = (4 * 8 + 4) * n + 8 * (NV.SUM_NCOUNT + NV.SUM_NMOMENT)

    # ---- peaks ---------------------------------------------------------------------------------
    peak_tf = engine.fp64_peak_tflops() if rank == 0 else 0.0
    # what a pure DFMA stream sustains at the screen kernel's own shape (16 warps per SM, two dependency chains per warp)
    peak_shape = engine.fp64_peak_tflops(chains=2, warps_per_sm=16, reps=3) if rank == 0 else 0.0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    secondary = None
    if not args.no_secondary:
        secondary = secondary_workloads(engine, NV, W, D, rank, world, dev, peak_tf, hbm_peak, n_events=int(args.events))
    if rank != 0:
        D.barrier()
        return
    # ---- roofline of the dominant kernel (k_screen_warp), FP64 vector pipe ----------------------
    # frac      = EXECUTED FP64 flop (2 DFMA + DMUL + DADD per evaluation, from the committed ncu opcode mix of this very
    #             kernel) x evaluations per launch / launch time (CUDA events, this run) / measured DFMA peak (this run)
    # pipe_active = sm__pipe_fp64_cycles_active of the same capture: the honest distance to the ceiling, because DMUL /
    #             DADD / DSETP occupy a pipe slot for one flop or none
    # algorithmic_ratio = the reference algorithm's 940 flop / evaluation on the same clock: how much
    #             cheaper than that count the kernel's instruction stream is (can exceed 1)
    prof = load_profile("k_screen_warp")
    evals = n * EVALS_PER_TRACE
    exe = prof["executed_flop_per_unit"] if prof else None
    achieved_tf = (evals * exe / scr_s / 1e12) if exe else None
    roofline = {"bound": "fp64", "kernel": "k_screen_warp<false, analytic, fused init> (sgp4init + SGP4 x2 + closed-form TCA search, one warp per trace)",
                "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": (achieved_tf / peak_tf) if exe else None,
                "pipe_active": (prof["pipe_active_pct"] / 100.0) if prof and prof["pipe_active_pct"] else None,
                "issue_active": (prof["issue_active_pct"] / 100.0) if prof and prof["issue_active_pct"] else None,
                "executed_flop_per_eval": exe, "fp64_instr_per_eval": prof["fp64_instr_per_unit"] if prof else None,
                "instr_per_eval": prof["instr_per_unit"] if prof else None,
                "algorithmic_flop_per_eval": FLOP_PER_EVAL, "algorithmic_ratio": evals * FLOP_PER_EVAL / scr_s / 1e12 / peak_tf,
                "evals_per_launch": evals, "launch_ms": scr_s * 1e3,
                "traffic": (prof["dram_bytes_per_trace"] * n) if prof and prof["dram_bytes_per_trace"] else None,
                "traffic_algorithmic": n * BYTES_PER_TRACE,
                "profile_files": prof["files"] if prof else None, "profile_matches_source": prof["matches_source"] if prof else None,
                "peak_source": "measured in this run: DFMA probe, 8 independent chains per thread at 64 warps per SM "
                               "(ksl_fp64_chain_kernel; profiles/r02_fp64_ilp.json has the chains x warps table); "
                               "MEASURED_PEAKS.json has no FP64 entry",
                "peak_at_kernel_shape": peak_shape,
                # The ceiling that matters for THIS instruction stream: a DFMA with three register operands issues every third
                # cycle on B200, everything else on the FP64 pipe every second (benchmarks/fp64_ilp.py; 24.6 instead of 36.8
                # TFLOP/s for a pure stream of them).  bound = (3 n3 + 2 n_other) pipe cycles per warp-evaluation from the ncu
                # capture; measured = launch time x SM clock x schedulers / warp-evaluations of the launch.
                "operand_limited": (lambda bound, meas: {"pipe_cycles_per_warp_eval_bound": bound, "pipe_cycles_per_warp_eval_measured": meas,
                                                         "frac": bound / meas,
                                                         "dfma_three_register_operands_per_eval": prof["dfma_three_register_operands_per_eval"]})(
                    prof["operand_limited_cycles_per_warp_eval"],
                    scr_s * (clocks["sm_mhz"] or 1965.0) * 1e6 * torch.cuda.get_device_properties(dev).multi_processor_count * 4 / (evals / 32.0))
                if prof and prof.get("operand_limited_cycles_per_warp_eval") and clocks else None,
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
                       "l2": "256 MiB flush between timed steps (outside the event pairs)",
                       "parallelism": f"dp{world} (samples sharded, no data-path collective; one packed all-gather of the summary)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s * 1e3 / args.steps, "matches_device_arm": same, "host_buffers": "pinned",
                    "pageable": {"value": world * n * args.steps / e2e_pg_s, "unit": UNIT, "ms_per_step": e2e_pg_s * 1e3 / args.steps,
                                 "host_buffers": "plain NumPy arrays in and out (no pinning)"}},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "summary": {"counts": counts_h, "moments": moments_h}, "secondary": secondary}
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
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary workloads (configs[2], configs[3], configs[4], K2, K5)")
    ap.add_argument("--events", type=float, default=1e5, help="configs[4] size (events) of the secondary workload; 0 skips it")
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
