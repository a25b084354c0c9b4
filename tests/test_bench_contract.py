"""bench.py contract on a CPU-only box: the reference arm runs (oracle port on the host cores) and prints one
JSON line with the keys the driver reads; the GPU arm refuses loudly without a GPU."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "traces/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["config"]["n_coarse"] == 6000 and line["config"]["n_fine"] == 600000
    assert line["sgp4_evals_per_s"] == pytest.approx(line["value"] * 12000)


def test_gpu_arm_fails_loudly_without_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--samples", "100"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode != 0
    assert "no CUDA device" in out.stderr or "KesslerCudaError" in out.stderr


def test_roofline_evidence_is_consistent():
    """bench.py derives the executed-flop figure of its roofline record from the committed ncu opcode mix
    (profiles/current.json -> CSV + metrics JSON).  The three must agree: the CSV's DFMA/DMUL/DADD rows reproduce the
    metrics file's executed_flop_per_eval, its per-evaluation column is warp_instructions x 32 / evaluations of the
    profiled launch, the traffic figures add up, and the fractions are physical."""
    sys.path.insert(0, ROOT)
    import bench
    man = json.load(open(os.path.join(ROOT, "profiles", "current.json")))["k_screen_warp"]
    prof = bench.load_profile("k_screen_warp")
    assert prof is not None and prof["files"] == [man["opcode_mix"], man["metrics"]]
    met = json.load(open(os.path.join(ROOT, "profiles", man["metrics"])))
    rows = [ln.split(",") for ln in open(os.path.join(ROOT, "profiles", man["opcode_mix"])).read().strip().splitlines()[1:]]
    evals = met["traces_per_launch"] * bench.EVALS_PER_TRACE
    for op, warp_instr, per_eval in rows:
        assert float(per_eval) == pytest.approx(int(warp_instr) * 32 / evals, abs=0.006), op
    assert prof["executed_flop_per_unit"] == pytest.approx(met["executed_flop_per_eval"], abs=0.05)
    assert prof["fp64_instr_per_unit"] == pytest.approx(met["fp64_instr_per_eval"], abs=0.05)
    assert sum(int(r[1]) for r in rows) == pytest.approx(met["warp_instructions"], rel=1e-3)
    assert met["dram_bytes_total"] == pytest.approx(met["dram_bytes_read"] + met["dram_bytes_write"])
    assert met["bytes_per_trace"] == pytest.approx(met["dram_bytes_total"] / met["traces_per_launch"])
    assert 0 < met["sm__pipe_fp64_cycles_active_pct"] <= 100 and 0 < met["smsp__issue_active_pct"] <= 100
    # physical: the FP64 pipe retires at most one DFMA (2 flop) per lane per slot -> executed flop <= 2 x FP64 instructions,
    # and the pipe's busy share implied by the instruction count cannot exceed what the counter says by more than rounding
    assert prof["executed_flop_per_unit"] <= 2 * prof["fp64_instr_per_unit"]
    cycles = met["gpu_time_ms_under_ncu"] * 1e-3 * 1.965e9 * 148 * 4                     # SMSP cycles of the launch
    implied = 2 * prof["fp64_instr_per_unit"] * evals / 32 / cycles * 100                # 2 cycles per warp instruction
    assert implied == pytest.approx(met["sm__pipe_fp64_cycles_active_pct"], rel=0.08)
    # does the record describe the kernel sources that are checked in?  bench.py reports this as roofline.profile_matches_source;
    # a stale capture is a warning here, not a failure (the numbers above are still self-consistent)
    if man["kernel_sources_sha16"] != bench.kernel_sources_sha():
        import warnings
        warnings.warn("profiles/current.json was captured from other kernel sources: re-capture (tools/make_profile_record.py)")
