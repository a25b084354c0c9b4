#!/usr/bin/env python
"""Turn one `ncu --set full --import-source on` capture of k_screen_warp (taken under gpurun, see profiles/README.md)
into the committed evidence bench.py reads:

    profiles/<tag>_k_screen_warp_sass_opcode_mix.csv   executed SASS by opcode, per SGP4 evaluation
    profiles/<tag>_k_screen_warp_metrics.json          time, registers, FP64-pipe / issue activity, stalls, DRAM bytes
    profiles/current.json                              manifest: which files describe the shipped kernel (+ source hash)

    python tools/make_profile_record.py gpurun_out/prof_screen_r02a.ncu-rep --tag r02a --traces 1e6 [--evals-per-trace 12000]
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    import ncu_opcode_mix
    import bench
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--tag", required=True)
    ap.add_argument("--traces", type=float, required=True)
    ap.add_argument("--evals-per-trace", type=float, default=12000)
    ap.add_argument("--kernel", default="k_screen_warp")
    ap.add_argument("--command", default="ncu --set full --clock-control none --import-source on -k regex:k_screen_warp -c 1 "
                                         "python bench.py --steps 1 --warmup 1 --no-cpu --no-secondary")
    a = ap.parse_args()
    units = a.traces * a.evals_per_trace
    counts = ncu_opcode_mix.opcode_counts(a.rep)
    mix = f"{a.tag}_{a.kernel}_sass_opcode_mix.csv"
    with open(os.path.join(ROOT, "profiles", mix), "w") as fh:
        fh.write("opcode,warp_instructions,per_sgp4_eval_x32\n")
        for op, c in sorted(counts.items(), key=lambda kv: -kv[1]):
            fh.write(f"{op},{c},{c * 32 / units:.2f}\n")
    summ = json.loads(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), a.rep], capture_output=True, text=True).stdout)
    rd, wr = summ.get("dram__bytes_read.sum", 0.0), summ.get("dram__bytes_write.sum", 0.0)
    met = {"source": a.command + "  (one B200 under gpurun)", "kernel": summ["kernel"], "traces_per_launch": int(a.traces),
           "gpu_time_ms_under_ncu": summ.get("gpu__time_duration.sum [ms]"),
           "registers_per_thread": summ.get("launch__registers_per_thread [register/thread]"),
           "block_size": summ.get("launch__block_size []"), "grid_size": summ.get("launch__grid_size []"),
           "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_total": rd + wr, "bytes_per_trace": (rd + wr) / a.traces,
           "algorithmic_bytes_per_trace": 144,
           "sm__pipe_fp64_cycles_active_pct": summ.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active [%]"),
           "smsp__issue_active_pct": summ.get("smsp__issue_active.avg.pct_of_peak_sustained_active [%]"),
           "sm__warps_active_pct": summ.get("sm__warps_active.avg.pct_of_peak_sustained_active [%]"),
           "warp_instructions": summ.get("smsp__inst_executed.sum [inst]"),
           "local_load_instructions": summ.get("smsp__inst_executed_op_local_ld.sum [inst]"),
           "local_store_instructions": summ.get("smsp__inst_executed_op_local_st.sum [inst]"),
           "stall_cycles_per_issue": summ.get("stall_cycles_per_issue"),
           "executed_flop_per_eval": (2 * counts.get("DFMA", 0) + counts.get("DMUL", 0) + counts.get("DADD", 0)) * 32 / units,
           "fp64_instr_per_eval": sum(counts.get(k, 0) for k in ("DFMA", "DMUL", "DADD", "DSETP")) * 32 / units,
           "instr_per_eval": sum(counts.values()) * 32 / units}
    shapes = ncu_opcode_mix.fp64_operand_shapes(a.rep)
    n3 = shapes.get(("DFMA", 3), 0) * 32 / units
    n_other = (sum(shapes.values()) - shapes.get(("DFMA", 3), 0)) * 32 / units
    met["dfma_three_register_operands_per_eval"] = n3
    met["fp64_other_per_eval"] = n_other
    # cycles of the FP64 pipe one warp-evaluation needs at best: 3 per three-register DFMA, 2 per anything else
    met["operand_limited_cycles_per_warp_eval"] = 3 * n3 + 2 * n_other
    metrics = f"{a.tag}_{a.kernel}_metrics.json"
    json.dump(met, open(os.path.join(ROOT, "profiles", metrics), "w"), indent=1)
    man_path = os.path.join(ROOT, "profiles", "current.json")
    man = json.load(open(man_path)) if os.path.exists(man_path) else {}
    man[a.kernel] = {"opcode_mix": mix, "metrics": metrics, "kernel_sources_sha16": bench.kernel_sources_sha(), "tag": a.tag}
    json.dump(man, open(man_path, "w"), indent=1)
    print(json.dumps(met, indent=1))


if __name__ == "__main__":
    main()
