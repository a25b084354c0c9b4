#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of metrics profiles/ keeps.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--json out.json] [--units-per-launch N --unit-name samples]
"""
import argparse
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum",
    "smsp__inst_executed_pipe_fp64.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed_op_local_ld.sum",
    "smsp__inst_executed_op_local_st.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_alu.sum",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__thread_inst_executed.sum",
]


def to_bytes(val, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(val) * mult.get(unit, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--json")
    ap.add_argument("--units-per-launch", type=float)
    ap.add_argument("--unit-name", default="units")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        rec = {"kernel": d.get("Kernel Name", "").split("(")[0]}
        for k in KEEP:
            if k in d and d[k] != "":
                try:
                    v = float(d[k].replace(",", ""))
                except ValueError:
                    continue
                if k.startswith("dram__bytes"):
                    v = to_bytes(v, u[k])
                rec[k + ("" if k.startswith("dram__bytes") else " [" + u[k] + "]")] = v
        stalls = {}
        for k in d:
            if "issue_stalled_" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k:
                try:
                    v = float(d[k])
                except ValueError:
                    continue
                if v >= 0.05:
                    stalls[k.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")] = round(v, 3)
        rec["stall_cycles_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))
        if a.units_per_launch:
            n = a.units_per_launch
            rec["per_" + a.unit_name] = {
                "warp_instructions_x32": rec.get("smsp__inst_executed.sum [inst]", 0) * 32 / n,
                "dram_bytes": (rec.get("dram__bytes_read.sum", 0) + rec.get("dram__bytes_write.sum", 0)) / n}
        out.append(rec)
    text = json.dumps(out if len(out) != 1 else out[0], indent=1)
    if a.json:
        open(a.json, "w").write(text + "\n")
    print(text)


if __name__ == "__main__":
    sys.exit(main())
