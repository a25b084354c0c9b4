#!/usr/bin/env python
"""Executed SASS instructions of one profiled kernel aggregated by opcode, from the source page of an .ncu-rep
(captured with --import-source on).  Writes the CSV that bench.py reads to derive the EXECUTED FP64 flop per unit.

    python tools/ncu_opcode_mix.py gpurun_out/prof.ncu-rep --units 2.4e8 --unit-name sgp4_eval -o profiles/<tag>_sass_opcode_mix.csv
"""
import argparse
import collections
import csv
import io
import subprocess
import sys


def opcode_counts(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr_i = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
    hdr = rows[hdr_i]
    si, ci = hdr.index("Source"), hdr.index("Instructions Executed")
    counts = collections.Counter()
    for r in rows[hdr_i + 1:]:
        if len(r) <= max(si, ci):
            continue
        text = r[si].strip()
        if text.startswith("@"):                       # predicate prefix
            text = text.split(None, 1)[1] if " " in text else text
        op = text.split()[0].split(".")[0] if text else ""
        try:
            counts[op] += int(float(r[ci]))
        except ValueError:
            continue
    return counts


def fp64_operand_shapes(rep):
    """Executed FP64-pipe instructions split by how many REGISTER source operands they read: {(opcode, n_reg): warp instructions}.
    On B200 a DFMA with three register operands issues every third cycle, everything else on the pipe every second
    (benchmarks/fp64_ilp.py), so the split decides what the pipe can deliver for a given instruction stream."""
    import re
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr_i = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
    hdr = rows[hdr_i]
    si, ci = hdr.index("Source"), hdr.index("Instructions Executed")
    out = collections.Counter()
    for r in rows[hdr_i + 1:]:
        if len(r) <= max(si, ci):
            continue
        text = re.sub(r"^@!?U?P\d+\s+", "", r[si].strip())
        if not text:
            continue
        op = text.split()[0].split(".")[0]
        if op not in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"):
            continue
        operands = [o.strip() for o in text[len(text.split()[0]):].split(",")][1:]
        n_reg = sum(1 for o in operands if re.match(r"^[-|~]*R\d+", o))
        try:
            out[(op, n_reg)] += int(float(r[ci]))
        except ValueError:
            continue
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--units", type=float, required=True, help="units (e.g. SGP4 evaluations) the profiled launch processed")
    ap.add_argument("--unit-name", default="unit")
    ap.add_argument("-o", "--out")
    a = ap.parse_args()
    counts = opcode_counts(a.rep)
    lines = [f"opcode,warp_instructions,per_{a.unit_name}_x32"]
    for op, c in sorted(counts.items(), key=lambda kv: -kv[1]):
        lines.append(f"{op},{c},{c * 32 / a.units:.2f}")
    text = "\n".join(lines) + "\n"
    if a.out:
        open(a.out, "w").write(text)
    tot = sum(counts.values())
    fp64 = {k: counts.get(k, 0) * 32 / a.units for k in ("DFMA", "DMUL", "DADD", "DSETP")}
    print(f"total {tot * 32 / a.units:.1f} thread instr / {a.unit_name}; FP64 pipe {sum(fp64.values()):.1f} "
          f"(DFMA {fp64['DFMA']:.1f} DMUL {fp64['DMUL']:.1f} DADD {fp64['DADD']:.1f} DSETP {fp64['DSETP']:.1f}); "
          f"executed flop {2 * fp64['DFMA'] + fp64['DMUL'] + fp64['DADD']:.1f}")


if __name__ == "__main__":
    sys.exit(main())
