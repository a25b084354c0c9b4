#!/usr/bin/env python
"""Executed warp instructions of one ncu capture (--import-source on) attributed to CUDA source lines.

    python tools/ncu_source_lines.py gpurun_out/prof.ncu-rep --units 1.2e10 [--top 40]

Prints file:line, instructions per unit (x32 = thread instructions), the sampled stall share and the source text."""
import argparse
import csv
import io
import os
import subprocess


def source_lines(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    out, fname, hdr = [], None, None
    for r in rows:
        if len(r) == 2 and r[0] in ("File Name", "File Path"):
            fname = r[1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            ci, wi = hdr.index("Instructions Executed"), hdr.index("# Samples")
            continue
        if hdr is None or len(r) <= ci or not r[0].strip().isdigit():
            continue
        try:
            out.append((fname, int(r[0]), r[1].strip(), int(float(r[ci])), int(float(r[wi]))))
        except ValueError:
            pass
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--units", type=float, required=True)
    ap.add_argument("--top", type=int, default=40)
    a = ap.parse_args()
    lines = source_lines(a.rep)
    tot = sum(l[3] for l in lines)
    samples = sum(l[4] for l in lines) or 1
    print(f"attributed: {tot * 32 / a.units:.1f} instructions per unit over {len(lines)} source lines")
    for f, n, text, instr, smp in sorted(lines, key=lambda l: -l[3])[:a.top]:
        print(f"{os.path.basename(f or "?")}:{n:<5d} {instr * 32 / a.units:7.2f}  {100.0 * smp / samples:5.1f}%  {text[:110]}")


if __name__ == "__main__":
    main()
