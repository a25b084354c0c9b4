"""Generates the constants of sincos_grid (kessler_b200/csrc/sgp4_device.cuh): the reduction step 2 pi / N split into
28 + 53 bits and the interleaved, correctly rounded (sin, cos)(2 pi k / N) table.  Needs mpmath.
    python tools/make_sincos_grid.py [N=512]   -> prints the C initialisers"""
import sys

import mpmath as mp

mp.mp.prec = 300
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
twopi = 2 * mp.pi
step = twopi / N
e = mp.floor(mp.log(step, 2))
unit = mp.mpf(2) ** (e - 27)                 # 28 significant bits: fn * step_hi is exact for fn < 2^25
step_hi = mp.floor(step / unit) * unit
step_lo = step - step_hi
assert mp.mpf(float(step_hi)) == step_hi
print("// [0] N / (2 pi)  [2] step_hi  [3] step_lo")
print("%.20e, %.20e, %.20e" % (float(N / twopi), float(step_hi), float(step_lo)))
row = []
for k in range(N):
    a = twopi * k / N
    s, c = float(mp.sin(a)), float(mp.cos(a))
    s = 0.0 if abs(s) < 1e-30 else s         # exact zeros at the multiples of pi / 2
    c = 0.0 if abs(c) < 1e-30 else c
    row.append("%r, %r" % (s, c))
    if len(row) == 2:
        print("    " + ", ".join(row) + ",")
        row = []
