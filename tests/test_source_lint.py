"""Static checks of the CUDA sources for a defect class the GPU cannot be trusted to expose.

A warp-synchronous intrinsic with a full mask must be executed by every lane of the warp.  Written behind a short-circuit
operator or inside a conditional expression (`ok && __shfl_xor_sync(...)`), the lanes whose left-hand side decides the
result skip it: undefined behaviour that nvcc compiles harmlessly or into a hang depending on the surrounding code
(k_mc_pc did exactly that until benchmarks/fuzz_mc.py met a cloud with mixed declines inside a warp)."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SYNC = r"__(?:shfl(?:_up|_down|_xor)?|any|all|ballot|reduce_(?:add|or|and|min|max|xor)|match_any|match_all)_sync\s*\("
# the one permitted form: both operands on the left are uniform over the warp by construction (a template parameter and a
# value computed from kernel parameters only)
ALLOWED = ("kLazy && lazy && __any_sync(",)


def _statements(src):
    src = re.sub(r"//[^\n]*", "", src)
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return re.split(r"[;{}]", src)


def test_no_warp_intrinsic_behind_a_short_circuit_or_conditional_operator():
    offenders = []
    for path in sorted(glob.glob(os.path.join(ROOT, "kessler_b200", "csrc", "*.cu*"))):
        for st in _statements(open(path).read()):
            for m in re.finditer(SYNC, st):
                before = st[:m.start()]
                flat = " ".join(st.split())
                if any(a in flat for a in ALLOWED):
                    continue
                # an unparenthesised && / || / ? to the left of the call within the same statement
                if re.search(r"(&&|\|\||\?)[^,]*$", before):
                    offenders.append(f"{os.path.basename(path)}: {flat[:160]}")
    assert not offenders, "\n".join(offenders)


def test_the_lint_catches_the_defect_it_was_written_for():
    bad = "const bool ok_pair = ok && (__shfl_xor_sync(0xffffffffu, ok ? 1 : 0, 1) != 0);"
    good = "const int other = __shfl_xor_sync(0xffffffffu, ok ? 1 : 0, 1); const bool ok_pair = ok && (other != 0);"
    hit = lambda text: any(re.search(r"(&&|\|\||\?)[^,]*$", st[:m.start()]) for st in _statements(text) for m in re.finditer(SYNC, st))
    assert hit(bad) and not hit(good)
