"""The C-ABI library loads on a CPU-only machine and exports every symbol that
include/kessler_b200.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

from kessler_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "kessler_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ksl_[a-z0-9_]+)\s*\(", txt)))


def test_library_is_built_in_tree():
    assert os.path.exists(N.LIB_PATH), "run __graft_entry__.build() first"
    assert os.path.commonpath([ROOT, N.LIB_PATH]) == ROOT


def test_exports_every_declared_symbol():
    names = _declared()
    assert len(names) >= 20
    raw = ctypes.CDLL(N.LIB_PATH)
    for nm in names:
        assert hasattr(raw, nm), f"{nm} declared in kessler_b200.h but not exported"
    assert sorted(N.EXPORTS) == names, "ctypes signatures out of sync with the header"


def test_metadata_calls_and_sm100a_cubin():
    lib = N.lib()
    assert lib.ksl_version() == 200
    assert lib.ksl_nconst() == N.NCONST == 36
    assert lib.ksl_device_count() >= 0
    out = os.popen(f"cuobjdump -lelf {N.LIB_PATH} 2>/dev/null").read()
    if out:
        assert "sm_100a" in out


def test_compute_entry_points_refuse_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from kessler_b200 import engine
    with pytest.raises(N.KesslerCudaError):
        engine.sgp4_init([[0.06]] * 7)
    import numpy as np
    with pytest.raises(N.KesslerCudaError):
        engine.screen_host(np.zeros((7, 1)), [0.0], np.zeros((7, 1)), [0.0], np.zeros(4), 16, 5e3)


def test_bad_arguments_return_error_codes():
    lib = N.lib()
    assert lib.ksl_sgp4_init(None, 5, None, None, None) == -1
    assert b"ksl_sgp4_init" in lib.ksl_last_error()
    assert lib.ksl_screen(None, None, None, 2, None, None, None, 3, None, 10, 100, 5e3, 0,
                          None, None, None, None, None, None, None) == -1
    assert lib.ksl_pc_count(None, -1, None, 0, 70.0, 0, None, None) == -1
