"""ctypes binding of libkessler_b200.so (include/kessler_b200.h).

The CUDA library is the product: there is no CPU fallback.  Importing this module
never needs a GPU (so the host logic is testable on a CPU box), but every compute
entry point raises ``KesslerCudaError`` when the library is missing or when no
CUDA device is visible.
"""
import ctypes
import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.environ.get("KSL_LIB_PATH") or os.path.join(_PKG, "lib", "libkessler_b200.so")   # env override: tuning experiments only
SRC = os.path.join(_PKG, "csrc", "kessler_b200.cu")
HEADERS = [os.path.join(_PKG, "csrc", "sgp4_device.cuh"), os.path.join(_PKG, "csrc", "screen_device.cuh"),
           os.path.join(_PKG, "csrc", "screen_kernels.cuh"), os.path.join(_PKG, "csrc", "sincos_grid_table.cuh"),
           os.path.join(_PKG, "csrc", "kepler_device.cuh"),
           os.path.join(_ROOT, "include", "kessler_b200.h")]

NCONST = 36
NMOMENT = 28
SUM_NCOUNT = 8
SUM_NMOMENT = 4
SCREEN_ANALYTIC = 0
SCREEN_BRUTE = 1
SCREEN_FORCE_CTA = 16
SCREEN_FORCE_WARP = 32
PAIR_ALL = 0
PAIR_ELEMENTWISE = 1
ERR_ECC, ERR_MEANMOTION, ERR_SEMILATUS, ERR_DECAYED, ERR_DEEPSPACE = 1, 2, 4, 8, 16
ERR_CHASER_SHIFT = 8

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-diag-suppress", "20091",
              "-Xcompiler", "-fPIC", "-shared"]


class KesslerCudaError(RuntimeError):
    pass


def build(force=False, verbose=False):
    """Compile csrc/kessler_b200.cu for sm_100a into lib/libkessler_b200.so (in-tree)."""
    deps = [SRC] + HEADERS
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise KesslerCudaError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


_lib = None

_vp, _i64, _i32, _f64, _int = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_double, ctypes.c_int

# name -> argtypes (restype is int unless listed in _RESTYPES); mirrors include/kessler_b200.h
_SIGNATURES = {
    "ksl_version": [],
    "ksl_nconst": [],
    "ksl_last_error": [],
    "ksl_device_count": [],
    "ksl_launch_count": [],
    "ksl_sgp4_init": [_vp, _i64, _vp, _vp, _vp],
    "ksl_sgp4_prop": [_vp, _i64, _vp, _i64, _int, _vp, _vp, _vp],
    "ksl_sgp4_prop_counted": [_vp, _i64, _vp, _i64, _int, _vp, _vp, _vp, _vp],
    "ksl_upsample": [_vp, _i64, _i64, _i64, _vp, _vp],
    "ksl_propagate_upsample": [_vp, _f64, _vp, _i64, _i64, _vp, _vp, _vp, _vp],
    "ksl_find_conjunction_ws": [_i64],
    "ksl_find_conjunction": [_vp, _vp, _i64, _i64, _f64, _vp, _vp, _vp, _vp],
    "ksl_screen_ws": [_i64, _i64, _i64],
    "ksl_screen": [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _i64, _f64, _int,
                   _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "ksl_screen_elems_ws": [_i64, _i64, _i64],
    "ksl_screen_elems": [_vp, _vp, _i64, _vp, _vp, _i64, _vp, _i64, _i64, _f64, _int,
                         _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "ksl_screen_summary_ws": [_i64],
    "ksl_screen_summary": [_vp, _vp, _vp, _i64, _f64, _vp, _vp, _vp, _vp],
    "ksl_tca_moments_ws": [_i64],
    "ksl_tca_moments": [_vp, _i64, _f64, _vp, _vp, _vp, _vp, _vp, _vp],
    "ksl_pc_count": [_vp, _i64, _vp, _i64, _f64, _int, _vp, _vp],
    "ksl_mc_pc_ws": [_i64],
    "ksl_mc_pc": [_vp, _vp, _f64, _f64, _vp, _vp, _vp, _f64, _f64, _vp, ctypes.c_uint64, _i64, _i64, _f64, _f64, _vp,
                  _vp, _vp, _vp, _vp, _i64, _vp, _vp],
    "ksl_newton_elements": [_vp, _vp, _i64, _int, _f64, _vp, _vp, _vp, _vp],
    "ksl_kepler_partials": [_vp, _i64, _f64, _vp, _vp, _vp],
    "ksl_states_at_fine": [_vp, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _i64, _vp, _vp, _vp],
    "ksl_pc_mvn_batched": [_vp, _vp, _i64, _i64, ctypes.c_uint64, _i64, _f64, _vp, _vp, _vp],
    "ksl_prior_param_bytes": [],
    "ksl_prior_pack": [_int, _int, _int, _f64, _f64, _vp, _vp, _vp, _vp, _vp, _vp],
    "ksl_sample_prior": [_vp, ctypes.c_uint64, _i64, _i64, _vp, _vp, _vp, _vp],
    "ksl_screen_host": [_int, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _i64, _i64, _f64, _int, _f64,
                        _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "ksl_host_cache_free": [],
    "ksl_fp64_peak_kernel": [_int, _int, _i64, _vp, _vp],
    "ksl_fp64_chain_kernel": [_int, _int, _i64, _int, _vp, _vp],
}
_RESTYPES = {"ksl_last_error": ctypes.c_char_p, "ksl_launch_count": _i64, "ksl_find_conjunction_ws": _i64,
             "ksl_screen_ws": _i64, "ksl_screen_elems_ws": _i64, "ksl_screen_summary_ws": _i64, "ksl_tca_moments_ws": _i64, "ksl_mc_pc_ws": _i64,
             "ksl_prior_param_bytes": _i64}

EXPORTS = tuple(_SIGNATURES)


def lib():
    """The loaded library (loads on first use; never builds implicitly)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise KesslerCudaError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(kessler_b200 has no CPU fallback)")
        l = ctypes.CDLL(LIB_PATH)
        for name, args in _SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = args
            fn.restype = _RESTYPES.get(name, ctypes.c_int)
        _lib = l
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().ksl_last_error()
        raise KesslerCudaError(f"{what or 'libkessler_b200'} failed (rc={rc}): {msg.decode() if msg else ''}")


def require_cuda():
    """Raise unless a CUDA device is usable (the product path never falls back)."""
    if lib().ksl_device_count() <= 0:
        raise KesslerCudaError("no CUDA device visible: kessler_b200 runs its hot path on B200 only (no CPU fallback)")


def ptr(t):
    """Device/host pointer of a torch tensor or numpy array, or None."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return ctypes.c_void_p(t.data_ptr())
    return ctypes.c_void_p(t.ctypes.data)


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
