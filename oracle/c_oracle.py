"""ctypes front-end to oracle/_build/liboracle.so (oracle/csrc/oracle.c).

TEST INFRASTRUCTURE (see oracle/__init__.py)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

_D = ctypes.POINTER(ctypes.c_double)
_I64 = ctypes.POINTER(ctypes.c_int64)
_I32 = ctypes.POINTER(ctypes.c_int32)


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "csrc", "oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


NCONST = 36


def sgp4_init(elems):
    """elems SoA [7,n] -> consts [NCONST,n], err [n]."""
    elems = _f64(elems)
    n = elems.shape[1]
    consts = np.zeros((NCONST, n))
    err = np.zeros(n, dtype=np.int32)
    lib().orc_sgp4_init(_p(elems, _D), ctypes.c_int64(n), _p(consts, _D), _p(err, _I32))
    return consts, err


def sgp4_prop(consts, tsince, per_sample_t=False):
    consts = _f64(consts)
    tsince = _f64(tsince)
    n = consts.shape[1]
    m = tsince.shape[-1]
    rv = np.zeros((n, m, 6))
    err = np.zeros(n, dtype=np.int32)
    lib().orc_sgp4_prop(_p(consts, _D), ctypes.c_int64(n), _p(tsince, _D), ctypes.c_int64(m),
                        ctypes.c_int(1 if per_sample_t else 0), _p(rv, _D), _p(err, _I32))
    return rv, err


def upsample(x, n_out):
    x = _f64(x)
    y = np.zeros((n_out, x.shape[1]))
    lib().orc_upsample(_p(x, _D), ctypes.c_int64(x.shape[0]), ctypes.c_int64(x.shape[1]), ctypes.c_int64(n_out), _p(y, _D))
    return y


def find_conjunction(tr0, tr1, thr):
    tr0, tr1 = _f64(tr0), _f64(tr1)
    i_min, i_conj = ctypes.c_int64(), ctypes.c_int64()
    d_min, d_conj = ctypes.c_double(), ctypes.c_double()
    lib().orc_find_conjunction(_p(tr0, _D), _p(tr1, _D), ctypes.c_int64(tr0.shape[0]), ctypes.c_double(thr),
                               ctypes.byref(i_min), ctypes.byref(d_min), ctypes.byref(i_conj), ctypes.byref(d_conj))
    return i_min.value, d_min.value, i_conj.value, d_conj.value


def screen(elems_t, epoch_t, elems_c, epoch_c, times_coarse, n_fine, thr):
    """Brute-force traces.  elems_* SoA [7,n_*]; returns dict of arrays [n]."""
    elems_t, elems_c = _f64(elems_t), _f64(elems_c)
    epoch_t, epoch_c = _f64(np.atleast_1d(epoch_t)), _f64(np.atleast_1d(epoch_c))
    times_coarse = _f64(times_coarse)
    n_t, n = elems_t.shape[1], elems_c.shape[1]
    assert n_t in (1, n) and epoch_t.shape[0] == n_t and epoch_c.shape[0] == n
    i_min = np.zeros(n, dtype=np.int64)
    i_conj = np.zeros(n, dtype=np.int64)
    d_min = np.zeros(n)
    d_conj = np.zeros(n)
    err = np.zeros(n, dtype=np.int32)
    rc = lib().orc_screen(_p(elems_t, _D), _p(epoch_t, _D), ctypes.c_int64(n_t), _p(elems_c, _D), _p(epoch_c, _D),
                          ctypes.c_int64(n), _p(times_coarse, _D), ctypes.c_int64(times_coarse.shape[0]),
                          ctypes.c_int64(n_fine), ctypes.c_double(thr),
                          _p(i_min, _I64), _p(d_min, _D), _p(i_conj, _I64), _p(d_conj, _D), _p(err, _I32))
    assert rc == 0
    return dict(i_min=i_min, d_min=d_min, i_conj=i_conj, d_conj=d_conj, err=err)


def tca_states(elems, tsince):
    elems = _f64(elems)
    n = elems.shape[1]
    st = np.zeros((n, 6))
    err = np.zeros(n, dtype=np.int32)
    lib().orc_tca_states(_p(elems, _D), ctypes.c_int64(n), ctypes.c_double(tsince), _p(st, _D), _p(err, _I32))
    return st, err


def mc_cov(states_m):
    states_m = _f64(states_m).reshape(-1, 6)
    mean = np.zeros(6)
    cov = np.zeros((6, 6))
    rc = lib().orc_mc_cov(_p(states_m, _D), ctypes.c_int64(states_m.shape[0]), _p(mean, _D), _p(cov, _D))
    assert rc == 0
    return mean, cov


def pc_count(t_xyz, c_xyz, thr, pairing=0):
    """t_xyz, c_xyz SoA [3,n]."""
    t_xyz, c_xyz = _f64(t_xyz), _f64(c_xyz)
    cnt = ctypes.c_uint64()
    lib().orc_pc_count(_p(t_xyz, _D), ctypes.c_int64(t_xyz.shape[1]), _p(c_xyz, _D), ctypes.c_int64(c_xyz.shape[1]),
                       ctypes.c_double(thr), ctypes.c_int(pairing), ctypes.byref(cnt))
    return cnt.value
