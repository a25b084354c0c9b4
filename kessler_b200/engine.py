"""Thin torch-facing wrappers over the C ABI (include/kessler_b200.h).

torch is plumbing here: it owns device memory and the current stream; every
function below hands raw pointers to libkessler_b200.so.  All tensors are FP64
(int64/int32 for indices/flags) on the current CUDA device.  There is no CPU path:
calling any of these without a GPU raises ``KesslerCudaError``.
"""
import ctypes

import numpy as np
import torch

from . import _native as N
from ._native import KesslerCudaError, ptr, stream_ptr  # noqa: F401

F64 = torch.float64


def _dev():
    N.require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def _cf64(x, dev):
    """contiguous FP64 tensor on dev"""
    if not torch.is_tensor(x):
        x = torch.as_tensor(np.asarray(x, dtype=np.float64))
    return x.to(device=dev, dtype=F64).contiguous()


# --- dsgp4.initialize_tle -------------------------------------------------------------
def sgp4_init(elems):
    """elems SoA [7, n] (no_kozai rad/min, ecco, inclo, nodeo, argpo, mo, bstar)
    -> (consts [36, n], err [n] int32)."""
    dev = _dev()
    elems = _cf64(elems, dev)
    assert elems.dim() == 2 and elems.shape[0] == 7
    n = elems.shape[1]
    consts = torch.empty((N.NCONST, n), dtype=F64, device=dev)
    err = torch.empty(n, dtype=torch.int32, device=dev)
    N.check(N.lib().ksl_sgp4_init(ptr(elems), n, ptr(consts), ptr(err), stream_ptr()), "ksl_sgp4_init")
    return consts, err


# --- dsgp4.propagate / propagate_batch ------------------------------------------------
def sgp4_prop(consts, tsince, per_sample_t=False, err=None):
    """consts [36, n]; tsince [m] (shared) or [n, m] -> rv [n, m, 6] km, km/s; err [n]."""
    dev = _dev()
    consts = _cf64(consts, dev)
    tsince = _cf64(tsince, dev)
    n = consts.shape[1]
    m = tsince.shape[-1] if tsince.dim() > 0 else 1
    if per_sample_t:
        assert tsince.numel() == n * m
    rv = torch.empty((n, m, 6), dtype=F64, device=dev)
    if err is None:
        err = torch.zeros(n, dtype=torch.int32, device=dev)
    N.check(N.lib().ksl_sgp4_prop(ptr(consts), n, ptr(tsince), m, 1 if per_sample_t else 0, ptr(rv), ptr(err),
                                  stream_ptr()), "ksl_sgp4_prop")
    return rv, err


def sgp4_prop_counted(consts, tsince, per_sample_t=False):
    """sgp4_prop + the number of evaluations that left the fast path: (rv [n, m, 6], err [n], n_exact int)."""
    dev = _dev()
    consts = _cf64(consts, dev)
    tsince = _cf64(tsince, dev)
    n = consts.shape[1]
    m = tsince.shape[-1] if tsince.dim() > 0 else 1
    rv = torch.empty((n, m, 6), dtype=F64, device=dev)
    err = torch.zeros(n, dtype=torch.int32, device=dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    N.check(N.lib().ksl_sgp4_prop_counted(ptr(consts), n, ptr(tsince), m, 1 if per_sample_t else 0, ptr(rv), ptr(err), ptr(cnt),
                                          stream_ptr()), "ksl_sgp4_prop_counted")
    return rv, err, int(cnt[0])


# --- util.upsample ---------------------------------------------------------------------
def upsample(x, n_out):
    dev = _dev()
    x = _cf64(x, dev)
    assert x.dim() == 2
    y = torch.empty((n_out, x.shape[1]), dtype=F64, device=dev)
    N.check(N.lib().ksl_upsample(ptr(x), x.shape[0], x.shape[1], n_out, ptr(y), stream_ptr()), "ksl_upsample")
    return y


# --- util.propagate_upsample -----------------------------------------------------------
def propagate_upsample(consts1, epoch_mjd, times_sub, n_out):
    """consts1 [36] or [36,1]; times_sub = times_mjd[::upsample_factor] -> ([n_out, 6] m, m/s; err int32[1])."""
    dev = _dev()
    consts1 = _cf64(consts1, dev).reshape(-1)
    assert consts1.numel() == N.NCONST
    times_sub = _cf64(times_sub, dev)
    n_sub = times_sub.numel()
    out = torch.empty((n_out, 6), dtype=F64, device=dev)
    ws = torch.empty(n_sub * 6, dtype=F64, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    N.check(N.lib().ksl_propagate_upsample(ptr(consts1), float(epoch_mjd), ptr(times_sub), n_sub, n_out, ptr(out),
                                           ptr(ws), ptr(err), stream_ptr()), "ksl_propagate_upsample")
    return out, err


# --- model.find_conjunction ------------------------------------------------------------
def find_conjunction(tr0, tr1, thr):
    """tr0, tr1 [n, >=3] device tensors (row stride taken from the tensor).
    Returns (out_i int64[2] = i_min, i_conj|-1 ; out_d f64[2] = d_min, d_conj|nan)."""
    dev = _dev()
    tr0 = tr0.to(device=dev, dtype=F64)
    tr1 = tr1.to(device=dev, dtype=F64)
    if tr0.stride(-1) != 1 or tr0.stride(0) != tr1.stride(0) or tr1.stride(-1) != 1:
        tr0, tr1 = tr0.contiguous(), tr1.contiguous()
    n, stride = tr0.shape[0], tr0.stride(0)
    if n == 1:
        stride = max(stride, 3)
    out_i = torch.empty(2, dtype=torch.int64, device=dev)
    out_d = torch.empty(2, dtype=F64, device=dev)
    ws = torch.empty(int(N.lib().ksl_find_conjunction_ws(n)), dtype=torch.uint8, device=dev)
    N.check(N.lib().ksl_find_conjunction(ptr(tr0), ptr(tr1), n, stride, float(thr), ptr(out_i), ptr(out_d), ptr(ws),
                                         stream_ptr()), "ksl_find_conjunction")
    return out_i, out_d


# --- the fused trace -------------------------------------------------------------------
def screen(consts_t, epoch_t, consts_c, epoch_c, times_coarse, n_fine, thr_m, mode=N.SCREEN_ANALYTIC,
           init_err_t=None, init_err_c=None):
    """Returns dict(i_min, d_min, i_conj, d_conj, err) of device tensors [n]."""
    dev = _dev()
    consts_t, consts_c = _cf64(consts_t, dev), _cf64(consts_c, dev)
    if consts_t.dim() == 1:
        consts_t = consts_t.reshape(N.NCONST, 1)
    epoch_t = _cf64(epoch_t, dev).reshape(-1)
    epoch_c = _cf64(epoch_c, dev).reshape(-1)
    times_coarse = _cf64(times_coarse, dev)
    n_t, n = consts_t.shape[1], consts_c.shape[1]
    assert epoch_t.numel() == n_t and epoch_c.numel() == n and n_t in (1, n)
    out = dict(i_min=torch.empty(n, dtype=torch.int64, device=dev), d_min=torch.empty(n, dtype=F64, device=dev),
               i_conj=torch.empty(n, dtype=torch.int64, device=dev), d_conj=torch.empty(n, dtype=F64, device=dev),
               err=torch.empty(n, dtype=torch.int32, device=dev))
    n_coarse = times_coarse.numel()
    ws = torch.empty(int(N.lib().ksl_screen_ws(n_t, n, n_coarse)), dtype=torch.uint8, device=dev)
    N.check(N.lib().ksl_screen(ptr(consts_t), ptr(epoch_t), ptr(init_err_t), n_t, ptr(consts_c), ptr(epoch_c),
                               ptr(init_err_c), n, ptr(times_coarse), n_coarse, int(n_fine), float(thr_m), int(mode),
                               ptr(out["i_min"]), ptr(out["d_min"]), ptr(out["i_conj"]), ptr(out["d_conj"]),
                               ptr(out["err"]), ptr(ws), stream_ptr()), "ksl_screen")
    return out


def screen_elems(elems_t, epoch_t, elems_c, epoch_c, times_coarse, n_fine, thr_m, mode=N.SCREEN_ANALYTIC):
    """The fused trace from the SGP4 inputs (sgp4init included): elems_t [7, n_t], elems_c [7, n] device tensors.
    Returns dict(i_min, d_min, i_conj, d_conj, err) of device tensors [n]."""
    dev = _dev()
    elems_t, elems_c = _cf64(elems_t, dev), _cf64(elems_c, dev)
    if elems_t.dim() == 1:
        elems_t = elems_t.reshape(7, 1)
    epoch_t = _cf64(epoch_t, dev).reshape(-1)
    epoch_c = _cf64(epoch_c, dev).reshape(-1)
    times_coarse = _cf64(times_coarse, dev)
    n_t, n = elems_t.shape[1], elems_c.shape[1]
    assert elems_t.shape[0] == 7 and elems_c.shape[0] == 7 and epoch_t.numel() == n_t and epoch_c.numel() == n and n_t in (1, n)
    out = dict(i_min=torch.empty(n, dtype=torch.int64, device=dev), d_min=torch.empty(n, dtype=F64, device=dev),
               i_conj=torch.empty(n, dtype=torch.int64, device=dev), d_conj=torch.empty(n, dtype=F64, device=dev),
               err=torch.empty(n, dtype=torch.int32, device=dev))
    n_coarse = times_coarse.numel()
    ws = torch.empty(int(N.lib().ksl_screen_elems_ws(n_t, n, n_coarse)), dtype=torch.uint8, device=dev)
    N.check(N.lib().ksl_screen_elems(ptr(elems_t), ptr(epoch_t), n_t, ptr(elems_c), ptr(epoch_c), n, ptr(times_coarse), n_coarse,
                                     int(n_fine), float(thr_m), int(mode), ptr(out["i_min"]), ptr(out["d_min"]), ptr(out["i_conj"]),
                                     ptr(out["d_conj"]), ptr(out["err"]), ptr(ws), stream_ptr()), "ksl_screen_elems")
    return out


def screen_summary(res, collision_thr_m):
    """res = dict from screen(); returns (counts uint64-as-int64 [8], moments f64 [4]) device tensors."""
    dev = _dev()
    n = res["i_min"].numel()
    counts = torch.zeros(N.SUM_NCOUNT, dtype=torch.int64, device=dev)
    moments = torch.zeros(N.SUM_NMOMENT, dtype=F64, device=dev)
    ws = torch.empty(int(N.lib().ksl_screen_summary_ws(n)), dtype=torch.uint8, device=dev)
    N.check(N.lib().ksl_screen_summary(ptr(res["i_conj"]), ptr(res["d_min"]), ptr(res["err"]), n, float(collision_thr_m),
                                       ptr(counts), ptr(moments), ptr(ws), stream_ptr()), "ksl_screen_summary")
    return counts, moments


# --- Monte Carlo covariance at TCA -------------------------------------------------------
def tca_moments(elems, tsince, ref_state_m, want_states=False):
    """elems SoA [7, n] -> (moments f64[28], err_count int64[1], states [n,6] m or None)."""
    dev = _dev()
    elems = _cf64(elems, dev)
    n = elems.shape[1]
    ref = _cf64(ref_state_m, dev).reshape(-1)
    assert ref.numel() == 6
    moments = torch.empty(N.NMOMENT, dtype=F64, device=dev)
    errc = torch.zeros(1, dtype=torch.int64, device=dev)
    states = torch.empty((n, 6), dtype=F64, device=dev) if want_states else None
    ws = torch.empty(int(N.lib().ksl_tca_moments_ws(n)), dtype=torch.uint8, device=dev)
    N.check(N.lib().ksl_tca_moments(ptr(elems), n, float(tsince), ptr(ref), ptr(states), ptr(moments), ptr(errc), ptr(ws),
                                    stream_ptr()), "ksl_tca_moments")
    return moments, errc, states


# --- Monte Carlo collision count ----------------------------------------------------------
def pc_count(t_xyz, c_xyz, thr_m, pairing=N.PAIR_ALL, count=None):
    """t_xyz SoA [3, nt], c_xyz SoA [3, nc] metres -> int64[1] device tensor (accumulated into `count`)."""
    dev = _dev()
    t_xyz, c_xyz = _cf64(t_xyz, dev), _cf64(c_xyz, dev)
    assert t_xyz.shape[0] == 3 and c_xyz.shape[0] == 3
    if count is None:
        count = torch.zeros(1, dtype=torch.int64, device=dev)
    N.check(N.lib().ksl_pc_count(ptr(t_xyz), t_xyz.shape[1], ptr(c_xyz), c_xyz.shape[1], float(thr_m), int(pairing),
                                 ptr(count), stream_ptr()), "ksl_pc_count")
    return count


# --- scale-out Monte Carlo collision probability (device-side draws) ---------------------------
def mc_pc_pack(mean_t, chol_t, ref_t, mean_c, chol_c, ref_c):
    """The host-side parameter block of mc_pc, converted once (contiguous float64: means [6], lower triangles [21], reference
    states [6]); pass the result as `packed=` to skip the per-call conversions."""
    host = lambda a, k: np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1)[:k])
    tril = lambda L: np.ascontiguousarray(np.asarray(L, dtype=np.float64)[np.tril_indices(6)])
    return (host(mean_t, 6), tril(chol_t), host(ref_t, 6), host(mean_c, 6), tril(chol_c), host(ref_c, 6))


def mc_pc(mean_t, chol_t, bstar_t, tsince_t, ref_t, mean_c, chol_c, bstar_c, tsince_c, ref_c, seed, sample_offset, n,
          thr_m, mu_m3s2=398600.5e9, n_dump=0, count=None, err_count=None, packed=None, scratch=None):
    """ksl_mc_pc.  mean_* [6] = (a [m], e, i, raan, argp, M); chol_* = lower-triangular Cholesky factor [6,6]
    of the element covariance; ref_* [6] reference states in metres.  Returns dict(count int64[1] (accumulated),
    moments_t/moments_c f64[28], err_count int64[1], dump [n_dump, 26] or None) of device tensors."""
    dev = _dev()
    mean_t, lt, ref_t, mean_c, lc, ref_c = packed if packed is not None else mc_pc_pack(mean_t, chol_t, ref_t, mean_c, chol_c, ref_c)
    if count is None:
        count = torch.zeros(1, dtype=torch.int64, device=dev)
    if err_count is None:
        err_count = torch.zeros(1, dtype=torch.int64, device=dev)
    moments = torch.empty(2 * N.NMOMENT, dtype=F64, device=dev)
    dump = torch.empty((int(n_dump), 26), dtype=F64, device=dev) if n_dump else None
    need = int(N.lib().ksl_mc_pc_ws(int(n)))
    if scratch is not None and scratch.get("ws") is not None and scratch["ws"].numel() >= need and scratch["ws"].device == dev:
        ws = scratch["ws"]                      # caller-held workspace: repeated launches of one size allocate nothing
    else:
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        if scratch is not None:
            scratch["ws"] = ws
    N.check(N.lib().ksl_mc_pc(ptr(mean_t), ptr(lt), float(bstar_t), float(tsince_t), ptr(ref_t), ptr(mean_c), ptr(lc),
                              float(bstar_c), float(tsince_c), ptr(ref_c), int(seed), int(sample_offset), int(n), float(thr_m),
                              float(mu_m3s2), ptr(count), ptr(moments), ptr(moments[N.NMOMENT:]), ptr(err_count), ptr(dump),
                              int(n_dump), ptr(ws), stream_ptr()), "ksl_mc_pc")
    return dict(count=count, moments_t=moments[:N.NMOMENT], moments_c=moments[N.NMOMENT:], err_count=err_count, dump=dump)


# --- dsgp4.newton_method / kessler.util Kepler partials, batched ---------------------------------
def newton_elements(target_km, bstar, max_iter=50, tol=1e-12):
    """target_km [n, 6] (km, km/s), bstar [n] -> (y [n, 6] mean elements, resid [n], iters [n] int32)."""
    dev = _dev()
    target_km = _cf64(target_km, dev).reshape(-1, 6)
    n = target_km.shape[0]
    bstar = _cf64(bstar, dev).reshape(-1)
    assert bstar.numel() == n
    y = torch.empty((n, 6), dtype=F64, device=dev)
    resid = torch.empty(n, dtype=F64, device=dev)
    iters = torch.empty(n, dtype=torch.int32, device=dev)
    N.check(N.lib().ksl_newton_elements(ptr(target_km), ptr(bstar), n, int(max_iter), float(tol), ptr(y), ptr(resid), ptr(iters),
                                        stream_ptr()), "ksl_newton_elements")
    return y, resid, iters


def kepler_partials(states, mu):
    """states [n, 6] -> (elements [n, 6], jac [n, 6, 6])."""
    dev = _dev()
    states = _cf64(states, dev).reshape(-1, 6)
    n = states.shape[0]
    el = torch.empty((n, 6), dtype=F64, device=dev)
    jac = torch.empty((n, 6, 6), dtype=F64, device=dev)
    N.check(N.lib().ksl_kepler_partials(ptr(states), n, float(mu), ptr(el), ptr(jac), stream_ptr()), "ksl_kepler_partials")
    return el, jac


# --- batched ground segment helpers --------------------------------------------------------------
def states_at_fine(consts, rec, epoch, fine_idx, times_coarse, n_fine):
    """consts [36, n_rec]; rec, epoch, fine_idx [n_items] -> (states [n_items, 6] m, err [n_items] int32)."""
    dev = _dev()
    consts = _cf64(consts, dev)
    rec = torch.as_tensor(np.asarray(rec, dtype=np.int64)).to(dev)
    fine_idx = torch.as_tensor(np.asarray(fine_idx, dtype=np.int64)).to(dev)
    epoch = _cf64(epoch, dev).reshape(-1)
    times_coarse = _cf64(times_coarse, dev)
    n_items = rec.numel()
    out = torch.empty((n_items, 6), dtype=F64, device=dev)
    err = torch.empty(n_items, dtype=torch.int32, device=dev)
    N.check(N.lib().ksl_states_at_fine(ptr(consts), consts.shape[1], ptr(rec), ptr(epoch), ptr(fine_idx), n_items, ptr(times_coarse),
                                       times_coarse.numel(), int(n_fine), ptr(out), ptr(err), stream_ptr()), "ksl_states_at_fine")
    return out, err


def pc_mvn_batched(mean, chol, n_pts, thr_m, seed=0, cdm_id0=0, return_points=False):
    """mean [n_cdm, 2, 3], chol [n_cdm, 2, 3, 3] (lower-triangular) -> counts int64 [n_cdm] (+ points [n_cdm,2,3,n_pts])."""
    dev = _dev()
    mean = _cf64(mean, dev).reshape(-1, 2, 3)
    n_cdm = mean.shape[0]
    chol = np.asarray(chol.cpu() if torch.is_tensor(chol) else chol, dtype=np.float64).reshape(n_cdm, 2, 3, 3)
    tri = np.ascontiguousarray(chol[:, :, [0, 1, 1, 2, 2, 2], [0, 0, 1, 0, 1, 2]])
    tri_d = _cf64(tri, dev)
    counts = torch.zeros(n_cdm, dtype=torch.int64, device=dev)
    done = 0
    pts_all = []
    while done < n_cdm:                       # bounded workspace: <= 4096 CDMs per launch
        m = min(4096, n_cdm - done)
        ws_flat = torch.empty(m * 2 * (3 * int(n_pts) + 6), dtype=F64, device=dev)
        ws = ws_flat[:m * 2 * 3 * int(n_pts)].view(m, 2, 3, int(n_pts))
        N.check(N.lib().ksl_pc_mvn_batched(ptr(mean[done:done + m].contiguous()), ptr(tri_d[done:done + m].contiguous()), m, int(n_pts),
                                           int(seed), int(cdm_id0) + done, float(thr_m), ptr(counts[done:]), ptr(ws_flat), stream_ptr()),
                "ksl_pc_mvn_batched")
        if return_points:
            pts_all.append(ws)
        done += m
    return (counts, torch.cat(pts_all) if pts_all else None) if return_points else counts


# --- prior sampling on the device --------------------------------------------------------------------
PRIOR_FIELDS = ("mean_motion", "mean_anomaly", "eccentricity", "inclination", "argument_of_perigee", "raan",
                "mean_motion_first_derivative", "b_star")


def pack_prior(fields):
    """fields: list of 8 dicts (PRIOR_FIELDS order) with kind 0 {'lo','hi'}, kind 1 {'probs','loc','scale'} or kind 2
    {'probs','loc','scale','lo','hi'}.  Returns the packed parameter block as a uint8 device tensor."""
    dev = _dev()
    nbytes = int(N.lib().ksl_prior_param_bytes())
    host = np.zeros(nbytes, dtype=np.uint8)
    arr = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1))
    for f, d in enumerate(fields):
        kind = int(d["kind"])
        if kind == 0:
            N.check(N.lib().ksl_prior_pack(f, 0, 0, float(d["lo"]), float(d["hi"]), None, None, None, None, None, ptr(host)), "ksl_prior_pack")
            continue
        probs, loc, scale = arr(d["probs"]), arr(d["loc"]), arr(d["scale"])
        lo, hi = float(d.get("lo", -np.inf)), float(d.get("hi", np.inf))
        cdf_a = cdf_b = None
        if kind == 2:
            nrm = torch.distributions.Normal(torch.zeros((), dtype=F64), torch.ones((), dtype=F64))
            cdf_a = arr(nrm.cdf(torch.from_numpy((lo - loc) / scale)).numpy())
            cdf_b = arr(nrm.cdf(torch.from_numpy((hi - loc) / scale)).numpy())
        N.check(N.lib().ksl_prior_pack(f, kind, len(probs), lo, hi, ptr(probs), ptr(loc), ptr(scale), ptr(cdf_a), ptr(cdf_b), ptr(host)),
                "ksl_prior_pack")
    return torch.from_numpy(host).to(dev)


def sample_prior(params, seed, sample_offset, n):
    """-> (raw [8, n], elems [7, n], flag [n] int32) device tensors."""
    dev = _dev()
    raw = torch.empty((len(PRIOR_FIELDS), int(n)), dtype=F64, device=dev)
    elems = torch.empty((7, int(n)), dtype=F64, device=dev)
    flag = torch.empty(int(n), dtype=torch.int32, device=dev)
    N.check(N.lib().ksl_sample_prior(ptr(params), int(seed), int(sample_offset), int(n), ptr(raw), ptr(elems), ptr(flag), stream_ptr()),
            "ksl_sample_prior")
    return raw, elems, flag


# --- end-to-end: host buffers in / out -----------------------------------------------------
def screen_host(elems_t, epoch_t, elems_c, epoch_c, times_coarse, n_fine, thr_m, mode=N.SCREEN_ANALYTIC,
                collision_thr_m=70.0, device=None, out=None, want_summary=True):
    """numpy (or pinned torch CPU) arrays in, numpy arrays out; all transfers inside."""
    N.require_cuda()
    if device is None:
        device = torch.cuda.current_device()
    as_np = lambda a: np.ascontiguousarray(a.numpy() if torch.is_tensor(a) else a, dtype=np.float64)
    elems_t, elems_c = as_np(elems_t), as_np(elems_c)
    if elems_t.ndim == 1:
        elems_t = elems_t.reshape(7, 1)
    epoch_t = as_np(np.atleast_1d(epoch_t))
    epoch_c = as_np(np.atleast_1d(epoch_c))
    times_coarse = as_np(times_coarse)
    n_t, n = elems_t.shape[1], elems_c.shape[1]
    if out is None:
        out = dict(i_min=np.empty(n, dtype=np.int64), d_min=np.empty(n), i_conj=np.empty(n, dtype=np.int64),
                   d_conj=np.empty(n), err=np.empty(n, dtype=np.int32))
    counts = np.zeros(N.SUM_NCOUNT, dtype=np.uint64) if want_summary else None
    moments = np.zeros(N.SUM_NMOMENT) if want_summary else None
    N.check(N.lib().ksl_screen_host(int(device), ptr(elems_t), ptr(epoch_t), n_t, ptr(elems_c), ptr(epoch_c), n,
                                    ptr(times_coarse), times_coarse.size, int(n_fine), float(thr_m), int(mode),
                                    float(collision_thr_m), ptr(out["i_min"]), ptr(out["d_min"]), ptr(out["i_conj"]),
                                    ptr(out["d_conj"]), ptr(out["err"]), ptr(counts), ptr(moments)), "ksl_screen_host")
    out = dict(out)
    out["counts"], out["moments"] = counts, moments
    return out


# --- measurement ---------------------------------------------------------------------------
def fp64_peak_tflops(iters=20000, reps=5, chains=8, warps_per_sm=64, mode=0):
    """Measured DFMA throughput of the current device, TFLOP/s: `chains` independent chains per thread at `warps_per_sm`
    resident warps (defaults: the shape that saturates the pipe, 36.8 of the 37.2 TFLOP/s a B200 can issue at 1965 MHz --
    benchmarks/fp64_ilp.py has the whole table).  mode: operand shape of the chain (0 constants, 1 three register operands, see
    include/kessler_b200.h).  Best of `reps`."""
    dev = _dev()
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    threads = min(1024, 32 * int(warps_per_sm))
    blocks = sms * max(1, (32 * int(warps_per_sm)) // threads)
    sink = torch.zeros(1, dtype=F64, device=dev)
    run = lambda it: N.check(N.lib().ksl_fp64_chain_kernel(blocks, threads, it, int(chains) + 16 * int(mode), ptr(sink), stream_ptr()), "fp64_peak")
    run(1000)
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(iters)
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1)
        flops = 2.0 * int(chains) * iters * blocks * threads
        best = max(best, flops / (ms * 1e-3) / 1e12)
    return best


def launch_count():
    return int(N.lib().ksl_launch_count())
