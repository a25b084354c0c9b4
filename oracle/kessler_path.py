"""CPU restatement of Kessler's own code on the Monte Carlo hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Each function cites the
reference lines it follows (/root/reference/kessler/...).  The structure
mirrors the reference on purpose -- one trace at a time, SGP4 on a [6000]
vector, ``torch.nn.functional.interpolate`` to the fine grid, three passes of
``find_conjunction`` -- because this module is also the timed "port" CPU
baseline of bench.py.
"""
import math

import numpy as np
import torch

from . import sgp4_ref as S

MU_EARTH_SI = 398600.5e9  # model.py:571-572: float(mu)*1e9, m^3/s^2


# --------------------------------------------------------------------------
# time grid (model.py:574) and coarse sub-sampling (util.py:573-575)
# --------------------------------------------------------------------------
def time_grid(time0, max_duration_days=7.0, time_resolution=6e5):
    delta = max_duration_days / time_resolution          # model.py:208
    return np.arange(time0, time0 + max_duration_days, delta)   # model.py:574


def coarse_tsince(times_mjd, epoch_mjd, upsample_factor):
    """util.py:573-575: every ``upsample_factor``-th epoch, minutes since the
    TLE epoch."""
    take_every = max(upsample_factor, 1)
    sub = np.asarray(times_mjd[::take_every], dtype=np.float64)
    return (sub - epoch_mjd) * 1440.0


# --------------------------------------------------------------------------
# linear up-sampling (util.py:541-555)
# --------------------------------------------------------------------------
def upsample_torch(s, target_resolution):
    """Literally util.upsample: F.interpolate(mode='linear', align_corners=True)."""
    s = torch.as_tensor(s).transpose(0, 1)
    s = torch.nn.functional.interpolate(s.unsqueeze(0), size=(target_resolution), mode="linear", align_corners=True)
    return s.squeeze(0).transpose(0, 1)


def upsample_index_weights(n_in, n_out):
    """Index/weight arithmetic of torch's CPU ``upsample_linear1d`` with
    align_corners=True for an FP64 input: the scale
    and source coordinate are FP64, but the integer part is taken through a
    FLOAT32 cast.  Returns (i0, i1, w0, w1) for every output index.

    [verified here against torch 2.11 CPU by tests/test_oracle_golden.py]"""
    j = np.arange(n_out, dtype=np.float64)
    scale = (n_in - 1) / (n_out - 1) if n_out > 1 else 0.0
    src = j * scale
    i0 = np.floor(src.astype(np.float32)).astype(np.int64)
    i0 = np.minimum(i0, n_in - 1)
    i1 = np.minimum(i0 + 1, n_in - 1)
    w1 = np.clip(src - i0, 0.0, 1.0)
    w0 = 1.0 - w1
    return i0, i1, w0, w1


def upsample_numpy(x, n_out):
    """x [n_in, c] -> [n_out, c], same arithmetic as upsample_torch."""
    x = np.asarray(x, dtype=np.float64)
    i0, i1, w0, w1 = upsample_index_weights(x.shape[0], n_out)
    return w0[:, None] * x[i0] + w1[:, None] * x[i1]


# --------------------------------------------------------------------------
# propagate_upsample (util.py:557-580)
# --------------------------------------------------------------------------
def propagate_upsample(consts, epoch_mjd, times_mjd, upsample_factor=1, use_torch=True):
    """Returns ([N,2,3] metres & m/s, err bitmask OR-ed over the epochs)."""
    if upsample_factor == 1:
        ts = (np.asarray(times_mjd, dtype=np.float64) - epoch_mjd) * 1440.0
        r, v, err = S.sgp4_propagate(consts, ts)
        return np.stack([r, v], axis=1) * 1e3, int(np.bitwise_or.reduce(err.reshape(-1))) if err.size else 0
    ts = coarse_tsince(times_mjd, epoch_mjd, upsample_factor)
    r, v, err = S.sgp4_propagate(consts, ts)
    ret = np.concatenate([r, v], axis=-1)                 # [n_coarse, 6] km, km/s
    if use_torch:
        ret = upsample_torch(torch.from_numpy(ret), len(times_mjd)).numpy()
    else:
        ret = upsample_numpy(ret, len(times_mjd))
    return ret.reshape(ret.shape[0], 2, 3) * 1e3, int(np.bitwise_or.reduce(err.reshape(-1)))


# --------------------------------------------------------------------------
# find_conjunction (model.py:23-52)
# --------------------------------------------------------------------------
def find_conjunction(tr0, tr1, miss_dist_threshold):
    tr0 = torch.as_tensor(tr0)
    tr1 = torch.as_tensor(tr1)
    d = tr0 - tr1
    squared_norm = torch.einsum("ix,ix->i", d, d)
    i_min = int(torch.argmin(squared_norm))
    d_min = squared_norm[i_min].sqrt()
    below = (squared_norm < miss_dist_threshold ** 2).nonzero().flatten()
    if below.nelement() == 0:
        return i_min, float(d_min), None, None
    i_conj = int(below[0])
    return i_min, float(d_min), i_conj, float(squared_norm[i_conj].sqrt())


def find_conjunction_numpy(tr0, tr1, miss_dist_threshold):
    """Same semantics without torch (NaN counts as the minimum, first index
    wins ties, as torch.argmin)."""
    d = np.asarray(tr0) - np.asarray(tr1)
    sq = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    nan = np.isnan(sq)
    i_min = int(np.argmax(nan)) if nan.any() else int(np.argmin(sq))
    below = np.nonzero(sq < miss_dist_threshold ** 2)[0]
    if below.size == 0:
        return i_min, float(np.sqrt(sq[i_min])), None, None
    i_conj = int(below[0])
    return i_min, float(np.sqrt(sq[i_min])), i_conj, float(np.sqrt(sq[i_conj]))


# --------------------------------------------------------------------------
# one trace of the space segment of Conjunction.forward (model.py:567-624)
# --------------------------------------------------------------------------
def altitude_prefilter(sma_t_m, ecc_t, sma_c_m, ecc_c, miss_dist_threshold):
    """model.py:567 -- True when the pair is rejected.  apogee/perigee_alt are
    dsgp4.TLE methods: a(1+e)-R and a(1-e)-R in metres with
    a = (mu/n^2)^(1/3)."""
    re_m = S.RE * 1e3
    apo_t, per_t = sma_t_m * (1 + ecc_t) - re_m, sma_t_m * (1 - ecc_t) - re_m
    apo_c, per_c = sma_c_m * (1 + ecc_c) - re_m, sma_c_m * (1 - ecc_c) - re_m
    return ((apo_t + miss_dist_threshold + 1000) < per_c) | ((apo_c + miss_dist_threshold + 1000) < per_t)


def screen_trace(el_t, epoch_t_mjd, el_c, epoch_c_mjd, times_mjd, upsample_factor=100,
                 miss_dist_threshold=5e3, use_torch=True):
    """el_* = (no_kozai, ecco, inclo, nodeo, argpo, mo, bstar).  Returns a dict
    with i_min, d_min, i_conj (-1 = None), d_conj (nan = None), err_t, err_c.
    Deep-space (init raises in dsgp4) -> prop error, model.py:580-599."""
    ct, e0 = S.sgp4_init(*el_t)
    cc, e1 = S.sgp4_init(*el_c)
    if int(e0) & S.ERR_DEEPSPACE or int(e1) & S.ERR_DEEPSPACE:
        return dict(i_min=-1, d_min=math.nan, i_conj=-1, d_conj=math.nan, err_t=int(e0), err_c=int(e1))
    st, et = propagate_upsample(ct, epoch_t_mjd, times_mjd, upsample_factor, use_torch)
    sc, ec = propagate_upsample(cc, epoch_c_mjd, times_mjd, upsample_factor, use_torch)
    fc = find_conjunction if use_torch else find_conjunction_numpy
    i_min, d_min, i_conj, d_conj = fc(st[:, 0], sc[:, 0], miss_dist_threshold)
    return dict(i_min=i_min, d_min=d_min, i_conj=-1 if i_conj is None else i_conj,
                d_conj=math.nan if d_conj is None else d_conj, err_t=et, err_c=ec)


def cdm_schedule(times, time0, time_min, cdm_update_every_hours=8.0):
    """model.py:632-638: max_ncdm and the snapped CDM issue times."""
    max_ncdm = max(1, int((time_min - time0) * 24.0 / cdm_update_every_hours))
    idx, tt = [], []
    for i in range(max_ncdm):
        t = time0 + (i * cdm_update_every_hours) / 24
        k = int(np.argmin(abs(times - t)))                # util.py:527-539
        idx.append(k)
        tt.append(float(times[k]))
    return max_ncdm, idx, tt


# --------------------------------------------------------------------------
# RTN rotation (util.py:304-341)
# --------------------------------------------------------------------------
def rotation_matrix(state):
    r, v = state[0], state[1]
    u = r / np.linalg.norm(r)
    w = np.cross(r, v)
    w = w / np.linalg.norm(w)
    t = np.cross(w, u)
    return np.vstack((u, t, w))


def cartesian_to_keplerian(r_vec, v_vec, mu):
    """util.py:187-264 in NumPy (elliptic branch): (a, e, i, Omega, omega, M)."""
    r_vec, v_vec = np.asarray(r_vec, dtype=np.float64), np.asarray(v_vec, dtype=np.float64)
    r, v = np.linalg.norm(r_vec), np.linalg.norm(v_vec)
    h_vec = np.cross(r_vec, v_vec)
    h = np.linalg.norm(h_vec)
    i = np.arccos(h_vec[2] / h) if h != 0 else 0.0
    n_vec = np.cross(np.array([0.0, 0.0, 1.0]), h_vec)
    n = np.linalg.norm(n_vec)
    e_vec = (1 / mu) * ((v ** 2 - mu / r) * r_vec - np.dot(r_vec, v_vec) * v_vec)
    e = np.linalg.norm(e_vec)
    energy = v ** 2 / 2 - mu / r
    a = -mu / (2 * energy) if abs(e - 1.0) > 1e-8 else np.inf
    raw = np.arccos(np.clip(n_vec[0] / n, -1.0, 1.0)) if n != 0 else 0.0
    Omega = (raw if n_vec[1] >= 0 else 2 * np.pi - raw) if n != 0 else 0.0
    raw = np.arccos(np.clip(np.dot(n_vec, e_vec) / (n * e + 1e-12), -1.0, 1.0))
    omega = (raw if e_vec[2] >= 0 else 2 * np.pi - raw) if (n != 0 and e != 0) else 0.0
    raw = np.arccos(np.clip(np.dot(e_vec, r_vec) / (e * r + 1e-12), -1.0, 1.0))
    nu = (raw if np.dot(r_vec, v_vec) >= 0 else 2 * np.pi - raw) if e != 0 else 0.0
    E = 2 * np.arctan(np.tan(nu / 2) * np.sqrt((1 - e) / (1 + e)))
    M = E - e * np.sin(E)
    wrap = lambda x: x - (2 * np.pi) * np.floor(x / (2 * np.pi))
    return a, e, i, wrap(Omega), wrap(omega), wrap(M)


# --------------------------------------------------------------------------
# Monte Carlo covariance at TCA (model.py:497-551), given the element samples
# --------------------------------------------------------------------------
def element_samples_to_sgp4(samples, mu_earth=MU_EARTH_SI):
    """model.py:500,523-533: column 0 (semi-major axis, m) -> mean motion
    [rad/s]; negative eccentricity -> 0.  Returns SGP4 inputs
    (no_kozai [rad/min], e, i, raan, argp, M), each [n]."""
    samples = np.asarray(samples, dtype=np.float64)
    n_rad_s = (mu_earth / samples[:, 0] ** 3) ** 0.5
    ecc = np.where(samples[:, 1] < 0.0, 0.0, samples[:, 1])
    return n_rad_s * 60.0, ecc, samples[:, 2], samples[:, 3], samples[:, 4], samples[:, 5]


def tca_states(no_kozai, ecco, inclo, nodeo, argpo, mo, bstar, tsince):
    """model.py:536-541: initialize_tle(list) + propagate_batch at one tsince;
    [n,2,3] metres."""
    c, err = S.sgp4_init(no_kozai, ecco, inclo, nodeo, argpo, mo, bstar)
    r, v, e2 = S.sgp4_propagate(c, np.broadcast_to(np.asarray(tsince, dtype=np.float64), np.shape(no_kozai)))
    return np.stack([r, v], axis=1) * 1e3, (err | e2)


def mc_covariance(states_xyz):
    """model.py:544-550: mean state, RTN rotation of the mean, per-sample
    rotation, np.cov (ddof=1).  states_xyz [n,2,3] m."""
    mean_state = states_xyz.mean(axis=0)
    R = rotation_matrix(mean_state)
    rtn = np.zeros_like(states_xyz)
    for i, s in enumerate(states_xyz):
        rtn[i, 0] = np.dot(R, s[0])
        rtn[i, 1] = np.dot(R, s[1])
    cov = np.cov(rtn.reshape(-1, 6).transpose())
    return mean_state, cov


# --------------------------------------------------------------------------
# Monte Carlo collision probability (model.py:424-437)
# --------------------------------------------------------------------------
def pc_count_all_pairs(t_xyz, c_xyz, thr, block=1024):
    """Number of (i,j) with ||t_i - c_j|| < thr.  The reference uses
    torch.cdist (matmul form for >25 rows, BLAS summation order unspecified);
    the contract for bit-exact counts is the direct-difference form
    ((dx*dx + dy*dy) + dz*dz) < thr*thr, no FMA."""
    t = np.asarray(t_xyz, dtype=np.float64)
    c = np.asarray(c_xyz, dtype=np.float64)
    thr2 = float(thr) * float(thr)
    total = 0
    for s in range(0, t.shape[0], block):
        tb = t[s:s + block]
        dx = tb[:, None, 0] - c[None, :, 0]
        dy = tb[:, None, 1] - c[None, :, 1]
        dz = tb[:, None, 2] - c[None, :, 2]
        sq = (dx * dx + dy * dy) + dz * dz
        total += int(np.count_nonzero(sq < thr2))
    return total


def pc_count_elementwise(t_xyz, c_xyz, thr):
    t = np.asarray(t_xyz, dtype=np.float64)
    c = np.asarray(c_xyz, dtype=np.float64)
    d = t - c
    sq = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    return int(np.count_nonzero(sq < float(thr) * float(thr)))


def pc_all_pairs_cdist(t_xyz, c_xyz, thr):
    """Literally model.py:432-437 (torch.cdist), for plausibility only."""
    d = torch.cdist(torch.as_tensor(t_xyz), torch.as_tensor(c_xyz))
    return int((d < thr).to(torch.int32).sum()), d.numel()
