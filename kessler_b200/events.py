"""Batched ground segment: CDM sequences for MANY accepted conjunctions at once.

`Conjunction.forward()` (model.py:630-692 of the reference) issues the CDMs of one event in a Python
loop: per CDM and observed object a Newton solve, six autograd passes, 100 TLE objects, a propagate, a
rotation loop, `np.cov`, then 2 x 10^4 Gaussian samples and a 10^4 x 10^4 `cdist`.  This module does the
same computation for all events of a `forward_batch()` at once:

  truth states at the CDM times     ksl_states_at_fine      (one launch for all events)
  osculating -> mean elements       ksl_sgp4_prop + ksl_newton_elements
  Kepler partials                   ksl_kepler_partials
  100-sample element clouds         ksl_sgp4_init + ksl_sgp4_prop on [items x 100] records
  collision probability             ksl_pc_mvn_batched      (device-side draws)

with the 6x6 algebra vectorised in NumPy between the launches.  The per-event semantics (CDM schedule,
observation draws, carry-forward of un-observed objects, the ITRF/RTN conventions incl. the reference's
quirks) are those of `Conjunction.forward()`; tests/test_gpu_events.py checks the deterministic parts
against it event by event.  Random streams differ from the sequential path (one vectorised draw per
stage instead of one per CDM), so the outputs are equal in distribution, not sample by sample.
"""
import math
import uuid

import numpy as np
import torch

from . import engine, util
from . import _native as N
from .cdm import ConjunctionDataMessage
from .model import _quantize_exp_field, _tdiv, quantize_sgp4_inputs, quantize_sgp4_inputs_torch  # noqa: F401
from .tle import MU_EARTH, epoch_to_datetime, from_datetime_to_mjd, from_mjd_to_datetime, from_mjd_to_epoch_days_after_1_jan


def _rotation_matrices(states):
    """util.rotation_matrix for states [n, 6] -> [n, 3, 3] (rows u, t, w)."""
    r, v = states[:, :3], states[:, 3:]
    u = r / np.linalg.norm(r, axis=1, keepdims=True)
    w = np.cross(r, v)
    w = w / np.linalg.norm(w, axis=1, keepdims=True)
    t = np.cross(w, u)
    return np.stack([u, t, w], axis=1)


def _blockdiag2(R):
    n = R.shape[0]
    T = np.zeros((n, 6, 6))
    T[:, :3, :3] = R
    T[:, 3:, 3:] = R
    return T


def _teme_to_itrf(states_m, jd):
    """util.from_TEME_to_ITRF vectorised: states [n, 6] metres, jd [n] (same arithmetic per row)."""
    jd = np.asarray(jd, dtype=np.float64)
    t = (jd - 2451545.0) / 36525.0
    g = 67310.54841 + (8640184.812866 + (0.093104 + (-6.2e-6) * t) * t) * t
    dg = 8640184.812866 + (0.093104 * 2.0 + (-6.2e-6 * 3.0) * t) * t
    theta = (np.mod(jd, 1.0) + np.mod(g / 86400.0, 1.0)) * (2.0 * math.pi)
    theta_dot = (1.0 + dg / 86400.0 / 36525.0) * (2.0 * math.pi)
    c, s = np.cos(theta), np.sin(theta)
    x, y, z = states_m[:, 0], states_m[:, 1], states_m[:, 2]
    vx, vy, vz = states_m[:, 3] * 86400.0, states_m[:, 4] * 86400.0, states_m[:, 5] * 86400.0
    rx, ry = c * x + s * y, -s * x + c * y
    # v' = R v + (0, 0, -theta_dot) x r'
    wx, wy = theta_dot * ry, -theta_dot * rx
    out = np.empty_like(states_m)
    out[:, 0], out[:, 1], out[:, 2] = rx, ry, z
    out[:, 3], out[:, 4], out[:, 5] = (c * vx + s * vy + wx) / 86400.0, (-s * vx + c * vy + wy) / 86400.0, vz / 86400.0
    return out


def _batched_cholesky(A):
    """Cholesky factors of a stack of matrices; rows that are not positive definite get ok = False (and L = 0)."""
    try:
        return np.linalg.cholesky(A), np.ones(len(A), dtype=bool)
    except np.linalg.LinAlgError:
        L, ok = np.zeros_like(A), np.ones(len(A), dtype=bool)
        for k in range(len(A)):
            try:
                L[k] = np.linalg.cholesky(A[k])
            except np.linalg.LinAlgError:
                ok[k] = False
        return L, ok


def obs_epoch_mjd(time_mjd):
    """Epoch of the TLE that newton_method builds for an observation at time_mjd: year + day-of-year written
    with 8 decimals into the TLE text, read back, truncated to the microsecond (tle.TLE semantics)."""
    year = from_mjd_to_datetime(time_mjd).year
    days = float("{:012.8f}".format(from_mjd_to_epoch_days_after_1_jan(time_mjd)))
    return from_datetime_to_mjd(epoch_to_datetime(year, days))


def generate_cdm_batch(model, batch, hits, seed=0, t_new_obs=None, c_new_obs=None, z_elements=None):
    """Ground segment for the accepted draws `hits` (indices into `batch` = model.forward_batch(n)).

    Returns a dict of arrays over events e (= hits) and CDM slots i < max_ncdm[e] (padded to the largest):
      max_ncdm [E]; issued [E, I] bool; time_cdm [I]; state_itrf_km [E, I, 2, 6]; cov_rtn [E, I, 2, 6, 6];
      pc [E, I]; tca_mjd [E]; plus the intermediates `items` (one row per new observation).
    t_new_obs / c_new_obs [E, I] and z_elements [n_items, S, 6] can be injected by tests."""
    hits = np.asarray(hits, dtype=np.int64)
    E = len(hits)
    times = model.time_grid()
    time0 = model._time0
    S = int(model._mc_samples)
    n_pc = int(model._mc_samples * model._mc_upsample_factor)
    time_min = batch["time_min"][hits]
    # ---- schedule (model.py:632-638): same grid times for every event -------------------------------
    max_ncdm = np.maximum(1, ((time_min - time0) * 24.0 / model._cdm_update_every_hours).astype(np.int64))
    I = int(max_ncdm.max())
    idx_cdm = np.empty(I, dtype=np.int64)
    for i in range(I):
        idx_cdm[i], _ = util.find_closest(times, time0 + (i * model._cdm_update_every_hours) / 24)
    time_cdm = times[idx_cdm]
    slot = np.arange(I)[None, :] < max_ncdm[:, None]
    # ---- observation draws (model.py:643-650) --------------------------------------------------------
    g = torch.Generator().manual_seed(int(seed))
    if t_new_obs is None:
        t_new_obs = torch.bernoulli(torch.full((E, I), float(model._t_prob_new_obs)), generator=g).numpy() > 0
    if c_new_obs is None:
        c_new_obs = torch.bernoulli(torch.full((E, I), float(model._c_prob_new_obs)), generator=g).numpy() > 0
    obs = np.stack([np.asarray(t_new_obs)[:, :I], np.asarray(c_new_obs)[:, :I]], axis=2).astype(bool)          # [E, I, 2]
    obs[:, 0, :] = True
    obs &= slot[:, :, None]
    ev, ci, ob = np.nonzero(obs)                                         # items: one per new observation
    n_items = len(ev)
    # ---- records: initialised target / chaser of every event ----------------------------------------
    # From here to the RTN covariances everything lives on the device as torch tensors (torch = plumbing: batched 6 x 6
    # algebra between the launches); one packed transfer at the end brings the per-item results back.
    dev = torch.device("cuda", torch.cuda.current_device())
    D = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    el = np.concatenate([batch["elems_t"][:, hits], batch["elems_c"][:, hits]], axis=1)      # [7, 2E]
    ep = np.concatenate([batch["epoch_t"][hits], batch["epoch_c"][hits]])
    consts, _ = engine.sgp4_init(el)
    rec = ob * E + ev
    rec_d = D(rec)
    tc = np.ascontiguousarray(times[::max(int(model._time_upsample_factor), 1)])
    truth_d, _ = engine.states_at_fine(consts, rec, ep[rec], idx_cdm[ci], tc, len(times))
    # instruments (model.py:655-676): mean over the object's instruments of (state + bias, diagonal RTN covariance)
    def inst(instruments):
        bias = np.mean([np.asarray(x._instrument_characteristics["bias_xyz"], dtype=np.float64).reshape(6) for x in instruments], axis=0)
        cov = np.mean([np.asarray(x._instrument_characteristics["covariance_rtn"], dtype=np.float64) for x in instruments], axis=0)
        return bias, cov
    bias_t, cov_t = inst(model._t_observing_instruments)
    bias_c, cov_c = inst(model._c_observing_instruments)
    is_c = D(ob.astype(np.float64))[:, None]
    state_obs_d = truth_d + (1.0 - is_c) * D(bias_t)[None, :] + is_c * D(bias_c)[None, :]
    cov_diag_d = (1.0 - is_c) * D(cov_t)[None, :] + is_c * D(cov_c)[None, :]
    # ---- newton_method (model.py:399,411): target = the TLE's own SGP4 state at the observation time -----
    tsince_obs = (time_cdm[ci] - ep[rec]) * 1440.0
    rv, _ = engine.sgp4_prop(consts[:, rec_d], D(tsince_obs).reshape(-1, 1), per_sample_t=True)
    bstar = el[6, rec]
    bstar_q = D(_quantize_exp_field(bstar))
    y_d, resid_d, _ = engine.newton_elements(rv[:, 0, :], D(bstar))
    two_pi = 2 * math.pi
    wrap = lambda a: torch.remainder(a, two_pi)
    raw_mm_d = _tdiv(y_d[:, 0], 60.0)            # IEEE division, as the host path's y[:, 0] / 60.0
    q_d = quantize_sgp4_inputs_torch(raw_mm_d, y_d[:, 1], wrap(y_d[:, 2]), wrap(y_d[:, 3]), wrap(y_d[:, 4]), wrap(y_d[:, 5]), bstar_q)   # TLE(dict)
    epoch_obs_by_slot = np.array([obs_epoch_mjd(t) for t in time_cdm])
    epoch_obs = epoch_obs_by_slot[ci]
    # ---- element covariance (model.py:461-478) -----------------------------------------------------------
    def rot(states):                               # util.rotation_matrix, batched: rows u, t, w
        r, v = states[:, :3], states[:, 3:]
        u = r / torch.linalg.norm(r, dim=1, keepdim=True)
        w = torch.linalg.cross(r, v)
        w = w / torch.linalg.norm(w, dim=1, keepdim=True)
        return torch.stack([u, torch.linalg.cross(w, u), w], dim=1)
    R = rot(state_obs_d)
    T = torch.zeros((n_items, 6, 6), dtype=torch.float64, device=dev)
    T[:, :3, :3] = R
    T[:, 3:, 3:] = R
    C_xyz = torch.einsum("nji,nj,njk->nik", T, cov_diag_d, T)
    _, J = engine.kepler_partials(state_obs_d, model._mu_earth)
    Cov = torch.einsum("nij,njk,nlk->nil", J, C_xyz, J)
    # model.py:474-477: "all eigenvalues > 1e-11, else symmetrise and add 1e-10 I".  Cov = J C J^T is symmetric up to
    # rounding, so the eigenvalues of its symmetric part decide (batched, on the device); non-finite rows count as failing
    sym = 0.5 * (Cov + Cov.transpose(1, 2))
    finite = torch.isfinite(sym).all(dim=2).all(dim=1)
    # (all eigenvalues of a symmetric matrix exceed 1e-11  <=>  the matrix minus 1e-11 I is positive definite  <=>  its
    # Cholesky factorisation succeeds: one batched factorisation instead of an eigen-decomposition)
    eye = torch.eye(6, dtype=torch.float64, device=dev)[None]
    _, info_pd = torch.linalg.cholesky_ex(torch.where(finite[:, None, None], sym, eye) - 1e-11 * eye)
    fix = ~((info_pd == 0) & finite)
    Cov = torch.where(fix[:, None, None], sym + 1e-10 * torch.eye(6, dtype=torch.float64, device=dev)[None], Cov)
    mean_els = torch.stack([(MU_EARTH / raw_mm_d ** 2) ** (1.0 / 3.0), q_d[1], q_d[2], q_d[3], q_d[4], q_d[5]], dim=1)
    # ---- element clouds (model.py:497-541) ------------------------------------------------------------------
    # not PD: the reference's "Expected parameter covariance_matrix" branch (the instrument's own covariance is kept)
    L, info = torch.linalg.cholesky_ex(torch.where(torch.isfinite(Cov), Cov, torch.zeros_like(Cov)))
    chol_ok = (info == 0) & torch.isfinite(Cov).all(dim=2).all(dim=1)
    if z_elements is None:
        gd = torch.Generator(device=dev).manual_seed(int(seed) + 104729)
        z_d = torch.randn((n_items, S, 6), dtype=torch.float64, device=dev, generator=gd)
    else:
        z_d = D(np.asarray(z_elements, dtype=np.float64))
    samples = mean_els[:, None, :] + torch.matmul(z_d, L.transpose(1, 2))
    samples = torch.where(chol_ok[:, None, None], samples, mean_els[:, None, :].expand(-1, S, -1))
    flat = samples.reshape(-1, 6)
    mm = (MU_EARTH / flat[:, 0] ** 3) ** 0.5
    el_s = quantize_sgp4_inputs_torch(mm, torch.where(flat[:, 1] < 0.0, torch.zeros_like(flat[:, 1]), flat[:, 1]), flat[:, 2], flat[:, 3],
                                      flat[:, 4], flat[:, 5], bstar_q.repeat_interleave(S))
    tsince_ca = D((time_min[ev] - epoch_obs) * 1440.0).repeat_interleave(S)
    c_s, _ = engine.sgp4_init(el_s)
    rv_s, _ = engine.sgp4_prop(c_s, tsince_ca.reshape(-1, 1), per_sample_t=True)
    del c_s
    mc = rv_s[:, 0, :].reshape(n_items, S, 6) * 1e3
    mean_state_d = mc.mean(dim=1)
    RmT = rot(mean_state_d).transpose(1, 2)
    rtn = torch.cat([torch.matmul(mc[:, :, :3], RmT), torch.matmul(mc[:, :, 3:], RmT)], dim=2)
    d = rtn - rtn.mean(dim=1, keepdim=True)
    cov_rtn_d = torch.matmul(d.transpose(1, 2), d) / (S - 1)
    cov_rtn_d = torch.where(chol_ok[:, None, None], cov_rtn_d, torch.diag_embed(cov_diag_d))
    # ---- one packed transfer of the per-item results -----------------------------------------------------------
    packed = torch.cat([truth_d, state_obs_d, y_d, resid_d[:, None], q_d.t(), Cov.reshape(n_items, 36), mean_state_d,
                        cov_rtn_d.reshape(n_items, 36)], dim=1).cpu().numpy()
    truth, state_obs, y, resid_h = packed[:, 0:6], packed[:, 6:12], packed[:, 12:18], packed[:, 18]
    q = np.ascontiguousarray(packed[:, 19:26].T)
    Cov_h = packed[:, 26:62].reshape(n_items, 6, 6)
    mean_state = np.ascontiguousarray(packed[:, 62:68])
    cov_rtn = packed[:, 68:104].reshape(n_items, 6, 6).copy()
    tca_jd = util.from_mjd_to_jd(time_min)
    state_itrf_km = _teme_to_itrf(mean_state, tca_jd[ev]) / 1e3
    # ---- carry-forward (model.py:345-346): an un-observed object keeps the previous CDM's numbers ------------
    st = np.full((E, I, 2, 6), np.nan)
    cv = np.full((E, I, 2, 6, 6), np.nan)
    st[ev, ci, ob] = state_itrf_km
    cv[ev, ci, ob] = cov_rtn
    for i in range(1, I):
        keep = ~obs[:, i, :]
        st[:, i][keep] = st[:, i - 1][keep]
        cv[:, i][keep] = cv[:, i - 1][keep]
    issued = obs.any(axis=2)
    # ---- collision probability of every issued CDM (model.py:424-437) ----------------------------------------
    e_i, c_i = np.nonzero(issued)
    s_m = st[e_i, c_i] * 1e3                                             # [n_cdm, 2, 6] metres (ITRF, sic)
    Rr = _rotation_matrices(s_m.reshape(-1, 6)).reshape(len(e_i), 2, 3, 3)
    mean_rtn = np.einsum("nkij,nkj->nki", Rr, s_m[:, :, :3])
    cov3 = cv[e_i, c_i][:, :, :3, :3]
    L3f, ok3 = _batched_cholesky(cov3.reshape(-1, 3, 3))
    L3, pc_ok = L3f.reshape(cov3.shape), ok3.reshape(-1, 2).all(axis=1)
    counts = engine.pc_mvn_batched(mean_rtn, L3, n_pc, model._collision_threshold, seed=int(seed) + 1).cpu().numpy()
    pc = np.full((E, I), np.nan)
    pc[e_i, c_i] = np.where(pc_ok, counts / float(n_pc * n_pc), np.nan)
    return dict(hits=hits, max_ncdm=max_ncdm, time_cdm=time_cdm, idx_cdm=idx_cdm, issued=issued, new_obs=obs,
                state_itrf_km=st, cov_rtn=cv, pc=pc, tca_mjd=time_min,
                items=dict(event=ev, slot=ci, obj=ob, truth_m=truth, state_obs_m=state_obs, newton_y=y, newton_resid=resid_h,
                           obs_elements=q, obs_epoch_mjd=epoch_obs, cov_tle=Cov_h, mean_state_tca_m=mean_state, cov_rtn=cov_rtn))


def to_cdms(model, res, e, t_tle=None, c_tle=None):
    """ConjunctionDataMessage objects of event e of a generate_cdm_batch() result (same fields as generate_cdm)."""
    cdms = []
    tca_jd = util.from_mjd_to_jd(res["tca_mjd"][e])
    for i in range(int(res["max_ncdm"][e])):
        if not res["issued"][e, i]:
            continue
        cdm = ConjunctionDataMessage()
        cdm.set_header("ORIGINATOR", "KESSLER_SOFTWARE")
        # object metadata as generate_cdm writes it (model.py:335-372): from the TLEs when the caller has them, the
        # KESSLER_SOFTWARE placeholders of a prior-sampled object otherwise -- every obligatory CCSDS key is set
        model._describe_object(cdm, 0, t_tle, t_tle is not None, "TARGET")
        model._describe_object(cdm, 1, c_tle, c_tle is not None, "CHASER")
        cdm.set_header("MESSAGE_ID", "KESSLER_SOFTWARE_{}".format(uuid.uuid1()))
        cdm.set_relative_metadata("TCA", util.from_jd_to_cdm_datetime_str(tca_jd))
        cdm.set_header("CREATION_DATE", util.from_jd_to_cdm_datetime_str(util.from_mjd_to_jd(res["time_cdm"][i])))
        for o in (0, 1):
            cdm.set_state(o, res["state_itrf_km"][e, i, o].reshape(2, 3))
            cdm.set_covariance(o, res["cov_rtn"][e, i, o])
        cdm.set_relative_metadata("COLLISION_PROBABILITY", float(res["pc"][e, i]))
        cdm.set_relative_metadata("COLLISION_PROBABILITY_METHOD", model._pc_method)
        cdms.append(cdm)
    return cdms


_RTN = ("R", "T", "N", "RDOT", "TDOT", "NDOT")
LSTM_FEATURES = (["__CREATION_DATE", "__TCA", "MISS_DISTANCE", "RELATIVE_SPEED", "RELATIVE_POSITION_R", "RELATIVE_POSITION_T",
                  "RELATIVE_POSITION_N", "RELATIVE_VELOCITY_R", "RELATIVE_VELOCITY_T", "RELATIVE_VELOCITY_N"] +
                 [f"OBJECT{o}_{k}" for o in (1, 2)
                  for k in ["X", "Y", "Z", "X_DOT", "Y_DOT", "Z_DOT"] + [f"C{_RTN[i]}_{_RTN[j]}" for i in range(6) for j in range(i + 1)]])


def feature_tensor(res):
    """Struct-of-arrays form of the CDM sequences of a generate_cdm_batch() result: the 64 default input features
    of the reference's LSTMPredictor (/root/reference/kessler/nn.py:60-123) for every issued CDM, without building
    CDM objects.  Returns (features [n_cdm, 64] float64 in LSTM_FEATURES order, event_ptr [E + 1]: the CDMs of event e
    are rows event_ptr[e]:event_ptr[e + 1], in issue order).  `__CREATION_DATE` / `__TCA` are days since the event's
    first CDM, through the CDM date strings (event.py:42-47)."""
    E, I = res["issued"].shape
    creation_str = [util.from_jd_to_cdm_datetime_str(util.from_mjd_to_jd(t)) for t in res["time_cdm"]]
    day0 = "2000-01-01T00:00:00.000000"
    creation_days = np.array([util.from_date_str_to_days(c, date0=day0) for c in creation_str])
    tca_days = np.array([util.from_date_str_to_days(util.from_jd_to_cdm_datetime_str(util.from_mjd_to_jd(t)), date0=day0)
                         for t in res["tca_mjd"]])
    e_i, c_i = np.nonzero(res["issued"])
    n = len(e_i)
    first = np.array([np.nonzero(res["issued"][e])[0][0] if res["issued"][e].any() else 0 for e in range(E)])
    feats = np.empty((n, len(LSTM_FEATURES)))
    feats[:, 0] = creation_days[c_i] - creation_days[first[e_i]]
    feats[:, 1] = tca_days[e_i] - creation_days[first[e_i]]
    st = res["state_itrf_km"][e_i, c_i]                                  # [n, 2, 6] km, km/s
    s1, s2 = st[:, 0] * 1e3, st[:, 1] * 1e3
    rot = _rotation_matrices(s1)
    rel_p = np.einsum("nij,nj->ni", rot, s2[:, :3] - s1[:, :3])
    rel_v = np.einsum("nij,nj->ni", rot, s2[:, 3:] - s1[:, 3:])
    feats[:, 2] = np.linalg.norm(s1[:, :3] - s2[:, :3], axis=1)
    feats[:, 3] = np.linalg.norm(rel_v, axis=1)
    feats[:, 4:7], feats[:, 7:10] = rel_p, rel_v
    tril = np.tril_indices(6)
    cv = res["cov_rtn"][e_i, c_i]
    for o in (0, 1):
        base = 10 + 27 * o
        feats[:, base:base + 6] = st[:, o]
        feats[:, base + 6:base + 27] = cv[:, o][:, tril[0], tril[1]]
    ptr = np.concatenate([[0], np.cumsum(res["issued"].sum(axis=1))])
    return feats, ptr


def to_dataframe(res, event_id_offset=0):
    """The CDM sequences of a generate_cdm_batch() result as one pandas frame, one row per issued CDM: `event_id`,
    the CDM date strings `CREATION_DATE` / `TCA`, `COLLISION_PROBABILITY`, and the 62 numeric CDM fields of
    LSTM_FEATURES under their CDM keys (what `EventDataset.to_dataframe()` holds for these columns in the reference,
    event.py:422-428) -- built from the struct-of-arrays result, no CDM objects."""
    import pandas as pd
    feats, ptr = feature_tensor(res)
    e_i, c_i = np.nonzero(res["issued"])
    creation = [util.from_jd_to_cdm_datetime_str(util.from_mjd_to_jd(t)) for t in res["time_cdm"]]
    tca = [util.from_jd_to_cdm_datetime_str(util.from_mjd_to_jd(t)) for t in res["tca_mjd"]]
    cols = {"event_id": e_i + int(event_id_offset), "CREATION_DATE": [creation[i] for i in c_i], "TCA": [tca[e] for e in e_i],
            "COLLISION_PROBABILITY": res["pc"][e_i, c_i]}
    for k, name in enumerate(LSTM_FEATURES):
        if not name.startswith("__"):
            cols[name] = feats[:, k]
    return pd.DataFrame(cols)


_BATCH_KEYS = ("elems_t", "elems_c", "epoch_t", "epoch_c", "time_min", "i_min", "d_min", "i_conj", "d_conj")


def _take_hits(batch, hits):
    """Compact copy of a forward_batch() result restricted to `hits` (what the ground segment needs)."""
    out = {}
    for k in _BATCH_KEYS:
        a = batch[k]
        out[k] = np.ascontiguousarray(a[:, hits] if a.ndim == 2 else a[hits])
    for who in ("t_raw", "c_raw"):
        out[who] = {k: (v[hits] if isinstance(v, np.ndarray) else v) for k, v in batch[who].items()}
    return out


def _concat_hits(parts):
    out = {}
    for k in _BATCH_KEYS:
        out[k] = np.concatenate([p[k] for p in parts], axis=1 if parts[0][k].ndim == 2 else 0)
    for who in ("t_raw", "c_raw"):
        out[who] = {k: (np.concatenate([p[who][k] for p in parts]) if isinstance(v, np.ndarray) else v)
                    for k, v in parts[0][who].items()}
    return out


def generate_events(model, n_events, batch_size=65536, seed=0, max_batches=100000, events_per_pass=8192, device_batches=True):
    """Rejection-sample conjunctions `batch_size` candidates at a time (one fused screen launch each), collect the
    accepted draws, and run the batched ground segment on up to `events_per_pass` of them at once.  Returns
    (list of generate_cdm_batch results -- each with its compact `batch` under key "batch" --, candidates screened)."""
    out, screened, found, pending, n_pending, passes = [], 0, 0, [], 0, 0

    def flush():
        nonlocal pending, n_pending, passes
        if not pending:
            return
        merged = _concat_hits(pending)
        res = generate_cdm_batch(model, merged, np.arange(len(merged["time_min"])), seed=seed + 7919 * passes)
        res["batch"] = merged
        out.append(res)
        pending, n_pending = [], 0
        passes += 1

    for _ in range(max_batches):
        if found >= n_events:
            break
        dev_batch = model.forward_batch_device(batch_size, max_hits=max(0, n_events - found)) if device_batches else None
        screened += batch_size
        if dev_batch is not None:       # drawn, pre-filtered, screened and compacted on the GPU: only the accepted draws came back
            compact, nh = dev_batch
            if nh == 0:
                continue
            pending.append(compact)
        else:
            batch = model.forward_batch(batch_size)
            hits = np.nonzero(batch["conj"])[0][:max(0, n_events - found)]
            nh = len(hits)
            if nh == 0:
                continue
            pending.append(_take_hits(batch, hits))
        n_pending += nh
        found += nh
        if n_pending >= events_per_pass:
            flush()
    flush()
    return out, screened
