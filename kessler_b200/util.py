"""Host-side mirror of the part of ``kessler.util`` that the Monte Carlo hot path uses
(/root/reference/kessler/util.py), same names, argument meaning and error behaviour.

Device work (``propagate_upsample``, ``upsample``) goes through the C ABI
(``engine``); the small 6x6 / scalar conversions stay on the host exactly as SURVEY.md
Sec. 8a rows 10-11 scope them.  Out-of-scope parts of kessler.util (plot helpers,
megaconstellation builders, CDM string utilities beyond what generate_cdm needs) are
not here.
"""
import datetime
import functools
import math
import random
import time
import warnings

import numpy as np
import torch

from . import tle as _tle

try:  # the reference builds its priors on pyro; run unmodified where it exists
    from pyro.distributions.torch_distribution import TorchDistribution as _DistBase
    import pyro.distributions as _pdist
    _Normal = _pdist.Normal
except Exception:  # pyro is absent in this image: plain torch distributions
    from torch.distributions import Distribution as _DistBase
    _Normal = torch.distributions.Normal

from torch.distributions import constraints

_random_seed = None


def seed(seed=None):
    """util.py:33-44."""
    if seed is None:
        seed = int((time.time() * 1e6) % 1e8)
    global _random_seed
    _random_seed = seed
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)


class TruncatedNormal(_DistBase):
    """util.py:46-160: truncated normal with inverse-CDF sampling (optionally batched)."""
    arg_constraints = {"_loc": constraints.real, "_scale": constraints.positive, "_low": constraints.real,
                       "_high": constraints.real}
    has_rsample = False

    def __init__(self, loc, scale, low, high):
        self._loc = torch.as_tensor(loc)
        self._scale = torch.as_tensor(scale)
        self._low = torch.as_tensor(low)
        self._high = torch.as_tensor(high)
        if self._loc.dim() == 0:
            self._batch_length = 0
        elif self._loc.dim() in (1, 2):
            self._batch_length = self._loc.size(0)
        else:
            raise RuntimeError("Expecting 1d or 2d (batched) probabilities.")
        self._standard_normal_dist = _Normal(torch.zeros_like(self._loc), torch.ones_like(self._scale))
        self._alpha = (self._low - self._loc) / self._scale
        self._beta = (self._high - self._loc) / self._scale
        self._standard_normal_cdf_alpha = self._standard_normal_dist.cdf(self._alpha)
        self._standard_normal_cdf_beta = self._standard_normal_dist.cdf(self._beta)
        self._Z = self._standard_normal_cdf_beta - self._standard_normal_cdf_alpha
        self._log_stddev_Z = torch.log(self._scale * self._Z)
        super().__init__(batch_shape=self._loc.shape, event_shape=torch.Size(), validate_args=False)

    @property
    def support(self):
        return constraints.real

    def log_prob(self, value):
        value = torch.as_tensor(value)
        lb = value.ge(self._low).type_as(self._low)
        ub = value.le(self._high).type_as(self._low)
        lp = (torch.log(lb.mul(ub)) + self._standard_normal_dist.log_prob((value - self._loc) / self._scale)
              - self._log_stddev_Z)
        if self._batch_length == 1:
            lp = lp.squeeze(0)
        if torch.any(torch.isnan(lp)) or torch.isinf(lp).any():
            warnings.warn("NaN, -Inf, or Inf encountered in TruncatedNormal log_prob.")
        return lp

    def sample(self, sample_shape=torch.Size()):
        shape = torch.Size(sample_shape) + self._low.shape      # util.py:135: low's shape, loc broadcasts
        if len(shape) > 0 and self._loc.dim() > self._low.dim():
            # the reference cannot draw a non-empty sample_shape from a batched instance (the
            # shapes do not broadcast); the batched front-end (model.forward_batch) needs it
            shape = torch.Size(sample_shape) + self._loc.shape
        attempt_count = 0
        ret = torch.full(shape, float("NaN"), dtype=self._low.dtype)
        outside_domain = True
        while torch.isnan(ret).any() or outside_domain:
            attempt_count += 1
            if attempt_count == 10000:
                warnings.warn("Trying to sample from the tail of a truncated normal distribution, which can take a long time.")
            rand = torch.rand(shape, dtype=self._low.dtype)
            ret = (self._standard_normal_dist.icdf(
                self._standard_normal_cdf_alpha + rand * (self._standard_normal_cdf_beta - self._standard_normal_cdf_alpha))
                * self._scale + self._loc)
            lb = ret.ge(self._low).type_as(self._low)
            ub = ret.le(self._high).type_as(self._low)
            outside_domain = (torch.sum(lb.mul(ub)) == 0)
        if self._batch_length == 1:
            ret = ret.squeeze(0)
        return ret


# ---- Cartesian <-> Kepler / RTN (util.py:187-358) ---------------------------------------------
def from_cartesian_to_keplerian(r_vec, v_vec, mu):
    """util.py:187-264, differentiable torch code; returns (a, e, i, Omega, omega, M)."""
    dt, dev = r_vec.dtype, r_vec.device
    r = torch.norm(r_vec)
    v = torch.norm(v_vec)
    h_vec = torch.linalg.cross(r_vec, v_vec)
    h = torch.norm(h_vec)
    zero = torch.tensor(0.0, device=dev, dtype=dt)
    i = torch.where(h != 0, torch.acos(h_vec[2] / h), zero)
    K = torch.tensor([0.0, 0.0, 1.0], device=dev, dtype=dt)
    n_vec = torch.linalg.cross(K, h_vec)
    n = torch.norm(n_vec)
    e_vec = (1 / mu) * ((v ** 2 - mu / r) * r_vec - torch.dot(r_vec, v_vec) * v_vec)
    e = torch.norm(e_vec)
    energy = v ** 2 / 2 - mu / r
    a = torch.where(torch.abs(e - 1.0) > 1e-8, -mu / (2 * energy), torch.tensor(float("inf"), device=dev, dtype=dt))
    cos_Omega = torch.clamp(n_vec[0] / n, -1.0, 1.0)
    raw_Omega = torch.acos(cos_Omega)
    Omega = torch.where(n_vec[1] >= 0, raw_Omega, 2 * torch.pi - raw_Omega)
    Omega = torch.where(n != 0, Omega, zero)
    cos_omega = torch.clamp(torch.dot(n_vec, e_vec) / (n * e + 1e-12), -1.0, 1.0)
    raw_omega = torch.acos(cos_omega)
    omega = torch.where(e_vec[2] >= 0, raw_omega, 2 * torch.pi - raw_omega)
    omega = torch.where((n != 0) & (e != 0), omega, zero)
    cos_nu = torch.clamp(torch.dot(e_vec, r_vec) / (e * r + 1e-12), -1.0, 1.0)
    raw_nu = torch.acos(cos_nu)
    nu = torch.where(torch.dot(r_vec, v_vec) >= 0, raw_nu, 2 * torch.pi - raw_nu)
    nu = torch.where(e != 0, nu, zero)
    if e < 1.0:
        E = 2 * torch.atan(torch.tan(nu / 2) * torch.sqrt((1 - e) / (1 + e)))
        M = E - e * torch.sin(E)
    elif e > 1.0:
        F = 2 * torch.atanh(torch.tan(nu / 2) * torch.sqrt((e - 1) / (e + 1)))
        M = e * torch.sinh(F) - F
    else:
        raise UnboundLocalError("parabolic orbit: mean anomaly undefined (as in the reference)")
    Omega = Omega - (2 * torch.pi) * torch.floor(Omega / (2 * torch.pi))
    omega = omega - (2 * torch.pi) * torch.floor(omega / (2 * torch.pi))
    M = M - (2 * torch.pi) * torch.floor(M / (2 * torch.pi))
    return a, e, i, Omega, omega, M


def keplerian_cartesian_partials(state, mu):
    """util.py:266-302: d(a,e,i,Omega,omega,M)/d(x,y,z,vx,vy,vz) as a [6,6] numpy array.
    The reference runs six separate forward+backward passes; one forward graph and six
    backward sweeps give the same derivatives."""
    st = torch.as_tensor(state).detach().clone().to(torch.float64).requires_grad_(True)
    els = from_cartesian_to_keplerian(st[0], st[1], mu=mu)
    rows = []
    for k, el in enumerate(els):
        (g,) = torch.autograd.grad(el, st, retain_graph=(k < 5))
        rows.append(g.flatten().numpy().copy())
    return np.stack(rows)


def keplerian_cartesian_partials_batch(states, mu):
    """Batched counterpart on the GPU (ksl_kepler_partials, forward-mode dual numbers): states [n, 2, 3] or
    [n, 6] -> numpy [n, 6, 6]; agrees with keplerian_cartesian_partials to rounding."""
    from . import engine
    _, jac = engine.kepler_partials(np.asarray(states, dtype=np.float64).reshape(-1, 6), mu)
    return jac.cpu().numpy()


def rotation_matrix(state):
    """util.py:304-320."""
    r, v = state[0], state[1]
    u = r / np.linalg.norm(r)
    w = np.cross(r, v)
    w = w / np.linalg.norm(w)
    v = np.cross(w, u)
    return np.vstack((u, v, w))


def from_cartesian_to_rtn(state, cartesian_to_rtn_rotation_matrix=None):
    """util.py:323-341."""
    if cartesian_to_rtn_rotation_matrix is None:
        cartesian_to_rtn_rotation_matrix = rotation_matrix(state)
    r, v = state[0], state[1]
    r_rtn = np.dot(cartesian_to_rtn_rotation_matrix, r)
    v_rtn = np.dot(cartesian_to_rtn_rotation_matrix, v)
    return np.stack([r_rtn, v_rtn]), cartesian_to_rtn_rotation_matrix


def from_rtn_to_cartesian(state_rtn, rtn_to_cartesian_rotation_matrix):
    """util.py:344-358."""
    r_rtn, v_rtn = state_rtn[0], state_rtn[1]
    return np.stack([np.matmul(rtn_to_cartesian_rotation_matrix, r_rtn),
                     np.matmul(rtn_to_cartesian_rotation_matrix, v_rtn)])


_T0 = 2451545.0
_TAU = 2.0 * math.pi


def _theta_gmst1982(jd_ut1, fraction_ut1=0.0):
    """IAU-82 Greenwich mean sidereal time (rad) and its rate (rad/day) -- the model
    skyfield.sgp4lib.TEME_to_ITRF uses (Vallado et al. 2006, eq. 1-2)."""
    t = (jd_ut1 - _T0 + fraction_ut1) / 36525.0
    g = 67310.54841 + (8640184.812866 + (0.093104 + (-6.2e-6) * t) * t) * t
    dg = 8640184.812866 + (0.093104 * 2.0 + (-6.2e-6 * 3.0) * t) * t
    theta = (jd_ut1 % 1.0 + fraction_ut1 + g / 86400.0 % 1.0) * _TAU
    theta_dot = (1.0 + dg / 86400.0 / 36525.0) * _TAU
    return theta, theta_dot


def from_TEME_to_ITRF(state, time):
    """util.py:361-388.  The reference delegates to skyfield.sgp4lib.TEME_to_ITRF (absent
    here); this is that transformation written out: rotation about z by -GMST(1982) and
    the omega x r term for the velocity, no polar motion (skyfield's xp = yp = 0 default).
    Pinned by /root/reference/tests/test_util.py:33-43 (tests/test_host_util.py)."""
    r, v = np.asarray(state[0], dtype=np.float64), np.asarray(state[1], dtype=np.float64)
    theta, theta_dot = _theta_gmst1982(time)
    c, s = math.cos(theta), math.sin(theta)
    R = np.array([[c, s, 0.0], [-s, c, 0.0], [0.0, 0.0, 1.0]])      # rot_z(-theta)
    r_new = R.dot(r)
    v_day = R.dot(v * 86400.0) + np.cross(np.array([0.0, 0.0, -theta_dot]), r_new)
    return np.stack([r_new, v_day / 86400.0])


# ---- dates (util.py:390-470) ----------------------------------------------------------------
def from_datetime_to_cdm_datetime_str(date):
    return date.strftime("%Y-%m-%dT%H:%M:%S.%f")


def from_mjd_to_jd(mjd_date):
    return mjd_date + 2400000.5


def from_jd_to_mjd(jd_date):
    return jd_date - 2400000.5


def from_jd_to_cdm_datetime_str(jd_date):
    return from_datetime_to_cdm_datetime_str(_tle.from_jd_to_datetime(jd_date))


@functools.lru_cache(maxsize=None)
def from_date_str_to_days(date, date0="2020-05-22T21:41:31.975", date_format="%Y-%m-%dT%H:%M:%S.%f"):
    date = datetime.datetime.strptime(date, date_format)
    date0 = datetime.datetime.strptime(date0, date_format)
    dd = date - date0
    return dd.days + (dd.seconds + dd.microseconds / 1e6) / (60 * 60 * 24)


def add_days_to_date_str(date0, days):
    date0 = datetime.datetime.strptime(date0, "%Y-%m-%dT%H:%M:%S.%f")
    return from_datetime_to_cdm_datetime_str(date0 + datetime.timedelta(days=days))


# ---- grid helpers (util.py:527-580) ------------------------------------------------------------
def find_closest(values, t):
    """util.py:527-539."""
    indx = np.argmin(abs(values - t))
    return indx, values[indx]


def upsample(s, target_resolution):
    """util.py:541-555 on the GPU (ksl_upsample): torch's linear / align_corners=True
    arithmetic, bit-equal to F.interpolate on CPU.  Returns a tensor on s's device."""
    from . import engine
    y = engine.upsample(s, int(target_resolution))
    return y if (torch.is_tensor(s) and s.is_cuda) else y.cpu()


def propagate_upsample(tle, times_mjd, upsample_factor=1):
    """util.py:557-580: numpy [N, 2, 3] in metres and m/s."""
    from . import sgp4
    return sgp4.propagate_upsample_device(tle, times_mjd, upsample_factor).reshape(-1, 2, 3).cpu().numpy()


def has_nan_or_inf(value):
    if torch.is_tensor(value):
        value = torch.sum(value)
        return (int(torch.isnan(value)) > 0) or (int(torch.isinf(value)) > 0)
    value = float(value)
    return math.isnan(value) or math.isinf(value)


def trace_to_event(trace):
    """util.py:612-621 (returns the CDM list wrapped as kessler_b200.cdm.Event)."""
    from .cdm import Event
    return Event(cdms=trace.nodes["cdms"]["infer"]["cdms"])
