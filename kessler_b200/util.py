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


_SQRT2 = math.sqrt(2.0)
_LOG_SQRT_2PI = 0.5 * math.log(2.0 * math.pi)


def _std_normal_cdf(x):
    return 0.5 * (1.0 + torch.erf(x / _SQRT2))


def _std_normal_icdf(p):
    return torch.erfinv(2.0 * p - 1.0) * _SQRT2


class TruncatedNormal(_DistBase):
    """Normal(loc, scale) conditioned on [low, high]; the interface of the reference's class (util.py:46-160: same
    constructor, `log_prob`, inverse-CDF `sample`, scalar or batched 1-d / 2-d parameters) behind an own, vectorised
    implementation.  With u ~ U(0, 1) a draw is loc + scale * Phi^-1(Phi(alpha) + u (Phi(beta) - Phi(alpha))): ONE
    torch.rand of the sample shape per attempt, like the reference, so seeded streams line up draw for draw; a second
    attempt is only needed when rounding pushes every draw of the batch outside [low, high]."""
    arg_constraints = {"_loc": constraints.real, "_scale": constraints.positive, "_low": constraints.real,
                       "_high": constraints.real}
    has_rsample = False
    support = constraints.real

    def __init__(self, loc, scale, low, high):
        self._loc, self._scale = torch.as_tensor(loc), torch.as_tensor(scale)
        self._low, self._high = torch.as_tensor(low), torch.as_tensor(high)
        if self._loc.dim() > 2:
            raise RuntimeError("Expecting 1d or 2d (batched) probabilities.")
        self._batch_length = 0 if self._loc.dim() == 0 else self._loc.size(0)
        self._cdf_low = _std_normal_cdf((self._low - self._loc) / self._scale)          # Phi(alpha)
        self._cdf_span = _std_normal_cdf((self._high - self._loc) / self._scale) - self._cdf_low   # Z = Phi(beta) - Phi(alpha)
        self._log_norm = torch.log(self._scale * self._cdf_span) + _LOG_SQRT_2PI
        super().__init__(batch_shape=self._loc.shape, event_shape=torch.Size(), validate_args=False)

    def _inside(self, x):
        return (x >= self._low) & (x <= self._high)

    def log_prob(self, value):
        value = torch.as_tensor(value)
        z = (value - self._loc) / self._scale
        out = -0.5 * z * z - self._log_norm
        out = torch.where(self._inside(value), out, torch.full_like(out, float("-inf")))    # zero density outside the support
        if self._batch_length == 1:
            out = out.squeeze(0)
        if not torch.isfinite(out).all():
            warnings.warn("NaN, -Inf, or Inf encountered in TruncatedNormal log_prob.")
        return out

    def sample(self, sample_shape=torch.Size()):
        # the reference sizes a draw by `low` (util.py:135), which only works for an empty sample_shape on a batched
        # instance; the batched front end (model.forward_batch) draws (n,) + batch_shape
        batched_draw = len(torch.Size(sample_shape)) > 0 and self._loc.dim() > self._low.dim()
        shape = torch.Size(sample_shape) + (self._loc.shape if batched_draw else self._low.shape)
        attempts = 0
        while True:
            attempts += 1
            if attempts == 10000:
                warnings.warn("Trying to sample from the tail of a truncated normal distribution, which can take a long time.")
            u = torch.rand(shape, dtype=self._low.dtype)
            draw = self._loc + self._scale * _std_normal_icdf(self._cdf_low + u * self._cdf_span)
            if not torch.isnan(draw).any() and bool(self._inside(draw).any()):       # the reference's acceptance rule
                break
        return draw.squeeze(0) if self._batch_length == 1 else draw


# ---- Cartesian <-> Kepler / RTN (util.py:187-358) ---------------------------------------------
def _angle_from_cos(cos_value, upper_half):
    """The angle in [0, 2 pi) whose cosine is cos_value (clamped to [-1, 1]) and which lies in [0, pi] when upper_half."""
    ang = torch.acos(torch.clamp(cos_value, -1.0, 1.0))
    return torch.where(upper_half, ang, 2 * torch.pi - ang)


def _wrap_two_pi(x):
    return x - (2 * torch.pi) * torch.floor(x / (2 * torch.pi))


def from_cartesian_to_keplerian(r_vec, v_vec, mu):
    """Osculating elements (a, e, i, Omega, omega, M) of a Cartesian state: the reference's function (util.py:187-264;
    same conventions: angles from arccos with quadrant fixes, 1e-12 guards in the denominators, zero for undefined
    angles, elliptic / hyperbolic mean anomaly, ranges [0, 2 pi)) written as differentiable torch code --
    keplerian_cartesian_partials takes its Jacobian by autograd, and tests/golden holds the reference's own outputs."""
    dt, dev = r_vec.dtype, r_vec.device
    zero = torch.zeros((), device=dev, dtype=dt)
    r, v = torch.linalg.norm(r_vec), torch.linalg.norm(v_vec)
    rv = torch.dot(r_vec, v_vec)
    h_vec = torch.linalg.cross(r_vec, v_vec)                                    # angular momentum
    h = torch.linalg.norm(h_vec)
    node_vec = torch.linalg.cross(torch.tensor([0.0, 0.0, 1.0], device=dev, dtype=dt), h_vec)
    node = torch.linalg.norm(node_vec)
    e_vec = ((v ** 2 - mu / r) * r_vec - rv * v_vec) * (1 / mu)                 # eccentricity (Laplace) vector
    e = torch.linalg.norm(e_vec)
    a = torch.where(torch.abs(e - 1.0) > 1e-8, -mu / (2 * (v ** 2 / 2 - mu / r)), torch.full((), float("inf"), device=dev, dtype=dt))
    inc = torch.where(h != 0, torch.acos(h_vec[2] / h), zero)
    Omega = torch.where(node != 0, _angle_from_cos(node_vec[0] / node, node_vec[1] >= 0), zero)
    omega = torch.where((node != 0) & (e != 0), _angle_from_cos(torch.dot(node_vec, e_vec) / (node * e + 1e-12), e_vec[2] >= 0), zero)
    nu = torch.where(e != 0, _angle_from_cos(torch.dot(e_vec, r_vec) / (e * r + 1e-12), rv >= 0), zero)
    half_tan = torch.tan(nu / 2)
    if e < 1.0:
        E = 2 * torch.atan(half_tan * torch.sqrt((1 - e) / (1 + e)))
        M = E - e * torch.sin(E)
    elif e > 1.0:
        F = 2 * torch.atanh(half_tan * torch.sqrt((e - 1) / (e + 1)))
        M = e * torch.sinh(F) - F
    else:
        raise UnboundLocalError("parabolic orbit: mean anomaly undefined (as in the reference)")
    return a, e, inc, _wrap_two_pi(Omega), _wrap_two_pi(omega), _wrap_two_pi(M)


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
