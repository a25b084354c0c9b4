"""Host-side mirror of /root/reference/kessler/model.py for the Monte Carlo hot path.

Same public surface -- ``find_conjunction``, ``default_prior``, ``Conjunction(**21 kwargs)`` with
``make_target/make_chaser/forward/get_conjunction/generate_cdm/propagate_uncertainty_monte_carlo``,
``ConjunctionSimplified(tles, exclude_object_name=None, ...)`` -- and the same trace site names
(SURVEY.md Appendix B), but every data-parallel step runs in libkessler_b200.so:

  model.py:581-593  initialize_tle + propagate_upsample   -> ksl_sgp4_init, ksl_propagate_upsample
  model.py:23-52    find_conjunction                      -> ksl_find_conjunction
  model.py:536-550  100 TLEs, propagate_batch, np.cov     -> ksl_tca_moments
  model.py:432-437  torch.cdist all-pairs count           -> ksl_pc_count

plus what the reference does not have: ``forward_batch(n)`` -- n prior draws screened in ONE fused
launch (ksl_screen) -- which ``get_conjunction`` uses to reject thousands of candidates per GPU call
and then replays only the accepted draw through the traced ``forward()``.
"""
import math
import uuid

import numpy as np
import torch

from . import engine, ppl, sgp4, util
from . import _native as N
from .cdm import ConjunctionDataMessage
from .observation_model import GNSS, Radar
from .ppl import dist
from .tle import TLE, MU_EARTH, R_EARTH, from_mjd_to_datetime, from_mjd_to_epoch_days_after_1_jan

Uniform, Normal, Bernoulli, Categorical, MixtureSameFamily = (dist.Uniform, dist.Normal, dist.Bernoulli,
                                                               dist.Categorical, dist.MixtureSameFamily)


# --------------------------------------------------------------------------------------------------
def find_conjunction(tr0, tr1, miss_dist_threshold):
    """model.py:23-52 on the GPU.  tr0, tr1: [N, 3] positions (torch CPU/CUDA tensors or numpy), metres.
    Returns (i_min:int, d_min:0-d tensor, i_conj:int|None, d_conj:0-d tensor|None); first index on
    ties and NaN-as-minimum exactly as torch.argmin."""
    oi, od = engine.find_conjunction(torch.as_tensor(tr0), torch.as_tensor(tr1), miss_dist_threshold)
    oi, od = oi.cpu(), od.cpu()
    i_min, i_conj = int(oi[0]), int(oi[1])
    if i_conj < 0:
        return i_min, od[0], None, None
    return i_min, od[0], i_conj, od[1]


# --------------------------------------------------------------------------------------------------
# priors (model.py:54-153): mixture parameters fitted by the Kessler authors to the May-2022 catalogue
_PRIOR_TABLE = {
    "mean_motion": dict(
        probs=[0.12375596165657043, 0.05202080309391022, 0.21220888197422028, 0.0373813770711422, 0.01674230769276619,
               0.5578906536102295],
        loc=[0.0010028142482042313, 0.00017592836171388626, 0.0010926761478185654, 0.0003353552892804146,
             0.0007777251303195953, 0.001032940074801445],
        scale=[0.00004670943133533001, 0.00003172305878251791, 0.000027726569678634405, 0.00007733114063739777,
               0.00013636205345392227, 0.00002651428570970893], low=0.0, high=0.004),
    "eccentricity": dict(
        probs=[0.5433819890022278, 0.04530993849039078, 0.08378008753061295, 0.02705608867108822, 0.03350389748811722,
               0.2669680118560791],
        loc=[0.0028987403493374586, 0.6150050163269043, 0.05085373669862747, 0.3420163094997406, 0.7167646288871765,
             0.013545362278819084],
        scale=[0.002526970813050866, 0.07872536778450012, 0.024748045951128006, 0.18968918919563293,
               0.011966796591877937, 0.0068586356937885284], low=0.0, high=0.8999999761581421),
    "inclination": dict(
        probs=[0.028989605605602264, 0.10272273421287537, 0.02265254408121109, 0.019256576895713806, 0.028676774352788925,
               0.06484941393136978, 0.13786117732524872, 0.0010146398562937975, 0.047179922461509705, 0.01607278548181057,
               0.020023610442876816, 0.06644929945468903, 0.4442509114742279],
        loc=[0.09954200685024261, 1.4393062591552734, 1.736578106880188, 1.0963480472564697, 0.48166394233703613,
             0.9063634872436523, 1.275956392288208, 2.5208728313446045, 1.5189905166625977, 0.3474450707435608,
             0.6648743152618408, 1.1465401649475098, 1.7207987308502197],
        scale=[0.04205162078142166, 0.012214339338243008, 0.11822951585054398, 0.010178830474615097, 0.04073172062635422,
               0.04156989976763725, 0.02754846028983593, 0.003279004478827119, 0.02461068518459797, 0.0433642603456974,
               0.11472384631633759, 0.014345825649797916, 0.012212350033223629], low=0.0, high=math.pi),
    "b_star": dict(
        probs=[0.9688150882720947, 0.0012630978599190712, 0.024090370163321495, 0.0009446446783840656, 0.0048867943696677685],
        loc=[0.0002002030232688412, 0.3039868175983429, 0.003936616238206625, -0.04726288095116615, 0.08823495358228683],
        scale=[0.0011279708705842495, 0.06403032690286636, 0.012939595617353916, 0.17036935687065125, 0.061987943947315216]),
}


def default_prior():
    """model.py:54-153: dictionary of priors over the TLE elements."""
    p = {}
    for key in ("mean_motion", "eccentricity", "inclination"):
        t = _PRIOR_TABLE[key]
        p[key + "_prior"] = MixtureSameFamily(
            mixture_distribution=Categorical(probs=torch.tensor(t["probs"])),
            component_distribution=util.TruncatedNormal(low=t["low"], high=t["high"], loc=torch.tensor(t["loc"]),
                                                        scale=torch.tensor(t["scale"])))
    t = _PRIOR_TABLE["b_star"]
    p["b_star_prior"] = MixtureSameFamily(mixture_distribution=Categorical(probs=torch.tensor(t["probs"])),
                                          component_distribution=Normal(loc=torch.tensor(t["loc"]), scale=torch.tensor(t["scale"])))
    for key in ("mean_anomaly", "argument_of_perigee", "raan"):
        p[key + "_prior"] = Uniform(low=0.0, high=2 * torch.pi)
    p["mean_motion_first_derivative_prior"] = Uniform(4.937096738377722e-13, 5.807570136601159e-13)
    return p


# --------------------------------------------------------------------------------------------------
class _Replay:
    """Distribution wrapper that yields a pre-drawn value once (used to replay the accepted
    candidate of a batched rejection step through the traced forward())."""

    def __init__(self, fn, value):
        self._fn, self._value = fn, value

    def sample(self, sample_shape=torch.Size()):
        return self._value

    def log_prob(self, value):
        return self._fn.log_prob(value)

    def __getattr__(self, k):
        return getattr(self._fn, k)


def _round_decimals(x, decimals):
    """float('{:.{d}f}'.format(v)) for every v, vectorised: rint(v * 10^d) / 10^d agrees with the correctly
    rounded decimal formatting except when v * 10^d sits within binary rounding noise of a half-integer; those
    (rare) entries are formatted one by one."""
    x = np.asarray(x, dtype=np.float64)
    scaled = x * (10.0 ** decimals)
    out = np.rint(scaled) / (10.0 ** decimals)
    frac = np.abs(scaled - np.floor(scaled) - 0.5)
    # x * 10^d carries one rounding (<= 1.2e-16 relative): only entries that close to a tie can differ
    risky = np.nonzero(~(frac > 1e-14 * np.maximum(1.0, np.abs(scaled))) | ~np.isfinite(scaled))[0]
    for i in risky:
        out[i] = float("{:.{d}f}".format(x[i], d=decimals))
    return out


def _quantize_exp_field(x):
    """TLE 'decimal point assumed' field (B*): 5-digit mantissa in [0.1, 1) and a one-digit exponent."""
    from .tle import _format_exp_field, _parse_exp_field
    x = np.asarray(x, dtype=np.float64)
    out = np.zeros_like(x)
    nz = np.nonzero((x != 0.0) & np.isfinite(x))[0]
    if len(nz):
        ax = np.abs(x[nz])
        ex = np.floor(np.log10(ax)).astype(np.int64) + 1
        scaled = ax / np.array([10.0 ** int(e) for e in range(-300, 301)])[np.clip(ex, -300, 300) + 300] * 1e5   # as _format_exp_field
        mant = np.rint(scaled)
        frac = np.abs(scaled - np.floor(scaled) - 0.5)
        # decade edges, half-way mantissas and exponents outside one digit go through the text formatter
        risky = (frac < 1e-9) | (mant >= 100000) | (mant < 10000) | (np.abs(ex) > 8)
        pow10 = np.array([10.0 ** e for e in range(-9, 10)])          # Python's pow, as tle._parse_exp_field uses
        val = np.sign(x[nz]) * (mant / 1e5) * pow10[np.clip(ex, -9, 9) + 9]
        for i in np.nonzero(risky)[0]:
            val[i] = _parse_exp_field(_format_exp_field(x[nz][i]))
        # (mant / 1e5) * 10^ex must equal float('0.' + digits) * 10^ex: same two roundings as the parser
        out[nz] = val
    return out


def quantize_sgp4_inputs(mean_motion_rad_s, ecc, incl, raan, argp, mean_anomaly, b_star):
    """What ``TLE(dict)`` does to the SGP4 inputs, vectorised: every element is written into the TLE
    text (8.4f degrees, 7-digit eccentricity, 11.8f rev/day, 5-digit B* mantissa) and parsed back
    (tests/test_model_host.py checks this against TLE(dict) element-wise).
    Returns SGP4 element SoA [7, n] (rad/min, rad)."""
    deg = math.pi / 180.0
    q4 = lambda x: _round_decimals(np.asarray(x, dtype=np.float64) / deg, 4) * deg
    rev = _round_decimals(np.asarray(mean_motion_rad_s, dtype=np.float64) * 86400.0 / (2 * math.pi), 8)
    e = _round_decimals(ecc, 7)
    e = np.where(e >= 1.0, 0.9999999, e)
    return np.stack([rev * (2 * math.pi) / 86400.0 * 60.0, e, q4(incl), q4(raan), q4(argp), q4(mean_anomaly),
                     _quantize_exp_field(b_star)])


def _tdiv(x, c):
    """x / c with IEEE division.  (torch's CUDA kernel for `tensor / python_float` multiplies by the rounded reciprocal,
    which is one ulp off for most divisors; a 0-dim tensor divisor takes the true-division kernel.)"""
    return x / torch.full((), float(c), dtype=x.dtype, device=x.device)


def _round_decimals_torch(x, decimals):
    """_round_decimals on a torch tensor (any device): (rounded, risky mask).  Same arithmetic -- rint(x 10^d) / 10^d --,
    the (rare) entries within binary noise of a decimal tie are reported instead of formatted."""
    p = 10.0 ** decimals
    scaled = x * p
    out = _tdiv(torch.round(scaled), p)
    frac = (scaled - torch.floor(scaled) - 0.5).abs()
    risky = ~(frac > 1e-14 * torch.clamp(scaled.abs(), min=1.0)) | ~torch.isfinite(scaled)
    return out, risky


def quantize_sgp4_inputs_torch(mean_motion_rad_s, ecc, incl, raan, argp, mean_anomaly, b_star_quantised):
    """quantize_sgp4_inputs on torch tensors that stay where they are (the GPU): bit-identical to the NumPy form; entries
    whose decimal rounding is a near-tie (probability ~1e-13 per field) are re-done by the text formatter on the host.
    `b_star_quantised`: B* already through _quantize_exp_field (it is per object, not per sample).  Returns [7, n]."""
    deg = math.pi / 180.0
    risky = None
    cols = []
    def q(x, d):
        nonlocal risky
        out, r = _round_decimals_torch(x, d)
        risky = r if risky is None else (risky | r)
        return out
    rev = q(_tdiv(mean_motion_rad_s * 86400.0, 2 * math.pi), 8)
    e = q(ecc, 7)
    e = torch.where(e >= 1.0, torch.full_like(e, 0.9999999), e)
    cols = [_tdiv(rev * (2 * math.pi), 86400.0) * 60.0, e] + [q(_tdiv(a, deg), 4) * deg for a in (incl, raan, argp, mean_anomaly)] + [b_star_quantised]
    out = torch.stack(cols)
    bad = torch.nonzero(risky).reshape(-1)
    if bad.numel():
        h = lambda t: t[bad].cpu().numpy()
        fixed = quantize_sgp4_inputs(h(mean_motion_rad_s), h(ecc), h(incl), h(raan), h(argp), h(mean_anomaly), np.zeros(bad.numel()))
        fixed[6] = h(b_star_quantised)
        out[:, bad] = torch.from_numpy(fixed).to(out.device)
    return out


class Conjunction:
    """Generative model of a conjunction event (model.py:156-755)."""

    def __init__(self, time0=58991.90384230018, max_duration_days=7.0, time_resolution=6e5, time_upsample_factor=100,
                 miss_dist_threshold=5e3, prior_dict=None, t_prob_new_obs=0.96, c_prob_new_obs=0.4,
                 cdm_update_every_hours=8., mc_samples=100, mc_upsample_factor=100, pc_method='MC', up_method='MC',
                 collision_threshold=70,
                 likelihood_t_stddev=[3.71068006e+02, 9.99999999e-02, 1.72560879e-01],
                 likelihood_c_stddev=[3.71068006e+02, 9.99999999e-02, 1.72560879e-01],
                 likelihood_time_to_tca_stddev=0.7, t_observing_instruments=None, c_observing_instruments=None,
                 t_tle=None, c_tle=None):
        self._time0 = time0
        self._max_duration_days = max_duration_days
        self._time_resolution = time_resolution
        self._time_upsample_factor = time_upsample_factor
        self._delta_time = max_duration_days / time_resolution
        self._miss_dist_threshold = miss_dist_threshold
        self._prior_dict = prior_dict or default_prior()
        self._t_prob_new_obs = t_prob_new_obs
        self._c_prob_new_obs = c_prob_new_obs
        self._cdm_update_every_hours = cdm_update_every_hours
        self._mc_samples = mc_samples
        self._mc_upsample_factor = mc_upsample_factor
        if pc_method not in ['MC', 'FOSTER-1992']:
            raise ValueError(f"Unknown method for probability of collision: {pc_method}. Currently, we only support MC and FOSTER-1992.")
        self._pc_method = pc_method
        if up_method not in ['MC', 'STM']:
            raise ValueError(f"Unknown method for uncertainty propagation: {up_method}. Currently, we only support MC and STM.")
        self._up_method = up_method
        self._collision_threshold = collision_threshold
        self._likelihood_t_stddev = likelihood_t_stddev
        self._likelihood_c_stddev = likelihood_c_stddev
        self._likelihood_time_to_tca_stddev = likelihood_time_to_tca_stddev
        zero_bias = np.array([[0., 0., 0.], [0., 0., 0.]])
        if t_observing_instruments is None:
            t_observing_instruments = [GNSS({'bias_xyz': zero_bias, 'covariance_rtn': np.array(
                [1e-9, 1.115849341564346, 0.059309835843067, 1e-9, 1e-9, 1e-9]) ** 2})]
            print(f"No observing instruments for target, using default one with diagonal covariance {t_observing_instruments[0]._instrument_characteristics['covariance_rtn']}")
        if c_observing_instruments is None:
            c_observing_instruments = [Radar({'bias_xyz': zero_bias, 'covariance_rtn': np.array(
                [1.9628939405514678, 2.2307686944695706, 0.9660907831563862, 1e-9, 1e-9, 1e-9]) ** 2})]
            print(f"No observing instruments for chaser, using default one with diagonal covariance {c_observing_instruments[0]._instrument_characteristics['covariance_rtn']}")
        if len(t_observing_instruments) == 0 or len(c_observing_instruments) == 0:
            raise ValueError("We need at least one observing instrument for target and chaser!")
        self._t_observing_instruments = t_observing_instruments
        self._c_observing_instruments = c_observing_instruments
        self._t_tle = t_tle
        self._c_tle = c_tle
        self._mu_earth = MU_EARTH
        self._replay = {}
        self._times = None

    # ---- sampling helpers ------------------------------------------------------------------------
    def _draw(self, name, fn):
        if name in self._replay:
            fn = _Replay(fn, self._replay[name])
        return ppl.sample(name, fn)

    def time_grid(self):
        """model.py:574 (cached): the fine grid of MJDs."""
        if self._times is None:
            self._times = np.arange(self._time0, self._time0 + self._max_duration_days, self._delta_time)
        return self._times

    _STATIC_TLE_FIELDS = dict(satellite_catalog_number=43437, classification='U', international_designator='18100A',
                              ephemeris_type=0, element_number=9996, revolution_number_at_epoch=56353)

    def _make_object(self, who, given_tle):
        """model.py:245-319: draw a TLE from the priors, or re-draw only the mean anomaly of a given one."""
        pr = self._prior_dict
        if given_tle is None:
            d = {}
            d['mean_motion'] = self._draw(f'{who}_mean_motion', pr['mean_motion_prior'])
            d['mean_anomaly'] = self._draw(f'{who}_mean_anomaly', pr['mean_anomaly_prior'])
            d['eccentricity'] = self._draw(f'{who}_eccentricity', pr['eccentricity_prior'])
            d['inclination'] = self._draw(f'{who}_inclination', pr['inclination_prior'])
            d['argument_of_perigee'] = self._draw(f'{who}_argument_of_perigee', pr['argument_of_perigee_prior'])
            d['raan'] = self._draw(f'{who}_raan', pr['raan_prior'])
            d['mean_motion_first_derivative'] = self._draw(f'{who}_mean_motion_first_derivative', pr['mean_motion_first_derivative_prior'])
            d['mean_motion_second_derivative'] = 0.0
            ppl.deterministic(f'{who}_mean_motion_second_derivative', torch.tensor(d['mean_motion_second_derivative']))
            d['b_star'] = self._draw(f'{who}_b_star', pr['b_star_prior'])
            d.update(self._STATIC_TLE_FIELDS)
            d['epoch_year'] = from_mjd_to_datetime(self._time0).year
            d['epoch_days'] = from_mjd_to_epoch_days_after_1_jan(self._time0)
            return TLE({k: (float(v) if torch.is_tensor(v) else v) for k, v in d.items()})
        mean_anomaly = self._draw(f'{who}_mean_anomaly', pr['mean_anomaly_prior'])
        tle = given_tle.copy()
        tle.update({'mean_anomaly': mean_anomaly})
        self._record_tle_sites(who, tle)
        return tle

    @staticmethod
    def _record_tle_sites(who, tle):
        for site, attr in (('mean_motion', 'mean_motion'), ('eccentricity', 'eccentricity'), ('inclination', 'inclination'),
                           ('argument_of_perigee', 'argument_of_perigee'), ('raan', 'raan'),
                           ('mean_motion_first_derivative', 'mean_motion_first_derivative'),
                           ('mean_motion_second_derivative', 'mean_motion_second_derivative'), ('b_star', 'b_star')):
            ppl.deterministic(f'{who}_{site}', torch.tensor(float(getattr(tle, attr)), dtype=torch.float64))

    def make_target(self):
        return self._make_object('t', self._t_tle)

    def make_chaser(self):
        return self._make_object('c', self._c_tle)

    # ---- ground segment ----------------------------------------------------------------------------
    def _describe_object(self, cdm, oid, tle, given, role):
        if given:
            cdm.set_object(oid, 'OBJECT_DESIGNATOR', tle.international_designator)
            cdm.set_object(oid, 'INTERNATIONAL_DESIGNATOR', tle.international_designator)
            cdm.set_object(oid, 'OBJECT_NAME', tle.name if hasattr(tle, "name") else 'KESSLER_SOFTWARE_' + role)
            cdm.set_object(oid, 'CATALOG_NAME', tle.satellite_catalog_number)
        else:
            cdm.set_object(oid, 'OBJECT_DESIGNATOR', 'KESSLER_SOFTWARE_' + str(uuid.uuid1()))
            cdm.set_object(oid, 'INTERNATIONAL_DESIGNATOR', 'UNKNOWN')
            cdm.set_object(oid, 'OBJECT_NAME', 'KESSLER_SOFTWARE_' + role)
            cdm.set_object(oid, 'CATALOG_NAME', 'UNKNOWN')
        for key, val in (('EPHEMERIS_NAME', 'NONE'), ('COVARIANCE_METHOD', 'CALCULATED'), ('MANEUVERABLE', 'N/A'),
                         ('REF_FRAME', 'ITRF'), ('ORBIT_CENTER', 'EARTH')):
            cdm.set_object(oid, key, val)

    def generate_cdm(self, t_state_new_obs, c_state_new_obs, time_obs_mjd, time_conj_mjd, t_tle, c_tle, previous_cdm):
        """model.py:321-443: one CDM from the latest observations (None when neither object was observed)."""
        if c_state_new_obs is None and t_state_new_obs is None:
            return None
        if previous_cdm:
            cdm = previous_cdm.copy()
        else:
            cdm = ConjunctionDataMessage()
            cdm.set_header('ORIGINATOR', 'KESSLER_SOFTWARE')
            from_population = hasattr(self, '_tles')
            self._describe_object(cdm, 0, t_tle, self._t_tle is not None or from_population, 'TARGET')
            self._describe_object(cdm, 1, c_tle, self._c_tle is not None or from_population, 'CHASER')
        cdm.set_header('MESSAGE_ID', 'KESSLER_SOFTWARE_{}'.format(str(uuid.uuid1())))
        tca_jd = util.from_mjd_to_jd(time_conj_mjd)
        obs_jd = util.from_mjd_to_jd(time_obs_mjd)
        cdm.set_relative_metadata('TCA', util.from_jd_to_cdm_datetime_str(tca_jd))
        cdm.set_header('CREATION_DATE', util.from_jd_to_cdm_datetime_str(obs_jd))
        for oid, state_obs, tle, cov_diag in ((0, t_state_new_obs, t_tle, getattr(self, '_t_cov_matrix_diagonal_obs_noise', None)),
                                              (1, c_state_new_obs, c_tle, getattr(self, '_c_cov_matrix_diagonal_obs_noise', None))):
            if state_obs is None:
                continue
            obs_tle, _ = sgp4.newton_method(tle, time_obs_mjd, verbose=False)
            if self._up_method == 'MC':
                mean_teme, cov_rtn = self.propagate_uncertainty_monte_carlo(state_xyz=state_obs, cov_matrix_diagonal_rtn=cov_diag,
                                                                            time_ca_mjd=time_conj_mjd, obs_tle=obs_tle)
                cdm.set_state(oid, util.from_TEME_to_ITRF(mean_teme, tca_jd) / 1e3)
                cdm.set_covariance(oid, cov_rtn)
        if self._pc_method == 'MC':
            t_state_tca_rtn, _ = util.from_cartesian_to_rtn(cdm.get_state(0) * 1e3)
            c_state_tca_rtn, _ = util.from_cartesian_to_rtn(cdm.get_state(1) * 1e3)
            n_pc = int(self._mc_samples * self._mc_upsample_factor)
            t_samples = torch.distributions.MultivariateNormal(torch.tensor(t_state_tca_rtn[0]), torch.tensor(cdm.get_covariance(0)[:3, :3])).sample((n_pc,))
            c_samples = torch.distributions.MultivariateNormal(torch.tensor(c_state_tca_rtn[0]), torch.tensor(cdm.get_covariance(1)[:3, :3])).sample((n_pc,))
            # model.py:432-437: all-vs-all miss distances below the threshold, counted on the GPU
            count = engine.pc_count(t_samples.T.contiguous(), c_samples.T.contiguous(), self._collision_threshold)
            probability_of_collision = int(count[0]) / float(n_pc * n_pc)
        else:
            raise NotImplementedError("pc_method 'FOSTER-1992' is accepted by the constructor but not implemented (as in the reference)")
        cdm.set_relative_metadata('COLLISION_PROBABILITY', float(probability_of_collision))
        cdm.set_relative_metadata('COLLISION_PROBABILITY_METHOD', self._pc_method)
        return cdm

    def tle_covariance(self, state_xyz, cov_matrix_diagonal_rtn):
        """model.py:461-478: RTN diagonal -> Cartesian -> (a, e, i, Omega, omega, M) covariance."""
        _, R = util.from_cartesian_to_rtn(state_xyz)
        T = np.zeros((6, 6))
        T[0:3, 0:3] = R
        T[3:, 3:] = R
        C_xyz = np.matmul(np.matmul(T.T, np.diag(cov_matrix_diagonal_rtn)), T)
        # util.keplerian_cartesian_partials (six autograd passes on the CPU) -> one dual-number kernel
        _, jac = engine.kepler_partials(np.asarray(state_xyz, dtype=np.float64).reshape(1, 6), self._mu_earth)
        dtle_dx = jac[0].cpu().numpy()
        cov = np.matmul(np.matmul(dtle_dx, C_xyz), dtle_dx.T)
        if ~torch.all(torch.from_numpy(np.real(np.linalg.eig(cov)[0])) > 1e-11):
            cov = 0.5 * (cov + cov.T)
            cov += 1e-10 * np.eye(cov.shape[0])
        return cov

    def propagate_uncertainty_monte_carlo(self, state_xyz, cov_matrix_diagonal_rtn, time_ca_mjd, obs_tle):
        """model.py:445-551.  Returns (mean state at TCA [2,3] m TEME, covariance [6,6] in the RTN frame
        of that mean).  The sample loop (100 TLE objects, propagate_batch, per-sample rotation, np.cov)
        is one ksl_tca_moments launch."""
        Cov_tle = self.tle_covariance(state_xyz, cov_matrix_diagonal_rtn)
        mean_tle_els = torch.tensor([float(obs_tle.semi_major_axis), obs_tle._ecco, obs_tle._inclo, obs_tle._nodeo,
                                     obs_tle._argpo, obs_tle._mo], dtype=torch.float64)
        epoch_mjd = sgp4.from_datetime_to_mjd(obs_tle._epoch)
        tsince = (time_ca_mjd - epoch_mjd) * 1440.
        ndot, bstar = obs_tle.mean_motion_first_derivative, obs_tle.b_star
        try:
            mvn = torch.distributions.MultivariateNormal(loc=mean_tle_els, covariance_matrix=torch.tensor(Cov_tle))
            samples = mvn.sample((self._mc_samples,))
            samples[:, 0] = (self._mu_earth / samples[:, 0] ** 3) ** (0.5)
        except Exception as e:
            if "Expected parameter covariance_matrix" not in str(e):
                raise e
            ecc = 0. if mean_tle_els[1] < 0. else float(mean_tle_els[1])
            el = quantize_sgp4_inputs([(self._mu_earth / float(mean_tle_els[0]) ** 3) ** 0.5], [ecc], [float(mean_tle_els[2])],
                                      [float(mean_tle_els[3])], [float(mean_tle_els[4])], [float(mean_tle_els[5])], [bstar])
            _, _, st = engine.tca_moments(el, tsince, np.zeros(6), want_states=True)
            return st.cpu().numpy().reshape(2, 3), np.diag(cov_matrix_diagonal_rtn)
        s = samples.numpy()
        ecc = np.where(s[:, 1] < 0., 0., s[:, 1])
        el = quantize_sgp4_inputs(s[:, 0], ecc, s[:, 2], s[:, 3], s[:, 4], s[:, 5], np.full(len(s), bstar))
        _ = ndot
        if len(s) <= 65536:
            _, errc, states = engine.tca_moments(el, tsince, np.zeros(6), want_states=True)
            mc_states = states.cpu().numpy().reshape(-1, 2, 3)
            mean_state = mc_states.mean(axis=0)
            R = util.rotation_matrix(mean_state)
            rtn = np.concatenate([mc_states[:, 0] @ R.T, mc_states[:, 1] @ R.T], axis=1)
            return mean_state, np.cov(rtn.transpose())
        ref_el = quantize_sgp4_inputs([(self._mu_earth / float(mean_tle_els[0]) ** 3) ** 0.5], [max(float(mean_tle_els[1]), 0.0)],
                                      [float(mean_tle_els[2])], [float(mean_tle_els[3])], [float(mean_tle_els[4])],
                                      [float(mean_tle_els[5])], [bstar])
        _, _, ref = engine.tca_moments(ref_el, tsince, np.zeros(6), want_states=True)
        ref = ref.cpu().numpy().reshape(6)
        mom, _, _ = engine.tca_moments(el, tsince, ref)
        from .stats import moments_to_mean_cov_rtn
        mean, cov = moments_to_mean_cov_rtn(mom.cpu().numpy(), ref)
        return mean.reshape(2, 3), cov

    # ---- the model -----------------------------------------------------------------------------------
    def _prefilter_rejects(self, t_tle, c_tle):
        """model.py:567: apogee/perigee shells that cannot come within the threshold."""
        margin = self._miss_dist_threshold + 1000
        return ((t_tle.apogee_alt() + margin) < c_tle.perigee_alt()) or ((c_tle.apogee_alt() + margin) < t_tle.perigee_alt())

    def _observe(self, instruments, time_cdm, state):
        states, covs = [], []
        for instrument in instruments:
            s, c = instrument.observe(time_cdm, {'state_xyz': state})
            states.append(s)
            covs.append(c)
        return np.stack(states).mean(0), np.stack(covs).mean(0).squeeze()

    def forward(self):
        """model.py:554-732, one trace."""
        t_tle = self.make_target()
        c_tle = self.make_chaser()
        if self._prefilter_rejects(t_tle, c_tle):
            ppl.deterministic('conj', torch.tensor(False))
            return
        times = self.time_grid()
        ppl.deterministic('time0', torch.tensor(self._time0))
        ppl.deterministic('max_duration_days', torch.tensor(self._max_duration_days))
        ppl.deterministic('delta_time', torch.tensor(self._delta_time))

        # space segment: both objects over the window (device-resident [N, 6] metres)
        states = {}
        for who, tle in (('t', t_tle), ('c', c_tle)):
            try:
                sgp4.initialize_tle(tle)
                states[who] = sgp4.propagate_upsample_device(tle, times, self._time_upsample_factor)
            except N.KesslerCudaError:
                raise                       # a missing GPU / library is not a propagation error
            except Exception:
                ppl.deterministic('conj', torch.tensor(False))
                return
            ppl.deterministic(f'{who}_prop_error', torch.tensor(False))
        t_states, c_states = states['t'], states['c']
        i_min, d_min, i_conj, d_conj = find_conjunction(t_states[:, :3], c_states[:, :3], self._miss_dist_threshold)
        time_min = times[i_min]
        conj = False
        if i_conj:                          # sic: index 0 is treated as "no conjunction" (model.py:617)
            conj = True
            ppl.deterministic('time_min', torch.tensor(time_min))
            ppl.deterministic('d_min', d_min)
            ppl.deterministic('i_conj', torch.tensor(i_conj))
            ppl.deterministic('time_conj', torch.tensor(times[i_conj]))
            ppl.deterministic('d_conj', d_conj)

        # ground segment: CDM issuance (model.py:630-692)
        cdms = []
        mc_prop_error = False
        if conj:
            max_ncdm = max(1, int((time_min - self._time0) * 24.0 / self._cdm_update_every_hours))
            ppl.deterministic('max_ncdm', torch.tensor(max_ncdm))
            for i in range(max_ncdm):
                i_cdm, time_cdm = util.find_closest(times, self._time0 + (i * self._cdm_update_every_hours) / 24)
                ppl.deterministic(f'time_cdm_{i}', torch.tensor(time_cdm))
                if i == 0:
                    t_new_obs, c_new_obs = True, True
                else:
                    t_new_obs = self._draw(f't_new_obs_{i}', Bernoulli(self._t_prob_new_obs))
                    c_new_obs = self._draw(f'c_new_obs_{i}', Bernoulli(self._c_prob_new_obs))
                t_state_new_obs = c_state_new_obs = None
                if t_new_obs:
                    truth = t_states[int(i_cdm)].cpu().numpy().reshape(2, 3)
                    t_state_new_obs, self._t_cov_matrix_diagonal_obs_noise = self._observe(self._t_observing_instruments, time_cdm, truth)
                if c_new_obs:
                    truth = c_states[int(i_cdm)].cpu().numpy().reshape(2, 3)
                    c_state_new_obs, self._c_cov_matrix_diagonal_obs_noise = self._observe(self._c_observing_instruments, time_cdm, truth)
                cdm = self.generate_cdm(t_state_new_obs=t_state_new_obs, c_state_new_obs=c_state_new_obs, time_obs_mjd=time_cdm,
                                        time_conj_mjd=time_min, t_tle=t_tle, c_tle=c_tle,
                                        previous_cdm=None if i == 0 else cdms[-1])
                if cdm:
                    cdms.append(cdm)
                ppl.deterministic(f'cdm_issued_{i}', torch.tensor(bool(cdm)))
        ppl.deterministic('mc_prop_error', torch.tensor(mc_prop_error))
        ppl.deterministic('prop_error', torch.tensor(False))
        ppl.deterministic('num_cdms', torch.tensor(len(cdms)))
        ppl.sample("cdms", dist.Delta(torch.tensor(0.)), infer={"cdms": cdms})
        ppl.deterministic('conj', torch.tensor(conj))
        if conj:
            ppl.sample("t_tle", dist.Delta(torch.tensor(0.)), infer={"t_tle": t_tle})
            ppl.sample("c_tle", dist.Delta(torch.tensor(0.)), infer={"c_tle": c_tle})
            self._record_likelihood(cdms[0])

    def _record_likelihood(self, cdm0):
        """model.py:710-732: Kepler elements of the first CDM's (ITRF, sic) states and their likelihood sites."""
        mu = self._mu_earth
        vals = {}
        for who, oid, std in (('t', 0, self._likelihood_t_stddev), ('c', 1, self._likelihood_c_stddev)):
            st = torch.tensor(cdm0.get_state(oid) * 1e3)
            sma, ecc, inc, _, _, _ = util.from_cartesian_to_keplerian(st[0], st[1], mu)
            for name, v, s in (('sma', sma, std[0]), ('ecc', ecc, std[1]), ('inc', inc, std[2])):
                vals[f'cdm0_{who}_{name}'] = (v, s)
        time_to_tca = util.from_date_str_to_days(cdm0['TCA'], date0=cdm0['CREATION_DATE'])
        vals['cdm0_time_to_tca'] = (time_to_tca, self._likelihood_time_to_tca_stddev)
        for k, (v, _) in vals.items():
            ppl.deterministic(k, torch.as_tensor(v, dtype=torch.float64))
        for k, (v, s) in vals.items():
            ppl.sample('obs_' + k, Normal(torch.as_tensor(v, dtype=torch.float64), s), obs=v)

    # ---- batched front end (not in the reference) ------------------------------------------------------
    def _sample_mixture(self, prior, n):
        return prior.sample((n,)).to(torch.float64).numpy()

    def sample_elements(self, who, n):
        """n prior draws for one object -> (dict of raw sampled fields [n], SGP4 element SoA [7,n], epoch_mjd [n])."""
        given = self._t_tle if who == 't' else self._c_tle
        pr = self._prior_dict
        ma = self._sample_mixture(pr['mean_anomaly_prior'], n)
        if given is not None:
            base = given.elements()
            raw = dict(mean_anomaly=ma, mean_motion=np.full(n, given.mean_motion), eccentricity=np.full(n, given.eccentricity),
                       inclination=np.full(n, given.inclination), raan=np.full(n, given.raan),
                       argument_of_perigee=np.full(n, given.argument_of_perigee), b_star=np.full(n, given.b_star))
            el = np.repeat(base[:, None], n, axis=1)
            deg = math.pi / 180.0
            el[5] = _round_decimals(ma / deg, 4) * deg
            raw['_sites'] = ['mean_anomaly']
            return raw, el, np.full(n, given.date_mjd)
        raw = dict(mean_motion=self._sample_mixture(pr['mean_motion_prior'], n), mean_anomaly=ma,
                   eccentricity=self._sample_mixture(pr['eccentricity_prior'], n),
                   inclination=self._sample_mixture(pr['inclination_prior'], n),
                   argument_of_perigee=self._sample_mixture(pr['argument_of_perigee_prior'], n),
                   raan=self._sample_mixture(pr['raan_prior'], n),
                   mean_motion_first_derivative=self._sample_mixture(pr['mean_motion_first_derivative_prior'], n),
                   b_star=self._sample_mixture(pr['b_star_prior'], n))
        el = quantize_sgp4_inputs(raw['mean_motion'], raw['eccentricity'], raw['inclination'], raw['raan'],
                                  raw['argument_of_perigee'], raw['mean_anomaly'], raw['b_star'])
        raw['_sites'] = ['mean_motion', 'mean_anomaly', 'eccentricity', 'inclination', 'argument_of_perigee', 'raan',
                         'mean_motion_first_derivative', 'b_star']
        probe = TLE(dict(self._STATIC_TLE_FIELDS, mean_motion=1e-3, mean_anomaly=0., eccentricity=0., inclination=0.,
                         argument_of_perigee=0., raan=0., mean_motion_first_derivative=0., mean_motion_second_derivative=0.,
                         b_star=0., epoch_year=from_mjd_to_datetime(self._time0).year,
                         epoch_days=from_mjd_to_epoch_days_after_1_jan(self._time0)))
        return raw, el, np.full(n, probe.date_mjd)

    def _device_prior(self):
        """Packed parameters of the prior for ksl_sample_prior, or None when the prior is not made of the
        default prior's families (uniform / mixture of normals / mixture of truncated normals)."""
        cached = getattr(self, "_device_prior_cache", None)
        if cached is not None:
            return None if cached is False else cached
        fields = []
        for name in engine.PRIOR_FIELDS:
            d = self._prior_dict.get(name + "_prior")
            if isinstance(d, Uniform):
                fields.append(dict(kind=0, lo=float(d.low), hi=float(d.high)))
            elif isinstance(d, MixtureSameFamily) and isinstance(d.component_distribution, util.TruncatedNormal):
                c = d.component_distribution
                fields.append(dict(kind=2, probs=d.mixture_distribution.probs.double().numpy(), loc=c._loc.double().numpy(),
                                   scale=c._scale.double().numpy(), lo=float(c._low), hi=float(c._high)))
            elif isinstance(d, MixtureSameFamily) and isinstance(d.component_distribution, Normal):
                c = d.component_distribution
                fields.append(dict(kind=1, probs=d.mixture_distribution.probs.double().numpy(), loc=c.loc.double().numpy(),
                                   scale=c.scale.double().numpy()))
            else:
                self._device_prior_cache = False
                return None
        if any(f["kind"] > 0 and len(f["probs"]) > 16 for f in fields):
            self._device_prior_cache = False
            return None
        self._device_prior_cache = engine.pack_prior(fields)
        return self._device_prior_cache

    def sample_elements_device(self, who, n, seed, offset=0):
        """Device-side counterpart of sample_elements for an object drawn from the full prior: the draws, the
        TLE-text quantisation and nothing else leave no footprint on the host.  Returns (raw dict of numpy [n] incl.
        '_sites', elems numpy [7, n], epoch [n])."""
        params = self._device_prior()
        raw_d, el_d, flag = engine.sample_prior(params, seed, offset, n)
        raw_h, el = raw_d.cpu().numpy(), el_d.cpu().numpy()
        raw = {name: raw_h[k] for k, name in enumerate(engine.PRIOR_FIELDS)}
        bad = np.nonzero(flag.cpu().numpy())[0]
        if len(bad):    # decimal ties (~1e-13 of the draws): let the text formatter decide
            el[:, bad] = quantize_sgp4_inputs(raw['mean_motion'][bad], raw['eccentricity'][bad], raw['inclination'][bad], raw['raan'][bad],
                                              raw['argument_of_perigee'][bad], raw['mean_anomaly'][bad], raw['b_star'][bad])
        raw['_sites'] = ['mean_motion', 'mean_anomaly', 'eccentricity', 'inclination', 'argument_of_perigee', 'raan',
                         'mean_motion_first_derivative', 'b_star']
        return raw, el, np.full(n, self._prior_epoch_mjd())

    def _prior_epoch_mjd(self):
        if getattr(self, "_prior_epoch_cache", None) is None:
            probe = TLE(dict(self._STATIC_TLE_FIELDS, mean_motion=1e-3, mean_anomaly=0., eccentricity=0., inclination=0.,
                             argument_of_perigee=0., raan=0., mean_motion_first_derivative=0., mean_motion_second_derivative=0.,
                             b_star=0., epoch_year=from_mjd_to_datetime(self._time0).year,
                             epoch_days=from_mjd_to_epoch_days_after_1_jan(self._time0)))
            self._prior_epoch_cache = probe.date_mjd
        return self._prior_epoch_cache

    def forward_batch(self, n, mode=N.SCREEN_ANALYTIC, device_sampling=True):
        """n independent draws of the space segment of forward() in one fused launch.  Returns a dict of
        numpy arrays [n]: conj (the reference's `if i_conj:` truthiness), prefiltered, prop_error, i_min,
        d_min, i_conj, d_conj, time_min, plus the raw draws (`t_raw`, `c_raw`) and SGP4 inputs."""
        def draw(who):
            given = self._t_tle if who == 't' else self._c_tle
            full_prior = given is None and type(self).sample_elements is Conjunction.sample_elements
            if device_sampling and full_prior and self._device_prior() is not None:
                seed = int(torch.randint(0, 2 ** 62, (1,)).item())      # torch.manual_seed governs reproducibility
                return self.sample_elements_device(who, n, seed)
            return self.sample_elements(who, n)
        t_raw, el_t, ep_t = draw('t')
        c_raw, el_c, ep_c = draw('c')
        sma = lambda raw: (MU_EARTH / np.maximum(raw['mean_motion'], 1e-300) ** 2) ** (1.0 / 3.0)
        a_t, a_c = sma(t_raw), sma(c_raw)
        apo = lambda a, raw: a * (1 + raw['eccentricity']) - R_EARTH
        per = lambda a, raw: a * (1 - raw['eccentricity']) - R_EARTH
        margin = self._miss_dist_threshold + 1000
        pre = ((apo(a_t, t_raw) + margin) < per(a_c, c_raw)) | ((apo(a_c, c_raw) + margin) < per(a_t, t_raw))
        times = self.time_grid()
        tc = np.ascontiguousarray(times[::max(int(self._time_upsample_factor), 1)])
        keep = np.nonzero(~pre)[0]
        out = dict(prefiltered=pre, prop_error=np.zeros(n, bool), conj=np.zeros(n, bool), i_min=np.full(n, -1, np.int64),
                   d_min=np.full(n, np.nan), i_conj=np.full(n, -1, np.int64), d_conj=np.full(n, np.nan),
                   time_min=np.full(n, np.nan), t_raw=t_raw, c_raw=c_raw, elems_t=el_t, elems_c=el_c, epoch_t=ep_t, epoch_c=ep_c)
        if len(keep):
            r = engine.screen_host(np.ascontiguousarray(el_t[:, keep]), ep_t[keep], np.ascontiguousarray(el_c[:, keep]), ep_c[keep],
                                   tc, len(times), self._miss_dist_threshold, mode, self._collision_threshold)
            deep = ((r['err'] | (r['err'] >> N.ERR_CHASER_SHIFT)) & N.ERR_DEEPSPACE) != 0
            out['prop_error'][keep] = deep
            out['i_min'][keep], out['d_min'][keep] = r['i_min'], r['d_min']
            out['i_conj'][keep], out['d_conj'][keep] = r['i_conj'], r['d_conj']
            out['conj'][keep] = (~deep) & (r['i_conj'] > 0)
            ok = keep[~deep]
            out['time_min'][ok] = times[out['i_min'][ok]]
            out['summary_counts'], out['summary_moments'] = r['counts'], r['moments']
        return out

    # ---- the same, with the candidates living on the GPU from the draw to the accepted few ------------------------------
    def _draw_device(self, who, n, dev):
        """n prior draws for one object on the device: dict(el [7, n], ep [n], raw {site: tensor [n]}, sites [...]),
        everything a cuda tensor.  Covers a given TLE (only the mean anomaly is drawn) and the full default-family prior
        (ksl_sample_prior); anything else returns None and the caller falls back to the host sampler."""
        given = self._t_tle if who == 't' else self._c_tle
        deg = math.pi / 180.0
        ma_prior = self._prior_dict['mean_anomaly_prior']
        if given is not None:
            if not isinstance(ma_prior, Uniform):
                return None
            lo, hi = float(ma_prior.low), float(ma_prior.high)
            ma = lo + (hi - lo) * torch.rand(n, dtype=torch.float64, device=dev)
            q, risky = _round_decimals_torch(_tdiv(ma, deg), 4)
            el = torch.from_numpy(given.elements()).to(dev)[:, None].repeat(1, n)
            el[5] = q * deg
            bad = torch.nonzero(risky).reshape(-1)
            if bad.numel():
                el[5, bad] = torch.from_numpy(_round_decimals(ma[bad].cpu().numpy() / deg, 4) * deg).to(dev)
            full = lambda v: torch.full((n,), float(v), dtype=torch.float64, device=dev)
            raw = dict(mean_anomaly=ma, mean_motion=full(given.mean_motion), eccentricity=full(given.eccentricity))
            return dict(el=el, ep=full(given.date_mjd), raw=raw, sites=['mean_anomaly'])
        if type(self).sample_elements is not Conjunction.sample_elements or self._device_prior() is None:
            return None
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        raw_d, el_d, flag = engine.sample_prior(self._device_prior(), seed, 0, n)
        bad = torch.nonzero(flag).reshape(-1)
        if bad.numel():
            r = raw_d[:, bad].cpu().numpy()
            el_d[:, bad] = torch.from_numpy(quantize_sgp4_inputs(r[0], r[2], r[3], r[5], r[4], r[1], r[7])).to(dev)
        raw = {name: raw_d[k] for k, name in enumerate(engine.PRIOR_FIELDS)}
        sites = ['mean_motion', 'mean_anomaly', 'eccentricity', 'inclination', 'argument_of_perigee', 'raan',
                 'mean_motion_first_derivative', 'b_star']
        return dict(el=el_d, ep=torch.full((n,), self._prior_epoch_mjd(), dtype=torch.float64, device=dev), raw=raw, sites=sites)

    def forward_batch_device(self, n, mode=N.SCREEN_ANALYTIC, max_hits=None):
        """forward_batch for the rejection loop: the n candidates are drawn, pre-filtered (model.py:567), screened and
        compacted ON THE DEVICE; only the accepted draws travel to the host.  Returns the compact dict the batched ground
        segment takes (kessler_b200.events: elems_t/c [7, h], epoch_t/c, time_min, i_min, d_min, i_conj, d_conj, t_raw,
        c_raw with '_sites') and the number of accepted draws; None when this model's priors cannot be sampled on the
        device (the caller then uses forward_batch).  Random streams: torch's CUDA generator (torch.manual_seed)."""
        dev = torch.device("cuda", torch.cuda.current_device())
        t = self._draw_device('t', n, dev)
        c = self._draw_device('c', n, dev) if t is not None else None
        if t is None or c is None:
            return None
        sma = lambda raw: (MU_EARTH / torch.clamp(raw['mean_motion'], min=1e-300) ** 2) ** (1.0 / 3.0)
        a_t, a_c = sma(t['raw']), sma(c['raw'])
        apo = lambda a, raw: a * (1 + raw['eccentricity']) - R_EARTH
        per = lambda a, raw: a * (1 - raw['eccentricity']) - R_EARTH
        margin = self._miss_dist_threshold + 1000
        pre = ((apo(a_t, t['raw']) + margin) < per(a_c, c['raw'])) | ((apo(a_c, c['raw']) + margin) < per(a_t, t['raw']))
        keep = torch.nonzero(~pre).reshape(-1)
        cache = getattr(self, "_grid_cache", None)
        if cache is None or cache[0] != (self._time0, self._max_duration_days, self._time_resolution, self._time_upsample_factor, dev):
            times = self.time_grid()
            tc = np.ascontiguousarray(times[::max(int(self._time_upsample_factor), 1)])
            cache = ((self._time0, self._max_duration_days, self._time_resolution, self._time_upsample_factor, dev), times,
                     torch.from_numpy(tc).to(dev), torch.from_numpy(times).to(dev))
            self._grid_cache = cache
        _, times, tc_d, times_d = cache
        empty = dict(elems_t=np.zeros((7, 0)), elems_c=np.zeros((7, 0)), epoch_t=np.zeros(0), epoch_c=np.zeros(0), time_min=np.zeros(0),
                     i_min=np.zeros(0, np.int64), d_min=np.zeros(0), i_conj=np.zeros(0, np.int64), d_conj=np.zeros(0),
                     t_raw={'_sites': t['sites']}, c_raw={'_sites': c['sites']})
        if keep.numel() == 0:
            return empty, 0
        g = lambda x: x[:, keep].contiguous() if x.dim() == 2 else x[keep].contiguous()
        r = engine.screen_elems(g(t['el']), g(t['ep']), g(c['el']), g(c['ep']), tc_d, len(times), self._miss_dist_threshold, mode)
        deep = ((r['err'] | (r['err'] >> N.ERR_CHASER_SHIFT)) & N.ERR_DEEPSPACE) != 0
        hk = torch.nonzero((~deep) & (r['i_conj'] > 0)).reshape(-1)
        if max_hits is not None:
            hk = hk[:max_hits]
        if hk.numel() == 0:
            return empty, 0
        hits = keep[hk]
        # one packed transfer: [rows, h] doubles (indices < 2^53 are exact in a double)
        rows = [t['el'][:, hits], c['el'][:, hits], t['ep'][hits][None], c['ep'][hits][None],
                r['i_min'][hk][None].to(torch.float64), r['d_min'][hk][None], r['i_conj'][hk][None].to(torch.float64), r['d_conj'][hk][None],
                times_d[r['i_min'][hk]][None]]
        rows += [t['raw'][k][hits][None].to(torch.float64) for k in t['raw']] + [c['raw'][k][hits][None].to(torch.float64) for k in c['raw']]
        host = torch.cat(rows).cpu().numpy()
        out = dict(elems_t=np.ascontiguousarray(host[0:7]), elems_c=np.ascontiguousarray(host[7:14]), epoch_t=host[14].copy(), epoch_c=host[15].copy(),
                   i_min=host[16].astype(np.int64), d_min=host[17].copy(), i_conj=host[18].astype(np.int64), d_conj=host[19].copy(),
                   time_min=host[20].copy())
        o = 21
        for who, d in (('t_raw', t), ('c_raw', c)):
            raw = {}
            for k in d['raw']:
                raw[k] = host[o].astype(np.int64) if k == 'sampled_index' else host[o].copy()
                o += 1
            raw['_sites'] = d['sites']
            out[who] = raw
        return out, int(hk.numel())

    def _replay_values(self, batch, k):
        vals = {}
        for who, raw in (('t', batch['t_raw']), ('c', batch['c_raw'])):
            for nm in raw['_sites']:
                vals[f'{who}_{nm}'] = torch.tensor(raw[nm][k])
        return vals

    def get_conjunction(self, batch_size=None):
        """model.py:734-754: rejection-sample until forward() yields a conjunction; returns
        (trace, iterations).  With ``batch_size`` (default 4096 when the model can be batched) the
        candidates are screened ``batch_size`` at a time on the GPU and only the first accepted one is
        replayed through the traced forward(); iterations counts candidates, as in the reference."""
        if batch_size is None:
            batch_size = 4096
        iteration = 0
        while True:
            if batch_size and batch_size > 1:
                batch = self.forward_batch(batch_size)
                hits = np.nonzero(batch['conj'])[0]
                if len(hits) == 0:
                    iteration += batch_size
                    continue
                k = int(hits[0])
                iteration += k + 1
                self._replay = self._replay_values(batch, k)
            else:
                iteration += 1
            try:
                traced_model = ppl.poutine.trace(self.forward).get_trace()
            finally:
                self._replay = {}
            if traced_model.nodes['conj']['value']:
                break
        print(f"After {iteration} iterations, generated event with {len(traced_model.nodes['cdms']['infer']['cdms'])} CDMs")
        return traced_model, iteration


class ConjunctionSimplified(Conjunction):
    """model.py:757-839: target and chaser are drawn from a TLE population; only the mean anomaly is
    re-sampled."""

    def __init__(self, tles, exclude_object_name=None, *args, **kwargs):
        self._tles = tles
        self._exclude_object_name = exclude_object_name
        if self._exclude_object_name is not None:
            self._include_idxs = [idx for idx, tle in enumerate(self._tles)
                                  if not tle.name.startswith(self._exclude_object_name)]
        super().__init__(*args, **kwargs)

    def _make_object(self, who, given_tle):
        mean_anomaly = self._draw(f'{who}_mean_anomaly', self._prior_dict['mean_anomaly_prior'])
        if given_tle is None:
            if (who == 'c' and self._exclude_object_name is not None and
                    getattr(self, '_t_tle_name', '').startswith(self._exclude_object_name)):
                index_prior = Categorical(torch.tensor(self._include_idxs))
            else:
                index_prior = Categorical(torch.tensor(range(len(self._tles))))
            tle_index = self._draw(f'{who}_sampled_index', index_prior)
            given_tle = self._tles[int(tle_index)]
        if who == 't':
            self._t_tle_name = getattr(given_tle, 'name', '') or ''
        tle = given_tle.copy()
        tle.update({'mean_anomaly': mean_anomaly})
        self._record_tle_sites(who, tle)
        return tle

    def _draw_device(self, who, n, dev):
        """Device-side counterpart of sample_elements: (index, mean anomaly) draws from the population."""
        given = self._t_tle if who == 't' else self._c_tle
        if given is not None:
            return super()._draw_device(who, n, dev)
        ma_prior = self._prior_dict['mean_anomaly_prior']
        if not isinstance(ma_prior, Uniform):
            return None
        tab = getattr(self, "_pop_table", None)
        if tab is None or tab['dev'] != dev:
            f = lambda attr: torch.tensor([float(getattr(t, attr)) for t in self._tles], dtype=torch.float64, device=dev)
            tab = dict(dev=dev, el=torch.from_numpy(np.stack([t.elements() for t in self._tles], axis=1)).to(dev),
                       mean_motion=f('mean_motion'), eccentricity=f('eccentricity'), date_mjd=f('date_mjd'),
                       probs=torch.arange(len(self._tles), dtype=torch.float64, device=dev),
                       excl=torch.tensor([(getattr(t, 'name', '') or '').startswith(self._exclude_object_name or '\0') for t in self._tles], device=dev))
            if self._exclude_object_name is not None:
                tab['probs_incl'] = torch.tensor(self._include_idxs, dtype=torch.float64, device=dev)
            self._pop_table = tab
        deg = math.pi / 180.0
        lo, hi = float(ma_prior.low), float(ma_prior.high)
        ma = lo + (hi - lo) * torch.rand(n, dtype=torch.float64, device=dev)
        idx = torch.multinomial(tab['probs'], n, replacement=True)      # Categorical(tensor(range(len(tles)))): the reference's weights
        if who == 't':
            self._batch_t_idx_dev = idx if self._t_tle is None else None
        elif self._exclude_object_name is not None:
            t_idx = getattr(self, '_batch_t_idx_dev', None)
            if t_idx is not None and t_idx.numel() == n:
                t_excl = tab['excl'][t_idx]
            else:
                t_excl = torch.full((n,), (getattr(self._t_tle, 'name', '') or '').startswith(self._exclude_object_name), device=dev)
            alt = torch.multinomial(tab['probs_incl'], n, replacement=True)
            idx = torch.where(t_excl, alt, idx)
        el = tab['el'][:, idx].contiguous()
        q, risky = _round_decimals_torch(_tdiv(ma, deg), 4)
        el[5] = q * deg
        bad = torch.nonzero(risky).reshape(-1)
        if bad.numel():
            el[5, bad] = torch.from_numpy(_round_decimals(ma[bad].cpu().numpy() / deg, 4) * deg).to(dev)
        raw = dict(mean_anomaly=ma, sampled_index=idx, mean_motion=tab['mean_motion'][idx], eccentricity=tab['eccentricity'][idx])
        return dict(el=el, ep=tab['date_mjd'][idx], raw=raw, sites=['mean_anomaly', 'sampled_index'])

    def sample_elements(self, who, n):
        """Batched counterpart of _make_object: n (index, mean anomaly) draws from the population."""
        given = self._t_tle if who == 't' else self._c_tle
        if given is not None:
            return super().sample_elements(who, n)
        ma = self._prior_dict['mean_anomaly_prior'].sample((n,)).to(torch.float64).numpy()
        idx = Categorical(torch.tensor(range(len(self._tles)))).sample((n,)).numpy()
        if who == 't':
            self._batch_t_idx = idx if self._t_tle is None else None
        elif self._exclude_object_name is not None:
            # model.py:813-814, per draw: when the draw's TARGET carries the excluded name, the chaser index comes from
            # Categorical(tensor(include_idxs)) (the reference hands the indices over as unnormalised probabilities and
            # uses the sampled value as the index: kept as is)
            t_idx = getattr(self, '_batch_t_idx', None)
            if t_idx is not None and len(t_idx) == n:
                t_excl = np.array([getattr(t, 'name', '').startswith(self._exclude_object_name) for t in self._tles])[t_idx]
            else:
                t_excl = np.full(n, getattr(self._t_tle, 'name', '').startswith(self._exclude_object_name))
            if t_excl.any():
                alt = Categorical(torch.tensor(self._include_idxs)).sample((n,)).numpy()
                idx = np.where(t_excl, alt, idx)
        base = np.stack([t.elements() for t in self._tles], axis=1)[:, idx]
        deg = math.pi / 180.0
        base[5] = _round_decimals(ma / deg, 4) * deg
        pick = lambda attr: np.array([getattr(t, attr) for t in self._tles], dtype=np.float64)[idx]
        raw = dict(mean_anomaly=ma, sampled_index=idx, mean_motion=pick('mean_motion'), eccentricity=pick('eccentricity'),
                   _sites=['mean_anomaly', 'sampled_index'])
        return raw, base, pick('date_mjd')

