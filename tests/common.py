"""Shared synthetic inputs for the parity tests (seeded; identical for oracle and GPU)."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EL_KEYS = ("_no_kozai", "_ecco", "_inclo", "_nodeo", "_argpo", "_mo", "_bstar")

# /root/reference/tests/test_model.py:20-26 -- the "default LEO pair" (pair A)
PAIR_A = (("1 44227U 19026C   22068.79876951  .00010731  00000-0  41303-3 0  9993",
           "2 44227  40.0221 252.2030 0008096   5.2961 354.7926 15.26135826158481"),
          ("1 44229U 19026E   22068.90017356  .00004812  00000-0  20383-3 0  9992",
           "2 44229  40.0180 261.5261 0008532 356.1827   3.8908 15.23652474158314"))
# /root/reference/tests/test_util.py:106-109 (COSMOS 2251 DEB); poliastro answers at :119-124
MU_POLIASTRO = 398600441800000.0
KEPLER_34427 = (7023679.5817881366237998, 0.0041649630912143, 1.2919744331609118, 5.3551396410293757,
                0.4409272281996022, 5.8458093777349722)
TLE_34427 = ("1 34427U 93036RU  22068.94647328  .00008100  00000-0  11455-2 0  9999",
             "2 34427  74.0145 306.8269 0033346  13.0723 347.1308 14.76870515693886")


def golden_trace():
    with open(os.path.join(ROOT, "tests", "golden", "trace_golden.json")) as fh:
        return json.load(fh)


def golden_pair():
    """(elems_t[7], epoch_t, elems_c[7], epoch_c, sites) of docs/notebooks/trace.pickle"""
    g = golden_trace()
    t, c = g["tles"]["t_tle"], g["tles"]["c_tle"]
    return (np.array([t[k] for k in EL_KEYS]), t["date_mjd"], np.array([c[k] for k in EL_KEYS]), c["date_mjd"], g["sites"])


def population_elems():
    """The 20 LEO TLEs of docs/notebooks/tles_sample_population.txt, as parsed numbers
    (fixture: tests/golden/population.json, made by make_golden.py)."""
    with open(os.path.join(ROOT, "tests", "golden", "population.json")) as fh:
        pop = json.load(fh)
    el = np.array([[p[k] for k in ("no_kozai", "ecco", "inclo", "nodeo", "argpo", "mo", "bstar")] for p in pop]).T.copy()
    ep = np.array([p["epoch_mjd"] for p in pop])
    return el, ep


def synthetic_catalog(n, seed=3, include_edge=True):
    """Prior-like synthetic catalog (SURVEY.md Sec. 8d C4 recipe, simplified): LEO-heavy
    mixture plus the branches no fixture exercises -- deep space (period >= 225 min),
    perigee < 156 / < 98 km (sfour adjustment), isimp, ecc <= 1e-4, high eccentricity."""
    rng = np.random.default_rng(seed)
    n_rad_s = np.where(rng.uniform(size=n) < 0.85, rng.normal(1.03e-3, 4e-5, n), rng.normal(7.8e-4, 1.4e-4, n))
    n_rad_s = np.clip(n_rad_s, 2.0e-4, 1.2e-3)
    ecc = np.where(rng.uniform(size=n) < 0.8, np.abs(rng.normal(0.0029, 0.0025, n)), rng.uniform(0.0, 0.3, n))
    inc = rng.uniform(0.0, np.pi, n)
    raan, argp, mo = rng.uniform(0, 2 * np.pi, (3, n))
    bstar = np.where(rng.uniform(size=n) < 0.95, rng.normal(2e-4, 1.1e-3, n), rng.normal(0.0, 0.05, n))
    el = np.stack([n_rad_s * 60.0, ecc, inc, raan, argp, mo, bstar])
    if include_edge and n >= 16:
        el[0, 0] = 1.76e-4 * 60.0          # deep space (period 595 min)
        el[0, 1] = 3.35e-4 * 60.0          # deep space (period 312 min)
        el[1, 2] = 0.0                     # ecc = 0 (cc3 = xmcof = 0)
        el[1, 3] = 5e-5                    # ecc <= 1e-4
        el[0, 4], el[1, 4] = 0.07182, 0.0005   # perigee ~ 147 km  (sfour = perige - 78)
        el[0, 5], el[1, 5] = 0.07300, 0.0005   # perigee ~ 77 km   (sfour = 20)
        el[0, 6], el[1, 6] = 0.07084, 0.0005   # perigee ~ 207 km  (isimp)
        el[0, 7], el[1, 7] = 0.03, 0.72        # prior's high-eccentricity component: perigee underground
        el[2, 8] = np.pi                   # retrograde equatorial: cosio = -1 (xlcof guard)
        el[6, 9] = 0.3                     # huge drag term
        el[1, 10] = 0.95                   # nearly parabolic
    return el


def perturbed_samples(nominal, n, seed, sig=(2e-7, 1e-6, 2e-6, 2e-6, 5e-5, 5e-5)):
    """n element sets around `nominal` [7]: relative noise on n, absolute on the rest; e clamped >= 0."""
    rng = np.random.default_rng(seed)
    el = np.repeat(np.asarray(nominal, dtype=np.float64)[:, None], n, 1)
    el[0] *= 1.0 + sig[0] * rng.standard_normal(n)
    for k in range(1, 6):
        el[k] += sig[k] * rng.standard_normal(n)
    el[1] = np.maximum(el[1], 0.0)
    return el
