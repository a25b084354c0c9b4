"""Synthetic inputs for BASELINE.json's configs: seeded, generated on the
host, identical for the oracle and the GPU path.  The nominal pair is the "default LEO pair"
of /root/reference/tests/test_model.py:20-26."""
import numpy as np
import torch

from . import util
from .tle import TLE, MU_EARTH

PAIR_A = (("0 ELECTRON KICK STAGE R/B",
           "1 44227U 19026C   22068.79876951  .00010731  00000-0  41303-3 0  9993",
           "2 44227  40.0221 252.2030 0008096   5.2961 354.7926 15.26135826158481"),
          ("0 HARBINGER",
           "1 44229U 19026E   22068.90017356  .00004812  00000-0  20383-3 0  9992",
           "2 44229  40.0180 261.5261 0008532 356.1827   3.8908 15.23652474158314"))

# model.py:231,235: default instrument diagonals (RTN standard deviations, squared)
T_COV_RTN = np.array([1e-9, 1.115849341564346, 0.059309835843067, 1e-9, 1e-9, 1e-9]) ** 2
C_COV_RTN = np.array([1.9628939405514678, 2.2307686944695706, 0.9660907831563862, 1e-9, 1e-9, 1e-9]) ** 2


def pair_a():
    return TLE(list(PAIR_A[0])), TLE(list(PAIR_A[1]))


def time_grid(time0, max_duration_days=7.0, time_resolution=6e5):
    """model.py:208,574."""
    return np.arange(time0, time0 + max_duration_days, max_duration_days / time_resolution)


def tle_covariance(state_xyz, cov_matrix_diagonal_rtn, mu_earth=MU_EARTH):
    """model.py:461-478: RTN diagonal -> Cartesian -> (a,e,i,Omega,omega,M) covariance."""
    _, R = util.from_cartesian_to_rtn(state_xyz)
    T = np.zeros((6, 6))
    T[0:3, 0:3] = R
    T[3:, 3:] = R
    C_xyz = np.matmul(np.matmul(T.T, np.diag(cov_matrix_diagonal_rtn)), T)
    J = util.keplerian_cartesian_partials(torch.tensor(state_xyz), mu_earth)
    cov = np.matmul(np.matmul(J, C_xyz), J.T)
    if not np.all(np.real(np.linalg.eig(cov)[0]) > 1e-11):
        cov = 0.5 * (cov + cov.T)
        cov += 1e-10 * np.eye(6)
    return cov


def mc_element_samples(tle, state_xyz_m, cov_matrix_diagonal_rtn, n, generator=None, mu_earth=MU_EARTH):
    """model.py:480-533 vectorised: n draws of (a,e,i,Omega,omega,M) ~ MVN(mean, Cov_tle), a ->
    mean motion, e < 0 -> 0.  Returns SGP4 element SoA [7, n] (rad/min, rad; bstar shared)."""
    cov = torch.tensor(tle_covariance(state_xyz_m, cov_matrix_diagonal_rtn, mu_earth))
    mean = torch.tensor([tle.semi_major_axis, tle._ecco, tle._inclo, tle._nodeo, tle._argpo, tle._mo], dtype=torch.float64)
    L = torch.linalg.cholesky(cov)
    z = torch.randn((n, 6), dtype=torch.float64, generator=generator)
    samples = mean + z @ L.T
    out = torch.empty((7, n), dtype=torch.float64)
    out[0] = (mu_earth / samples[:, 0] ** 3) ** 0.5 * 60.0
    out[1] = torch.clamp(samples[:, 1], min=0.0)
    out[2:6] = samples[:, 2:6].T
    out[6] = tle._bstar
    return out


def config2_batch(n, seed=1, rank=0, state_t=None, state_c=None):
    """BASELINE config 2: n perturbed element sets of pair A for target and chaser around their
    epoch states, default 7-day window starting 3.5 d before the target epoch.  `state_*` are the
    [2,3] metre states at epoch (tsince = 0), obtained from the GPU by the caller."""
    t_tle, c_tle = pair_a()
    g = torch.Generator().manual_seed(seed * 1000003 + rank)
    el_t = mc_element_samples(t_tle, state_t, T_COV_RTN, n, g)
    el_c = mc_element_samples(c_tle, state_c, C_COV_RTN, n, g)
    time0 = t_tle.date_mjd - 3.5
    times = time_grid(time0)
    return dict(elems_t=el_t, elems_c=el_c, epoch_t=t_tle.date_mjd, epoch_c=c_tle.date_mjd, time0=time0, times=times,
                times_coarse=np.ascontiguousarray(times[::100]), n_fine=len(times), t_tle=t_tle, c_tle=c_tle)
