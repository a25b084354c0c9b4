"""Scale-out Monte Carlo collision probability for one conjunction pair (BASELINE configs[2];
SURVEY.md Sec. 8d C3): N element samples per object, sharded over the GPUs of one box by contiguous
global-index ranges, one fused kernel per rank (draws + sgp4init + sgp4 + miss-distance count +
moments), and ONE small all-reduce at the end: the int64 count (exact) and the shifted FP64 moments.

The per-sample semantics are the reference's (model.py:461-541): element covariance from the RTN
observation noise through the Kepler partials, MVN draws, a -> mean motion, e < 0 -> 0, SGP4 to the
TCA; the count is model.py:432-437 in its elementwise form (sample i of the target against sample i
of the chaser).  The all-pairs form on 10^4 x 10^4 positions is `all_pairs_probability`."""
import numpy as np
import torch

from . import distributed as D
from . import engine
from .stats import moments_to_mean_cov_rtn
from .tle import MU_EARTH
from .workloads import tle_covariance


def element_distribution(tle, state_xyz_m, cov_matrix_diagonal_rtn, mu_earth=MU_EARTH):
    """(mean [6], Cholesky factor [6,6]) of the element distribution of model.py:461-497."""
    cov = tle_covariance(state_xyz_m, cov_matrix_diagonal_rtn, mu_earth)
    mean = np.array([float(tle.semi_major_axis), tle._ecco, tle._inclo, tle._nodeo, tle._argpo, tle._mo], dtype=np.float64)
    return mean, np.linalg.cholesky(0.5 * (cov + cov.T))


def nominal_state_m(tle, tsince):
    """State [6] in metres at tsince of the un-perturbed element set (reference point of the moments)."""
    c, _ = engine.sgp4_init(tle.elements().reshape(7, 1))
    rv, _ = engine.sgp4_prop(c, [float(tsince)])
    return rv[0, 0].cpu().numpy() * 1e3


def prepare_pair(t_tle, t_state_m, t_cov_rtn, c_tle, c_state_m, c_cov_rtn, time_ca_mjd):
    """Everything of a collision-probability run that does not depend on the samples: element distributions
    (model.py:461-497), times since epoch, reference states.  Host algebra + two tiny launches; done once."""
    mean_t, L_t = element_distribution(t_tle, t_state_m, t_cov_rtn)
    mean_c, L_c = element_distribution(c_tle, c_state_m, c_cov_rtn)
    ts_t = (time_ca_mjd - t_tle.date_mjd) * 1440.0
    ts_c = (time_ca_mjd - c_tle.date_mjd) * 1440.0
    pp = dict(mean_t=mean_t, L_t=L_t, bstar_t=t_tle._bstar, ts_t=ts_t, ref_t=nominal_state_m(t_tle, ts_t),
              mean_c=mean_c, L_c=L_c, bstar_c=c_tle._bstar, ts_c=ts_c, ref_c=nominal_state_m(c_tle, ts_c))
    pp["packed"] = engine.mc_pc_pack(pp["mean_t"], pp["L_t"], pp["ref_t"], pp["mean_c"], pp["L_c"], pp["ref_c"])
    pp["scratch"] = {}
    return pp


def run_prepared(pp, n_total, collision_threshold=70.0, seed=2, n_dump=0, chunk=1 << 28, to_host=True):
    """This rank's shard of [0, n_total) through ksl_mc_pc, then the ONE collective (counts exact, moments added in
    rank order).  With to_host=False returns the device tensors (packed int64 [count, err], moments f64 [56]) without
    any host synchronisation -- what bench.py times; otherwise the finished dict."""
    rank, world = (torch.distributed.get_rank(), torch.distributed.get_world_size()) if (
        torch.distributed.is_available() and torch.distributed.is_initialized()) else (0, 1)
    lo, hi = D.shard_range(n_total, rank, world)
    count = err = None
    mom_t = mom_c = None
    dump = None
    for start in range(lo, hi, chunk):
        n = min(chunk, hi - start)
        r = engine.mc_pc(pp["mean_t"], pp["L_t"], pp["bstar_t"], pp["ts_t"], pp["ref_t"], pp["mean_c"], pp["L_c"], pp["bstar_c"],
                         pp["ts_c"], pp["ref_c"], seed, start, n, collision_threshold, n_dump=(n_dump if start == lo else 0),
                         count=count, err_count=err, packed=pp.get("packed"), scratch=pp.get("scratch"))
        count, err = r["count"], r["err_count"]
        mom_t = r["moments_t"] if mom_t is None else mom_t + r["moments_t"]
        mom_c = r["moments_c"] if mom_c is None else mom_c + r["moments_c"]
        if r["dump"] is not None:
            dump = r["dump"]
    dev = torch.device("cuda", torch.cuda.current_device())
    if count is None:   # empty shard
        count, err = torch.zeros(1, dtype=torch.int64, device=dev), torch.zeros(1, dtype=torch.int64, device=dev)
        mom_t, mom_c = torch.zeros(28, dtype=torch.float64, device=dev), torch.zeros(28, dtype=torch.float64, device=dev)
    packed, mom, _ = D.exchange_packed(torch.cat([count, err]), torch.cat([mom_t, mom_c]))
    if not to_host:
        return packed, mom
    host = torch.cat([packed.to(torch.float64), mom]).cpu().numpy()       # ONE device -> host read
    mean_state_t, cov_t = moments_to_mean_cov_rtn(host[2:30], pp["ref_t"])
    mean_state_c, cov_c = moments_to_mean_cov_rtn(host[30:58], pp["ref_c"])
    cnt, nerr = int(host[0]), int(host[1])          # (counts < 2^53: exact in a double)
    out = dict(pc=cnt / float(n_total), count=cnt, n=int(n_total), err_count=nerr, mean_t=mean_state_t.reshape(2, 3),
               cov_rtn_t=cov_t, mean_c=mean_state_c.reshape(2, 3), cov_rtn_c=cov_c, shard=(lo, hi))
    if dump is not None:
        out["dump"] = dump
    return out


def collision_probability(t_tle, t_state_m, t_cov_rtn, c_tle, c_state_m, c_cov_rtn, time_ca_mjd, n_total,
                          collision_threshold=70.0, seed=2, n_dump=0, chunk=1 << 28):
    """Returns dict(pc, count, n, mean_t, cov_rtn_t, mean_c, cov_rtn_c, err_count[, dump]).  Call from every rank
    of an initialised process group (or from a single process): each rank runs its shard, the result is
    identical on all ranks and -- because the generator is keyed by the global sample index -- the count does
    not depend on the number of ranks."""
    pp = prepare_pair(t_tle, t_state_m, t_cov_rtn, c_tle, c_state_m, c_cov_rtn, time_ca_mjd)
    return run_prepared(pp, n_total, collision_threshold, seed, n_dump, chunk)


def all_pairs_probability(t_xyz, c_xyz, collision_threshold=70.0):
    """model.py:432-437 with target rows sharded over the ranks: count(||t_i - c_j|| < thr) / (nt * nc)."""
    rank, world = (torch.distributed.get_rank(), torch.distributed.get_world_size()) if (
        torch.distributed.is_available() and torch.distributed.is_initialized()) else (0, 1)
    t_xyz = torch.as_tensor(t_xyz, dtype=torch.float64)
    c_xyz = torch.as_tensor(c_xyz, dtype=torch.float64)
    lo, hi = D.shard_range(t_xyz.shape[1], rank, world)
    cnt = engine.pc_count(t_xyz[:, lo:hi].contiguous(), c_xyz, collision_threshold) if hi > lo else torch.zeros(
        1, dtype=torch.int64, device=torch.device("cuda", torch.cuda.current_device()))
    D.allreduce_counts(cnt)
    return int(cnt[0]) / float(t_xyz.shape[1] * c_xyz.shape[1]), int(cnt[0])
