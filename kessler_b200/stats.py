"""Host-side finish of the per-sample reductions (tiny 6x6 algebra).

The kernels return shifted raw moments (n, sum d, sum d d^T with d = x - ref) -- the
quantities that add across thread blocks and across GPUs.  These helpers turn them into
what /root/reference/kessler/model.py:544-550 returns: the mean state and the RTN
covariance ``np.cov(rotated_states, ddof=1)``.  A rotation is linear, so rotating the
Cartesian covariance by blockdiag(R, R) equals the covariance of the rotated samples.
"""
import numpy as np

NMOMENT = 28


def unpack_moments(mom):
    """mom[28] -> (n, s[6], S[6,6] symmetric)."""
    mom = np.asarray(mom, dtype=np.float64)
    n = mom[0]
    s = mom[1:7].copy()
    S = np.zeros((6, 6))
    iu = np.triu_indices(6)
    S[iu] = mom[7:28]
    S = S + S.T - np.diag(np.diag(S))
    return n, s, S


def rotation_matrix(state):
    """util.py:304-320 (state [2,3] -> rows u, t, w)."""
    r, v = np.asarray(state[0], dtype=np.float64), np.asarray(state[1], dtype=np.float64)
    u = r / np.linalg.norm(r)
    w = np.cross(r, v)
    w = w / np.linalg.norm(w)
    t = np.cross(w, u)
    return np.vstack((u, t, w))


def moments_to_mean_cov_xyz(mom, ref_state):
    n, s, S = unpack_moments(mom)
    ref = np.asarray(ref_state, dtype=np.float64).reshape(6)
    mean = ref + s / n
    cov = (S - np.outer(s, s) / n) / (n - 1.0)
    return mean, cov


def moments_to_mean_cov_rtn(mom, ref_state):
    """-> (mean state [6] in the input frame, covariance [6,6] in the RTN frame of the mean)."""
    mean, cov = moments_to_mean_cov_xyz(mom, ref_state)
    R = rotation_matrix(mean.reshape(2, 3))
    T = np.zeros((6, 6))
    T[:3, :3] = R
    T[3:, 3:] = R
    return mean, T @ cov @ T.T


def merge_moments(moms):
    """Moments about the SAME reference add (what the NCCL all-reduce does)."""
    return np.sum(np.asarray(moms, dtype=np.float64), axis=0)
