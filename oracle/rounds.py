"""Oracle (test infrastructure): the host tracker's association rounds.

numpy / scipy restatement of
  KalmanFilter.initiate / multi_predict / project / update   (adapters/CenterTrack/src/lib/utils/mot_online/kalman_filter.py:54-225;
                                                             the ByteTrack and TransCenter copies are identical),
  matching.iou_distance, fuse_score, linear_assignment        (adapters/ByteTrack/yolox/tracker/matching.py:39-50, 73-91, 165-180),
  remove_duplicate_stracks                                    (adapters/ByteTrack/yolox/tracker/byte_tracker.py:685-698).
Pinned by tests/golden/rounds.npz: outputs of the unmodified reference functions on seeded inputs (tests/golden/make_golden.py rounds).
``lap`` (the C++ Jonker-Volgenant solver behind linear_assignment) is a third-party package that is not in /root/reference and not
installable here - PARITY UNPINNED for it: restated from its published semantics (lapjv(cost, extend_cost=True, cost_limit=L) = the
optimum of the matrix extended to (n+m) x (n+m) with L/2 'unassigned' blocks), checked against brute force on small problems.
"""
import itertools

import numpy as np
import scipy.linalg
from scipy.optimize import linear_sum_assignment

from . import geometry as ogeo

W_POS, W_VEL = 1.0 / 20, 1.0 / 160
F = np.eye(8)
for _i in range(4):
    F[_i, 4 + _i] = 1.0
H = np.eye(4, 8)


def kf_initiate(z):
    """kalman_filter.py:54-86."""
    z = np.asarray(z, np.float64)
    mean = np.r_[z, np.zeros_like(z)]
    std = [2 * W_POS * z[3], 2 * W_POS * z[3], 1e-2, 2 * W_POS * z[3], 10 * W_VEL * z[3], 10 * W_VEL * z[3], 1e-5, 10 * W_VEL * z[3]]
    return mean, np.diag(np.square(std))


def kf_multi_predict(mean, cov, tracked=None):
    """kalman_filter.py:154-191 preceded by STrack.multi_predict's `mean[7] = 0` for tracks that are not Tracked (byte_tracker.py:50-61)."""
    mean = np.array(mean, np.float64).reshape(-1, 8)
    cov = np.asarray(cov, np.float64).reshape(-1, 8, 8)
    if tracked is not None:
        mean[~np.asarray(tracked, bool), 7] = 0
    h = mean[:, 3]
    std = np.r_[[W_POS * h, W_POS * h, 1e-2 * np.ones_like(h), W_POS * h], [W_VEL * h, W_VEL * h, 1e-5 * np.ones_like(h), W_VEL * h]]
    sqr = np.square(std).T
    q = np.asarray([np.diag(sqr[i]) for i in range(len(mean))]).reshape(-1, 8, 8)
    left = np.dot(F, cov).transpose((1, 0, 2))
    return np.dot(mean, F.T), np.dot(left, F.T) + q


def kf_update(mean, cov, z):
    """kalman_filter.py:126-152 (project) and 193-225 (update)."""
    std = [W_POS * mean[3], W_POS * mean[3], 1e-1, W_POS * mean[3]]
    pm = np.dot(H, mean)
    pc = np.linalg.multi_dot((H, cov, H.T)) + np.diag(np.square(std))
    chol, lower = scipy.linalg.cho_factor(pc, lower=True, check_finite=False)
    gain = scipy.linalg.cho_solve((chol, lower), np.dot(cov, H.T).T, check_finite=False).T
    return mean + np.dot(z - pm, gain.T), cov - np.linalg.multi_dot((gain, pc, gain.T))


def iou_distance(a_tlbr, b_tlbr):
    """matching.py:73-91: 1 - IoU with the +1 pixel convention; empty inputs give an empty [len(a), len(b)] matrix."""
    a, b = np.asarray(a_tlbr, np.float64).reshape(-1, 4), np.asarray(b_tlbr, np.float64).reshape(-1, 4)
    if len(a) == 0 or len(b) == 0:
        return np.zeros((len(a), len(b)))
    return 1 - ogeo.bbox_overlaps(a, b)


def fuse_score(cost, scores):
    """matching.py:165-180."""
    if cost.size == 0:
        return cost
    sim = 1 - cost
    det = np.expand_dims(np.asarray(scores), axis=0).repeat(cost.shape[0], axis=0)
    return 1 - sim * det


def linear_assignment(cost, thresh):
    """matching.py:39-50 -> (x [n] column or -1, y [m] row or -1) of lap.lapjv(cost, extend_cost=True, cost_limit=thresh)."""
    cost = np.asarray(cost, np.float64)
    n, m = cost.shape
    x, y = np.full(n, -1, int), np.full(m, -1, int)
    if cost.size == 0:
        return x, y
    ext = np.full((n + m, n + m), thresh / 2.0)
    ext[n:, m:] = 0.0
    ext[:n, :m] = cost
    for r, c in zip(*linear_sum_assignment(ext)):
        if r < n and c < m:
            x[r], y[c] = c, r
    return x, y


def assignment_objective(cost, x, thresh):
    """What the extended problem minimises, up to a constant: the sum over matched pairs of (c_ij - thresh)."""
    return float(sum(cost[i, j] - thresh for i, j in enumerate(x) if j >= 0))


def brute_force_assignment(cost, thresh):
    """Exhaustive optimum of the same objective (small problems only): the minimum value."""
    n, m = cost.shape
    best = 0.0
    for k in range(1, min(n, m) + 1):
        for rows in itertools.combinations(range(n), k):
            for cols in itertools.permutations(range(m), k):
                best = min(best, sum(cost[r, c] - thresh for r, c in zip(rows, cols)))
    return best


def remove_duplicates(a_tlbr, a_age, b_tlbr, b_age, thresh=0.15):
    """byte_tracker.py:685-698 -> (drop_a [na] bool, drop_b [nb] bool); age = frame_id - start_frame."""
    d = iou_distance(a_tlbr, b_tlbr)
    da, db = np.zeros(d.shape[0], bool), np.zeros(d.shape[1], bool)
    for p, q in zip(*np.where(d < thresh)):
        if a_age[p] > b_age[q]:
            db[q] = True
        else:
            da[p] = True
    return da, db
