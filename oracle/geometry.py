"""Oracle (test infrastructure): motion proposals, centre distance, IoU, candidate selection.

numpy fp64, written operation-by-operation so that the CUDA kernel (which uses the *_rn intrinsics,
i.e. no FMA contraction) can be bit-exact against it.
"""
import numpy as np


def kalman_predict_mean(mean, tracked):
    """STrack.multi_predict -> KalmanFilter.multi_predict, mean only
    (adapters/ByteTrack/yolox/tracker/byte_tracker.py:50-61;
    adapters/CenterTrack/src/lib/utils/mot_online/kalman_filter.py:154-191).
    ``mean[7] = 0`` for tracks that are not Tracked, then mean' = mean @ F.T with F = I + shift,
    which is bit-equal to pos += vel."""
    m = np.array(mean, dtype=np.float64, copy=True)
    m[~np.asarray(tracked, bool), 7] = 0.0
    out = m.copy()
    out[:, :4] = m[:, :4] + m[:, 4:]
    return out


def mean_to_tlwh(mean):
    """STrack.tlwh (byte_tracker.py:140-151): (cx, cy, a, h) -> (x, y, w, h)."""
    r = np.array(mean[:, :4], dtype=np.float64, copy=True)
    r[:, 2] = r[:, 2] * r[:, 3]
    r[:, 0] = r[:, 0] - r[:, 2] / 2
    r[:, 1] = r[:, 1] - r[:, 3] / 2
    return r


def tlwh_to_tlbr(tlwh):
    """STrack.tlbr (byte_tracker.py:153-161)."""
    r = np.array(tlwh, dtype=np.float64, copy=True)
    r[:, 2:] = r[:, 2:] + r[:, :2]
    return r


def center_distance(atlbr, btlbr):
    """busca/tracking.py:23-60 with weight_size=False: scipy cdist 'euclidean' on box centres
    (per pair: s = dx*dx; s += dy*dy; sqrt(s), all IEEE fp64, no FMA)."""
    a = np.asarray(atlbr, np.float64).reshape(-1, 4)
    b = np.asarray(btlbr, np.float64).reshape(-1, 4)
    if len(a) == 0 or len(b) == 0:
        return np.zeros((len(a), len(b)), np.float64)
    ac = (a[:, :2] + a[:, 2:]) / 2.0
    bc = (b[:, :2] + b[:, 2:]) / 2.0
    dx = ac[:, None, 0] - bc[None, :, 0]
    dy = ac[:, None, 1] - bc[None, :, 1]
    return np.sqrt(dx * dx + dy * dy)


def bbox_overlaps(boxes, query):
    """cython_bbox.bbox_overlaps as called by matching.ious
    (adapters/ByteTrack/yolox/tracker/matching.py:53-70); the +1-pixel convention is pinned by
    trackers/ByteTrack/tutorials/trades/tracker.py:266-285 and
    adapters/GHOST/src/tracking_utils.py:176-205."""
    a = np.asarray(boxes, np.float64).reshape(-1, 4)
    q = np.asarray(query, np.float64).reshape(-1, 4)
    out = np.zeros((len(a), len(q)), np.float64)
    if out.size == 0:
        return out
    a_area = (a[:, 2] - a[:, 0] + 1) * (a[:, 3] - a[:, 1] + 1)
    q_area = (q[:, 2] - q[:, 0] + 1) * (q[:, 3] - q[:, 1] + 1)
    iw = np.minimum(a[:, None, 2], q[None, :, 2]) - np.maximum(a[:, None, 0], q[None, :, 0]) + 1
    ih = np.minimum(a[:, None, 3], q[None, :, 3]) - np.maximum(a[:, None, 1], q[None, :, 1]) + 1
    inter = iw * ih
    ua = a_area[:, None] + q_area[None, :] - inter
    ok = (iw > 0) & (ih > 0)
    out[ok] = inter[ok] / ua[ok]
    return out


def iou_distance(atlbr, btlbr):
    """matching.iou_distance (matching.py:73-91): cost = 1 - IoU."""
    return 1 - bbox_overlaps(atlbr, btlbr)


def select_candidates(dists, num_candidates, use_kalman):
    """network.py:324-380: per track the ``num_candidates`` nearest detections (np.argsort; exact
    ties are broken by LOWER INDEX here - the reference's introsort leaves them unspecified,
    SURVEY.md Appendix C.3), padded with -1; the Kalman proposal overwrites slot min(D, C-1) with
    index D+t.  Returns (idx [T,C] int32, num_available)."""
    T, D = dists.shape
    C = num_candidates
    idx = np.full((T, C), -1, np.int32)
    k = min(D, C)
    if k:
        idx[:, :k] = np.argsort(dists, axis=1, kind="stable")[:, :k]
    n_avail = min(D, C)
    if use_kalman:
        n_avail = min(D + 1, C)
        idx[:, min(D, C - 1)] = D + np.arange(T)
    return idx, n_avail
