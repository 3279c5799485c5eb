"""Stand-in for the third-party ``cython_bbox`` package (not in the mount), used ONLY by tests/golden/make_golden.py.
IoU with the +1-pixel convention in float64, as pinned by the reference's two in-tree restatements
(trackers/ByteTrack/tutorials/trades/tracker.py:266-285, adapters/GHOST/src/tracking_utils.py:176-205)."""
import numpy as np


def bbox_overlaps(boxes, query_boxes):
    b = np.asarray(boxes, dtype=np.float64)
    q = np.asarray(query_boxes, dtype=np.float64)
    out = np.zeros((len(b), len(q)), dtype=np.float64)
    for k in range(len(q)):
        qa = (q[k, 2] - q[k, 0] + 1) * (q[k, 3] - q[k, 1] + 1)
        for n in range(len(b)):
            iw = min(b[n, 2], q[k, 2]) - max(b[n, 0], q[k, 0]) + 1
            if iw > 0:
                ih = min(b[n, 3], q[k, 3]) - max(b[n, 1], q[k, 1]) + 1
                if ih > 0:
                    ua = (b[n, 2] - b[n, 0] + 1) * (b[n, 3] - b[n, 1] + 1) + qa - iw * ih
                    out[n, k] = iw * ih / ua
    return out
