"""Oracle (test infrastructure): the detection-coverage gate of Step 3b.

numpy restatement of ``BYTETracker.get_detection_coverage`` and ``is_reliable``
(adapters/ByteTrack/yolox/tracker/byte_tracker.py:574-623 and :459-465; the CenterTrack / TransCenter copies are identical): filled
rectangles are drawn on a black canvas of the frame's size at the ``int()``-truncated corners of every active track's ``tlbr * scale``
(cv2.rectangle, thickness -1: BOTH corners inclusive, corners in either order, clipped to the canvas), the non-black pixels are counted,
and the gate compares the covered fraction with the covered fraction per object.  Pinned by tests/golden/coverage.npz (outputs of the
unmodified reference method) and against cv2 itself (tests/test_oracle_golden.py).
"""
import numpy as np


def rectangle_mask(H, W, boxes):
    """Union of the filled rectangles as a bool [H, W] canvas (cv2.rectangle(img, (int(x1), int(y1)), (int(x2), int(y2)), ..., thickness=-1))."""
    canvas = np.zeros((H, W), bool)
    for b in np.asarray(boxes, np.float64).reshape(-1, 4):
        x1, y1, x2, y2 = (int(v) for v in b)                     # truncation toward zero, as Python's int()
        xa, xb, ya, yb = min(x1, x2), max(x1, x2), min(y1, y2), max(y1, y2)
        xa, ya, xb, yb = max(xa, 0), max(ya, 0), min(xb, W - 1), min(yb, H - 1)
        if xa <= xb and ya <= yb:
            canvas[ya:yb + 1, xa:xb + 1] = True
    return canvas


def bbox_areas(H, W, boxes):
    """byte_tracker.py:589 per box: max(min(((x2 - x1) / shape[0]) * ((y2 - y1) / shape[1]), 1.0), 0.0) - the reference divides the WIDTH by
    the frame height and the HEIGHT by the frame width (sic); IEEE fp64, one operation at a time."""
    out = []
    for b in np.asarray(boxes, np.float64).reshape(-1, 4):
        v = ((b[2] - b[0]) / H) * ((b[3] - b[1]) / W)
        out.append(max(min(v, 1.0), 0.0))
    return out


def detection_coverage(frame_shape, boxes):
    """get_detection_coverage(frame, active_stracks, inactive_stracks=[]) for boxes = [t.tlbr * t.scale for t in active_stracks]."""
    H, W = int(frame_shape[0]), int(frame_shape[1])
    boxes = np.asarray(boxes, np.float64).reshape(-1, 4)
    n = len(boxes)
    areas = bbox_areas(H, W, boxes)
    count = int(np.count_nonzero(rectangle_mask(H, W, boxes)))
    covered = count / (H * W)
    if n > 0:
        per_obj = covered / n
        average = np.sqrt(np.array(areas)).mean() ** 2
    else:
        per_obj, average = 0.0, 0.0
    return {"area_covered": covered, "area_covered_per_obj": per_obj, "max_bbox_area": max([0.0] + areas), "average_bbox_area": average,
            "bbox_areas": areas, "nonzero": count}


def is_reliable(frame_shape, boxes, p):
    """byte_tracker.py:459-465."""
    c = detection_coverage(frame_shape, boxes)
    return bool(c["area_covered"] > c["area_covered_per_obj"] * p[0] + p[1])
