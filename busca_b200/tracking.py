"""Host mirror of the reference's busca/tracking.py: same names, argument meaning and error behaviour;
the arithmetic runs in libbusca_b200.so.

  center_distance        tracking.py:23-60   -> busca_center_distance (fp64, bit-exact vs scipy cdist)
  get_bbox_crop          tracking.py:62-78   -> busca_crop (integer-exact vs cv2.resize INTER_LINEAR)
  _cutout_with_pad       tracking.py:80-113  -> fused into the gather (no cut-out is materialised)
  missing_candidate_bbox tracking.py:7-20    -> host constant (and its device twin in transformer.cu)
"""
from __future__ import annotations

import numpy as np

from . import engine as _engine

_default = None


def default_engine() -> "_engine.Engine":
    """Module-level functions (the adapters call ``busca.tracking.center_distance(tracks, dets)`` with no handle,
    byte_tracker.py:489) run on the context of the tracker's own BUSCA instance: the most recently constructed Engine of
    this process.  Only if there is none a small context is created, on the GPU this process was given (LOCAL_RANK under
    torchrun, else device 0) - never a second context on GPU 0 from every rank."""
    global _default
    live = _engine.last_engine()
    if live is not None:
        return live
    if _default is None or _default.h is None:
        import os
        _default = _engine.Engine(device=int(os.environ.get("LOCAL_RANK", "0")), bank_slots=8)
    return _default


def missing_candidate_bbox(seq_len=None, flavour="ltrb", legacy_float64=True):
    """The 'very unrealistic' filler box.  Under the reference's pinned numpy 1.23.5 the array is float64
    (``np.float32 / 100.0`` promotes); under numpy>=2 it is float32.  ``legacy_float64`` picks which
    (default: the pinned environment)."""
    fmin = np.finfo("float32").min
    if legacy_float64:
        lo, q = float(fmin), float(fmin) / 100.0
        dt = np.float64
    else:
        lo, q = fmin, np.float32(fmin) / np.float32(100.0)
        dt = np.float32
    if flavour == "ltrb":
        bbox = np.array([lo, lo, q, q], dtype=dt)
    elif flavour == "ltwh":
        bbox = np.array([lo, lo, -q, -q], dtype=dt)
    else:
        raise ValueError("Unknown flavour: {}".format(flavour))
    if seq_len is not None:
        bbox = np.tile(bbox, (seq_len, 1))
    return bbox


def _as_tlbr(tracks):
    if len(tracks) > 0 and isinstance(tracks[0], np.ndarray):
        return np.asarray(tracks, dtype=np.float64).reshape(-1, 4)
    return np.array([t.tlbr for t in tracks], dtype=np.float64).reshape(-1, 4)


def center_distance(atracks, btracks, weight_size=False, engine=None):
    """Centre-to-centre distances, float64 [len(a), len(b)].  Accepts lists of objects with ``.tlbr`` or
    arrays of tlbr rows (the reference crashes on a *list* of arrays, tracking.py:45; accepted here)."""
    a, b = _as_tlbr(atracks), _as_tlbr(btracks)
    if len(a) == 0 or len(b) == 0:
        return np.zeros((len(atracks), len(btracks)), dtype=np.float64)
    d = (engine or default_engine()).center_distance(a, b)
    if weight_size:                                     # unused by every adapter; host post-scale as tracking.py:50-58
        sa = np.sqrt((a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]))
        sb = np.sqrt((b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1]))
        sa = np.tile(sa, (len(sb), 1)).T
        sb = np.tile(sb, (len(sa), 1))
        d = d * np.maximum(sa / sb, sb / sa)
    return d


def iou_distance(atlbrs, btlbrs, engine=None):
    """matching.iou_distance (adapters/ByteTrack/yolox/tracker/matching.py:73-91): 1 - IoU with the +1 pixel
    convention of cython_bbox.  Offered here because north_star puts the IoU matrix on the path."""
    a, b = _as_tlbr(atlbrs), _as_tlbr(btlbrs)
    if len(a) == 0 or len(b) == 0:
        return np.zeros((len(a), len(b)), dtype=np.float64)
    return 1 - (engine or default_engine()).iou(a, b)


def get_detection_coverage(frame_shape, active_stracks, inactive_stracks=(), engine=None):
    """BYTETracker.get_detection_coverage (adapters/ByteTrack/yolox/tracker/byte_tracker.py:574-623): the fraction of the frame covered by
    the union of the tracks' boxes (filled cv2.rectangle at the int()-truncated corners of ``tlbr * scale``), the same per object, and the
    per-box relative areas.  ``frame_shape`` = ``frame.shape`` (only H and W are used); tracks are objects with ``tlbr`` / ``scale`` or
    rows (x1, y1, x2, y2) already multiplied by the scale.  The raster count runs on the device; the scalar bookkeeping is the reference's
    numpy arithmetic on the same values."""
    H, W = int(frame_shape[0]), int(frame_shape[1])
    rows = []
    for t in list(active_stracks) + list(inactive_stracks):
        rows.append(np.asarray(t.tlbr, np.float64) * t.scale if hasattr(t, "tlbr") else np.asarray(t, np.float64).reshape(4))
    boxes = np.stack(rows) if rows else np.zeros((0, 4))
    count, areas = (engine or default_engine()).detection_coverage(boxes, H, W)
    bbox_areas = [float(a) for a in areas]
    percentage_covered = count / (H * W)
    if len(rows) > 0:
        avg_area_covered = percentage_covered / len(rows)
        average_bbox_area = np.sqrt(np.array(bbox_areas)).mean() ** 2
    else:
        avg_area_covered, average_bbox_area = 0.0, 0.0
    return {"area_covered": percentage_covered, "area_covered_per_obj": avg_area_covered, "max_bbox_area": max([0.0] + bbox_areas),
            "average_bbox_area": average_bbox_area, "bbox_areas": bbox_areas}


def is_reliable(frame_shape, active_stracks, p, engine=None):
    """BYTETracker.is_reliable (byte_tracker.py:459-465): the gate of the whole Step 3b in the MOT17 configurations (``reliable_thresh``)."""
    c = get_detection_coverage(frame_shape, active_stracks, (), engine=engine)
    return bool(c["area_covered"] > c["area_covered_per_obj"] * p[0] + p[1])


def get_bbox_crop(im, bbox_real_scale, output_size=(128, 384), normalize=True, ghost_normalize=True, engine=None):
    """One crop (x1,y1,x2,y2 in ``im`` pixels) -> [384,128,3]; uint8 BGR, or float32 normalised when
    ``normalize`` (host LUT, bit-equal to the reference's float arithmetic)."""
    assert im is not None, "Image is None"
    if tuple(output_size) != (128, 384):
        raise NotImplementedError("busca_b200 crops are fixed at 128x384 (ReID_Encoder.PRETRAINED_SIZE)")
    eng = engine or default_engine()
    eng.upload_frame(im)
    slots = eng.alloc_slots(1)
    try:
        crop = eng.crop(np.asarray(bbox_real_scale, dtype=np.float64).reshape(1, 4), slots)[0]
    finally:
        eng.free_slots(slots)
    if normalize:
        crop = normalize_crop(crop, ghost_normalize)
    return crop


def normalize_crop(crop_u8, ghost_normalize=True):
    std_r = 0.299 if ghost_normalize else 0.229
    mean = np.array([0.406, 0.456, 0.485])
    std = np.array([0.225, 0.224, std_r])
    out = crop_u8.astype(np.float32) / 255.0
    out -= mean
    out /= std
    return out


def _cutout_with_pad(im, bbox):
    raise NotImplementedError("the padded cut-out is never materialised: the gather reads the frame directly "
                              "(busca_b200/csrc/crop.cu); use get_bbox_crop")
