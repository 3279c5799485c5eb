"""Oracle (test infrastructure): frame ingest and the MOT result format.

numpy restatement of the evaluator's de-normalisation (adapters/ByteTrack/yolox/evaluators/mot_evaluator.py:198-204) and of
``write_results`` (:30-40).  Pinned by tests/golden/ingest.npz: the reference's own statements executed by tests/golden/make_golden.py.
"""
import numpy as np


def denormalize_frame(chw, means, std):
    """[3,H,W] float32 RGB normalised -> [H,W,3] uint8 BGR: x*std + mean in fp32 (two roundings), BGR flip, clip, *255, truncate."""
    x = np.transpose(np.asarray(chw, np.float32), (1, 2, 0))
    v = (x * np.asarray(std, np.float32)).astype(np.float32) + np.asarray(means, np.float32)
    v = v[..., [2, 1, 0]]
    v = np.clip(v, 0.0, 1.0)
    return (v * np.float32(255.0)).astype(np.uint8)


def mot_lines(results):
    """write_results: ``results`` = [(frame_id, tlwhs, track_ids, scores)] -> the text the reference writes."""
    out = []
    for frame_id, tlwhs, ids, scores in results:
        for tlwh, tid, s in zip(tlwhs, ids, scores):
            if tid < 0:
                continue
            x1, y1, w, h = tlwh
            out.append(f"{frame_id},{tid},{round(x1, 1)},{round(y1, 1)},{round(w, 1)},{round(h, 1)},{round(s, 2)},-1,-1,-1\n")
    return "".join(out)
