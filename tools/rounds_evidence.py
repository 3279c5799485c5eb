#!/usr/bin/env python
"""Drives the SURVEY.md 8(f) kernels and the IoU / centre-distance matrices at sizes where their bound shows, for an ncu capture
(tools/gpu_rounds_evidence.sh) and for plain CUDA-event timings (printed as JSON: kernel -> ms, algorithmic bytes, GB/s)."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from busca_b200 import synth
from busca_b200.engine import Engine

e = Engine(device=0, bank_slots=8)
rng = np.random.default_rng(0)
out = {}


def timed(name, fn, alg_bytes, reps=3):
    fn()
    e.set_profiling(True)
    best = None
    for _ in range(reps):
        fn()
        p = e.last_profile()
        ms = p[name]["ms"] / max(1, p[name]["launches"])
        best = ms if best is None else min(best, ms)
    e.set_profiling(False)
    out[name] = {"ms_per_launch": round(best, 5), "algorithmic_MB": round(alg_bytes / 1e6, 3), "GBps": round(alg_bytes / best / 1e6, 1)}


def boxes(n):
    b = synth.random_boxes(rng, n)
    b[:, 2:] += b[:, :2]
    return b


# a2 / a3 at a size where the matrix no longer fits the launch latency: 4096 x 4096 pairs = 134 MB of fp64 per matrix
a, b = boxes(4096), boxes(4096)
timed("iou_matrix", lambda: e.iou(a, b), 4096 * 4096 * 8 + 2 * 4096 * 32)
timed("center_distance", lambda: e.center_distance(a, b), 4096 * 4096 * 8 + 2 * 4096 * 32)
# the per-frame kernel over a batch of frames (64 sequences x one MOT20-scale frame): 2 x 30.7 MB of fp64 matrices + the candidate tables
Bn, Tn, Dn = 64, 200, 300
gm = np.concatenate([rng.uniform(0, 1900, (Bn, Tn, 2)), rng.uniform(0.2, 0.8, (Bn, Tn, 1)), rng.uniform(60, 300, (Bn, Tn, 1)), rng.normal(0, 3, (Bn, Tn, 4))], axis=2)
gd = np.stack([boxes(Dn) for _ in range(Bn)])
timed("frame_geometry", lambda: e.frame_geometry_batch(gm, None, gd, 5), Bn * Tn * Dn * 16 + Bn * (Tn * 64 + Dn * 32 + Tn * (64 + 20)))
out["frame_geometry_batch64"] = out.pop("frame_geometry")
# 8f row 1
ta, tb = boxes(500), np.concatenate([boxes(500)[:350] + rng.normal(0, 4, (350, 4)), boxes(50)])
sc = rng.uniform(0.1, 1, len(tb))
timed("match_cost", lambda: e.match_round(ta, tb, sc, 0.9), 500 * 400 * 8)
timed("assignment", lambda: e.match_round(ta, tb, sc, 0.9), 500 * 400 * 8)
n = 4096
mean = np.concatenate([rng.uniform(0, 1900, (n, 2)), rng.uniform(0.2, 0.8, (n, 1)), rng.uniform(60, 300, (n, 1)), rng.normal(0, 3, (n, 4))], axis=1)
cov = np.stack([np.diag(rng.uniform(0.5, 4.0, 8)) for _ in range(n)])
timed("kalman_predict", lambda: e.kalman_predict(mean, cov, None), n * (64 + 512) * 2)
timed("kalman_update", lambda: e.kalman_update(mean, cov, mean[:, :4] + 1.0), n * (64 + 512) * 2 + n * 32)
timed("duplicate_tracks", lambda: e.duplicate_tracks(a[:2000], np.arange(2000), b[:2000], np.arange(2000)), 2000 * 2000)
# 8f row 2
cb = boxes(300)
timed("detection_coverage", lambda: e.detection_coverage(cb, 1080, 1920), 1080 * 1920 / 8)
# 8f row 4
for H, W in ((1080, 1920), (2160, 3840)):
    chw = synth.make_detector_tensor(1, H, W)
    p = e.dev_alloc(chw.nbytes)
    e.h2d(p, chw)
    timed("frame_ingest", lambda: e.ingest_frame(p, synth.YOLOX_MEANS, synth.YOLOX_STD, H, W, to_host=False), H * W * 15)
    out[f"frame_ingest_{H}p"] = out.pop("frame_ingest")
    e.dev_free(p)
# 8f row 3
f1 = synth.make_frame(21)
f2 = synth.make_moved_frame(f1, 0.003, 2.2, -1.3, 21)
timed("ecc_iteration", lambda: e.camera_motion(f1, f2), 1080 * 1920 * 4 * 4)     # template + image + two gradient planes read once per iteration
timed("ecc_prepare", lambda: e.camera_motion(f1, f2), 1080 * 1920 * (3 + 4 * 5))
print(json.dumps(out, indent=1))
