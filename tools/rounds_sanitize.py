#!/usr/bin/env python
"""Small invocations of every SURVEY 8(f) kernel for compute-sanitizer (memcheck / racecheck / synccheck), checked against the oracle."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from busca_b200 import synth
from busca_b200.engine import Engine
from oracle import coverage as ocov
from oracle import ecc as oecc
from oracle import ingest as oing
from oracle import rounds as ornd

e = Engine(device=0, bank_slots=8)
rng = np.random.default_rng(0)


def boxes(n):
    b = synth.random_boxes(rng, n)
    b[:, 2:] += b[:, :2]
    return b


a = boxes(37)
b = np.concatenate([a[:20] + rng.normal(0, 4, (20, 4)), boxes(33)])
sc = rng.uniform(0.1, 1, len(b))
x, y, cost = e.match_round(a, b, sc, 0.9, want_cost=True)
assert np.array_equal(cost, ornd.fuse_score(ornd.iou_distance(a, b), sc)) and np.array_equal(x, ornd.linear_assignment(cost, 0.9)[0])
big = rng.uniform(0, 1.3, (300, 260))
assert np.array_equal(e.linear_assignment(big, 0.7)[0], ornd.linear_assignment(big, 0.7)[0])
n = 70
mean = np.concatenate([rng.uniform(0, 1900, (n, 2)), rng.uniform(0.2, 0.8, (n, 1)), rng.uniform(60, 300, (n, 1)), rng.normal(0, 3, (n, 4))], axis=1)
cov = np.stack([np.diag(rng.uniform(0.5, 4.0, 8)) for _ in range(n)])
mp, cp = e.kalman_predict(mean, cov, rng.uniform(size=n) < 0.5)
mu, cu = e.kalman_update(mp, cp, mp[:, :4] + 1.0)
assert np.isfinite(mu).all() and np.isfinite(cu).all()
da, db = e.duplicate_tracks(a, np.arange(len(a)), b, np.arange(len(b))[::-1].copy())
wa, wb = ornd.remove_duplicates(a, np.arange(len(a)), b, np.arange(len(b))[::-1])
assert np.array_equal(da, wa) and np.array_equal(db, wb)
cnt, _ = e.detection_coverage(boxes(90) * 0.4, 300, 1100)
chw = synth.make_detector_tensor(3, 61, 77)
assert np.array_equal(e.ingest_frame(chw, synth.YOLOX_MEANS, synth.YOLOX_STD), oing.denormalize_frame(chw, synth.YOLOX_MEANS, synth.YOLOX_STD))
f1 = synth.make_frame(9, 120, 160)
f2 = synth.make_moved_frame(f1, 0.002, 0.8, -0.6, 9)
warp, rho, it = e.camera_motion(f1, f2)
rho_o, warp_o = oecc.camera_motion(f1, f2)
assert np.abs(warp - warp_o).max() < 1e-3, (warp, warp_o)
e.upload_frame(f1)
slots = e.alloc_slots(3)
e.crop(np.array([[3.5, 4.5, 60.2, 100.9], [-5, -5, 30, 40], [100, 50, 159.5, 119.5]]), slots)
print("rounds_sanitize OK", it, "ECC iterations")
