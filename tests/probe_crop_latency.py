"""Probe (not a test): latency of the adapter's single-box get_image_crops call (byte_tracker.py:468-479) and of the other short
synchronous entry points, with and without the polling stream wait (BUSCA_SPIN=0)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from busca_b200 import synth
from busca_b200.engine import Engine

e = Engine(device=0, bank_slots=4096)
frame = synth.make_frame(1)
e.upload_frame(frame)
rng = np.random.default_rng(0)
b = synth.random_boxes(rng, 2000)
b[:, 2:] += b[:, :2]
slots = e.alloc_slots(2000)
for i in range(200):
    e.crop_owned(b[i:i + 1], slots[i:i + 1])
t0 = time.perf_counter()
keep = []
for i in range(2000):
    keep.append(e.crop_owned(b[i:i + 1], slots[i:i + 1]))
dt = (time.perf_counter() - t0) / 2000 * 1e6
e.set_profiling(True)
e.crop_owned(b[:1], slots[:1])
k = e.last_profile()["crop_resize"]["ms"] * 1e3
e.set_profiling(False)
a4, b4 = b[:200], b[200:500]
t0 = time.perf_counter()
for _ in range(200):
    e.center_distance(a4, b4)
cd = (time.perf_counter() - t0) / 200 * 1e6
t0 = time.perf_counter()
for _ in range(200):
    e.match_round(a4, b4, None, 0.9)
mr = (time.perf_counter() - t0) / 200 * 1e6
print(f"BUSCA_SPIN={os.environ.get('BUSCA_SPIN', '1')}: single-box crop_owned {dt:.1f} us per call (kernel {k:.1f} us); center_distance 200x300 {cd:.1f} us; match_round 200x300 {mr:.1f} us")
