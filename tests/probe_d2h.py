"""Probe (not a test): cost of a 300-box get_image_crops call when the caller keeps the arrays (adapter pattern: pageable destination once
the page-locked pool is exhausted) and when it drops them (pool recycles: page-locked destination), BUSCA_D2H_PIPELINE=0/1."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from busca_b200 import synth
from busca_b200.engine import Engine

e = Engine(device=0, bank_slots=65536)
frame = synth.make_frame(1)
e.upload_frame(frame)
rng = np.random.default_rng(0)
b = synth.random_boxes(rng, 300)
b[:, 2:] += b[:, :2]


def run(keep, n=40):
    held, ts = [], []
    for i in range(n):
        slots = e.alloc_slots(300)
        t0 = time.perf_counter()
        out = e.crop_owned(b, slots)
        ts.append(time.perf_counter() - t0)
        if keep:
            held.append(out)
        else:
            e.free_slots(slots)
            del out
    return np.median(ts[5:]) * 1e3, np.median(ts[-8:]) * 1e3


d = run(False)
k = run(True)
print(f"BUSCA_D2H_PIPELINE={os.environ.get('BUSCA_D2H_PIPELINE', '1')}: 300-box crop call (44 MB to the host): dropped arrays {d[0]:.2f} ms; "
      f"kept arrays {k[0]:.2f} ms (last 8 calls, pool exhausted: {k[1]:.2f} ms)")
