"""Diagnostic (not a test): ReID embedding and Transformer at growing batch sizes with progress output."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from busca_b200 import synth  # noqa: E402
from busca_b200.engine import Engine  # noqa: E402

w = synth.make_weights(0)
e = Engine(precision="bf16", bank_slots=512)
e.load_state_dict({k: v for k, v in w.items() if "running" not in k and "num_batches" not in k})
rng = np.random.default_rng(0)
patches = rng.integers(0, 255, (256, 384, 128, 3), dtype=np.uint8)
slots = e.alloc_slots(256)
e.bank_upload(patches, slots)
for T in (4, 16, 50):
    print("transformer T =", T, end=" ", flush=True)
    me = rng.standard_normal((T, 11, 512)).astype(np.float32)
    ce = rng.standard_normal((T, 5, 512)).astype(np.float32)
    mb = np.tile(np.array([100.0, 100.0, 50.0, 120.0]), (T, 11, 1))
    cb = np.tile(np.array([110.0, 100.0, 50.0, 120.0]), (T, 5, 1))
    out = e.transformer(me, ce, mb, cb)
    print("ok", float(out["probs"].sum()), flush=True)
e.set_profiling(True)
for n in [int(a) for a in sys.argv[1:]] or (8, 40, 80, 176, 256):
    print("reid_embed N =", n, end=" ", flush=True)
    emb = e.reid_embed(slots[:n])
    print("ok", np.isfinite(emb).all(), flush=True)
print("done")
