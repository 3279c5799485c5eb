"""Where the bf16 path's probability error comes from (run on a B200): the MOT20-scale golden scene with (a) everything bf16,
(b) bf16 ReID + fp32 Decision Transformer, several runs each (the fp32 reductions are atomics, so runs differ)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from busca_b200.scene import Scene  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_gpu_scene import build  # noqa: E402

g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scene_mot20_cond.npz"))
seed, T, D, L, C = (int(v) for v in g["meta"])
m = build("bf16")
sc = Scene(T, D, L, C, seed=seed)
sc.setup_resident(m, busca_thresh=0.3)
for tr_tc in (1, 0):
    m.engine.set_option("tr_tc", tr_tc)
    for run in range(4):
        sc.step_resident()
        m.engine.sync()
        dp = np.abs(sc.read_resident()["probs"] - g["probs"])
        print(f"tr_tc={tr_tc} run {run}: max {dp.max():.4f}  p99 {np.percentile(dp, 99):.4f}  p95 {np.percentile(dp, 95):.4f}  median {np.median(dp):.4f}")
