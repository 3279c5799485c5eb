"""Diagnostic (not a test): how far apart are the bf16 ReID embeddings of (a) two runs of the same stacked batch,
(b) the stacked batch and the duplicate-eliminated batch with multiplicity-weighted statistics, (c) each of them and
the fp32 SIMT path on the stacked batch (the parity reference)?  Separates a weighting bug from the chaotic
amplification of summation-order noise by an untrained batch-statistic ResNet (tests/analysis_bf16_error.py)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from busca_b200 import synth  # noqa: E402
from busca_b200.engine import Engine  # noqa: E402


def cos(a, b):
    return (a * b).sum(1) / (np.linalg.norm(a, axis=1) * np.linalg.norm(b, axis=1))


def main():
    w = {k: v for k, v in synth.make_weights(0).items() if "running" not in k and "num_batches" not in k}
    frame = synth.make_frame(3)
    H, W = frame.shape[:2]
    rng = np.random.default_rng(11)
    nd = int(sys.argv[1]) if len(sys.argv) > 1 else 9
    boxes = synth.random_boxes(rng, nd, H, W)
    boxes[:, 2:] += boxes[:, :2]
    engines = {}
    for prec in ("bf16", "fp32"):
        e = Engine(precision=prec, bank_slots=256)
        e.load_state_dict(w)
        e.upload_frame(frame)
        slots = e.alloc_slots(nd)
        e.crop(boxes, slots, to_host=False)
        engines[prec] = (e, slots)
    mult = (np.arange(nd) % 8) + 1
    for name, m, zeros in (("x2 uniform", np.full(nd, 2), 0), ("1..8 + 4 zero images", mult, 4), ("1..8, no zero image", mult, 0)):
        out = {}
        for prec, (e, base) in engines.items():
            slots = np.concatenate([np.repeat(base, m), np.full(zeros, -1, np.int32)]).astype(np.int32)
            if prec == "bf16":
                e.set_option("dedup", 0)
                out["off1"] = e.reid_embed(slots)
                out["off2"] = e.reid_embed(slots)
                e.set_option("dedup", 1)
                out["on"] = e.reid_embed(slots)
            else:
                out["fp32"] = e.reid_embed(slots)
        print(f"[{name}] stacked {len(slots)} images, {nd + (zeros > 0)} distinct")
        for a, b in (("off1", "off2"), ("on", "off1"), ("off1", "fp32"), ("on", "fp32")):
            c = cos(out[a], out[b])
            print(f"   cos({a:4s},{b:4s}): min {c.min():.6f}  mean {c.mean():.6f}   max|d| {np.abs(out[a] - out[b]).max():.2e}")
    print("done")


if __name__ == "__main__":
    main()
