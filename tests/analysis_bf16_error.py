"""Analysis script (not a test; test infrastructure): where does the bf16 ReID path lose accuracy against the fp32 oracle?

Emulates on the CPU, with the oracle's functional ResNet-50, the roundings the CUDA bf16 path performs (weights,
stem input, raw conv outputs, post-BN activations, block outputs) one at a time and reports the embedding error
versus the unrounded fp32 forward on the same seeded patches.  Used to set / justify the bf16 tolerances in
tests/test_gpu_parity.py and DESIGN.md.

    python tests/analysis_bf16_error.py [n_tracks]
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from busca_b200 import synth  # noqa: E402
from oracle import crop as ocrop  # noqa: E402
from oracle import network as onet  # noqa: E402


def r16(x):
    return x.to(torch.bfloat16).to(torch.float32)


def forward(sd, x, round_w=False, round_raw=False, round_act=False, round_blk=False, stats_from_rounded=True):
    r = "reid_encoder.model."
    W = (lambda k: r16(onet._t(sd, k))) if round_w else (lambda k: onet._t(sd, k))

    def bn(p, raw):
        g, b = onet._t(sd, p + ".weight"), onet._t(sd, p + ".bias")
        src = r16(raw) if (round_raw and stats_from_rounded) else raw
        mean = src.double().mean(dim=(0, 2, 3))
        var = src.double().var(dim=(0, 2, 3), unbiased=False)
        a = (g.double() / torch.sqrt(var + 1e-5)).float()
        sh = (b.double() - mean * a.double()).float()
        xin = r16(raw) if round_raw else raw
        return xin * a[None, :, None, None] + sh[None, :, None, None]

    act = (lambda t: r16(F.relu(t))) if round_act else F.relu
    blk = (lambda t: r16(F.relu(t))) if round_blk else F.relu
    with torch.no_grad():
        x = F.conv2d(x, W(r + "conv1.weight"), stride=2, padding=3)
        x = F.max_pool2d(act(bn(r + "bn1", x)), 3, 2, 1)
        for li, (planes, blocks, stride) in enumerate(synth.RESNET_LAYERS, start=1):
            for b in range(blocks):
                p = f"{r}layer{li}.{b}"
                s = stride if b == 0 else 1
                idt = x
                o = act(bn(p + ".bn1", F.conv2d(x, W(p + ".conv1.weight"))))
                o = act(bn(p + ".bn2", F.conv2d(o, W(p + ".conv2.weight"), stride=s, padding=1)))
                o = bn(p + ".bn3", F.conv2d(o, W(p + ".conv3.weight")))
                if b == 0:
                    idt = bn(p + ".downsample.1", F.conv2d(x, W(p + ".downsample.0.weight"), stride=s))
                x = blk(o + idt)
        x = torch.amax(x, dim=(2, 3))
        x = F.linear(x, onet._t(sd, r + "red.weight"), onet._t(sd, r + "red.bias"))
        return F.normalize(x, p=2, dim=1)


def report(name, a, b):
    a, b = a.numpy(), b.numpy()
    cos = (a * b).sum(1)
    l2 = np.linalg.norm(a - b, axis=1)
    print(f"{name:55s} cos min {cos.min():.5f} med {np.median(cos):.5f}   relL2 max {l2.max():.4f} med {np.median(l2):.4f}")


if __name__ == "__main__":
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    torch.set_num_threads(os.cpu_count())
    sd = synth.make_weights(0)
    case = synth.make_assoc_case(11, T, 2 * T, 11, crop_fn=ocrop.get_image_crops)
    patches = np.stack([c for tr in case.tracks for c in tr.images_mem[-11:]])
    x = onet.normalize_patches(patches)
    ref = forward(sd, x)
    report("fp32 vs oracle.reid_forward", forward(sd, x), onet.reid_forward(sd, x))
    report("weights bf16", forward(sd, x, round_w=True), ref)
    report("stem input bf16", forward(sd, r16(x)), ref)
    report("raw conv outputs bf16", forward(sd, x, round_raw=True), ref)
    report("raw conv outputs bf16, stats from unrounded", forward(sd, x, round_raw=True, stats_from_rounded=False), ref)
    report("post-BN activations bf16", forward(sd, x, round_act=True), ref)
    report("block outputs bf16", forward(sd, x, round_blk=True), ref)
    report("all but stem input", forward(sd, x, True, True, True, True), ref)
    report("all incl. stem input (current CUDA bf16 path)", forward(sd, r16(x), True, True, True, True), ref)
