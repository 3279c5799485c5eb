"""Analysis script (not a test; test infrastructure): design of the CONDITIONED synthetic weight set.

For a weight profile of busca_b200.synth.make_weights it reports, with the CPU oracle on seeded synthetic tracks:
  (1) how far a CPU emulation of the CUDA bf16 ReID path lands from the fp32 forward (embedding cosine / rel-L2),
  (2) how far bf16-rounded Transformer GEMM operands move logits / probabilities,
  (3) the decision statistics of the fp32 reference algorithm: which of the C+2 outputs wins, how often the Kalman
      proposal clears busca_thresh, the top-1/top-2 margins.

    python tests/analysis_weights.py [profile] [n_tracks] [seed]
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
from oracle import geometry as ogeo  # noqa: E402
from oracle import network as onet  # noqa: E402


def r16(x):
    return x.to(torch.bfloat16).to(torch.float32)


def reid_bf16_emulation(sd, x):
    """The roundings of the tensor-core path: bf16 stem input, weights as bf16(W*|s_in|), raw conv1 /
    conv2 / stem outputs stored in bf16 (statistics from the rounded values), conv3 / downsample kept in fp32 through their BN,
    block outputs in bf16, the pooled features exact, red linear with bf16 operands."""
    r = "reid_encoder.model."
    W = lambda k: onet._t(sd, k)

    def stats(raw):
        mean = raw.double().mean(dim=(0, 2, 3))
        var = raw.double().var(dim=(0, 2, 3), unbiased=False)
        return mean, var

    def scale_shift(p, raw):
        g, b = onet._t(sd, p + ".weight").double(), onet._t(sd, p + ".bias").double()
        mean, var = stats(raw)
        a = g / torch.sqrt(var + 1e-5)
        return a.float(), (b - mean * a).float()

    def conv_folded(xraw, a_in, sh_in, w, **kw):
        """consumer conv of relu(a*x+sh): weights bf16(W*|a|), input max(sgn*x, theta) exact, '+t' dropped (removed by the BN)."""
        if a_in is None:
            return F.conv2d(xraw, r16(w), **kw)
        aa = a_in.abs().clamp_min(1e-30)
        theta = r16(-sh_in / aa)
        xin = torch.maximum(xraw * torch.sign(a_in)[None, :, None, None], theta[None, :, None, None])
        # padding must read theta: emulate by shifting so that padding (0) is theta, i.e. conv(xin - theta) + const; const is removed by BN
        return F.conv2d(xin - theta[None, :, None, None], r16(w * aa[None, :, None, None]), **kw)

    with torch.no_grad():
        x = r16(x)
        raw = r16(F.conv2d(x, r16(W(r + "conv1.weight")), stride=2, padding=3))
        a, sh = scale_shift(r + "bn1", raw)
        x = r16(F.max_pool2d(F.relu(raw * a[None, :, None, None] + sh[None, :, None, None]), 3, 2, 1))
        for li, (planes, blocks, stride) in enumerate(synth.RESNET_LAYERS, start=1):
            for b in range(blocks):
                p = f"{r}layer{li}.{b}"
                s = stride if b == 0 else 1
                r1 = r16(conv_folded(x, None, None, W(p + ".conv1.weight")))
                a1, s1 = scale_shift(p + ".bn1", r1)
                r2 = r16(conv_folded(r1, a1, s1, W(p + ".conv2.weight"), stride=s, padding=1))
                a2, s2 = scale_shift(p + ".bn2", r2)
                r3 = conv_folded(r2, a2, s2, W(p + ".conv3.weight"))
                a3, s3 = scale_shift(p + ".bn3", r3)
                o = r3 * a3[None, :, None, None] + s3[None, :, None, None]
                if b == 0:
                    rd = F.conv2d(x, r16(W(p + ".downsample.0.weight")), stride=s)
                    ad, sd_ = scale_shift(p + ".downsample.1", rd)
                    idt = rd * ad[None, :, None, None] + sd_[None, :, None, None]
                else:
                    idt = x
                x = r16(F.relu(o + idt))
        x = torch.amax(x, dim=(2, 3))
        x = F.linear(x, r16(W(r + "red.weight")), W(r + "red.bias"))
        return F.normalize(x, p=2, dim=1)


def transformer_bf16(sd, *a, **k):
    """Transformer with every linear's operands rounded to bf16 (fp32 accumulate), as launch_linear_tc runs them."""
    orig = F.linear

    def lin(x, w, b=None):
        return orig(r16(x), r16(w), b)

    F.linear = lin
    try:
        return onet.transformer_forward(sd, *a, **k)
    finally:
        F.linear = orig


def cosrel(a, b):
    a, b = a.numpy().astype(np.float64), b.numpy().astype(np.float64)
    cos = (a * b).sum(1) / np.linalg.norm(a, axis=1) / np.linalg.norm(b, axis=1)
    l2 = np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)
    return cos, l2


def search_decoder(sd, rows, kslot, n=4000, gains=(2.0, 3.0, 4.0, 6.0)):
    """Among n random decoder directions (the draw make_weights would make with decoder_seed = i) pick the one whose
    decisions on the calibration rows are the most varied: winners spread over the outputs, Kalman-slot probability on
    both sides of busca_thresh.  rows: cand_rows [T, C+2, 512] of the fp32 reference."""
    import math
    y = F.layer_norm(torch.from_numpy(rows), (512,), onet._t(sd, "decoder.0.weight"), onet._t(sd, "decoder.0.bias"), 1e-5).numpy()
    best = []
    for i in range(n):
        rng = np.random.default_rng(1_000_003 + i)
        w0 = rng.uniform(-1.0, 1.0, size=512).astype(np.float32) / math.sqrt(512)
        for g in gains:
            lg = y @ (w0 * g)
            p = np.exp(lg - lg.max(1, keepdims=True))
            p /= p.sum(1, keepdims=True)
            win = np.bincount(p.argmax(1), minlength=p.shape[1]) / len(p)
            ent = -(win[win > 0] * np.log(win[win > 0])).sum()
            fk = (p[:, kslot] > 0.3).mean()
            score = ent - 4.0 * abs(fk - 0.4) - 2.0 * max(0.0, win.max() - 0.5)
            best.append((score, i, g, fk, win.round(2)))
    best.sort(key=lambda t: -t[0])
    for b in best[:8]:
        print("  decoder_seed", b[1], "gain", b[2], "score %.3f" % b[0], "frac p_k>0.3 %.2f" % b[3], "winners", b[4])
    return best[0]


def main():
    if "--search" in sys.argv:
        sys.argv.remove("--search")
        do_search = True
    else:
        do_search = False
    profile = sys.argv[1] if len(sys.argv) > 1 else "conditioned"
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    torch.set_num_threads(os.cpu_count())
    sd = synth.make_weights(seed, profile=profile)
    L, C, D = 11, 5, 3 * T
    case = synth.make_assoc_case(31, T, D, L, crop_fn=ocrop.get_image_crops, short_history=1)
    dists = ogeo.center_distance([t.tlbr for t in case.tracks], [d.tlbr for d in case.dets])
    mem_img, mem_box, can_img, can_box, idx, n_avail, reliable = onet.gather_inputs(case.tracks, case.dets, dists, L, C, True, case.kalman)
    xm = onet.normalize_patches(mem_img.reshape(T * L, 384, 128, 3))
    xc = onet.normalize_patches(can_img.reshape(T * C, 384, 128, 3))
    em, ec = onet.reid_forward(sd, xm), onet.reid_forward(sd, xc)
    em16, ec16 = reid_bf16_emulation(sd, xm), reid_bf16_emulation(sd, xc)
    for name, a, b in (("mem", em16, em), ("can", ec16, ec)):
        cos, l2 = cosrel(a, b)
        print(f"[{profile}] ReID bf16 emulation vs fp32 ({name}): cos min {cos.min():.5f} med {np.median(cos):.5f}  relL2 max {l2.max():.4f} med {np.median(l2):.4f}")
    g = em.numpy() @ em.numpy().T
    print(f"[{profile}] cosine between DIFFERENT memory patches: med {np.median(g[np.triu_indices(len(g), 1)]):.4f} min {g.min():.4f}")
    taps = {}
    lg = onet.transformer_forward(sd, em.view(T, L, -1), ec.view(T, C, -1), mem_box, can_box, True, taps=taps)
    if do_search:
        np.save("/tmp/cand_rows_%s.npy" % profile, taps["cand_rows"])
        search_decoder(sd, taps["cand_rows"], min(D, C - 1))
    lg_t16 = transformer_bf16(sd, em.view(T, L, -1), ec.view(T, C, -1), mem_box, can_box, True)
    lg_16 = transformer_bf16(sd, em16.view(T, L, -1), ec16.view(T, C, -1), mem_box, can_box, True)
    p, p_t16, p_16 = (torch.softmax(x, -1).numpy() for x in (lg, lg_t16, lg_16))
    print(f"[{profile}] logits std over rows {lg.std(dim=0).numpy().round(2)}  mean {lg.mean(dim=0).numpy().round(2)}")
    print(f"[{profile}] bf16 Transformer only : max|dlogit| {float((lg_t16 - lg).abs().max()):.4f}  max|dp| {np.abs(p_t16 - p).max():.4f}")
    print(f"[{profile}] bf16 ReID+Transformer : max|dlogit| {float((lg_16 - lg).abs().max()):.4f}  max|dp| {np.abs(p_16 - p).max():.4f}  med|dp| {np.median(np.abs(p_16 - p)):.5f}")
    win = p.argmax(1)
    srt = np.sort(p, 1)
    margin = srt[:, -1] - srt[:, -2]
    kslot = min(D, C - 1)
    print(f"[{profile}] winners (slot histogram 0..{C + 1}; Kalman slot {kslot}): {np.bincount(win, minlength=C + 2)}")
    print(f"[{profile}] p_kalman: {np.round(p[:, kslot], 3)}")
    print(f"[{profile}] p_kalman > 0.3: {(p[:, kslot] > 0.3).sum()}/{T}; |p_k-0.3| < 0.02: {(np.abs(p[:, kslot] - 0.3) < 0.02).sum()}; margin<2e-2: {(margin < 2e-2).sum()}  flips bf16: {(p_16.argmax(1) != win).sum()}")


if __name__ == "__main__":
    main()
