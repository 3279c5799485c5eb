"""Seeded synthetic inputs shared by the golden generator, the tests and bench.py.

Everything here is derived from ``numpy.random.default_rng`` (PCG64, stream-stable across
numpy versions) so the golden fixtures generated in the build container can be regenerated
bit-for-bit on the GPU box, where ``/root/reference`` does not exist.

Shapes follow SURVEY.md section 8(d): 1920x1080 BGR uint8 frames, track boxes w~U(30,90),
h~U(90,250), velocity N(0,3) px/frame, detector score U(0.7,0.95).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

PATCH_H, PATCH_W = 384, 128


# ----------------------------------------------------------------------------------------
# frames
# ----------------------------------------------------------------------------------------
def _upsample_linear(field_lr: np.ndarray, H: int, W: int) -> np.ndarray:
    """Separable linear interpolation of a [h,w,3] field to [H,W,3] (numpy only)."""
    h, w, _ = field_lr.shape
    ys = np.linspace(0.0, h - 1.0, H)
    xs = np.linspace(0.0, w - 1.0, W)
    y0 = np.floor(ys).astype(np.int64).clip(0, h - 2)
    x0 = np.floor(xs).astype(np.int64).clip(0, w - 2)
    fy = (ys - y0)[:, None, None]
    fx = (xs - x0)[None, :, None]
    rows = field_lr[y0] * (1.0 - fy) + field_lr[y0 + 1] * fy          # [H,w,3]
    return rows[:, x0] * (1.0 - fx) + rows[:, x0 + 1] * fx             # [H,W,3]


def make_frame(seed: int, H: int = 1080, W: int = 1920, noise: float = 8.0) -> np.ndarray:
    """Low-frequency colour field + Gaussian noise, uint8 BGR [H,W,3]."""
    rng = np.random.default_rng(seed)
    lr = rng.uniform(20.0, 235.0, size=(32, 32, 3))
    img = _upsample_linear(lr, H, W)
    img += rng.normal(0.0, noise, size=img.shape)
    return np.ascontiguousarray(np.clip(np.rint(img), 0, 255).astype(np.uint8))     # C order, as cv2 / a detector hands frames over


def next_frame(prev: np.ndarray, seed: int, noise: float = 4.0) -> np.ndarray:
    """Previous frame shifted by one pixel in x plus fresh noise (cheap 'video')."""
    rng = np.random.default_rng(seed)
    out = np.roll(prev, 1, axis=1).astype(np.int16)
    out += np.rint(rng.normal(0.0, noise, size=out.shape)).astype(np.int16)
    return np.ascontiguousarray(np.clip(out, 0, 255).astype(np.uint8))


# ----------------------------------------------------------------------------------------
# weights in the model_busca.pth layout
# ----------------------------------------------------------------------------------------
RESNET_LAYERS = ((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2))  # (planes, blocks, stride)


def reid_conv_specs():
    """Yield (key_prefix, Cin, Cout, k, stride) for every conv of the ReID ResNet-50 in
    forward order, plus the name of the BatchNorm that follows it."""
    specs = [("conv1", "bn1", 3, 64, 7, 2)]
    inplanes = 64
    for li, (planes, blocks, stride) in enumerate(RESNET_LAYERS, start=1):
        for b in range(blocks):
            p = f"layer{li}.{b}"
            s = stride if b == 0 else 1
            specs.append((f"{p}.conv1", f"{p}.bn1", inplanes, planes, 1, 1))
            specs.append((f"{p}.conv2", f"{p}.bn2", planes, planes, 3, s))
            specs.append((f"{p}.conv3", f"{p}.bn3", planes, planes * 4, 1, 1))
            if b == 0:
                specs.append((f"{p}.downsample.0", f"{p}.downsample.1", inplanes, planes * 4, 1, s))
            inplanes = planes * 4
    return specs


# Knobs of the "conditioned" profile (designed with tests/analysis_weights.py, see make_weights)
CONDITIONED = dict(bn3_gain=0.2, decoder_gain=6.0, decoder_seed=108, branch_gain=0.5)


def make_weights(seed: int = 0, d_model: int = 512, ff: int = 1024, nlayer: int = 4,
                 decoder_gain: float = 8.0, neg_bn_frac: float = 0.03, profile: str = "chaotic") -> Dict[str, np.ndarray]:
    """Random-init state dict with the reference's key names and shapes (fc.* omitted: the
    reference drops them with ``ignore_reid_fc=True``, network.py:445-448).

    ``profile="chaotic"`` (round-1 fixtures): plain random init.  An untrained batch-statistic-BN ResNet-50 with unit
    residual-branch gains amplifies any perturbation ~50-100x over its 16 blocks, and a gain-8 decoder on random rows
    makes one learned token win every row - good for stressing fp32 parity, useless for judging decisions or bf16.
    ``profile="conditioned"``: the same draws, then (a) the last BatchNorm of every bottleneck gets a small gain
    (|gamma| x 0.2: the residual branch perturbs the identity path instead of replacing it - what zero-init-residual
    training produces, and what trained weights look like), (b) the Transformer's residual branches are damped (out_proj,
    linear2 x 0.5), and (c) the decoder direction is the one of 4000 random draws (``decoder_seed``, gain 6; searched by
    tests/analysis_weights.py --search on a calibration frame) for which the learned NON / BAD tokens do not dominate: the
    C+2 logits then depend on the candidate embeddings, the winner varies from track to track and the Kalman-slot
    probability falls on both sides of busca_thresh."""
    if profile not in ("chaotic", "conditioned"):
        raise ValueError(profile)
    if profile == "conditioned":
        decoder_gain = CONDITIONED["decoder_gain"]
    rng = np.random.default_rng(seed)
    sd: Dict[str, np.ndarray] = {}
    f32 = np.float32

    def normal(shape, std):
        return (rng.standard_normal(shape) * std).astype(f32)

    def uniform(shape, bound):
        return rng.uniform(-bound, bound, size=shape).astype(f32)

    for name in ("sep_token", "non_token", "bad_token"):
        sd[name] = normal((d_model,), 1.0)
    sd["encoder.weight"] = uniform((d_model, d_model), 1.0 / math.sqrt(d_model))
    sd["encoder.bias"] = uniform((d_model,), 1.0 / math.sqrt(d_model))
    for l in range(nlayer):
        p = f"transformer_encoder.layers.{l}"
        sd[f"{p}.self_attn.in_proj_weight"] = uniform((3 * d_model, d_model), math.sqrt(6.0 / (4 * d_model)))
        sd[f"{p}.self_attn.in_proj_bias"] = normal((3 * d_model,), 0.02)
        sd[f"{p}.self_attn.out_proj.weight"] = uniform((d_model, d_model), 1.0 / math.sqrt(d_model))
        sd[f"{p}.self_attn.out_proj.bias"] = normal((d_model,), 0.02)
        sd[f"{p}.linear1.weight"] = uniform((ff, d_model), 1.0 / math.sqrt(d_model))
        sd[f"{p}.linear1.bias"] = uniform((ff,), 1.0 / math.sqrt(d_model))
        sd[f"{p}.linear2.weight"] = uniform((d_model, ff), 1.0 / math.sqrt(ff))
        sd[f"{p}.linear2.bias"] = uniform((d_model,), 1.0 / math.sqrt(ff))
        for n in ("norm1", "norm2"):
            sd[f"{p}.{n}.weight"] = (1.0 + 0.1 * rng.standard_normal(d_model)).astype(f32)
            sd[f"{p}.{n}.bias"] = normal((d_model,), 0.1)
    sd["decoder.0.weight"] = (1.0 + 0.1 * rng.standard_normal(d_model)).astype(f32)
    sd["decoder.0.bias"] = normal((d_model,), 0.1)
    sd["decoder.1.weight"] = uniform((1, d_model), decoder_gain / math.sqrt(d_model))
    sd["decoder.1.bias"] = uniform((1,), 1.0 / math.sqrt(d_model))

    def bn(prefix, c):
        g = rng.uniform(0.5, 1.5, size=c)
        flip = rng.uniform(size=c) < neg_bn_frac          # a few negative scales on purpose
        g = np.where(flip, -g, g)
        sd[f"{prefix}.weight"] = g.astype(f32)
        sd[f"{prefix}.bias"] = normal((c,), 0.2)
        sd[f"{prefix}.running_mean"] = np.zeros(c, f32)
        sd[f"{prefix}.running_var"] = np.ones(c, f32)
        sd[f"{prefix}.num_batches_tracked"] = np.zeros((), np.int64)

    r = "reid_encoder.model."
    for conv, bnn, cin, cout, k, _s in reid_conv_specs():
        sd[f"{r}{conv}.weight"] = normal((cout, cin, k, k), math.sqrt(2.0 / (cout * k * k)))
        bn(f"{r}{bnn}", cout)
    sd[f"{r}red.weight"] = uniform((512, 2048), 1.0 / math.sqrt(2048))
    sd[f"{r}red.bias"] = uniform((512,), 1.0 / math.sqrt(2048))
    if profile == "conditioned":
        for k in list(sd):
            if k.endswith(".bn3.weight"):
                sd[k] = (sd[k] * CONDITIONED["bn3_gain"]).astype(f32)
            elif k.endswith("self_attn.out_proj.weight") or k.endswith("linear2.weight"):
                sd[k] = (sd[k] * CONDITIONED["branch_gain"]).astype(f32)
        drng = np.random.default_rng(1_000_003 + CONDITIONED["decoder_seed"])
        sd["decoder.1.weight"] = (drng.uniform(-1.0, 1.0, size=d_model).astype(f32) / f32(math.sqrt(d_model))
                                  * f32(CONDITIONED["decoder_gain"])).reshape(1, d_model).astype(f32)
    return sd


# ----------------------------------------------------------------------------------------
# duck-typed tracks / detections (what adapters hand to BUSCA)
# ----------------------------------------------------------------------------------------
class SynthTrack:
    """Minimal stand-in for the adapters' STrack as seen by BUSCA (byte_tracker.py:23-161):
    ``images_mem`` (uint8 HWC BGR crops), ``tlwh_mem`` (fp64 ltwh, original-image coords),
    ``scale``, ``tlwh``, ``tlbr``."""

    def __init__(self, tlwh, scale=1.0, score=0.9):
        self._tlwh = np.asarray(tlwh, dtype=np.float64).copy()
        self.scale = scale
        self.score = score
        self.images_mem: List[np.ndarray] = []
        self.tlwh_mem: List[np.ndarray] = []

    @property
    def tlwh(self):
        return self._tlwh.copy()

    @property
    def tlbr(self):
        r = self._tlwh.copy()
        r[2:] += r[:2]
        return r


def random_boxes(rng, n, H=1080, W=1920, border_frac=0.02):
    """ltwh fp64 boxes; a small fraction straddles the image border."""
    w = rng.uniform(30, 90, n)
    h = rng.uniform(90, 250, n)
    x = rng.uniform(0, W - w)
    y = rng.uniform(0, H - h)
    straddle = rng.uniform(size=n) < border_frac
    side = rng.uniform(size=n) < 0.5
    x = np.where(straddle, np.where(side, -0.5 * w, W - 0.5 * w), x)
    return np.stack([x, y, w, h], axis=1)


@dataclass
class AssocCase:
    """One ``associate_embeddings`` invocation worth of inputs."""
    frames: List[np.ndarray]
    tracks: List[SynthTrack]
    dets: List[SynthTrack]
    kalman: List[SynthTrack]
    frame: np.ndarray = field(default=None)


def make_assoc_case(seed: int, T: int, D: int, L: int = 11, crop_fn=None, H: int = 1080, W: int = 1920,
                    hist_frames: Optional[int] = None, short_history: int = 0, scale: float = 1.0) -> AssocCase:
    """T unmatched tracks with >= L observed crops each (except ``short_history`` of them),
    D current-frame detections and one Kalman proposal per track.

    ``crop_fn(frame, boxes_x1y1x2y2) -> uint8 [N,384,128,3]`` supplies the crops (the reference's
    ``get_image_crops`` when generating goldens, ours when testing)."""
    rng = np.random.default_rng(seed)
    hist_frames = hist_frames or (L + 2)
    frames = [make_frame(seed * 1000 + 1, H, W)]
    for i in range(1, hist_frames + 1):
        frames.append(next_frame(frames[-1], seed * 1000 + 1 + i))
    cur = frames[-1]

    box0 = random_boxes(rng, T, H, W)
    vel = rng.normal(0.0, 3.0, size=(T, 2))
    tracks = [SynthTrack(box0[t], scale=scale) for t in range(T)]
    for t, tr in enumerate(tracks):
        n_obs = hist_frames if t >= short_history else max(1, L - 1 - t)
        start = hist_frames - n_obs
        for f in range(start, hist_frames):
            b = box0[t].copy()
            b[:2] += vel[t] * f
            b[2:] *= (1.0 + 0.01 * rng.standard_normal(2))
            tr.tlwh_mem.append(b / scale)
        tr._tlwh = tr.tlwh_mem[-1].copy()
    # crops: one call per history frame, like the adapters do
    for f in range(hist_frames):
        idx = [t for t, tr in enumerate(tracks) if len(tr.tlwh_mem) >= hist_frames - f]
        if not idx:
            continue
        boxes = []
        for t in idx:
            b = tracks[t].tlwh_mem[f - (hist_frames - len(tracks[t].tlwh_mem))] * scale
            boxes.append([b[0], b[1], b[0] + b[2], b[1] + b[3]])
        crops = crop_fn(frames[f], np.asarray(boxes))
        for j, t in enumerate(idx):
            tracks[t].images_mem.append(crops[j])

    # current-frame detections: half near tracks (jittered), half random
    n_near = min(D, T) // 2
    det_boxes = random_boxes(rng, D, H, W)
    for j in range(n_near):
        b = box0[j].copy()
        b[:2] += vel[j] * hist_frames + rng.normal(0, 6.0, 2)
        det_boxes[j] = b
    dets = []
    if D > 0:
        x = det_boxes.copy()
        x[:, 2:] += x[:, :2]
        dcrops = crop_fn(cur, x.astype(np.float32))        # detector boxes are fp32 in the adapters
        for j in range(D):
            d = SynthTrack(det_boxes[j] / scale, scale=scale, score=float(rng.uniform(0.15, 0.95)))
            d.tlwh_mem.append(d._tlwh.copy())
            d.images_mem.append(dcrops[j])
            dets.append(d)

    kalman = []
    for t, tr in enumerate(tracks):
        b = tr.tlwh_mem[-1] * scale
        b = b.copy()
        b[:2] += vel[t] + rng.normal(0, 1.0, 2)
        k = SynthTrack(b / scale, scale=scale, score=0.10000001)
        k.tlwh_mem.append(k._tlwh.copy())
        kb = k.tlbr * scale
        k.images_mem.append(crop_fn(cur, [kb])[0])
        kalman.append(k)
    return AssocCase(frames=frames, tracks=tracks, dets=dets, kalman=kalman, frame=cur)


# ----------------------------------------------------------------------------------------
# detector output for a whole sequence (what a host tracker's update() receives per frame)
# ----------------------------------------------------------------------------------------
@dataclass
class Sequence:
    """frames[f]: uint8 BGR [H,W,3]; dets[f]: float32 [n_f,5] rows (x1, y1, x2, y2, score) in frame pixels."""
    frames: List[np.ndarray]
    dets: List[np.ndarray]
    H: int
    W: int


def make_sequence(seed: int, n_frames: int, n_objects: int, H: int = 1080, W: int = 1920, warm: int = 13, miss: float = 0.2,
                  low_score: float = 0.1, clutter: float = 0.5, frame_ring: int = 0) -> Sequence:
    """``n_objects`` boxes on constant-velocity paths (+ jitter).  During the first ``warm`` frames every object is
    detected with a high score (so every track collects >= seq_len confident observations); afterwards each object is
    missed with probability ``miss``, detected with a low score (second-round material, 0.15..0.55) with probability
    ``low_score``, and ``clutter`` false positives per frame (Poisson) are added.  ``frame_ring`` > 0 synthesises only
    that many distinct frames and cycles through them (long benchmark sequences)."""
    rng = np.random.default_rng(seed)
    n_img = frame_ring if frame_ring else n_frames
    frames = [make_frame(seed * 7919 + 1, H, W)]
    for i in range(1, n_img):
        frames.append(next_frame(frames[-1], seed * 7919 + 1 + i))
    box = random_boxes(rng, n_objects, H, W, border_frac=0.0)
    vel = rng.normal(0.0, 2.5, size=(n_objects, 2))
    dets = []
    for f in range(n_frames):
        rows = []
        for o in range(n_objects):
            b = box[o].copy()
            b[:2] += vel[o] * f + rng.normal(0.0, 0.7, 2)
            b[2:] *= 1.0 + 0.01 * rng.standard_normal(2)
            u = rng.uniform()
            score = rng.uniform(0.75, 0.95)
            if f >= warm:
                if u < miss:
                    continue
                if u < miss + low_score:
                    score = rng.uniform(0.15, 0.55)
            rows.append([b[0], b[1], b[0] + b[2], b[1] + b[3], score])
        if f >= warm:
            for _ in range(rng.poisson(clutter)):
                b = random_boxes(rng, 1, H, W, border_frac=0.0)[0]
                rows.append([b[0], b[1], b[0] + b[2], b[1] + b[3], rng.uniform(0.15, 0.9)])
        dets.append(np.asarray(rows, dtype=np.float32).reshape(-1, 5))
    return Sequence(frames=[frames[f % n_img] for f in range(n_frames)], dets=dets, H=H, W=W)


YOLOX_MEANS, YOLOX_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def make_detector_tensor(seed: int, H: int, W: int, means=YOLOX_MEANS, std=YOLOX_STD) -> np.ndarray:
    """A detector input as the evaluators hold it (mot_evaluator.py:198-204): float32 [3,H,W], RGB, normalised; a few percent of the
    values de-normalise outside [0, 1] (the clip matters) and the first pixels sit on rounding edges."""
    rng = np.random.default_rng(seed)
    img = rng.uniform(-0.15, 1.15, (H, W, 3)).astype(np.float32)
    img[0, :8] = np.array([0.0, 1.0, 0.5, 1.0 / 255, 254.999 / 255, 0.999999, 1e-8, 0.25], np.float32)[:, None]
    chw = ((img - np.array(means, np.float32)) / np.array(std, np.float32)).astype(np.float32)
    return np.ascontiguousarray(chw.transpose(2, 0, 1))


def make_moved_frame(frame: np.ndarray, theta: float, tx: float, ty: float, seed: int) -> np.ndarray:
    """``frame`` seen from a camera that moved: every pixel (x, y) shows the source at the Euclidean map
    (cos x - sin y + tx, sin x + cos y + ty), bilinear, border replicated, plus +-3 levels of fresh noise (numpy only)."""
    H, W = frame.shape[:2]
    ys, xs = np.mgrid[0:H, 0:W].astype(np.float64)
    sx = np.clip(np.cos(theta) * xs - np.sin(theta) * ys + tx, 0, W - 1.001)
    sy = np.clip(np.sin(theta) * xs + np.cos(theta) * ys + ty, 0, H - 1.001)
    x0, y0 = sx.astype(np.int64), sy.astype(np.int64)
    fx, fy = (sx - x0)[..., None], (sy - y0)[..., None]
    f = frame.astype(np.float64)
    out = (f[y0, x0] * (1 - fx) + f[y0, x0 + 1] * fx) * (1 - fy) + (f[y0 + 1, x0] * (1 - fx) + f[y0 + 1, x0 + 1] * fx) * fy
    rng = np.random.default_rng(seed)
    return np.clip(np.rint(out) + rng.integers(-3, 4, out.shape), 0, 255).astype(np.uint8)
