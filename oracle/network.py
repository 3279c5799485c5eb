"""Oracle (test infrastructure): the floating-point network of the hot path and the host logic
of ``BUSCA.associate_embeddings``, restated functionally from a state dict.

fp32 PyTorch-CPU functional ops (F.conv2d, F.batch_norm, matmul): the "plain fp32 reference" for
the floating-point kernels.  Follows busca/network.py:176-244 (forward), :247-279 (_get_track_mem),
:282-429 (associate_embeddings), :470-478 (normalisation); busca/reid/resnet.py:85-128, 266-322;
busca/custom_layers.py:30-41; busca/encodings.py:43-94.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import encoding as enc
from . import geometry as geo

MEAN_BGR = np.array([0.406, 0.456, 0.485])
STD_BGR = np.array([0.225, 0.224, 0.299])     # 0.299 on R is the reference's own constant (network.py:472)
PATCH_H, PATCH_W = 384, 128


def _t(sd, k):
    v = sd[k]
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v))


def normalize_lut():
    """network.py:470-478 evaluated for all 256 byte values: v = float32(u8)/255 (fp32);
    v = float32(double(v) - mean_c); v = float32(double(v) / std_c).  Returns [256,3] fp32, BGR."""
    v = (np.arange(256, dtype=np.float32) / np.float32(255.0)).astype(np.float32)
    lut = np.empty((256, 3), np.float32)
    for c in range(3):
        a = (v.astype(np.float64) - MEAN_BGR[c]).astype(np.float32)
        lut[:, c] = (a.astype(np.float64) / STD_BGR[c]).astype(np.float32)
    return lut


def normalize_patches(u8_bgr_nhwc):
    """uint8 [N,384,128,3] BGR -> fp32 [N,3,384,128] RGB (network.py:313-316, 397-398)."""
    lut = normalize_lut()
    x = np.empty(u8_bgr_nhwc.shape, np.float32)
    for c in range(3):
        x[..., c] = lut[u8_bgr_nhwc[..., c], c]
    return torch.from_numpy(np.ascontiguousarray(x[..., ::-1].transpose(0, 3, 1, 2)))


def _bn(sd, p, x):
    """BatchNorm2d in TRAINING mode (batch statistics, biased variance, eps 1e-5): the reference
    forces train() on the ReID encoder (network.py:553-556)."""
    return F.batch_norm(x, None, None, _t(sd, p + ".weight"), _t(sd, p + ".bias"), training=True, eps=1e-5)


def reid_forward(sd, x, prefix="reid_encoder.model.", taps=None):
    """resnet.py:266-322 with pool='max', red=4, output_option='plain'; x fp32 [N,3,384,128].
    Returns unit-norm embeddings [N,512].  ``taps`` (dict) receives named intermediates."""
    r = prefix
    with torch.no_grad():
        x = F.conv2d(x, _t(sd, r + "conv1.weight"), stride=2, padding=3)
        if taps is not None:
            taps["stem_raw"] = x
        x = F.relu(_bn(sd, r + "bn1", x))
        x = F.max_pool2d(x, 3, 2, 1)
        if taps is not None:
            taps["pool"] = x
        for li, (planes, blocks, stride) in enumerate(((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2)), start=1):
            for b in range(blocks):
                p = f"{r}layer{li}.{b}"
                s = stride if b == 0 else 1
                idt = x
                o = F.relu(_bn(sd, p + ".bn1", F.conv2d(x, _t(sd, p + ".conv1.weight"))))
                o = F.relu(_bn(sd, p + ".bn2", F.conv2d(o, _t(sd, p + ".conv2.weight"), stride=s, padding=1)))
                o = _bn(sd, p + ".bn3", F.conv2d(o, _t(sd, p + ".conv3.weight")))
                if b == 0:
                    idt = _bn(sd, p + ".downsample.1", F.conv2d(x, _t(sd, p + ".downsample.0.weight"), stride=s))
                x = F.relu(o + idt)
                if taps is not None:
                    taps[f"layer{li}.{b}"] = x
        x = torch.amax(x, dim=(2, 3))
        x = F.linear(x, _t(sd, r + "red.weight"), _t(sd, r + "red.bias"))
        return F.normalize(x, p=2, dim=1)


def _layer(sd, p, x, nhead, activation="relu"):
    """custom_layers.py:30-41 (post-LN) with nn.MultiheadAttention semantics: packed in_proj rows
    [Q;K;V], heads of d/nhead, scale 1/sqrt(d_head), softmax over keys, no mask, out_proj.

    ACTIVATION: the YAML says ``gelu`` but the reference EXECUTES ReLU.  TransformerEncoder clones its
    layer with copy.deepcopy (custom_layers.py:44-45); deepcopy calls the layer's __setstate__
    (custom_layers.py:24-27), which finds no 'activation' key in the instance dict (an nn.Module
    attribute lives in ``_modules``) and injects ``F.relu`` as an instance attribute that shadows the
    nn.GELU sub-module.  Pinned by tests/golden/assoc_*.npz (GELU is off by 0.25 on layer 0)."""
    act = {"relu": F.relu, "gelu": F.gelu}[activation]
    B, S, Dm = x.shape
    dh = Dm // nhead
    qkv = F.linear(x, _t(sd, p + ".self_attn.in_proj_weight"), _t(sd, p + ".self_attn.in_proj_bias"))
    q, k, v = qkv.split(Dm, dim=-1)
    q = q.view(B, S, nhead, dh).transpose(1, 2)
    k = k.view(B, S, nhead, dh).transpose(1, 2)
    v = v.view(B, S, nhead, dh).transpose(1, 2)
    att = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(dh), dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, S, Dm)
    o = F.linear(o, _t(sd, p + ".self_attn.out_proj.weight"), _t(sd, p + ".self_attn.out_proj.bias"))
    x = F.layer_norm(x + o, (Dm,), _t(sd, p + ".norm1.weight"), _t(sd, p + ".norm1.bias"), 1e-5)
    f = F.linear(act(F.linear(x, _t(sd, p + ".linear1.weight"), _t(sd, p + ".linear1.bias"))),
                 _t(sd, p + ".linear2.weight"), _t(sd, p + ".linear2.bias"))
    return F.layer_norm(x + f, (Dm,), _t(sd, p + ".norm2.weight"), _t(sd, p + ".norm2.bias"), 1e-5)


def transformer_forward(sd, mem_emb, can_emb, mem_ltrb32, can_ltrb32, sentinel_fp64, nhead=4, nlayer=4, taps=None):
    """network.py:203-232: encoder linear x sqrt(d), token assembly, PE add, 4 layers, decoder on the
    C+2 candidate rows.  mem_emb [T,L,512], can_emb [T,C,512] fp32 tensors.  Returns logits [T,C+2]."""
    with torch.no_grad():
        T, L, Dm = mem_emb.shape
        C = can_emb.shape[1]
        W, b = _t(sd, "encoder.weight"), _t(sd, "encoder.bias")
        mem = F.linear(mem_emb, W, b) * np.sqrt(Dm)
        can = F.linear(can_emb, W, b) * np.sqrt(Dm)
        sep, non, bad = _t(sd, "sep_token"), _t(sd, "non_token"), _t(sd, "bad_token")
        toks = []
        for k in range(C):
            toks += [sep.expand(T, Dm), can[:, k]]
        toks += [sep.expand(T, Dm), non.expand(T, Dm), sep.expand(T, Dm), bad.expand(T, Dm)]
        can_seq = torch.stack(toks, dim=1)                                  # [T, 2(C+2), 512]
        mem_t, can_t = enc.temporal_ids(L, C + 2)
        mem_xy, mem_sz, can_xy, can_sz = enc.spatial_ids(mem_ltrb32, can_ltrb32, sentinel_fp64)
        tx, ty, tz = (torch.from_numpy(t) for t in enc.pe_tables(Dm))

        def lookup(xy, sz, t):
            xy, sz = torch.from_numpy(xy), torch.from_numpy(sz)
            t = torch.from_numpy(np.broadcast_to(t, xy.shape).copy())
            return torch.cat([tx[xy], ty[sz], tz[t]], dim=-1)            # fp16 [T,n,512]

        x = torch.cat([mem + lookup(mem_xy, mem_sz, mem_t), can_seq + lookup(can_xy, can_sz, can_t)], dim=1)
        if taps is not None:
            taps.update(mem_xy=mem_xy, mem_size=mem_sz, can_xy=can_xy, can_size=can_sz, mem_t=mem_t, can_t=can_t,
                        input_seq=x.numpy().copy())
        for l in range(nlayer):
            x = _layer(sd, f"transformer_encoder.layers.{l}", x, nhead)
        rows = x[:, [L + 1 + 2 * k for k in range(C + 2)]]
        if taps is not None:
            taps.update(trans_out=x.numpy().copy(), cand_rows=rows.numpy().copy(), mem_logits=x[:, :L].mean(dim=1).numpy())
        y = F.layer_norm(rows, (Dm,), _t(sd, "decoder.0.weight"), _t(sd, "decoder.0.bias"), 1e-5)
        return F.linear(y, _t(sd, "decoder.1.weight"), _t(sd, "decoder.1.bias"))[:, :, 0]


def sample_memory(n_obs, seq_len, use_broader_memory):
    """network.py:247-279: which history entries a track contributes.  Returns a list of indices
    into the track's memory (may be shorter than seq_len when the history is incomplete)."""
    if use_broader_memory and not (seq_len == 1 and n_obs >= 1) and n_obs >= seq_len:
        sep = float(n_obs - 1) / float(seq_len - 1)
        return [int(i * sep) for i in range(seq_len)]
    return list(range(max(0, n_obs - seq_len), n_obs))


def gather_inputs(tracks, dets, dists, seq_len, num_candidates, use_broader_memory, kalman):
    """Host side of associate_embeddings (network.py:293-394): returns uint8 patch batches, fp32
    ltrb boxes, the candidate index table and the 'reliable' mask."""
    T, D, L, C = len(tracks), len(dets), seq_len, num_candidates
    mem_img = np.zeros((T, L, PATCH_H, PATCH_W, 3), np.uint8)
    mem_box = np.zeros((T, L, 4), np.float64)
    reliable = np.zeros(T, bool)
    for t, tr in enumerate(tracks):
        sel = sample_memory(len(tr.images_mem), L, use_broader_memory)
        if len(sel) == L:
            reliable[t] = True
            for i, j in enumerate(sel):
                mem_img[t, i] = tr.images_mem[j]
                mem_box[t, i] = np.asarray(tr.tlwh_mem[j], np.float64) * tr.scale
        else:
            mem_box[t] = np.array([250.0, 250.0, 500.0, 500.0])
    idx, n_avail = geo.select_candidates(np.asarray(dists, np.float64).reshape(T, D), C, len(kalman) > 0)
    can_img = np.zeros((T, C, PATCH_H, PATCH_W, 3), np.uint8)
    can_box = np.tile(enc.sentinel_ltwh(True), (T, C, 1))
    for t in range(T):
        for k in range(C):
            j = idx[t, k]
            if j < 0:
                continue
            d = dets[j] if j < D else kalman[j - D]
            can_img[t, k] = d.images_mem[-1]
            can_box[t, k] = (np.asarray(d.tlwh_mem[-1], np.float64) if j < D else np.asarray(d.tlwh, np.float64)) * d.scale
    return mem_img, enc.boxes_to_ltrb32(mem_box), can_img, enc.boxes_to_ltrb32(can_box), idx, n_avail, reliable


def scatter_probs(probs, idx, n_avail, n_cols, select_highest_candidate=False,
                  highest_candidate_minimum_thresh=None, keep_highest_value=False):
    """network.py:407-425."""
    T = probs.shape[0]
    out = np.zeros((T, n_cols), np.float64)
    for t in range(T):
        p = probs[t]
        if select_highest_candidate:
            q = np.zeros_like(p)
            thr = highest_candidate_minimum_thresh
            if thr is None or thr == 0 or (thr > 0.0 and np.max(p) >= thr):
                q[np.argmax(p)] = np.max(p) if keep_highest_value else 1.0
            p = q
        out[t, idx[t, :n_avail]] = p[:n_avail]
    return out


def associate(sd, tracks, dets, dists, seq_len, num_candidates, use_broader_memory, select_highest_candidate=False,
              highest_candidate_minimum_thresh=None, keep_highest_value=False, kalman=(), sentinel_fp64=True,
              taps=None):
    """BUSCA.associate_embeddings (network.py:282-429), normalize_ims=True."""
    if len(tracks) == 0 or (len(dets) == 0 and len(kalman) == 0):
        return None, None
    T, D, L, C = len(tracks), len(dets), seq_len, num_candidates
    mem_img, mem_box, can_img, can_box, idx, n_avail, reliable = gather_inputs(
        tracks, dets, dists, L, C, use_broader_memory, kalman)
    mem_emb = reid_forward(sd, normalize_patches(mem_img.reshape(T * L, PATCH_H, PATCH_W, 3))).view(T, L, -1)
    can_emb = reid_forward(sd, normalize_patches(can_img.reshape(T * C, PATCH_H, PATCH_W, 3))).view(T, C, -1)
    logits = transformer_forward(sd, mem_emb, can_emb, mem_box, can_box, sentinel_fp64, taps=taps)
    probs = torch.softmax(logits, dim=-1).numpy()
    if taps is not None:
        taps.update(mem_emb=mem_emb.numpy(), can_emb=can_emb.numpy(), logits=logits.numpy(), probs=probs, cand_idx=idx)
    pm = scatter_probs(probs, idx, n_avail, D + len(kalman), select_highest_candidate,
                       highest_candidate_minimum_thresh, keep_highest_value)
    return pm, reliable
