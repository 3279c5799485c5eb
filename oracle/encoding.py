"""Oracle (test infrastructure): spatio-temporal positional encoding of busca/encodings.py.

Index triples (xy-distance bin, size bin, time bin) per token and the separable fp16 table.
Token layout (flavour MEM-SEP-CAN-BAD, encode_separator_as_reference=True), SURVEY.md A.3:
    [MEM x L] [SEP CAN_0] ... [SEP CAN_{C-1}] [SEP NON] [SEP BAD]
"""
import numpy as np

MAX_XY = 105
MAX_SIZE = 105
MAX_T = 30
FMIN32 = np.finfo(np.float32).min


def sentinel_ltwh(fp64):
    """busca/tracking.py:7-20 with flavour='ltwh'.  Under numpy 1.23.5 (pinned) the array is
    float64 and the quotient is a float64 division; under numpy>=2 everything stays float32."""
    if fp64:
        m = float(FMIN32)
        return np.array([m, m, -m / 100.0, -m / 100.0], dtype=np.float64)
    q = np.float32(-FMIN32) / np.float32(100.0)
    return np.array([FMIN32, FMIN32, q, q], dtype=np.float32)


def pe_tables(d_model=512):
    """encodings.py:23-32 + positional_encodings 6.0.x (PARITY UNPINNED, see oracle/__init__.py):
    the 211x211x61xd table is separable, pe[i,j,k] = cat(code(i)[:ch], code(j)[:ch], code(k))[:d]
    with code(p)[2m] = sin(p f_m), code(p)[2m+1] = cos(p f_m), f_m = 10000^(-2m/ch); computed in
    fp32 and cast to fp16.  Returns (tab_xy [211,ch], tab_size [211,ch], tab_t [61,d-2ch]) fp16."""
    import torch
    ch = int(np.ceil(d_model / 6) * 2)
    ch += ch % 2
    inv_freq = 1.0 / (10000 ** (torch.arange(0, ch, 2).float() / ch))

    def code(n):
        s = torch.einsum("i,j->ij", torch.arange(n, dtype=torch.float32), inv_freq)
        return torch.flatten(torch.stack((s.sin(), s.cos()), dim=-1), -2, -1)

    tx = code(2 * MAX_XY + 1).to(torch.float16).numpy()
    ty = code(2 * MAX_SIZE + 1).to(torch.float16).numpy()
    tz = code(2 * MAX_T + 1)[:, : d_model - 2 * ch].to(torch.float16).numpy()
    return tx, ty, tz


def temporal_ids(L, C_total):
    """encodings.py:150-180: MEM_i -> clamp(2(i-(L-1)),-30,30)+30 ; pair tokens -> 32, 34."""
    mem = np.clip(2.0 * np.arange(-L + 1, 1), -MAX_T, MAX_T).astype(np.int64) + MAX_T
    can = np.clip(2.0 * np.array([1, 2] * C_total), -MAX_T, MAX_T).astype(np.int64) + MAX_T
    return mem, can


def _distance_values(box, ref, dt):
    """encodings.py:239-272, evaluated in dtype ``dt`` (np.float32 or np.float64)."""
    box = box.astype(dt)
    ref = ref.astype(dt)
    one, half, eps = dt(1), dt(0.5), dt(1e-3)
    with np.errstate(all="ignore"):
        wr = ref[..., 2] - ref[..., 0] + one
        hr = ref[..., 3] - ref[..., 1] + one
        cxr = half * (ref[..., 0] + ref[..., 2])
        cyr = half * (ref[..., 1] + ref[..., 3])
        w = box[..., 2] - box[..., 0] + one
        h = box[..., 3] - box[..., 1] + one
        cx = half * (box[..., 0] + box[..., 2])
        cy = half * (box[..., 1] + box[..., 3])
        dx = (cx - cxr) / w
        dy = (cy - cyr) / h
        xy = np.log(np.sqrt(dx * dx + dy * dy) + eps)
        size = np.log(w / wr + eps) + np.log(h / hr + eps)
    return xy, size


def _to_bin(v, lim, dt):
    """clamp(v*15, -lim, lim).to(long) + lim  (encodings.py:212-233); truncation toward zero."""
    with np.errstate(all="ignore"):
        s = np.clip(v * dt(15.0), dt(-lim), dt(lim))
    return np.trunc(s).astype(np.int64) + lim


def spatial_ids(mem_ltrb32, can_ltrb32, sentinel_fp64):
    """encodings.py:97-148 + 183-235.

    mem_ltrb32 [T,L,4], can_ltrb32 [T,C,4]: float32 ltrb boxes exactly as network.py:318-319,
    388-394 produce them.  ``sentinel_fp64`` selects the numpy-1.23.5 behaviour (candidate-side
    arithmetic in float64 because the float64 sentinel promotes the concatenation,
    SURVEY.md Appendix C.1); False is the numpy>=2 behaviour (everything float32).
    Returns (mem_xy, mem_size, can_xy, can_size) with can_* of length 2*(C+2)."""
    T, L, _ = mem_ltrb32.shape
    C = can_ltrb32.shape[1]
    f32 = np.float32
    ref = mem_ltrb32[:, -1:, :]
    mxy, msz = _distance_values(mem_ltrb32, np.broadcast_to(ref, mem_ltrb32.shape), f32)
    mem_xy, mem_size = _to_bin(mxy, MAX_XY, f32), _to_bin(msz, MAX_SIZE, f32)

    dt = np.float64 if sentinel_fp64 else np.float32
    bad = sentinel_ltwh(sentinel_fp64).astype(dt)                 # used as if it were ltrb (encodings.py:21,124)
    toks = np.empty((T, 2 * (C + 2), 4), dt)
    for k in range(C):
        toks[:, 2 * k] = ref[:, 0]
        toks[:, 2 * k + 1] = can_ltrb32[:, k]
    toks[:, 2 * C] = ref[:, 0]
    toks[:, 2 * C + 1] = ref[:, 0]
    toks[:, 2 * C + 2] = bad
    toks[:, 2 * C + 3] = bad
    cxy, csz = _distance_values(toks, np.broadcast_to(ref.astype(dt), toks.shape), dt)
    return mem_xy, mem_size, _to_bin(cxy, MAX_XY, dt), _to_bin(csz, MAX_SIZE, dt)


def boxes_to_ltrb32(ltwh64):
    """network.py:318-319 / 388-389 / 393-394: fp64 ltwh -> .float() -> ltwh_to_ltrb in fp32."""
    b = np.asarray(ltwh64, np.float64).astype(np.float32)
    out = b.copy()
    with np.errstate(all="ignore"):
        out[..., 2:] = b[..., 2:] + b[..., :2]
    return out
