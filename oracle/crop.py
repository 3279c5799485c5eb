"""Oracle (test infrastructure): crop + pad + 8-bit bilinear resize, integer-exact.

Follows busca/tracking.py:62-113 (get_bbox_crop, _cutout_with_pad) and restates what
``cv2.resize(cutout, (128, 384), interpolation=cv2.INTER_LINEAR)`` does for uint8 images
(OpenCV imgproc resize.cpp: 11-bit fixed-point coefficients, HResizeLinear / VResizeLinear,
and the silent INTER_AREA switch at an exact 2x down-scale).  SURVEY.md Appendix A.1.
"""
import math

import numpy as np

OUT_W, OUT_H = 128, 384
COEF_BITS = 11
COEF_SCALE = 1 << COEF_BITS


def box_to_window(box, H, W):
    """tracking.py:80-99: floor/ceil the box, clip to the image.  Returns the integer cut-out
    (X1,Y1,X2,Y2) and its clipped version."""
    x1, y1, x2, y2 = (float(v) for v in box)
    X1, Y1, X2, Y2 = int(math.floor(x1)), int(math.floor(y1)), int(math.ceil(x2)), int(math.ceil(y2))
    Y1c, Y2c = min(max(Y1, 0), H), min(max(Y2, 0), H)
    X1c, X2c = min(max(X1, 0), W), min(max(X2, 0), W)
    return (X1, Y1, X2, Y2), (X1c, Y1c, X2c, Y2c)


def cutout_with_pad(im, box):
    """tracking.py:80-113.  The pad constant is uint8(trunc(mean over the clipped window, all 3
    channels)); an empty window gives NaN -> 0 on x86.  A zero-extent cut-out becomes 1x1x3 zeros."""
    H, W = im.shape[:2]
    (X1, Y1, X2, Y2), (X1c, Y1c, X2c, Y2c) = box_to_window(box, H, W)
    win = im[Y1c:Y2c, X1c:X2c]
    n = win.shape[0] * win.shape[1] * 3
    padval = int(win.sum(dtype=np.int64)) // n if n > 0 else 0
    pt, pb, pl, pr = abs(Y1c - Y1), abs(Y2c - Y2), abs(X1c - X1), abs(X2c - X2)
    out = np.full((win.shape[0] + pt + pb, win.shape[1] + pl + pr, 3), padval, dtype=np.uint8)
    out[pt:pt + win.shape[0], pl:pl + win.shape[1]] = win
    if out.shape[0] == 0 or out.shape[1] == 0:
        out = np.zeros((1, 1, 3), dtype=np.uint8)
    return out


def _axis_x(sw, dw):
    """Horizontal taps/coefs: the COEFFICIENT is clamped at the borders."""
    scale = 1.0 / (dw / sw)
    dx = np.arange(dw, dtype=np.float64)
    fx = ((dx + 0.5) * scale - 0.5).astype(np.float32)
    sx = np.floor(fx).astype(np.int64)
    fx = (fx - sx.astype(np.float32)).astype(np.float32)
    lo = sx < 0
    sx[lo] = 0
    fx[lo] = 0
    hi = sx >= sw - 1
    sx[hi] = sw - 1
    fx[hi] = 0
    a0 = np.rint(((np.float32(1.0) - fx) * np.float32(COEF_SCALE)).astype(np.float32)).astype(np.int64)
    a1 = np.rint((fx * np.float32(COEF_SCALE)).astype(np.float32)).astype(np.int64)
    sx1 = np.minimum(sx + 1, sw - 1)
    return sx, sx1, a0, a1


def _axis_y(sh, dh):
    """Vertical taps/coefs: the ROW INDEX is clamped, both weights are kept."""
    scale = 1.0 / (dh / sh)
    dy = np.arange(dh, dtype=np.float64)
    fy = ((dy + 0.5) * scale - 0.5).astype(np.float32)
    sy = np.floor(fy).astype(np.int64)
    fy = (fy - sy.astype(np.float32)).astype(np.float32)
    b0 = np.rint(((np.float32(1.0) - fy) * np.float32(COEF_SCALE)).astype(np.float32)).astype(np.int64)
    b1 = np.rint((fy * np.float32(COEF_SCALE)).astype(np.float32)).astype(np.int64)
    r0 = np.clip(sy, 0, sh - 1)
    r1 = np.clip(sy + 1, 0, sh - 1)
    return r0, r1, b0, b1


def resize_linear_u8(src, dw=OUT_W, dh=OUT_H):
    """cv2.resize(src, (dw, dh), INTER_LINEAR) for uint8 HWC, bit-exact."""
    sh, sw = src.shape[:2]
    if sw == 2 * dw and sh == 2 * dh:  # cv2 silently switches to INTER_AREA
        s = src.astype(np.int64)
        return ((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    sx, sx1, a0, a1 = _axis_x(sw, dw)
    r0, r1, b0, b1 = _axis_y(sh, dh)
    s = src.astype(np.int64)
    Hrow = s[:, sx, :] * a0[None, :, None] + s[:, sx1, :] * a1[None, :, None]      # [sh, dw, 3]
    top = (b0[:, None, None] * (Hrow[r0] >> 4)) >> 16
    bot = (b1[:, None, None] * (Hrow[r1] >> 4)) >> 16
    return ((top + bot + 2) >> 2).astype(np.uint8)


def get_bbox_crop(im, box):
    """tracking.py:62-78 with normalize=False (what every adapter passes)."""
    return resize_linear_u8(cutout_with_pad(im, box))


def get_image_crops(im, boxes):
    """network.py:492-507 with normalize=False; the empty case returns the reference's float64
    [0,128,384,3] array (dims swapped, network.py:503)."""
    crops = [get_bbox_crop(im, b) for b in boxes]
    if len(crops) == 0:
        return np.zeros([0, OUT_W, OUT_H, 3])
    return np.stack(crops, axis=0)


def crop_direct(im, box, dw=OUT_W, dh=OUT_H):
    """Same result computed straight from the frame, the way the CUDA gather does it (no
    materialised cut-out): taps outside the clipped window read the pad scalar."""
    H, W = im.shape[:2]
    (X1, Y1, X2, Y2), (X1c, Y1c, X2c, Y2c) = box_to_window(box, H, W)
    out = np.zeros((dh, dw, 3), np.uint8)
    if X2c <= X1c or Y2c <= Y1c:            # empty clipped window -> all zeros (NaN pad -> 0)
        return out
    sw, sh = X2 - X1, Y2 - Y1
    n = (X2c - X1c) * (Y2c - Y1c) * 3
    pad = int(im[Y1c:Y2c, X1c:X2c].sum(dtype=np.int64)) // n

    def fetch(rows, cols):
        fy, fx = rows + Y1, cols + X1
        inside = ((fy >= Y1c) & (fy < Y2c))[:, None] & ((fx >= X1c) & (fx < X2c))[None, :]
        v = im[np.clip(fy, 0, H - 1)][:, np.clip(fx, 0, W - 1)].astype(np.int64)
        return np.where(inside[:, :, None], v, pad)

    if sw == 2 * dw and sh == 2 * dh:
        r, c = np.arange(dh) * 2, np.arange(dw) * 2
        return ((fetch(r, c) + fetch(r, c + 1) + fetch(r + 1, c) + fetch(r + 1, c + 1) + 2) >> 2).astype(np.uint8)
    sx, sx1, a0, a1 = _axis_x(sw, dw)
    r0, r1, b0, b1 = _axis_y(sh, dh)
    h0 = fetch(r0, sx) * a0[None, :, None] + fetch(r0, sx1) * a1[None, :, None]
    h1 = fetch(r1, sx) * a0[None, :, None] + fetch(r1, sx1) * a1[None, :, None]
    top = (b0[:, None, None] * (h0 >> 4)) >> 16
    bot = (b1[:, None, None] * (h1 >> 4)) >> 16
    return ((top + bot + 2) >> 2).astype(np.uint8)
