"""Oracle (test infrastructure): camera-motion compensation.

numpy restatement of ``BYTETracker.camera_motion_compensation`` (adapters/ByteTrack/yolox/tracker/byte_tracker.py:626-657): gray
conversion of both frames (cv2.COLOR_BGR2GRAY), ``cv2.findTransformECC(template = previous, input = current, eye(2,3), MOTION_EUCLIDEAN,
(EPS | COUNT, 100, 1e-5))`` and ``STrack.apply_camera_motion`` on every pooled track.  OpenCV (``opencv_python==4.7.0.72`` in the
reference's requirements) is third-party code outside /root/reference: the algorithm is restated from the published implementation
(modules/video/src/ecc.cpp: 5x5 Gaussian smoothing, central-difference gradients, inverse-map bilinear warp with OpenCV's 1/32-pixel
coordinate quantisation, forward-additive ECC update with float32 Hessian / projections) and PINNED AGAINST cv2 ITSELF: cv2 4.13 is
installed in the build container, tests/golden/make_golden.py records its outputs (tests/golden/ecc.npz) and tests/test_host_rounds.py
also calls it live.  Agreement is to ~1e-3 px / 1e-6 in the rotation terms, not bit-wise (cv2's SIMD summation order is unspecified).
"""
import numpy as np

AB_BITS, INTER_BITS = 10, 5
AB_SCALE, TAB = 1 << AB_BITS, 1 << INTER_BITS


def bgr2gray(img):
    """cv2.cvtColor(COLOR_BGR2GRAY) on uint8: (B*3735 + G*19235 + R*9798 + 16384) >> 15 (OpenCV 4's 15-bit coefficients; checked
    against cv2 4.13 on 16.7 M random colours) - integer-exact."""
    i = img.astype(np.int32)
    return ((i[..., 0] * 3735 + i[..., 1] * 19235 + i[..., 2] * 9798 + 16384) >> 15).astype(np.uint8)


def _reflect101(n, idx):
    idx = np.abs(idx)
    return np.where(idx >= n, 2 * (n - 1) - idx, idx)


def gaussian5(img):
    """cv2.GaussianBlur(float32, (5,5), 0): separable [1,4,6,4,1]/16, BORDER_REFLECT_101, float32 arithmetic (rows, then columns)."""
    k0, k1, k2 = np.float32(0.375), np.float32(0.25), np.float32(0.0625)
    H, W = img.shape
    x = np.arange(W)
    c = img
    r = c * k0 + (c[:, _reflect101(W, x - 1)] + c[:, _reflect101(W, x + 1)]) * k1 + (c[:, _reflect101(W, x - 2)] + c[:, _reflect101(W, x + 2)]) * k2
    y = np.arange(H)
    return (r * k0 + (r[_reflect101(H, y - 1)] + r[_reflect101(H, y + 1)]) * k1 + (r[_reflect101(H, y - 2)] + r[_reflect101(H, y + 2)]) * k2).astype(np.float32)


def gradients(img):
    """filter2D with (-0.5, 0, 0.5) along x and along y, BORDER_REFLECT_101 (so the border gradient is 0)."""
    H, W = img.shape
    x, y = np.arange(W), np.arange(H)
    gx = (img[:, _reflect101(W, x + 1)] - img[:, _reflect101(W, x - 1)]) * np.float32(0.5)
    gy = (img[_reflect101(H, y + 1)] - img[_reflect101(H, y - 1)]) * np.float32(0.5)
    return gx.astype(np.float32), gy.astype(np.float32)


def _fixed_coords(M, H, W, round_delta):
    """cv::warpAffine's integer source coordinates (WARP_INVERSE_MAP): 10 fractional bits, rounded (half to even) per term."""
    M = np.asarray(M, np.float64)
    x = np.arange(W, dtype=np.float64)
    y = np.arange(H, dtype=np.float64)
    ad = np.rint(M[0, 0] * x * AB_SCALE).astype(np.int64)
    bd = np.rint(M[1, 0] * x * AB_SCALE).astype(np.int64)
    X0 = np.rint((M[0, 1] * y + M[0, 2]) * AB_SCALE).astype(np.int64) + round_delta
    Y0 = np.rint((M[1, 1] * y + M[1, 2]) * AB_SCALE).astype(np.int64) + round_delta
    return X0[:, None] + ad[None, :], Y0[:, None] + bd[None, :]


def warp_linear(planes, M):
    """warpAffine(INTER_LINEAR | WARP_INVERSE_MAP), BORDER_CONSTANT 0, for a list of float32 planes of one size."""
    H, W = planes[0].shape
    X, Y = _fixed_coords(M, H, W, AB_SCALE // TAB // 2)
    X >>= AB_BITS - INTER_BITS
    Y >>= AB_BITS - INTER_BITS
    sx, sy = X >> INTER_BITS, Y >> INTER_BITS
    fx, fy = (X & (TAB - 1)).astype(np.float32) / np.float32(TAB), (Y & (TAB - 1)).astype(np.float32) / np.float32(TAB)
    w00, w01, w10, w11 = (1 - fy) * (1 - fx), (1 - fy) * fx, fy * (1 - fx), fy * fx
    out = []
    for p in planes:
        def tap(yy, xx):
            ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
            return np.where(ok, p[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)], np.float32(0))
        out.append((tap(sy, sx) * w00 + tap(sy, sx + 1) * w01 + tap(sy + 1, sx) * w10 + tap(sy + 1, sx + 1) * w11).astype(np.float32))
    return out


def warp_mask(M, H, W):
    """warpAffine(ones, INTER_NEAREST | WARP_INVERSE_MAP): 1 where the rounded source coordinate is inside the image."""
    X, Y = _fixed_coords(M, H, W, AB_SCALE // 2)
    X >>= AB_BITS
    Y >>= AB_BITS
    return (X >= 0) & (X < W) & (Y >= 0) & (Y < H)


def _inv3_f32(h):
    """cv::Mat::inv of a 3x3 CV_32F matrix: cofactors over the determinant in double, stored as float32 (zero matrix when singular)."""
    a = h.astype(np.float64)
    c = np.array([[a[1, 1] * a[2, 2] - a[1, 2] * a[2, 1], a[0, 2] * a[2, 1] - a[0, 1] * a[2, 2], a[0, 1] * a[1, 2] - a[0, 2] * a[1, 1]],
                  [a[1, 2] * a[2, 0] - a[1, 0] * a[2, 2], a[0, 0] * a[2, 2] - a[0, 2] * a[2, 0], a[0, 2] * a[1, 0] - a[0, 0] * a[1, 2]],
                  [a[1, 0] * a[2, 1] - a[1, 1] * a[2, 0], a[0, 1] * a[2, 0] - a[0, 0] * a[2, 1], a[0, 0] * a[1, 1] - a[0, 1] * a[1, 0]]])
    d = a[0, 0] * c[0, 0] + a[0, 1] * c[1, 0] + a[0, 2] * c[2, 0]
    return (c / d).astype(np.float32) if d != 0 else np.zeros((3, 3), np.float32)


def ecc_sums(tmpl, img, gx, gy, M):
    """Everything one ECC iteration needs from the pixels, as fp64 sums over the template grid (one pass): the masked moments of the
    warped image a and the template t, and the Jacobian products.  J = (gxw*hatX + gyw*hatY, gxw, gyw) with hatX = -X sin - Y cos,
    hatY = X cos - Y sin (image_jacobian_euclidean_ECC)."""
    H, W = tmpl.shape
    a, gxw, gyw = warp_linear([img, gx, gy], M)
    m = warp_mask(M, H, W)
    h0, h1 = np.float32(M[0, 0]), np.float32(M[1, 0])
    Xg, Yg = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))
    hatX = -(Xg * h1) - (Yg * h0)
    hatY = (Xg * h0) - (Yg * h1)
    J = [(gxw * hatX + gyw * hatY).astype(np.float32), gxw, gyw]
    d = lambda u, v: float(np.dot(u.astype(np.float64).ravel(), v.astype(np.float64).ravel()))
    mf = m.astype(np.float32)
    s = {"n": int(m.sum()), "Sa": d(a, mf), "Saa": d(a * mf, a), "St": d(tmpl, mf), "Stt": d(tmpl * mf, tmpl), "Sta": d(tmpl * mf, a)}
    s["H"] = np.array([[d(J[i], J[j]) for j in range(3)] for i in range(3)])
    s["Ja"] = np.array([d(J[i], a) for i in range(3)])
    s["Jm"] = np.array([d(J[i], mf) for i in range(3)])
    s["Jmt"] = np.array([d(J[i], tmpl * mf) for i in range(3)])
    return s


def ecc_step(s, M):
    """From the sums to (rho, new map): ecc.cpp's scalar algebra with its float32 containers (Hessian, its inverse, the projections)."""
    n = s["n"]
    mu_a, mu_t = s["Sa"] / n, s["St"] / n
    var_a, var_t = s["Saa"] / n - mu_a * mu_a, s["Stt"] / n - mu_t * mu_t
    img_norm, tmp_norm = np.sqrt(n * var_a), np.sqrt(n * var_t)
    mu_af, mu_tf = np.float64(np.float32(mu_a)), np.float64(np.float32(mu_t))          # subtract(Mat32f, Scalar): the scalar is used in fp32
    corr = s["Sta"] - mu_tf * s["Sa"] - mu_af * s["St"] + n * mu_af * mu_tf
    Hf = s["H"].astype(np.float32)
    Hinv = _inv3_f32(Hf)
    ip = (s["Ja"] - mu_af * s["Jm"]).astype(np.float32)
    tp = (s["Jmt"] - mu_tf * s["Jm"]).astype(np.float32)
    rho = corr / (img_norm * tmp_norm)
    iph = (Hinv.astype(np.float64) @ ip.astype(np.float64)).astype(np.float32)
    lam_n = img_norm * img_norm - float(np.dot(ip.astype(np.float64), iph.astype(np.float64)))
    lam_d = corr - float(np.dot(tp.astype(np.float64), iph.astype(np.float64)))
    if lam_d <= 0.0:
        raise RuntimeError("ECC: the correlation is going to be minimised (cv2 raises StsNoConv here)")
    lam = lam_n / lam_d
    ep = (lam * tp.astype(np.float64) - ip.astype(np.float64)).astype(np.float32)
    dp = (Hinv.astype(np.float64) @ ep.astype(np.float64)).astype(np.float32)
    M = M.copy()
    theta = np.arcsin(np.float64(M[1, 0])) + np.float64(dp[0])
    M[0, 2] += dp[1]
    M[1, 2] += dp[2]
    M[0, 0] = M[1, 1] = np.float32(np.cos(theta))
    M[1, 0] = np.float32(np.sin(theta))
    M[0, 1] = -M[1, 0]
    return rho, M


def find_transform_ecc(template_gray, input_gray, iterations=100, eps=1e-5):
    """cv2.findTransformECC(template, input, eye(2,3,float32), MOTION_EUCLIDEAN, (EPS | COUNT, iterations, eps)) -> (rho, warp 2x3 f32)."""
    tmpl = gaussian5(template_gray.astype(np.float32))
    img = gaussian5(input_gray.astype(np.float32))
    gx, gy = gradients(img)
    M = np.eye(2, 3, dtype=np.float32)
    rho, last = -1.0, -eps
    it = 1
    while it <= iterations and abs(rho - last) >= eps:
        last = rho
        rho, M = ecc_step(ecc_sums(tmpl, img, gx, gy, M), M)
        it += 1
    return rho, M


def camera_motion(previous_bgr, current_bgr, iterations=100, eps=1e-5):
    """byte_tracker.py:626-651 up to the warp matrix."""
    return find_transform_ecc(bgr2gray(previous_bgr), bgr2gray(current_bgr), iterations, eps)
