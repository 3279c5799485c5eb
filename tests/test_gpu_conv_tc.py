"""GPU: the tcgen05/TMEM/TMA implicit-GEMM convolution (conv_tc.cu) against the SIMT fp32-accumulate kernel on the same
bf16 inputs, for every distinct convolution shape of the ReID ResNet-50, through the C ABI."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from busca_b200 import synth


def bf16_round(x):
    """float32 -> bf16 bits (round-to-nearest-even) as uint16, and the rounded float32 values."""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32)
    r = ((u >> 16) & 1) + 0x7FFF
    b = ((u + r) >> 16).astype(np.uint16)
    return b, (b.astype(np.uint32) << 16).view(np.float32)


def bf16_to_f32(b):
    return (b.astype(np.uint32) << 16).view(np.float32)


@pytest.fixture(scope="module")
def engine(weights):
    from busca_b200.engine import Engine
    e = Engine(precision="bf16", bank_slots=8)
    e.load_state_dict({k: v for k, v in weights.items() if "running" not in k and "num_batches" not in k})
    return e


def conv_shapes():
    """(conv index, input H, W) for every conv; one representative per distinct (cin,cout,k,stride,H)."""
    specs = synth.reid_conv_specs()
    H, W = 96, 32
    seen, out = set(), []
    i = 1
    for planes, blocks, stride in synth.RESNET_LAYERS:
        for b in range(blocks):
            s = stride if b == 0 else 1
            items = [(i, H, W), (i + 1, H, W), (i + 2, H // s, W // s)] + ([(i + 3, H, W)] if b == 0 else [])
            for idx, h, w in items:
                _c, _b, cin, cout, k, st = specs[idx]
                key = (cin, cout, k, st, h)
                if key not in seen:
                    seen.add(key)
                    out.append((idx, h, w))
            i += 4 if b == 0 else 3
            H, W = H // s, W // s
    return out


SHAPES = conv_shapes()


@pytest.mark.parametrize("idx,H,W", SHAPES)
@pytest.mark.parametrize("N", [3, 8])
def test_conv_tc_matches_simt(engine, idx, H, W, N):
    L = engine.L
    info = (C.c_int32 * 4)()
    assert L.busca_conv_info(engine.h, idx, info) == 0
    cin, cout, k, stride = list(info)
    rng = np.random.default_rng(idx * 100 + N)
    x = np.maximum(rng.standard_normal((N, H, W, cin)).astype(np.float32), 0) + 0.1 * rng.standard_normal((N, H, W, cin)).astype(np.float32)
    xb, _ = bf16_round(x)
    Ho, Wo = H // stride, W // stride
    outs, stats = [], []
    for use_tc in (0, 1):
        o = np.empty((N, Ho, Wo, cout), np.uint16)
        st = np.empty(2 * cout, np.float64)
        rc = L.busca_debug_conv(engine.h, idx, xb.ctypes.data_as(C.c_void_p), N, H, W, use_tc, o.ctypes.data_as(C.c_void_p), st.ctypes.data_as(C.c_void_p))
        assert rc == 0, L.busca_last_error().decode()
        outs.append(bf16_to_f32(o))
        stats.append(st)
    ref, tc = outs
    assert np.isfinite(tc).all()
    scale = np.abs(ref).max()
    err = np.abs(tc - ref).max() / scale
    assert err < 1e-2, f"conv {idx} cin={cin} cout={cout} k={k} s={stride} {H}x{W} N={N}: max rel err {err}"   # <= 1 bf16 ulp of the largest value
    assert np.mean(tc != ref) < 0.05                      # the two fp32 accumulation orders round differently only rarely
    assert np.allclose(stats[1], stats[0], rtol=2e-3, atol=2e-3 * np.abs(stats[0]).max())


@pytest.mark.parametrize("N", [1, 5])
def test_stem_tc_matches_simt(engine, N):
    """7x7 stride-2 stem as an overlapping-window TMA GEMM vs the direct SIMT kernel (fp32 weights / inputs)."""
    from oracle import crop as ocrop
    rng = np.random.default_rng(N)
    frame = synth.make_frame(3 + N)
    b = synth.random_boxes(rng, N)
    b[:, 2:] += b[:, :2]
    patches = ocrop.get_image_crops(frame, b)
    slots = engine.alloc_slots(N)
    engine.bank_upload(patches, slots)
    sl = np.ascontiguousarray(slots, np.int32)
    if N > 1:
        sl[-1] = -1                                   # the all-zero filler image
    outs, stats = [], []
    for use_tc in (0, 1):
        o = np.empty((N, 192, 64, 64), np.uint16)
        st = np.empty(128, np.float64)
        rc = engine.L.busca_debug_stem(engine.h, sl.ctypes.data_as(C.c_void_p), N, use_tc, o.ctypes.data_as(C.c_void_p), st.ctypes.data_as(C.c_void_p))
        assert rc == 0, engine.L.busca_last_error().decode()
        outs.append(bf16_to_f32(o))
        stats.append(st)
    engine.free_slots(slots)
    ref, tc = outs
    assert np.isfinite(tc).all()
    err = np.abs(tc - ref).max() / np.abs(ref).max()
    assert err < 2e-2, err
    assert np.allclose(stats[1], stats[0], rtol=2e-2, atol=2e-2 * np.abs(stats[0]).max())


# --------------------------------------------------------------------------------------------------------------------
# the fused modes of the tensor-core convolution (busca_debug_conv_ex)
# --------------------------------------------------------------------------------------------------------------------
def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def run_conv(engine, idx, xb, N, H, W, use_tc, mode=0, in_scale=None, in_shift=None, e_scale=None, e_shift=None, idt=None,
             ds_index=-1, ds_in=None, ds_H=0, ds_W=0, ds_scale=None, ds_shift=None, img_w=None):
    from busca_b200._lib import DebugConvArgs
    L = engine.L
    info = (C.c_int32 * 4)()
    assert L.busca_conv_info(engine.h, idx, info) == 0
    cin, cout, k, stride = list(info)
    Ho, Wo = H // stride, W // stride
    out = np.full((N, Ho, Wo, cout), 0xAAAA, np.uint16)
    st = np.zeros(2 * cout, np.float64)
    keep = [np.ascontiguousarray(a, np.float32) if a is not None else None for a in (in_scale, in_shift, e_scale, e_shift, ds_scale, ds_shift, img_w)]
    d = DebugConvArgs(conv_index=idx, N=N, H=H, W=W, use_tc=use_tc, mode=mode, in_bf16=_ptr(xb), in_scale=_ptr(keep[0]), in_shift=_ptr(keep[1]),
                      e_scale=_ptr(keep[2]), e_shift=_ptr(keep[3]), idt_bf16=_ptr(idt), ds_index=ds_index, ds_H=ds_H, ds_W=ds_W,
                      ds_in_bf16=_ptr(ds_in), ds_scale=_ptr(keep[4]), ds_shift=_ptr(keep[5]), out_bf16=_ptr(out), stats_out=_ptr(st), img_w=_ptr(keep[6]))
    rc = L.busca_debug_conv_ex(engine.h, C.byref(d))
    assert rc == 0, L.busca_last_error().decode()
    return out, st, (cin, cout, k, stride)


def bn_params(rng, c):
    scale = rng.uniform(0.5, 1.5, c).astype(np.float32) * np.where(rng.uniform(size=c) < 0.05, -1, 1).astype(np.float32)
    shift = (0.3 * rng.standard_normal(c)).astype(np.float32)
    return scale, shift


def conv_roles():
    """(name, conv index, input H, W) for every conv, in forward order."""
    specs = synth.reid_conv_specs()
    out = []
    H, W = 96, 32
    i = 1
    for planes, blocks, stride in synth.RESNET_LAYERS:
        for b in range(blocks):
            s = stride if b == 0 else 1
            out.append((specs[i][0], i, H, W))
            out.append((specs[i + 1][0], i + 1, H, W))
            out.append((specs[i + 2][0], i + 2, H // s, W // s))
            if b == 0:
                out.append((specs[i + 3][0], i + 3, H, W))
            i += 4 if b == 0 else 3
            H, W = H // s, W // s
    return out


ROLES = conv_roles()
XFORM = [(n, i, h, w) for n, i, h, w in ROLES if (n.endswith("conv2") or n.endswith("conv3")) and n.split(".")[1] in ("0", "1")]


def fold_const(weights, name, shift):
    """sum_k W[o,k,taps] * shift[k]: the per-output-channel constant the max-transform drops (conv_tc.cu header): the kernel computes
    conv(relu(BN(x))) - const, which the batch-statistic BN after the conv cannot see."""
    w = np.asarray(weights["reid_encoder.model." + name + ".weight"], np.float64)       # [cout, cin, k, k]
    return w.sum(axis=(2, 3)) @ shift.astype(np.float64)


@pytest.mark.parametrize("name,idx,H,W", XFORM)
@pytest.mark.parametrize("N", [3, 9])
def test_conv_tc_input_bn_relu_in_smem(engine, weights, name, idx, H, W, N):
    """RAW and STATS modes with the producer's BN+ReLU applied to the A tile in shared memory as the exact max-transform (incl.
    the padding of 3x3 taps and of images beyond N) == the SIMT kernel fed the pre-activated bf16 tensor, up to the dropped
    per-channel constant."""
    rng = np.random.default_rng(idx * 10 + N)
    info = (C.c_int32 * 4)()
    engine.L.busca_conv_info(engine.h, idx, info)
    cin = info[0]
    xb, xr = bf16_round(rng.standard_normal((N, H, W, cin)).astype(np.float32))
    sc, sh = bn_params(rng, cin)
    if idx % 2 == 0:
        sc[3] = 0.0                                                  # degenerate BN scale: relu(shift) everywhere
    act = np.maximum(xr.astype(np.float64) * sc + sh, 0).astype(np.float32)
    ab, _ = bf16_round(act)
    ref_o, _, _ = run_conv(engine, idx, ab, N, H, W, use_tc=0)
    tc_o, tc_st, _ = run_conv(engine, idx, xb, N, H, W, use_tc=1, mode=0, in_scale=sc, in_shift=sh)
    ref, tc = bf16_to_f32(ref_o).astype(np.float64), bf16_to_f32(tc_o).astype(np.float64)
    assert np.isfinite(tc).all()
    got = tc + fold_const(weights, name, sh)
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert err < 2e-2, (name, err)
    assert np.abs(got - ref).mean() / np.abs(ref).mean() < 5e-3
    cout = tc.shape[-1]
    want_st = np.concatenate([tc.reshape(-1, cout).sum(0), (tc.reshape(-1, cout) ** 2).sum(0)])
    assert np.allclose(tc_st, want_st, rtol=1e-3, atol=1e-3 * np.abs(want_st).max())
    so_o, so_st, _ = run_conv(engine, idx, xb, N, H, W, use_tc=1, mode=1, in_scale=sc, in_shift=sh)
    assert (so_o == 0xFFFF).all()                                   # statistics-only pass writes nothing
    assert np.allclose(so_st, tc_st, rtol=2e-3, atol=2e-3 * np.abs(tc_st).max())   # fp32 accumulators vs the bf16-rounded stored values


FINALS = [(n, i, h, w) for n, i, h, w in ROLES if n.endswith("conv3") and n.split(".")[1] in ("0", "1")]


@pytest.mark.parametrize("name,idx,H,W", FINALS)
@pytest.mark.parametrize("N", [2, 7])
def test_conv_tc_final_epilogue(engine, weights, name, idx, H, W, N):
    """FINAL mode: out = relu(BN3(conv3(relu(BN2(raw2)))) + identity), identity = a tensor (blocks 1..) or BN_ds(downsample conv(x))
    accumulated in a second TMEM accumulator (block 0), against numpy on the SIMT kernel's raw outputs."""
    rng = np.random.default_rng(idx * 7 + N)
    info = (C.c_int32 * 4)()
    engine.L.busca_conv_info(engine.h, idx, info)
    cin, cout = info[0], info[1]
    xb, xr = bf16_round(rng.standard_normal((N, H, W, cin)).astype(np.float32))
    sc, sh = bn_params(rng, cin)
    ab, _ = bf16_round(np.maximum(xr.astype(np.float64) * sc + sh, 0).astype(np.float32))
    raw3 = bf16_to_f32(run_conv(engine, idx, ab, N, H, W, use_tc=0)[0]).astype(np.float64) - fold_const(weights, name, sh)
    es, et = bn_params(rng, cout)
    if name.split(".")[1] == "0":
        ds_idx = idx + 1
        engine.L.busca_conv_info(engine.h, ds_idx, info)
        dcin, dstride = info[0], info[3]
        dH, dW = H * dstride, W * dstride
        db, _ = bf16_round(np.maximum(rng.standard_normal((N, dH, dW, dcin)), 0).astype(np.float32))
        rawd = bf16_to_f32(run_conv(engine, ds_idx, db, N, dH, dW, use_tc=0)[0]).astype(np.float64)
        dsc, dsh = bn_params(rng, cout)
        want = np.maximum(raw3 * es + et + rawd * dsc + dsh, 0)
        got, _, _ = run_conv(engine, idx, xb, N, H, W, use_tc=1, mode=2, in_scale=sc, in_shift=sh, e_scale=es, e_shift=et,
                             ds_index=ds_idx, ds_in=db, ds_H=dH, ds_W=dW, ds_scale=dsc, ds_shift=dsh)
    else:
        ib, ir = bf16_round(np.maximum(rng.standard_normal((N, H, W, cout)), 0).astype(np.float32))
        want = np.maximum(raw3 * es + et + ir, 0)
        got, _, _ = run_conv(engine, idx, xb, N, H, W, use_tc=1, mode=2, in_scale=sc, in_shift=sh, e_scale=es, e_shift=et, idt=ib)
    got = bf16_to_f32(got)
    assert np.isfinite(got).all()
    err = np.abs(got - want).max() / np.abs(want).max()
    assert err < 2e-2, (name, err)                                  # raw3/rawd reference values are bf16-rounded, the kernel's are fp32
    assert np.abs(got - want).mean() / np.abs(want).mean() < 5e-3


# --------------------------------------------------------------------------------------------------------------------
# independent references: torch.nn.functional.conv2d on the CPU, and stacked-vs-deduplicated statistics
# --------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("idx,H,W", SHAPES)
def test_conv_tc_matches_torch_conv2d(engine, weights, idx, H, W):
    """Every distinct convolution shape of the ReID network on the tensor cores against F.conv2d (fp32, CPU) on the SAME bf16-rounded
    input and bf16-rounded weights: the only differences left are the fp32 accumulation order and the final rounding to bf16."""
    import torch
    import torch.nn.functional as F
    name, _bn, cin, cout, k, stride = synth.reid_conv_specs()[idx]
    N = 3
    rng = np.random.default_rng(idx * 31 + 7)
    x = np.maximum(rng.standard_normal((N, H, W, cin)).astype(np.float32), 0) + 0.1 * rng.standard_normal((N, H, W, cin)).astype(np.float32)
    xb, xr = bf16_round(x)
    _wb, wr = bf16_round(np.asarray(weights["reid_encoder.model." + name + ".weight"], np.float32))
    ref = F.conv2d(torch.from_numpy(np.ascontiguousarray(xr.transpose(0, 3, 1, 2))).double(), torch.from_numpy(wr).double(), stride=stride, padding=k // 2)
    ref = ref.permute(0, 2, 3, 1).numpy()                                     # fp64 accumulate: the exact value of the bf16 x bf16 products' sum
    out, st, _ = run_conv(engine, idx, xb, N, H, W, use_tc=1)
    tc = bf16_to_f32(out).astype(np.float64)
    assert np.isfinite(tc).all()
    _rb, ref16 = bf16_round(ref.astype(np.float32))
    scale = np.abs(ref).max()
    assert np.abs(tc - ref).max() / scale < 4e-3                              # half a bf16 ulp of the largest value, plus accumulation noise
    assert np.mean(tc != ref16) < 0.03                                        # same bf16 value except where the fp32 sum sits on a rounding boundary
    want = np.concatenate([tc.reshape(-1, cout).sum(0), (tc.reshape(-1, cout) ** 2).sum(0)])
    assert np.allclose(st, want, rtol=1e-4, atol=1e-4 * np.abs(want).max())   # statistics of exactly the stored values


@pytest.mark.parametrize("name,idx,H,W", [r for r in XFORM if r[0].startswith("layer1.0") or r[0].startswith("layer3.0")])
def test_statistics_of_a_deduplicated_batch(engine, name, idx, H, W):
    """The duplicate elimination of a BatchNorm batch: running each DISTINCT image once with its multiplicity as
    the weight of its rows gives the per-channel sums of the stacked batch - here checked at fp32 rounding (1e-5 relative) on the first
    convolutions after a transform, RAW and statistics-only modes."""
    rng = np.random.default_rng(idx)
    info = (C.c_int32 * 4)()
    engine.L.busca_conv_info(engine.h, idx, info)
    cin, cout = info[0], info[1]
    mult = np.array([1, 3, 2, 5, 1], np.int64)
    U = len(mult)
    xb, _ = bf16_round(rng.standard_normal((U, H, W, cin)).astype(np.float32))
    sc, sh = bn_params(rng, cin)
    stacked = np.repeat(xb, mult, axis=0)
    for mode in (0, 1):
        _o, st_stacked, _ = run_conv(engine, idx, stacked, int(mult.sum()), H, W, use_tc=1, mode=mode, in_scale=sc, in_shift=sh)
        _o, st_dedup, _ = run_conv(engine, idx, xb, U, H, W, use_tc=1, mode=mode, in_scale=sc, in_shift=sh, img_w=mult.astype(np.float32))
        ref = np.abs(st_stacked).reshape(2, cout).max(axis=1).repeat(cout)
        assert np.abs(st_dedup - st_stacked).max() < 1e-5 * ref.max(), (mode, np.abs(st_dedup - st_stacked).max() / ref.max())
        assert np.all(np.abs(st_dedup - st_stacked) < 2e-5 * ref)


GRAM = [(n, i, h, w) for n, i, h, w in ROLES if (n.endswith("conv3") or "downsample" in n) and n.split(".")[1] == "0"
        and synth.reid_conv_specs()[i][2] <= 256]


@pytest.mark.parametrize("name,idx,H,W", GRAM)
@pytest.mark.parametrize("N", [2, 8, 21])
def test_gram_matrix_statistics(engine, name, idx, H, W, N):
    """Batch statistics of a 1x1 convolution from the Gram matrix of its (transformed) input - sum_p y = W m, sum_p y^2 = W^T G W, G on the
    tensor cores with MN-major operands, quadratic forms in fp64 - against the statistics-only GEMM pass they replace (mode 1), for the
    last convolution of a bottleneck (input transformed in shared memory) and the strided downsample convolution (no transform)."""
    rng = np.random.default_rng(idx * 13 + N)
    info = (C.c_int32 * 4)()
    engine.L.busca_conv_info(engine.h, idx, info)
    cin, cout, k, stride = list(info)
    assert k == 1
    xb, _ = bf16_round(rng.standard_normal((N, H, W, cin)).astype(np.float32) + (0.3 if "downsample" in name else 0.0))
    kw = {}
    if "downsample" not in name:
        sc, sh = bn_params(rng, cin)
        kw = dict(in_scale=sc, in_shift=sh)
    _o, st_direct, _ = run_conv(engine, idx, xb, N, H, W, use_tc=1, mode=1, **kw)
    _o, st_gram, _ = run_conv(engine, idx, xb, N, H, W, use_tc=1, mode=3, **kw)
    s_ref, q_ref = st_direct[:cout], st_direct[cout:]
    assert np.isfinite(st_gram).all()
    assert np.abs(st_gram[:cout] - s_ref).max() < 2e-4 * np.abs(s_ref).max() + 1e-3
    assert np.abs(st_gram[cout:] - q_ref).max() < 2e-4 * np.abs(q_ref).max()
    # what BatchNorm makes of them: mean and variance per channel
    cnt = N * (H // stride) * (W // stride)
    var_d = q_ref / cnt - (s_ref / cnt) ** 2
    var_g = st_gram[cout:] / cnt - (st_gram[:cout] / cnt) ** 2
    assert np.abs(var_g - var_d).max() < 1e-3 * np.abs(var_d).max()


@pytest.mark.parametrize("shape", [(3, 192, 64, 64), (2, 16, 8, 32), (1, 192, 64, 64)])
def test_bn_relu_maxpool_kernels_agree(engine, shape):
    """relu(BN(x)) + 3x3/2 max-pool of the stem: the strip kernel (row reuse, stem shape) and the monotone kernel (pool the packed bf16
    max / min first, one fma per channel) against the tap-by-tap kernel - equal values - and against numpy on the same bf16 input."""
    N, H, W, Cc = shape
    rng = np.random.default_rng(H + Cc)
    xb, xr = bf16_round(rng.standard_normal(shape).astype(np.float32) * 3)
    scale = (rng.uniform(0.2, 2.0, Cc) * np.where(rng.uniform(size=Cc) < 0.3, -1, 1)).astype(np.float32)
    scale[1] = 0.0
    shift = (0.5 * rng.standard_normal(Cc)).astype(np.float32)
    outs = []
    for mono in (0, 1):
        out = np.empty((N, H // 2, W // 2, Cc), np.uint16)
        engine.set_option("pool_mono", mono)
        rc = engine.L.busca_debug_maxpool(engine.h, _ptr(xb), N, H, W, Cc, _ptr(scale), _ptr(shift), _ptr(out))
        engine.set_option("pool_mono", 1)
        assert rc == 0, engine.L.busca_last_error().decode()
        outs.append(bf16_to_f32(out))
    assert np.array_equal(outs[0], outs[1])
    act = np.maximum(np.float32(xr * scale + shift), 0)            # fp32 fma vs mul+add can differ in the last bit before the bf16 rounding
    pad = np.full((N, H + 2, W + 2, Cc), -np.inf, np.float32)
    pad[:, 1:-1, 1:-1] = act
    want = np.max(np.stack([pad[:, dy:dy + H:2, dx:dx + W:2] for dy in range(3) for dx in range(3)]), axis=0)
    _b, want16 = bf16_round(want)
    assert np.mean(outs[1] != want16) < 2e-3
    assert np.abs(outs[1] - want16).max() <= 2.0 ** -7 * np.abs(want16).max()


# --------------------------------------------------------------------------------------------------------------------
# paired (weight-multicast) variant of the streamed-weight kernels: the same MMAs in the same order, so the same bits
# --------------------------------------------------------------------------------------------------------------------
def _role(name):
    return next(r for r in ROLES if r[0] == name)


@pytest.mark.parametrize("name,mode", [("layer3.1.conv1", 0), ("layer3.0.conv2", 0), ("layer4.1.conv2", 0), ("layer4.0.conv3", 1),
                                       ("layer3.1.conv3", 2), ("layer3.0.conv3", 2), ("layer4.0.conv3", 2)])
@pytest.mark.parametrize("N", [5, 16])
def test_paired_multicast_variant_is_bit_identical(engine, name, mode, N):
    """conv_tc_kernel<.., MC=1> (clusters of two CTAs on adjacent pixel tiles, every weight tile loaded half by each and multicast to
    both) against the one-CTA variant: RAW (1x1, stride-2 and stride-1 3x3 through the tap loop), statistics-only (transposed operands),
    FINAL with an identity tensor and FINAL with the downsample conv, odd and even numbers of pixel tiles."""
    _n, idx, H, W = _role(name)
    rng = np.random.default_rng(idx * 3 + N + mode)
    info = (C.c_int32 * 4)()
    engine.L.busca_conv_info(engine.h, idx, info)
    cin, cout, k, stride = list(info)
    xb, _ = bf16_round(rng.standard_normal((N, H, W, cin)).astype(np.float32))
    kw = {}
    if name.endswith("conv2") or name.endswith("conv3"):
        sc, sh = bn_params(rng, cin)
        kw.update(in_scale=sc, in_shift=sh)
    if mode == 2:
        es, et = bn_params(rng, cout)
        kw.update(e_scale=es, e_shift=et)
        if name.split(".")[1] == "0":
            engine.L.busca_conv_info(engine.h, idx + 1, info)
            dcin, dstride = info[0], info[3]
            db, _ = bf16_round(np.maximum(rng.standard_normal((N, H * dstride, W * dstride, dcin)), 0).astype(np.float32))
            dsc, dsh = bn_params(rng, cout)
            kw.update(ds_index=idx + 1, ds_in=db, ds_H=H * dstride, ds_W=W * dstride, ds_scale=dsc, ds_shift=dsh)
        else:
            ib, _ = bf16_round(np.maximum(rng.standard_normal((N, H // stride, W // stride, cout)), 0).astype(np.float32))
            kw.update(idt=ib)
    outs = []
    try:
        engine.set_profiling(True)
        for min_tiles in (1 << 30, 1):
            engine.set_option("mc_min_tiles", min_tiles)
            o, st, _ = run_conv(engine, idx, xb, N, H, W, use_tc=1, mode=mode, **kw)
            outs.append((o.copy(), st.copy(), engine.last_profile()["conv_tc"]["kernel"]))
    finally:
        engine.set_option("mc_min_tiles", -1)
        engine.set_profiling(False)
    (o0, s0, k0), (o1, s1, k1) = outs
    targs = lambda k: [int(v) for v in k[k.index("<") + 1:k.index(">")].split(",")]
    assert targs(k0)[4] == 0 and targs(k1)[4] == 1, (k0, k1)           # ... and the second run really took the paired kernel (template argument MC)
    assert np.array_equal(o0, o1)
    assert np.allclose(s0, s1, rtol=1e-6, atol=1e-6 * max(1.0, np.abs(s0).max()))


# --------------------------------------------------------------------------------------------------------------------
# cta_group::2 variant: one MMA over a CTA pair (M = 256), each CTA holding its own pixel tile and HALF of the weight tile
# --------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,mode", [("layer3.1.conv1", 0), ("layer3.0.conv2", 0), ("layer4.1.conv2", 0),
                                       ("layer3.1.conv3", 2), ("layer3.0.conv3", 2), ("layer4.0.conv3", 2)])
@pytest.mark.parametrize("N", [5, 16])
def test_cta_pair_mma_variant_is_bit_identical(engine, name, mode, N):
    """conv_tc_kernel<.., MC=1, S4=0, CG2=1> (tcgen05.mma.cta_group::2 issued by the leader CTA of a pair, completion multicast to both)
    against the one-CTA kernel: the same products accumulated in the same order, so the same bits - RAW (1x1, stride-2 and stride-1 3x3
    through the tap loop), FINAL with an identity tensor and FINAL with the downsample conv, odd and even numbers of pixel tiles."""
    _n, idx, H, W = _role(name)
    rng = np.random.default_rng(idx * 5 + N + mode)
    info = (C.c_int32 * 4)()
    engine.L.busca_conv_info(engine.h, idx, info)
    cin, cout, k, stride = list(info)
    xb, _ = bf16_round(rng.standard_normal((N, H, W, cin)).astype(np.float32))
    kw = {}
    if name.endswith("conv2") or name.endswith("conv3"):
        sc, sh = bn_params(rng, cin)
        kw.update(in_scale=sc, in_shift=sh)
    if mode == 2:
        es, et = bn_params(rng, cout)
        kw.update(e_scale=es, e_shift=et)
        if name.split(".")[1] == "0":
            engine.L.busca_conv_info(engine.h, idx + 1, info)
            dcin, dstride = info[0], info[3]
            db, _ = bf16_round(np.maximum(rng.standard_normal((N, H * dstride, W * dstride, dcin)), 0).astype(np.float32))
            dsc, dsh = bn_params(rng, cout)
            kw.update(ds_index=idx + 1, ds_in=db, ds_H=H * dstride, ds_W=W * dstride, ds_scale=dsc, ds_shift=dsh)
        else:
            ib, _ = bf16_round(np.maximum(rng.standard_normal((N, H // stride, W // stride, cout)), 0).astype(np.float32))
            kw.update(idt=ib)
    outs = []
    try:
        engine.set_profiling(True)
        for min_tiles in (-1, 1):
            engine.set_option("cg2_min_tiles", min_tiles)
            o, st, _ = run_conv(engine, idx, xb, N, H, W, use_tc=1, mode=mode, **kw)
            outs.append((o.copy(), st.copy(), engine.last_profile()["conv_tc"]["kernel"]))
    finally:
        engine.set_option("cg2_min_tiles", -1)
        engine.set_profiling(False)
    (o0, s0, k0), (o1, s1, k1) = outs
    targs = lambda k: [int(v) for v in k[k.index("<") + 1:k.index(">")].split(",")]
    assert targs(k0)[6] == 0 and targs(k1)[6] == 1, (k0, k1)     # the second run took the cta_group::2 kernel (template argument CG2)
    assert np.array_equal(o0, o1)
    assert np.allclose(s0, s1, rtol=1e-6, atol=1e-6 * max(1.0, np.abs(s0).max()))
