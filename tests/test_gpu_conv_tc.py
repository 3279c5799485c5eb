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
