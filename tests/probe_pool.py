"""Diagnostic (not collected by pytest): the experimental monotone max-pool kernel (pool before BN + ReLU, reid.cu) against
the shipped one, bit for bit, on the stem's shape, then timing.

    timeout 60 python tests/probe_pool.py
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
from busca_b200.engine import Engine  # noqa: E402
from test_gpu_conv_tc import bf16_round, bf16_to_f32  # noqa: E402


def pool(e, xb, scale, shift, mono):
    N, H, W, Cc = xb.shape
    out = np.empty((N, H // 2, W // 2, Cc), np.uint16)
    e.set_option("pool_mono", mono)
    rc = e.L.busca_debug_maxpool(e.h, xb.ctypes.data_as(C.c_void_p), N, H, W, Cc, scale.ctypes.data_as(C.c_void_p),
                                 shift.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    e.set_option("pool_mono", 0)
    assert rc == 0, e.L.busca_last_error().decode()
    return out


def main():
    e = Engine(precision="bf16", bank_slots=8)
    rng = np.random.default_rng(5)
    ok = True
    for N, H, W, Cc in ((3, 192, 64, 64), (2, 16, 8, 32)):
        xb, xr = bf16_round(rng.standard_normal((N, H, W, Cc)).astype(np.float32) * 3)
        scale = (rng.uniform(0.2, 2.0, Cc) * np.where(rng.uniform(size=Cc) < 0.3, -1, 1)).astype(np.float32)
        scale[1] = 0.0
        shift = (0.5 * rng.standard_normal(Cc)).astype(np.float32)
        a, b = pool(e, xb, scale, shift, 0), pool(e, xb, scale, shift, 1)
        # values must be equal; the only admissible bit difference is the sign of a zero
        va, vb = bf16_to_f32(a), bf16_to_f32(b)
        same_bits = np.array_equal(a, b)
        same_vals = np.array_equal(va, vb)
        ok &= same_vals
        print(f"[{N}x{H}x{W}x{Cc}] bitwise equal: {same_bits}; values equal: {same_vals}; differing words: {(a != b).sum()}")
    N = 512
    xb, _ = bf16_round(rng.standard_normal((N, 192, 64, 64)).astype(np.float32))
    scale = rng.uniform(0.5, 1.5, 64).astype(np.float32)
    shift = np.zeros(64, np.float32)
    for mono in (0, 1):
        pool(e, xb, scale, shift, mono)
        e.set_profiling(True)
        pool(e, xb, scale, shift, mono)
        ms = sum(v["ms"] for v in e.last_profile().values())
        e.set_profiling(False)
        gb = N * 192 * 64 * 64 * 2 * 1.25 / 1e9
        print(f"N={N} pool_mono={mono}: {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s (compulsory bytes)")
    print("POOL", "OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
