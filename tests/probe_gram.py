"""Diagnostic (not collected by pytest): A^T A on the tensor cores with both operands read MN-major from K-major-stored, 128-byte-
swizzled pixel tiles (busca_debug_gram) against numpy - the hardware assumption of the Gram-matrix batch statistics.

    timeout 60 python tests/probe_gram.py"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from busca_b200.engine import Engine  # noqa: E402


def bf16_round(x):
    u = np.ascontiguousarray(x, np.float32).view(np.uint32)
    r = ((u >> 16) & 1) + 0x7FFF
    b = ((u + r) >> 16).astype(np.uint16)
    return b, (b.astype(np.uint32) << 16).view(np.float32)


def main():
    e = Engine(precision="bf16", bank_slots=8)
    ok = True
    for nb in (1, 2, 4):
        Cc = 64 * nb
        rng = np.random.default_rng(nb)
        for fill in ("index", "random"):
            if fill == "index":
                a = (np.arange(128)[:, None] % 7 + 1.0) * ((np.arange(Cc)[None, :] % 5) + 1.0) * (np.arange(Cc)[None, :] // 64 + 1)
            else:
                a = rng.standard_normal((128, Cc))
            ab, ar = bf16_round(a.astype(np.float32))
            out = np.empty((Cc, Cc), np.float32)
            rc = e.L.busca_debug_gram(e.h, ab.ctypes.data_as(C.c_void_p), nb, out.ctypes.data_as(C.c_void_p))
            assert rc == 0, e.L.busca_last_error().decode()
            want = ar.astype(np.float64).T @ ar.astype(np.float64)
            err = np.abs(out - want).max() / np.abs(want).max()
            good = np.isfinite(out).all() and err < 1e-5
            ok &= bool(good)
            print(f"nb={nb} ({Cc} channels) fill={fill}: max rel err {err:.2e} {'OK' if good else 'FAIL'}")
            if not good:
                bad = np.argwhere(np.abs(out - want) > 1e-4 * np.abs(want).max())[:6]
                print("   first mismatches (i, j, got, want):", [(int(i), int(j), float(out[i, j]), float(want[i, j])) for i, j in bad])
    print("GRAM", "OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
