"""Diagnostic (not collected by pytest): does a tcgen05 A descriptor that starts a whole number of 128-byte rows into a
1024-byte-aligned SWIZZLE_128B tile read the rows and chunks conv3x3_halo_kernel expects?  Run FIRST in round 2:

    timeout 60 python tests/probe_umma.py

For shifts 0..70 (the halo kernel uses r*(W+2)+q <= 70) and both settings of descriptor bits 49-51 it reports whether
D[m][n] == A[m+shift][n] for the row-index fill and the column-index fill."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from busca_b200.engine import Engine  # noqa: E402


def main():
    e = Engine(precision="bf16", bank_slots=8)
    out = np.empty((128, 64), np.float32)
    m = np.arange(128, dtype=np.float32)[:, None]
    n = np.arange(64, dtype=np.float32)[None, :]
    verdict = {}
    for use_bo in (1, 0):
        bad_rows, bad_cols = [], []
        for shift in list(range(0, 19)) + [33, 34, 35, 36, 37, 42, 68, 69, 70]:
            for fill in (0, 1):
                rc = e.L.busca_debug_umma_rowshift(e.h, shift, fill, use_bo, out.ctypes.data_as(C.c_void_p))
                assert rc == 0, e.L.busca_last_error().decode()
                want = np.broadcast_to(m + shift if fill == 0 else n, out.shape)
                if not np.array_equal(out, want):
                    (bad_rows if fill == 0 else bad_cols).append(shift)
                    if len(bad_rows) + len(bad_cols) <= 3:
                        d = np.argwhere(out != want)[:4]
                        print(f"  base_offset={use_bo} shift={shift} fill={fill}: first mismatches (m, n, got, want):",
                              [(int(a), int(b), float(out[a, b]), float(want[a, b])) for a, b in d])
        verdict[use_bo] = (bad_rows, bad_cols)
        print(f"base_offset={'(start>>7)&7' if use_bo else '0'}: row mapping wrong at shifts {bad_rows or 'none'}; chunk de-swizzle wrong at shifts {bad_cols or 'none'}")
    # measured on a B200 (profiles/r02a_probe_umma.log): clean with base_offset 0, chunk order wrong with (start >> 7) & 7 -
    # the swizzle XOR follows the address bits of each row, which is what conv3x3_halo_kernel now relies on
    ok = not verdict[0][0] and not verdict[0][1]
    print("conv3x3_halo_kernel's descriptor assumption (plain descriptor, base_offset 0)", "HOLDS" if ok else "DOES NOT HOLD (see above)")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
