"""Diagnostic (NOT collected by pytest: an untested tcgen05 kernel must never run in the graded `-m gpu` suite).
First thing to run on a B200 in round 2:

    timeout 120 python tests/probe_halo.py            # parity of conv3x3_halo_kernel vs the SIMT kernel, then timing vs the tap kernel

For every stride-1 3x3 convolution shape the halo kernel covers (layer1/2/3 conv2 at 96x32, 48x16, 24x8) it runs the
same check as tests/test_gpu_conv_tc.py::test_conv_tc_input_bn_relu_in_smem with busca_set_option("halo", 1): output
against the SIMT kernel fed the pre-activated tensor (up to the dropped per-channel constant), statistics against the
stored values - and then against the tap-by-tap tensor-core kernel (halo off), which must agree to accumulation order."""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from busca_b200 import synth  # noqa: E402
from busca_b200.engine import Engine  # noqa: E402
import test_gpu_conv_tc as T  # noqa: E402


def main():
    weights = synth.make_weights(0)
    e = Engine(precision="bf16", bank_slots=8)
    e.load_state_dict({k: v for k, v in weights.items() if "running" not in k and "num_batches" not in k})
    shapes = [(n, i, h, w) for n, i, h, w in T.ROLES if n.endswith("conv2") and w in (32, 16, 8)]
    seen, todo = set(), []
    for n, i, h, w in shapes:
        info = (T.C.c_int32 * 4)()
        e.L.busca_conv_info(e.h, i, info)
        cin, cout, k, stride = list(info)
        if k == 3 and stride == 1 and (cin, cout, h) not in seen:
            seen.add((cin, cout, h))
            todo.append((n, i, h, w, cin, cout))
    ok = True
    for name, idx, H, W, cin, cout in todo:
        for N in (1, 3, 9):
            rng = np.random.default_rng(idx * 10 + N)
            xb, xr = T.bf16_round(rng.standard_normal((N, H, W, cin)).astype(np.float32))
            sc, sh = T.bn_params(rng, cin)
            act = np.maximum(xr.astype(np.float64) * sc + sh, 0).astype(np.float32)
            ab, _ = T.bf16_round(act)
            ref_o, _, _ = T.run_conv(e, idx, ab, N, H, W, use_tc=0)
            e.set_option("halo", 0)
            tap_o, tap_st, _ = T.run_conv(e, idx, xb, N, H, W, use_tc=1, mode=0, in_scale=sc, in_shift=sh)
            e.set_option("halo", 1)
            print(f"{name} [{cin}>{cout} {H}x{W}] N={N}: launching halo kernel ...", end=" ", flush=True)
            halo_o, halo_st, _ = T.run_conv(e, idx, xb, N, H, W, use_tc=1, mode=0, in_scale=sc, in_shift=sh)
            e.set_option("halo", 0)
            ref = T.bf16_to_f32(ref_o).astype(np.float64)
            tap = T.bf16_to_f32(tap_o).astype(np.float64)
            halo = T.bf16_to_f32(halo_o).astype(np.float64)
            got = halo + T.fold_const(weights, name, sh)
            err_ref = np.abs(got - ref).max() / np.abs(ref).max()
            err_tap = np.abs(halo - tap).max() / np.abs(tap).max()
            want_st = np.concatenate([halo.reshape(-1, cout).sum(0), (halo.reshape(-1, cout) ** 2).sum(0)])
            st_ok = np.allclose(halo_st, want_st, rtol=1e-3, atol=1e-3 * np.abs(want_st).max())
            good = np.isfinite(halo).all() and err_ref < 2e-2 and err_tap < 1e-2 and st_ok
            ok &= bool(good)
            print(f"finite={np.isfinite(halo).all()} err_vs_simt={err_ref:.2e} err_vs_tap={err_tap:.2e} stats_ok={st_ok} {'OK' if good else 'FAIL'}")
    print("PARITY", "OK" if ok else "FAILED")
    if not ok:
        return 1
    # timing at MOT20 scale of one layer1 conv2 (N = 2680 distinct patches would need 22 GB; 512 is enough for steady state)
    N = 512
    for name, idx, H, W, cin, cout in todo:
        rng = np.random.default_rng(1)
        xb, _ = T.bf16_round(rng.standard_normal((N, H, W, cin)).astype(np.float32))
        sc, sh = T.bn_params(rng, cin)
        for halo in (0, 1):
            e.set_option("halo", halo)
            T.run_conv(e, idx, xb, N, H, W, use_tc=1, mode=0, in_scale=sc, in_shift=sh)
            e.set_profiling(True)
            T.run_conv(e, idx, xb, N, H, W, use_tc=1, mode=0, in_scale=sc, in_shift=sh)
            prof = e.last_profile()
            e.set_profiling(False)
            ms = sum(v["ms"] for k, v in prof.items() if k.startswith("conv"))
            fl = 2.0 * N * H * W * cout * cin * 9
            print(f"{name} [{cin}>{cout} {H}x{W}] N={N} halo={halo}: {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s", flush=True)
        e.set_option("halo", 0)
    return 0


if __name__ == "__main__":
    t0 = time.time()
    rc = main()
    print(f"done in {time.time() - t0:.1f} s")
    sys.exit(rc)
