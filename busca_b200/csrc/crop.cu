// Crop-and-resize gather: frame (uint8 BGR HWC, resident in HBM) -> 384x128x3 uint8 patches in the patch bank.
//
// Reproduces, bit for bit, busca/tracking.py:62-113 (floor/ceil cut-out, clip, pad with the truncated mean of the
// clipped window) followed by cv2.resize(..., (128,384), INTER_LINEAR) for 8-bit images: 11-bit fixed-point
// coefficients, x clamps the coefficient / y clamps the row index, and the silent INTER_AREA switch at an exact
// 2x down-scale (oracle/crop.py is the CPU restatement).  No cut-out is materialised:
// taps that fall outside the clipped window read the pad scalar.
//
// CROP_PARTS CTAs per crop (96 output rows each; the window sum is recomputed by each - it hits L2).  Phase 1: window sum (for the pad value) with 128-bit loads where alignment allows.  Phase 2: coefficient tables
// in shared memory and the partition of the 384 output rows into BANDS whose source rows fit a 16 KB staging buffer.  Phase 3, per
// band: the source rows of the clipped window are STAGED in shared memory by the bulk-copy (TMA) engine - one cp.async.bulk per row
// over the 16-byte-aligned span that covers it, completion on an mbarrier, two buffers so the copies of band b+1 run under the blend
// of band b - and every thread blends 16 consecutive output bytes per iteration from shared memory (four byte taps each, no global
// load in the loop) and stores them with one 128-bit store (rows are 384 B = 24 x 16 B, patches are 16 B aligned).  Windows wider than
// the staging buffer allows (> 2700 pixels) take the direct path (taps read through the read-only cache).
#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int CROP_THREADS = 256;
constexpr float COEF_SCALE = 2048.f;
constexpr int CROP_STAGE_BYTES = 16384;          // per staging buffer (two of them)
constexpr int CROP_PARTS = 4;                    // CTAs per crop (blockIdx.y): 96 output rows each - a frame's few hundred crops fill the machine
constexpr int CROP_ROWS = PATCH_H / CROP_PARTS;
constexpr int CROP_PARTS_SMALL = 16;             // single-box calls (the adapters' Kalman proposals): 24 rows per CTA - the launch is pure latency

__device__ __forceinline__ uint32_t smem_u32_(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init_(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32_(bar)), "r"(count) : "memory");
}

struct CropWin {
    int X1, Y1, sw, sh;          // integer cut-out origin and size (may extend outside the frame)
    int X1c, Y1c, X2c, Y2c;      // clipped to the frame
    int pad;                     // uint8 pad value
    int empty;                   // clipped window empty or zero extent -> all-zero patch
    int area2x;                  // exact 2x down-scale -> INTER_AREA
};

__device__ __forceinline__ int clamp_coord(double v) {
    v = fmin(fmax(v, -1073741824.0), 1073741824.0);
    return (int)v;
}

__device__ __forceinline__ void crop_body(const uint8_t *__restrict__ frame, int H, int W, long long row_stride, const double *b,
                                          int slot, uint8_t *__restrict__ bank, int rows_per_cta) {
    __shared__ CropWin win;
    __shared__ unsigned long long ssum[CROP_THREADS / 32];
    __shared__ int xs0[PATCH_W], xs1[PATCH_W];
    __shared__ short xa0[PATCH_W], xa1[PATCH_W];
    __shared__ int yr0[PATCH_H], yr1[PATCH_H];
    __shared__ short yb0[PATCH_H], yb1[PATCH_H];
    __shared__ __align__(128) uint8_t stage[2][CROP_STAGE_BYTES];
    __shared__ uint64_t sbar[2];
    __shared__ short bstart[PATCH_H + 1];
    __shared__ int4 xt[PATCH_W];
    __shared__ int nbands;

    if (slot < 0) return;
    const int tid = threadIdx.x;
    const int dy_lo = blockIdx.y * rows_per_cta, dy_hi = dy_lo + rows_per_cta;      // this CTA's output rows
    uint8_t *out = bank + (size_t)slot * PATCH_BYTES;

    if (tid == 0) {
        CropWin w;
        int X1 = clamp_coord(floor(b[0])), Y1 = clamp_coord(floor(b[1]));
        int X2 = clamp_coord(ceil(b[2])), Y2 = clamp_coord(ceil(b[3]));
        w.X1 = X1; w.Y1 = Y1;
        w.sw = X2 - X1; w.sh = Y2 - Y1;
        w.X1c = min(max(X1, 0), W); w.X2c = min(max(X2, 0), W);
        w.Y1c = min(max(Y1, 0), H); w.Y2c = min(max(Y2, 0), H);
        w.empty = (w.X2c <= w.X1c) || (w.Y2c <= w.Y1c) || w.sw <= 0 || w.sh <= 0;
        w.area2x = (w.sw == 2 * PATCH_W) && (w.sh == 2 * PATCH_H);
        w.pad = 0;
        win = w;
    }
    __syncthreads();

    if (win.empty) {                                  // np.mean of an empty crop is NaN -> pad casts to 0
        uint4 z = make_uint4(0, 0, 0, 0);
        for (int q = dy_lo * (PATCH_W * 3 / 16) + tid; q < dy_hi * (PATCH_W * 3 / 16); q += CROP_THREADS) reinterpret_cast<uint4 *>(out)[q] = z;
        return;
    }

    // ---- phase 1: sum of the clipped window over all three channels - only when the cut-out leaves the frame (no tap reads the pad
    // value otherwise; most boxes are inside, and for the adapters' single-box calls this dependent-load chain was a third of the launch)
    const bool clipped = win.X1c != win.X1 || win.Y1c != win.Y1 || win.X2c != win.X1 + win.sw || win.Y2c != win.Y1 + win.sh;
    if (!clipped) {
        if ((tid & 31) == 0) ssum[tid >> 5] = 0;
    } else {
        const int wbytes = (win.X2c - win.X1c) * 3, rows = win.Y2c - win.Y1c;
        unsigned long long acc = 0;
        for (int r = tid >> 5; r < rows; r += CROP_THREADS / 32) {         // one warp per row
            const uint8_t *p = frame + (size_t)(win.Y1c + r) * row_stride + (size_t)win.X1c * 3;
            unsigned int a32 = 0;
            const int lane = tid & 31;
            // head up to 16-byte alignment, then 128-bit body, then tail
            int head = (int)((16 - ((uintptr_t)p & 15)) & 15);
            if (head > wbytes) head = wbytes;
            for (int k = lane; k < head; k += 32) a32 += p[k];
            const int body = (wbytes - head) >> 4;
            const uint4 *pv = reinterpret_cast<const uint4 *>(p + head);
            for (int k = lane; k < body; k += 32) {
                uint4 v = __ldg(pv + k);
                a32 += __vsadu4(v.x, 0) + __vsadu4(v.y, 0) + __vsadu4(v.z, 0) + __vsadu4(v.w, 0);
            }
            for (int k = head + (body << 4) + lane; k < wbytes; k += 32) a32 += p[k];
            acc += a32;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if ((tid & 31) == 0) ssum[tid >> 5] = acc;
    }
    // ---- phase 2: coefficient tables (cv2 resize.cpp: HResizeLinear / VResizeLinear set-up)
    if (tid < PATCH_W) {
        const int sw = win.sw;
        const double scale = 1.0 / ((double)PATCH_W / (double)sw);
        float fx = (float)__dsub_rn(__dmul_rn((double)tid + 0.5, scale), 0.5);
        int sx = (int)floorf(fx);
        fx = __fsub_rn(fx, (float)sx);
        if (sx < 0) { sx = 0; fx = 0.f; }
        if (sx >= sw - 1) { sx = sw - 1; fx = 0.f; }
        xs0[tid] = sx;
        xs1[tid] = min(sx + 1, sw - 1);
        xa0[tid] = (short)__float2int_rn(__fmul_rn(__fsub_rn(1.f, fx), COEF_SCALE));
        xa1[tid] = (short)__float2int_rn(__fmul_rn(fx, COEF_SCALE));
        // the same, packed for the staged path: byte offsets of the two taps inside a window row (-1 = outside the clipped window)
        const int f0 = xs0[tid] + win.X1, f1 = xs1[tid] + win.X1;
        const int o0 = (f0 < win.X1c || f0 >= win.X2c) ? -1 : (f0 - win.X1c) * 3, o1 = (f1 < win.X1c || f1 >= win.X2c) ? -1 : (f1 - win.X1c) * 3;
        xt[tid] = make_int4(o0, o1, (int)(((uint32_t)(uint16_t)xa0[tid]) | ((uint32_t)(uint16_t)xa1[tid] << 16)), 0);
    }
    for (int dy = dy_lo + tid; dy < dy_hi; dy += CROP_THREADS) {
        const int sh = win.sh;
        const double scale = 1.0 / ((double)PATCH_H / (double)sh);
        float fy = (float)__dsub_rn(__dmul_rn((double)dy + 0.5, scale), 0.5);
        int sy = (int)floorf(fy);
        fy = __fsub_rn(fy, (float)sy);
        yr0[dy] = min(max(sy, 0), sh - 1);
        yr1[dy] = min(max(sy + 1, 0), sh - 1);
        yb0[dy] = (short)__float2int_rn(__fmul_rn(__fsub_rn(1.f, fy), COEF_SCALE));
        yb1[dy] = (short)__float2int_rn(__fmul_rn(fy, COEF_SCALE));
    }
    __syncthreads();
    if (tid == 0) {
        unsigned long long tot = 0;
        for (int k = 0; k < CROP_THREADS / 32; ++k) tot += ssum[k];
        unsigned long long cnt = (unsigned long long)(win.X2c - win.X1c) * (win.Y2c - win.Y1c) * 3ull;
        win.pad = (int)(tot / cnt);                   // uint8(trunc(np.mean(window)))
    }
    __syncthreads();

    const int X1 = win.X1, Y1 = win.Y1, X1c = win.X1c, X2c = win.X2c, Y1c = win.Y1c, Y2c = win.Y2c, pad = win.pad;
    const bool area = win.area2x != 0;
    const int wbytes = (X2c - X1c) * 3;
    const int pitch = (wbytes + 15 + 15) & ~15;                 // staged row: the aligned span that covers [X1c*3, X2c*3) of a frame row
    const int rows_max = CROP_STAGE_BYTES / pitch;
    const bool staged = rows_max >= 4;
    const uint8_t *win0 = frame + (size_t)X1c * 3;               // + frame row * row_stride = first byte of the window in that row

    if (staged) {
        // ---- band plan: output rows [bstart[b], bstart[b+1]) read cut-out rows lo_b .. hi_b with hi_b - lo_b + 1 <= rows_max
        if (tid == 0) {
            int nb = 0, lo = 0;
            for (int dy = dy_lo; dy < dy_hi; ++dy) {
                const int r0 = area ? 2 * dy : yr0[dy], r1 = area ? 2 * dy + 1 : yr1[dy];
                if (dy == dy_lo || r1 - lo + 1 > rows_max) { bstart[nb++] = (short)dy; lo = r0; }
            }
            bstart[nb] = (short)dy_hi;
            nbands = nb;
            mbar_init_(&sbar[0], 1);
            mbar_init_(&sbar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        const int nb = nbands;
        // source rows of band b, clipped to the window (cut-out row r is frame row r + Y1)
        auto band_rows = [&](int b, int &lo, int &hi) {
            const int dy0 = bstart[b], dy1 = bstart[b + 1] - 1;
            lo = max(area ? 2 * dy0 : yr0[dy0], Y1c - Y1);
            hi = min(area ? 2 * dy1 + 1 : yr1[dy1], Y2c - Y1 - 1);
        };
        auto issue = [&](int b) {                                 // warp 0: one bulk copy per source row of band b into buffer b & 1
            int lo, hi;
            band_rows(b, lo, hi);
            const int lane = tid & 31;
            uint32_t bytes = 0;
            for (int r = lo + lane; r <= hi; r += 32) {
                const uintptr_t a = (uintptr_t)(win0 + (size_t)(r + Y1) * row_stride);
                bytes += (uint32_t)((((a & 15) + wbytes + 15) & ~(uintptr_t)15));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, o);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the buffer was read by ordinary loads two bands ago
            if (lane == 0) {
                if (bytes) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32_(&sbar[b & 1])), "r"(bytes) : "memory");
                else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32_(&sbar[b & 1])) : "memory");
            }
            __syncwarp();
            for (int r = lo + lane; r <= hi; r += 32) {
                const uintptr_t a = (uintptr_t)(win0 + (size_t)(r + Y1) * row_stride);
                const uint32_t n = (uint32_t)((((a & 15) + wbytes + 15) & ~(uintptr_t)15));
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32_(&stage[b & 1][(r - lo) * pitch])),
                             "l"((uint64_t)(a & ~(uintptr_t)15)), "r"(n), "r"(smem_u32_(&sbar[b & 1]))
                             : "memory");
            }
        };
        if (tid < 32) issue(0);
        for (int b = 0; b < nb; ++b) {
            if (tid < 32 && b + 1 < nb) issue(b + 1);
            int lo, hi;
            band_rows(b, lo, hi);
            {
                const uint32_t parity = (uint32_t)(b >> 1) & 1u;
                uint32_t ok = 0;
                while (!ok)
                    asm volatile(
                        "{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}"
                        : "=r"(ok)
                        : "r"(smem_u32_(&sbar[b & 1])), "r"(parity)
                        : "memory");
            }
            const uint8_t *sb = stage[b & 1];
            const uint32_t mis0 = (uint32_t)((uintptr_t)win0 & 15), stride15 = (uint32_t)(row_stride & 15);
            auto row_base = [&](int r) -> int {                 // staged offset of cut-out row r's first window byte, -1 outside the window
                if (r < lo || r > hi) return -1;
                return (r - lo) * pitch + (int)((mis0 + (uint32_t)(r + Y1) * stride15) & 15u);
            };
            // one unit = 16 output pixels of one row (48 bytes, three 128-bit stores): the per-pixel work (table entry, bounds, row bases) is
            // shared by the three channels
            const int u0 = bstart[b] * (PATCH_W / 16), u1 = bstart[b + 1] * (PATCH_W / 16);
            for (int u = u0 + tid; u < u1; u += CROP_THREADS) {
                const int dy = u / (PATCH_W / 16), px0 = (u - dy * (PATCH_W / 16)) * 16;
                uint32_t wds[12];
                if (!area) {
                    const int base0 = row_base(yr0[dy]), base1 = row_base(yr1[dy]), wb0 = yb0[dy], wb1 = yb1[dy];
#pragma unroll
                    for (int i = 0; i < 12; ++i) wds[i] = 0;
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const int4 t = xt[px0 + k];                 // byte offsets of the two taps inside a window row (-1: outside), coefficients
                        const int a0 = (short)(t.z & 0xffff), a1 = t.z >> 16;
                        const int i00 = (base0 < 0 || t.x < 0) ? -1 : base0 + t.x, i01 = (base0 < 0 || t.y < 0) ? -1 : base0 + t.y;
                        const int i10 = (base1 < 0 || t.x < 0) ? -1 : base1 + t.x, i11 = (base1 < 0 || t.y < 0) ? -1 : base1 + t.y;
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const int p00 = i00 < 0 ? pad : (int)sb[i00 + c], p01 = i01 < 0 ? pad : (int)sb[i01 + c];
                            const int p10 = i10 < 0 ? pad : (int)sb[i10 + c], p11 = i11 < 0 ? pad : (int)sb[i11 + c];
                            const int h0 = p00 * a0 + p01 * a1, h1 = p10 * a0 + p11 * a1;
                            const int v = ((((wb0 * (h0 >> 4)) >> 16) + ((wb1 * (h1 >> 4)) >> 16) + 2) >> 2);
                            const int byte = k * 3 + c;
                            wds[byte >> 2] |= (uint32_t)(v & 0xff) << (8 * (byte & 3));
                        }
                    }
                } else {
                    const int base0 = row_base(2 * dy), base1 = row_base(2 * dy + 1);
#pragma unroll
                    for (int i = 0; i < 12; ++i) wds[i] = 0;
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const int x0 = 2 * (px0 + k) + X1, x1 = x0 + 1;                     // frame columns of the 2x2 block
                        const int o0 = (x0 < X1c || x0 >= X2c) ? -1 : (x0 - X1c) * 3, o1 = (x1 < X1c || x1 >= X2c) ? -1 : (x1 - X1c) * 3;
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const int p00 = (base0 < 0 || o0 < 0) ? pad : (int)sb[base0 + o0 + c], p01 = (base0 < 0 || o1 < 0) ? pad : (int)sb[base0 + o1 + c];
                            const int p10 = (base1 < 0 || o0 < 0) ? pad : (int)sb[base1 + o0 + c], p11 = (base1 < 0 || o1 < 0) ? pad : (int)sb[base1 + o1 + c];
                            const int v = (p00 + p01 + p10 + p11 + 2) >> 2;
                            const int byte = k * 3 + c;
                            wds[byte >> 2] |= (uint32_t)(v & 0xff) << (8 * (byte & 3));
                        }
                    }
                }
                uint4 *o4 = reinterpret_cast<uint4 *>(out + ((size_t)dy * PATCH_W + px0) * 3);
                o4[0] = make_uint4(wds[0], wds[1], wds[2], wds[3]);
                o4[1] = make_uint4(wds[4], wds[5], wds[6], wds[7]);
                o4[2] = make_uint4(wds[8], wds[9], wds[10], wds[11]);
            }
            __syncthreads();                                      // buffer b & 1 may be refilled (band b + 2)
        }
        return;
    }

    // ---- direct path (window rows too wide to stage): taps through the read-only cache
    auto fetch = [&](int r, int x, int c) -> int {   // cut-out coordinates
        const int fy = r + Y1, fxp = x + X1;
        if (fy < Y1c || fy >= Y2c || fxp < X1c || fxp >= X2c) return pad;
        return (int)__ldg(frame + (size_t)fy * row_stride + (size_t)fxp * 3 + c);
    };
    for (int q = dy_lo * (PATCH_W * 3 / 16) + tid; q < dy_hi * (PATCH_W * 3 / 16); q += CROP_THREADS) {
        const int dy = q / (PATCH_W * 3 / 16);
        const int b0 = (q - dy * (PATCH_W * 3 / 16)) * 16;
        uint32_t wds[4];
        if (!area) {
            const int r0 = yr0[dy], r1 = yr1[dy], wb0 = yb0[dy], wb1 = yb1[dy];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                uint32_t wv = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int b = b0 + g * 4 + j;
                    const int dx = b / 3, c = b - dx * 3;
                    const int s0 = xs0[dx], s1 = xs1[dx], a0 = xa0[dx], a1 = xa1[dx];
                    const int h0 = fetch(r0, s0, c) * a0 + fetch(r0, s1, c) * a1;
                    const int h1 = fetch(r1, s0, c) * a0 + fetch(r1, s1, c) * a1;
                    const int v = ((((wb0 * (h0 >> 4)) >> 16) + ((wb1 * (h1 >> 4)) >> 16) + 2) >> 2);
                    wv |= (uint32_t)(v & 0xff) << (8 * j);
                }
                wds[g] = wv;
            }
        } else {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                uint32_t wv = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int b = b0 + g * 4 + j;
                    const int dx = b / 3, c = b - dx * 3;
                    const int v = (fetch(2 * dy, 2 * dx, c) + fetch(2 * dy, 2 * dx + 1, c) + fetch(2 * dy + 1, 2 * dx, c) +
                                   fetch(2 * dy + 1, 2 * dx + 1, c) + 2) >> 2;
                    wv |= (uint32_t)(v & 0xff) << (8 * j);
                }
                wds[g] = wv;
            }
        }
        reinterpret_cast<uint4 *>(out)[q] = make_uint4(wds[0], wds[1], wds[2], wds[3]);
    }
}

__global__ void __launch_bounds__(CROP_THREADS) crop_resize_kernel(const uint8_t *__restrict__ frame, int H, int W,
                                                                   long long row_stride, const double *__restrict__ boxes,
                                                                   int n, const int32_t *__restrict__ slots,
                                                                   uint8_t *__restrict__ bank) {
    const int i = blockIdx.x;
    if (i >= n) return;
    crop_body(frame, H, W, row_stride, boxes + 4 * (size_t)i, slots[i], bank, CROP_ROWS);
}

// up to 4 boxes passed BY VALUE in the kernel parameters: the adapters crop the Kalman proposal of every unmatched track with
// one single-box call each (byte_tracker.py:468-479) - no host->device copy of boxes / slots for those calls
__global__ void __launch_bounds__(CROP_THREADS) crop_resize_small_kernel(const uint8_t *__restrict__ frame, int H, int W, long long row_stride,
                                                                         const __grid_constant__ CropSmall s, uint8_t *__restrict__ bank) {
    const int i = blockIdx.x;
    if (i >= s.n) return;
    crop_body(frame, H, W, row_stride, &s.boxes[4 * i], s.slots[i], bank, PATCH_H / CROP_PARTS_SMALL);
}

}  // namespace

cudaError_t launch_crop_resize_small(const uint8_t *frame, int H, int W, int64_t row_stride, const CropSmall &sm, uint8_t *bank, cudaStream_t s) {
    if (sm.n <= 0) return cudaSuccess;
    crop_resize_small_kernel<<<dim3(sm.n, CROP_PARTS_SMALL), CROP_THREADS, 0, s>>>(frame, H, W, (long long)row_stride, sm, bank);
    return cudaGetLastError();
}

cudaError_t launch_crop_resize(const uint8_t *frame, int H, int W, int64_t row_stride, const double *boxes, int n,
                               const int32_t *slots, uint8_t *bank, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    crop_resize_kernel<<<dim3(n, CROP_PARTS), CROP_THREADS, 0, s>>>(frame, H, W, (long long)row_stride, boxes, n, slots, bank);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------------
// Frame ingest: the detector's input tensor [3,H,W] float32 (RGB, normalised) -> the uint8 BGR HWC frame the
// crops read, without the frame leaving HBM.  Reference (adapters/ByteTrack/yolox/evaluators/mot_evaluator.py:198-204):
//   v = x * std + mean (two fp32 operations), RGB -> BGR, clip to [0, 1], (v * 255.0) in fp32, astype(uint8) = truncation.
// HBM bound: 12 B read + 3 B written per pixel.  One thread per 16 pixels: four 128-bit loads per plane, three 128-bit stores.
// ------------------------------------------------------------------------------------------------------------------
namespace {
__device__ __forceinline__ unsigned int ingest_px(float x, float sd, float mu) {
    float v = __fadd_rn(__fmul_rn(x, sd), mu);
    v = fminf(fmaxf(v, 0.0f), 1.0f);                 // np.clip (NaN stays NaN in numpy and casts to 0 here: not a pixel value)
    return (unsigned int)__fmul_rn(v, 255.0f);       // truncation toward zero
}

__global__ void __launch_bounds__(256) frame_ingest_kernel(const float *__restrict__ chw, long long npix, float3 mean, float3 sd,
                                                           uint8_t *__restrict__ bgr) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long p0 = g * 16;
    if (p0 >= npix) return;
    const float *R = chw, *G = chw + npix, *B = chw + 2 * npix;
    if (p0 + 16 <= npix && (npix & 3) == 0) {
        unsigned int w[12];                                  // 16 pixels = 48 bytes = 12 words, packed in registers
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 r = __ldg(reinterpret_cast<const float4 *>(R + p0 + 4 * q));
            const float4 gg = __ldg(reinterpret_cast<const float4 *>(G + p0 + 4 * q));
            const float4 b = __ldg(reinterpret_cast<const float4 *>(B + p0 + 4 * q));
            const unsigned int B0 = ingest_px(b.x, sd.z, mean.z), G0 = ingest_px(gg.x, sd.y, mean.y), R0 = ingest_px(r.x, sd.x, mean.x);
            const unsigned int B1 = ingest_px(b.y, sd.z, mean.z), G1 = ingest_px(gg.y, sd.y, mean.y), R1 = ingest_px(r.y, sd.x, mean.x);
            const unsigned int B2 = ingest_px(b.z, sd.z, mean.z), G2 = ingest_px(gg.z, sd.y, mean.y), R2 = ingest_px(r.z, sd.x, mean.x);
            const unsigned int B3 = ingest_px(b.w, sd.z, mean.z), G3 = ingest_px(gg.w, sd.y, mean.y), R3 = ingest_px(r.w, sd.x, mean.x);
            w[3 * q + 0] = B0 | (G0 << 8) | (R0 << 16) | (B1 << 24);
            w[3 * q + 1] = G1 | (R1 << 8) | (B2 << 16) | (G2 << 24);
            w[3 * q + 2] = R2 | (B3 << 8) | (G3 << 16) | (R3 << 24);
        }
        uint4 *dst = reinterpret_cast<uint4 *>(bgr + p0 * 3);                      // p0 * 3 = 48 g: 16-byte aligned
        dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
        dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
        dst[2] = make_uint4(w[8], w[9], w[10], w[11]);
    } else {
        for (long long p = p0; p < npix && p < p0 + 16; ++p) {
            bgr[p * 3 + 0] = (unsigned char)ingest_px(B[p], sd.z, mean.z);
            bgr[p * 3 + 1] = (unsigned char)ingest_px(G[p], sd.y, mean.y);
            bgr[p * 3 + 2] = (unsigned char)ingest_px(R[p], sd.x, mean.x);
        }
    }
}
}  // namespace

cudaError_t launch_frame_ingest(const float *chw, int H, int W, const float mean[3], const float sd[3], uint8_t *bgr, cudaStream_t s) {
    const long long npix = (long long)H * W;
    if (npix <= 0) return cudaErrorInvalidValue;
    frame_ingest_kernel<<<ceil_div(ceil_div(npix, 16), 256), 256, 0, s>>>(chw, npix, make_float3(mean[0], mean[1], mean[2]),
                                                                          make_float3(sd[0], sd[1], sd[2]), bgr);
    return cudaGetLastError();
}
