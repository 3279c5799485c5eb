// Crop-and-resize gather: frame (uint8 BGR HWC, resident in HBM) -> 384x128x3 uint8 patches in the patch bank.
//
// Reproduces, bit for bit, busca/tracking.py:62-113 (floor/ceil cut-out, clip, pad with the truncated mean of the
// clipped window) followed by cv2.resize(..., (128,384), INTER_LINEAR) for 8-bit images: 11-bit fixed-point
// coefficients, x clamps the coefficient / y clamps the row index, and the silent INTER_AREA switch at an exact
// 2x down-scale (oracle/crop.py is the CPU restatement).  No cut-out is materialised:
// taps that fall outside the clipped window read the pad scalar.
//
// One CTA per crop.  Phase 1: window sum (for the pad value) with 128-bit loads where alignment allows.
// Phase 2: coefficient tables in shared memory.  Phase 3: every thread produces 16 consecutive output bytes
// per iteration and stores them with one 128-bit store (rows are 384 B = 24 x 16 B, patches are 16 B aligned).
#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int CROP_THREADS = 256;
constexpr float COEF_SCALE = 2048.f;

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
                                          int slot, uint8_t *__restrict__ bank) {
    __shared__ CropWin win;
    __shared__ unsigned long long ssum[CROP_THREADS / 32];
    __shared__ int xs0[PATCH_W], xs1[PATCH_W];
    __shared__ short xa0[PATCH_W], xa1[PATCH_W];
    __shared__ int yr0[PATCH_H], yr1[PATCH_H];
    __shared__ short yb0[PATCH_H], yb1[PATCH_H];

    if (slot < 0) return;
    const int tid = threadIdx.x;
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
        for (int q = tid; q < PATCH_BYTES / 16; q += CROP_THREADS) reinterpret_cast<uint4 *>(out)[q] = z;
        return;
    }

    // ---- phase 1: sum of the clipped window over all three channels
    {
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
    }
    for (int dy = tid; dy < PATCH_H; dy += CROP_THREADS) {
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
    auto fetch = [&](int r, int x, int c) -> int {   // cut-out coordinates
        const int fy = r + Y1, fxp = x + X1;
        if (fy < Y1c || fy >= Y2c || fxp < X1c || fxp >= X2c) return pad;
        return (int)__ldg(frame + (size_t)fy * row_stride + (size_t)fxp * 3 + c);
    };

    // ---- phase 3: 16 output bytes per thread per iteration
    const bool area = win.area2x != 0;
    for (int q = tid; q < PATCH_BYTES / 16; q += CROP_THREADS) {
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
    crop_body(frame, H, W, row_stride, boxes + 4 * (size_t)i, slots[i], bank);
}

// up to 4 boxes passed BY VALUE in the kernel parameters: the adapters crop the Kalman proposal of every unmatched track with
// one single-box call each (byte_tracker.py:468-479) - no host->device copy of boxes / slots for those calls
__global__ void __launch_bounds__(CROP_THREADS) crop_resize_small_kernel(const uint8_t *__restrict__ frame, int H, int W, long long row_stride,
                                                                         const __grid_constant__ CropSmall s, uint8_t *__restrict__ bank) {
    const int i = blockIdx.x;
    if (i >= s.n) return;
    crop_body(frame, H, W, row_stride, &s.boxes[4 * i], s.slots[i], bank);
}

}  // namespace

cudaError_t launch_crop_resize_small(const uint8_t *frame, int H, int W, int64_t row_stride, const CropSmall &sm, uint8_t *bank, cudaStream_t s) {
    if (sm.n <= 0) return cudaSuccess;
    crop_resize_small_kernel<<<sm.n, CROP_THREADS, 0, s>>>(frame, H, W, (long long)row_stride, sm, bank);
    return cudaGetLastError();
}

cudaError_t launch_crop_resize(const uint8_t *frame, int H, int W, int64_t row_stride, const double *boxes, int n,
                               const int32_t *slots, uint8_t *bank, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    crop_resize_kernel<<<n, CROP_THREADS, 0, s>>>(frame, H, W, (long long)row_stride, boxes, n, slots, bank);
    return cudaGetLastError();
}
