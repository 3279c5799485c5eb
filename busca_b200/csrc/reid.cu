// ReID ResNet-50 with batch-statistic BatchNorm (busca/reid/resnet.py:85-128, 266-322; network.py:542-570),
// exact-arithmetic path: SIMT fp32 implicit-GEMM convolutions.  This is the fp32 parity mode and the on-device
// reference the tcgen05 bf16 path (conv_tc.cu) is validated against.
//
// Layout: activations NHWC ([N*H*W, C] row-major), weights [Cout][kh][kw][Cin] (K-major), so every convolution is
// C[M, Cout] = A[M, K] * W[Cout, K]^T with M = N*Ho*Wo, K = kh*kw*Cin and the A rows gathered on the fly.
// BatchNorm in training mode needs the statistics of the WHOLE batch before anything can be normalised, so every
// conv kernel (a) writes the raw output, (b) accumulates per-channel sum / sum-of-squares in its epilogue, and the
// consumer applies y = relu(x*scale + shift) while loading its A operand ("deferred BN").  Residual joins are
// materialised by bn_add_relu.
#include "common.cuh"
#include <cstdlib>

#include "kernels.h"

namespace {

// ------------------------------------------------------------------------------------------------
// generic implicit-GEMM conv / linear, SIMT fp32 accumulate
// ------------------------------------------------------------------------------------------------
struct IGemmParams {
    const void *in;
    void *out;
    const float *w;                  // [Cout][K]
    int N, H, W, Cin, Ho, Wo, Cout, ksz, stride, pad;
    long long M;
    int K;
    const float *in_scale, *in_shift;
    const float *bias;
    const float *residual;           // fp32 [M,Cout]
    float alpha;
    int act;
    double *stats;
};

constexpr int BM = 128, BK = 16, IG_THREADS = 256;

template <typename TIn, typename TOut, int BN>
__global__ void __launch_bounds__(IG_THREADS, 2) igemm_simt_kernel(IGemmParams p) {
    constexpr int TN = BN / 16;
    constexpr int AS = BM + 4, BS = BN + 4;
    __shared__ __align__(16) float As[2][BK][AS];
    __shared__ __align__(16) float Bs[2][BK][BS];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const TIn *in = reinterpret_cast<const TIn *>(p.in);

    // ---- A-load bookkeeping
    constexpr bool kF32 = sizeof(TIn) == 4;
    constexpr int AROWS = kF32 ? 2 : 1;
    int a_row[AROWS], a_hi0[AROWS], a_wi0[AROWS];
    long long a_img[AROWS];
    bool a_ok[AROWS];
    const int a_kq = kF32 ? (tid & 3) : (tid & 1);           // which 4-float (or 8-bf16) group of the 16-wide k tile
#pragma unroll
    for (int i = 0; i < AROWS; ++i) {
        a_row[i] = kF32 ? ((tid >> 2) + i * 64) : (tid >> 1);
        long long m = m0 + a_row[i];
        a_ok[i] = m < p.M;
        long long mm = a_ok[i] ? m : 0;
        int wo = (int)(mm % p.Wo);
        long long t = mm / p.Wo;
        int ho = (int)(t % p.Ho);
        a_img[i] = (t / p.Ho) * (long long)p.H * p.W;
        a_hi0[i] = ho * p.stride - p.pad;
        a_wi0[i] = wo * p.stride - p.pad;
    }
    // ---- B-load bookkeeping: BN x 16 floats = BN*4 float4
    constexpr int BLOADS = BN * 4 / IG_THREADS;
    float areg[AROWS][kF32 ? 4 : 8];
    float4 breg[BLOADS];

    auto load_tiles = [&](int kt) {
        const int kbase = kt * BK;
        const int tap = kbase / p.Cin;
        const int c0 = kbase - tap * p.Cin + a_kq * (kF32 ? 4 : 8);
        const int r = tap / p.ksz, s = tap - r * p.ksz;
#pragma unroll
        for (int i = 0; i < AROWS; ++i) {
            const int hi = a_hi0[i] + r, wi = a_wi0[i] + s;
            const bool ok = a_ok[i] && hi >= 0 && hi < p.H && wi >= 0 && wi < p.W;
            constexpr int NV = kF32 ? 4 : 8;
            if (ok) {
                const TIn *src = in + ((a_img[i] + (long long)hi * p.W + wi) * p.Cin + c0);
                if constexpr (kF32) {
                    float4 v = __ldg(reinterpret_cast<const float4 *>(src));
                    areg[i][0] = v.x; areg[i][1] = v.y; areg[i][2] = v.z; areg[i][3] = v.w;
                } else {
                    uint4 v = __ldg(reinterpret_cast<const uint4 *>(src));
                    const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&v);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float2 f = __bfloat1622float2(h[j]);
                        areg[i][2 * j] = f.x; areg[i][2 * j + 1] = f.y;
                    }
                }
                if (p.in_scale) {
#pragma unroll
                    for (int j = 0; j < NV; ++j)
                        areg[i][j] = fmaxf(fmaf(areg[i][j], __ldg(p.in_scale + c0 + j), __ldg(p.in_shift + c0 + j)), 0.f);
                }
            } else {
#pragma unroll
                for (int j = 0; j < NV; ++j) areg[i][j] = 0.f;
            }
        }
#pragma unroll
        for (int i = 0; i < BLOADS; ++i) {
            const int idx = tid + i * IG_THREADS;
            const int col = idx >> 2, kq = idx & 3;
            breg[i] = (n0 + col < p.Cout)
                          ? __ldg(reinterpret_cast<const float4 *>(p.w + (long long)(n0 + col) * p.K + kbase + kq * 4))
                          : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int i = 0; i < AROWS; ++i) {
            constexpr int NV = kF32 ? 4 : 8;
#pragma unroll
            for (int j = 0; j < NV; ++j) As[buf][a_kq * NV + j][a_row[i]] = areg[i][j];
        }
#pragma unroll
        for (int i = 0; i < BLOADS; ++i) {
            const int idx = tid + i * IG_THREADS;
            const int col = idx >> 2, kq = idx & 3;
            Bs[buf][kq * 4 + 0][col] = breg[i].x;
            Bs[buf][kq * 4 + 1][col] = breg[i].y;
            Bs[buf][kq * 4 + 2][col] = breg[i].z;
            Bs[buf][kq * 4 + 3][col] = breg[i].w;
        }
    };

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int nk = p.K / BK;
    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) load_tiles(kt + 1);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[8], b[TN];
            *reinterpret_cast<float4 *>(&a[0]) = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 8]);
            *reinterpret_cast<float4 *>(&a[4]) = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 8 + 4]);
#pragma unroll
            for (int j = 0; j < TN; j += 4)
                *reinterpret_cast<float4 *>(&b[j]) = *reinterpret_cast<const float4 *>(&Bs[buf][k][tx * TN + j]);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) store_tiles(buf ^ 1);
        __syncthreads();
    }

    // ---- per-channel batch statistics of the RAW output (rows beyond M are exact zeros)
    if (p.stats) {
        float *red = &As[0][0][0];                      // [16][BN] sums, then [16][BN] squares (fits: 2*16*BN <= 2*16*132)
        float *red2 = &Bs[0][0][0];
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            float s = 0.f, q = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) { s += acc[i][j]; q = fmaf(acc[i][j], acc[i][j], q); }
            red[ty * BN + tx * TN + j] = s;
            red2[ty * BN + tx * TN + j] = q;
        }
        __syncthreads();
        if (tid < BN && n0 + tid < p.Cout) {
            double s = 0.0, q = 0.0;
#pragma unroll
            for (int r = 0; r < 16; ++r) { s += (double)red[r * BN + tid]; q += (double)red2[r * BN + tid]; }
            atomicAdd(p.stats + n0 + tid, s);
            atomicAdd(p.stats + p.Cout + n0 + tid, q);
        }
    }

    // ---- epilogue
    TOut *out = reinterpret_cast<TOut *>(p.out);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long m = m0 + ty * 8 + i;
        if (m >= p.M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + tx * TN + j;
            if (n >= p.Cout) continue;
            float v = acc[i][j];
            if (p.bias) v += __ldg(p.bias + n);
            v *= p.alpha;
            if (p.act == 1) v = fmaxf(v, 0.f);
            else if (p.act == 2) v = 0.5f * v * (1.f + erff(v * 0.70710678118654752f));
            if (p.residual) v += __ldg(p.residual + m * p.Cout + n);
            ActIO<TOut>::st(out + m * p.Cout + n, v);
        }
    }
}

template <typename TIn, typename TOut>
cudaError_t run_igemm(const IGemmParams &p, cudaStream_t s) {
    if (p.K % BK != 0 || p.Cin % (sizeof(TIn) == 4 ? 4 : 8) != 0 || p.Cin % BK != 0) return cudaErrorInvalidValue;
    if (p.M <= 0) return cudaSuccess;
    if (p.Cout <= 64) {
        dim3 grid(ceil_div(p.M, BM), ceil_div(p.Cout, 64));
        igemm_simt_kernel<TIn, TOut, 64><<<grid, IG_THREADS, 0, s>>>(p);
    } else {
        dim3 grid(ceil_div(p.M, BM), ceil_div(p.Cout, 128));
        igemm_simt_kernel<TIn, TOut, 128><<<grid, IG_THREADS, 0, s>>>(p);
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// stem: 7x7 stride-2 conv straight from the uint8 BGR patch (normalisation LUT, BGR->RGB and the layout
// change of network.py:470-478, 397-398 are folded into the load), raw output + statistics
// ------------------------------------------------------------------------------------------------
constexpr int STEM_ROWS = 4;                               // output rows per CTA (x 64 output columns)
constexpr int STEM_IN_ROWS = STEM_ROWS * 2 + 5;            // 13
constexpr int STEM_IN_COLS = 134;                          // 2*63 + 7 = 133, padded
constexpr int STEM_K = 147;
constexpr int STEM_IN_FLOATS = (STEM_IN_ROWS * STEM_IN_COLS * 3 + 3) / 4 * 4;   // keeps the weight tile 16-byte aligned
constexpr int STEM_SMEM = (STEM_IN_FLOATS + STEM_K * 64 + 2 * 8 * 64) * 4;

template <typename TOut>
__global__ void __launch_bounds__(256) stem_kernel(const uint8_t *__restrict__ bank, const int32_t *__restrict__ slots,
                                                   const float *__restrict__ lut, const float *__restrict__ w /*[147][64]*/,
                                                   TOut *__restrict__ out, double *__restrict__ stats) {
    extern __shared__ __align__(16) float smem[];
    float *sin = smem;                                              // [13][134][3]
    float *sw = smem + STEM_IN_FLOATS;                              // [147][64]
    float *sred = sw + STEM_K * 64;                                 // [2][8][64]
    __shared__ float slut[256 * 3];

    const int tid = threadIdx.x;
    const int n = blockIdx.y, oy0 = blockIdx.x * STEM_ROWS;
    const int slot = slots[n];
    for (int i = tid; i < 768; i += 256) slut[i] = lut[i];
    for (int i = tid; i < STEM_K * 64 / 4; i += 256) reinterpret_cast<float4 *>(sw)[i] = __ldg(reinterpret_cast<const float4 *>(w) + i);
    __syncthreads();
    const uint8_t *patch = slot >= 0 ? bank + (size_t)slot * PATCH_BYTES : nullptr;
    const int iy0 = oy0 * 2 - 3;
    for (int i = tid; i < STEM_IN_ROWS * STEM_IN_COLS * 3; i += 256) {
        const int c = i % 3, t = i / 3;
        const int col = t % STEM_IN_COLS, row = t / STEM_IN_COLS;
        const int iy = iy0 + row, ix = col - 3;
        float v = 0.f;                                              // conv zero padding (normalised domain)
        if (iy >= 0 && iy < PATCH_H && ix >= 0 && ix < PATCH_W) {
            const int u = patch ? (int)__ldg(patch + ((size_t)iy * PATCH_W + ix) * 3 + c) : 0;
            v = slut[u * 3 + c];
        }
        sin[i] = v;
    }
    __syncthreads();

    const int prow = tid >> 6, pcol = tid & 63;
    float acc[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) acc[j] = 0.f;
    for (int ky = 0; ky < 7; ++ky) {
        const float *irow = sin + ((prow * 2 + ky) * STEM_IN_COLS + pcol * 2) * 3;
        const float *wrow = sw + ky * 21 * 64;
#pragma unroll 3
        for (int t = 0; t < 21; ++t) {                              // (kx, c) flattened: contiguous in both arrays
            const float v = irow[t];
            const float4 *w4 = reinterpret_cast<const float4 *>(wrow + t * 64);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float4 ww = w4[j];
                acc[4 * j + 0] = fmaf(v, ww.x, acc[4 * j + 0]);
                acc[4 * j + 1] = fmaf(v, ww.y, acc[4 * j + 1]);
                acc[4 * j + 2] = fmaf(v, ww.z, acc[4 * j + 2]);
                acc[4 * j + 3] = fmaf(v, ww.w, acc[4 * j + 3]);
            }
        }
    }
    // raw output, NHWC [N,192,64,64]
    TOut *o = out + (((size_t)n * 192 + oy0 + prow) * 64 + pcol) * 64;
#pragma unroll
    for (int j = 0; j < 64; ++j) ActIO<TOut>::st(o + j, acc[j]);
    // statistics
    const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
    for (int j = 0; j < 64; ++j) {
        float s = warp_sum(acc[j]);
        float q = warp_sum(acc[j] * acc[j]);
        if (lane == 0) { sred[warp * 64 + j] = s; sred[512 + warp * 64 + j] = q; }
    }
    __syncthreads();
    if (tid < 128) {
        const int j = tid & 63, which = tid >> 6;
        double s = 0.0;
#pragma unroll
        for (int wv = 0; wv < 8; ++wv) s += (double)sred[which * 512 + wv * 64 + j];
        atomicAdd(stats + which * 64 + j, s);
    }
}

__global__ void bn_finalize_kernel(const double *__restrict__ stats, const float *__restrict__ gamma,
                                   const float *__restrict__ beta, int C, double inv_count, float *__restrict__ scale,
                                   float *__restrict__ shift) {
    pdl_trigger();
    pdl_wait();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double mean = stats[c] * inv_count;
    double var = stats[C + c] * inv_count - mean * mean;           // biased variance, as F.batch_norm(training=True)
    var = var > 0.0 ? var : 0.0;
    const double a = (double)gamma[c] / sqrt(var + 1e-5);
    scale[c] = (float)a;
    shift[c] = (float)((double)beta[c] - mean * a);
}

// relu(s*x + t) = |s| * max(sgn(s)*x, -t/|s|) + t : theta / sign masks for the A-tile transform and |s| folded into the weights
__global__ void __launch_bounds__(256) bn_fold_kernel(const float *__restrict__ scale, const float *__restrict__ shift, int C,
                                                       const float4 *__restrict__ w32, uint2 *__restrict__ w16s, uint16_t *__restrict__ xf,
                                                       long long total4) {
    __shared__ float sabs[512];
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float sc = scale[c];
        if (!(fabsf(sc) >= 1e-20f)) sc = 1e-20f;                   // scale == 0: relu(t) is reproduced by a huge |theta|
        const float a = fabsf(sc);
        sabs[c] = a;
        if (blockIdx.x == 0) {
            const __nv_bfloat16 th = __float2bfloat16_rn(-shift[c] / a);
            xf[c] = *reinterpret_cast<const uint16_t *>(&th);
            xf[C + c] = sc < 0.f ? 0x8000u : 0u;
        }
    }
    __syncthreads();
    const int C4 = C / 4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4) * 4;
        const float4 w = w32[i];
        __nv_bfloat162 lo = __floats2bfloat162_rn(w.x * sabs[c], w.y * sabs[c + 1]), hi = __floats2bfloat162_rn(w.z * sabs[c + 2], w.w * sabs[c + 3]);
        w16s[i] = make_uint2(*reinterpret_cast<uint32_t *>(&lo), *reinterpret_cast<uint32_t *>(&hi));
    }
}

// bn_finalize_kernel + bn_fold_kernel in one launch: every block finalises all C (<= 512) channels of the producer redundantly
// (same fp64 arithmetic, so scale / shift are bit-equal to the two-launch path), block 0 publishes them, all blocks fold.
__global__ void __launch_bounds__(256) bn_finalize_fold_kernel(const double *__restrict__ stats, const float *__restrict__ gamma, const float *__restrict__ beta,
                                                                int C, double inv_count, float *__restrict__ scale, float *__restrict__ shift,
                                                                const float4 *__restrict__ w32, uint2 *__restrict__ w16s, uint16_t *__restrict__ xf, long long total4) {
    __shared__ float sabs[512];
    pdl_trigger();
    pdl_wait();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const double mean = stats[c] * inv_count;
        double var = stats[C + c] * inv_count - mean * mean;
        var = var > 0.0 ? var : 0.0;
        const double a64 = (double)gamma[c] / sqrt(var + 1e-5);
        float sc = (float)a64;
        const float sh = (float)((double)beta[c] - mean * a64);
        if (blockIdx.x == 0) { scale[c] = sc; shift[c] = sh; }
        if (!(fabsf(sc) >= 1e-20f)) sc = 1e-20f;
        const float a = fabsf(sc);
        sabs[c] = a;
        if (blockIdx.x == 0) {
            const __nv_bfloat16 th = __float2bfloat16_rn(-sh / a);
            xf[c] = *reinterpret_cast<const uint16_t *>(&th);
            xf[C + c] = sc < 0.f ? 0x8000u : 0u;
        }
    }
    __syncthreads();
    const int C4 = C / 4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4) * 4;
        const float4 w = w32[i];
        __nv_bfloat162 lo = __floats2bfloat162_rn(w.x * sabs[c], w.y * sabs[c + 1]), hi = __floats2bfloat162_rn(w.z * sabs[c + 2], w.w * sabs[c + 3]);
        w16s[i] = make_uint2(*reinterpret_cast<uint32_t *>(&lo), *reinterpret_cast<uint32_t *>(&hi));
    }
}

// relu(bn(x)) followed by MaxPool2d(3, stride 2, pad 1)           resnet.py:271-272
template <typename T>
__global__ void bn_relu_maxpool_kernel(const T *__restrict__ raw, T *__restrict__ out, int N, int H, int W, int C,
                                       const float *__restrict__ scale, const float *__restrict__ shift) {
    const int Ho = H / 2, Wo = W / 2, C4 = C / 4;
    const long long total = (long long)N * Ho * Wo * C4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4) * 4;
        long long t = i / C4;
        const int ox = (int)(t % Wo);
        t /= Wo;
        const int oy = (int)(t % Ho);
        const long long n = t / Ho;
        float sc[4], sh[4], best[4] = {0.f, 0.f, 0.f, 0.f};      // relu output >= 0 and the window is never empty
#pragma unroll
        for (int j = 0; j < 4; ++j) { sc[j] = scale[c + j]; sh[j] = shift[c + j]; }
        for (int dy = -1; dy <= 1; ++dy) {
            const int iy = 2 * oy + dy;
            if (iy < 0 || iy >= H) continue;
            for (int dx = -1; dx <= 1; ++dx) {
                const int ix = 2 * ox + dx;
                if (ix < 0 || ix >= W) continue;
                const T *p = raw + ((n * H + iy) * W + ix) * C + c;
#pragma unroll
                for (int j = 0; j < 4; ++j) best[j] = fmaxf(best[j], fmaf(ActIO<T>::ld(p + j), sc[j], sh[j]));
            }
        }
        T *o = out + ((n * Ho + oy) * Wo + ox) * C + c;
#pragma unroll
        for (int j = 0; j < 4; ++j) ActIO<T>::st(o + j, best[j]);
    }
}

// block output: relu(bn3(raw) + identity) with identity either an activated tensor or bn_ds(raw_ds)   resnet.py:107-128
template <typename T>
__global__ void bn_add_relu_kernel(const T *__restrict__ raw, const float *__restrict__ scale, const float *__restrict__ shift,
                                   const T *idt, const float *__restrict__ iscale, const float *__restrict__ ishift, T *out,
                                   long long rows, int C) {
    const int C4 = C / 4;
    const long long total = rows * C4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4) * 4;
        const long long off = (i / C4) * C + c;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v = fmaf(ActIO<T>::ld(raw + off + j), scale[c + j], shift[c + j]);
            float d = ActIO<T>::ld(idt + off + j);
            if (iscale) d = fmaf(d, iscale[c + j], ishift[c + j]);
            ActIO<T>::st(out + off + j, fmaxf(v + d, 0.f));
        }
    }
}

// relu(bn(x)) in place: materialises the activated tensor for the TMA-fed tensor-core convolutions
template <typename T>
__global__ void bn_relu_inplace_kernel(T *x, const float *__restrict__ scale, const float *__restrict__ shift, long long rows, int C) {
    const int C8 = C / 8;
    const long long total = rows * C8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C8) * 8;
        T *p = x + (i / C8) * C + c;
        if constexpr (sizeof(T) == 2) {
            uint4 v = *reinterpret_cast<const uint4 *>(p);
            __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&v);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float2 f = __bfloat1622float2(h[j]);
                f.x = fmaxf(fmaf(f.x, scale[c + 2 * j], shift[c + 2 * j]), 0.f);
                f.y = fmaxf(fmaf(f.y, scale[c + 2 * j + 1], shift[c + 2 * j + 1]), 0.f);
                h[j] = __floats2bfloat162_rn(f.x, f.y);
            }
            *reinterpret_cast<uint4 *>(p) = v;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) ActIO<T>::st(p + j, fmaxf(fmaf(ActIO<T>::ld(p + j), scale[c + j], shift[c + j]), 0.f));
        }
    }
}

// fp32 -> bf16 copy (A operand of the tensor-core linears)
__global__ void __launch_bounds__(256) cast_bf16_kernel(const float4 *__restrict__ in, uint2 *__restrict__ out, long long n4) {
    pdl_trigger();
    pdl_wait();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = in[i];
        __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        out[i] = make_uint2(*reinterpret_cast<uint32_t *>(&a), *reinterpret_cast<uint32_t *>(&b));
    }
}

template <typename T>
__global__ void global_maxpool_kernel(const T *__restrict__ x, float *__restrict__ out, int N, int HW, int C) {
    pdl_trigger();
    pdl_wait();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)N * C) return;
    const int c = (int)(i % C);
    const long long n = i / C;
    float m = -CUDART_INF_F;
    for (int p = 0; p < HW; ++p) m = fmaxf(m, ActIO<T>::ld(x + (n * HW + p) * C + c));
    out[i] = m;
}

// F.normalize(x, p=2, dim=1): x / max(||x||_2, 1e-12)              resnet.py:319-322
__global__ void l2norm_rows_kernel(float *x, int rows, int cols) {
    pdl_trigger();
    pdl_wait();
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    float *p = x + (long long)row * cols;
    float s = 0.f;
    for (int c = lane; c < cols; c += 32) s = fmaf(p[c], p[c], s);
    s = warp_sum(s);
    const float inv = 1.f / fmaxf(sqrtf(s), 1e-12f);
    for (int c = lane; c < cols; c += 32) p[c] *= inv;
}

// ---- bf16 fast paths: 8 channels (16 bytes) per thread ---------------------------------------------------------
__device__ __forceinline__ void unpack8(const uint4 &v, float (&f)[8]) {
    const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 t = __bfloat1622float2(h[j]); f[2 * j] = t.x; f[2 * j + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 v;
    __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
    return v;
}
__device__ __forceinline__ void load8(const float *__restrict__ p, float (&f)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4 *>(p)), b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

__global__ void __launch_bounds__(256) bn_add_relu_bf16_kernel(const uint4 *__restrict__ raw, const float *__restrict__ scale,
                                                                const float *__restrict__ shift, const uint4 *idt, const float *__restrict__ iscale,
                                                                const float *__restrict__ ishift, uint4 *out, long long total8, int C8) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C8) * 8;
        float x[8], d[8], sc[8], sh[8];
        unpack8(raw[i], x);
        unpack8(idt[i], d);
        load8(scale + c, sc);
        load8(shift + c, sh);
        if (iscale) {
            float a[8], b[8];
            load8(iscale + c, a);
            load8(ishift + c, b);
#pragma unroll
            for (int j = 0; j < 8; ++j) d[j] = fmaf(d[j], a[j], b[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = fmaxf(fmaf(x[j], sc[j], sh[j]) + d[j], 0.f);
        out[i] = pack8(x);
    }
}

__global__ void __launch_bounds__(256) bn_relu_maxpool_bf16_kernel(const uint4 *__restrict__ raw, uint4 *__restrict__ out, int N, int H, int W, int C8,
                                                                    const float *__restrict__ scale, const float *__restrict__ shift) {
    const int Ho = H / 2, Wo = W / 2;
    const long long total = (long long)N * Ho * Wo * C8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % C8);
        long long t = i / C8;
        const int ox = (int)(t % Wo);
        t /= Wo;
        const int oy = (int)(t % Ho);
        const long long n = t / Ho;
        float sc[8], sh[8], best[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        load8(scale + c8 * 8, sc);
        load8(shift + c8 * 8, sh);
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
            const int iy = 2 * oy + dy;
            if (iy < 0 || iy >= H) continue;
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                const int ix = 2 * ox + dx;
                if (ix < 0 || ix >= W) continue;
                float x[8];
                unpack8(__ldg(raw + ((n * H + iy) * W + ix) * C8 + c8), x);
#pragma unroll
                for (int j = 0; j < 8; ++j) best[j] = fmaxf(best[j], fmaf(x[j], sc[j], sh[j]));
            }
        }
        out[i] = pack8(best);
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
cudaError_t launch_stem(const uint8_t *bank, const int32_t *slots, int N, const float *lut, const ConvLayer &L, void *out,
                        int bf16, cudaStream_t s) {
    if (N <= 0) return cudaSuccess;
    dim3 grid(192 / STEM_ROWS, N);
    cudaError_t e;
    if (bf16) {
        e = cudaFuncSetAttribute(stem_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, STEM_SMEM);
        if (e != cudaSuccess) return e;
        stem_kernel<__nv_bfloat16><<<grid, 256, STEM_SMEM, s>>>(bank, slots, lut, L.w32, (__nv_bfloat16 *)out, L.stats);
    } else {
        e = cudaFuncSetAttribute(stem_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, STEM_SMEM);
        if (e != cudaSuccess) return e;
        stem_kernel<float><<<grid, 256, STEM_SMEM, s>>>(bank, slots, lut, L.w32, (float *)out, L.stats);
    }
    return cudaGetLastError();
}

cudaError_t launch_conv_simt(const ConvLayer &L, const ConvArgs &a, int bf16, cudaStream_t s) {
    IGemmParams p{};
    p.in = a.in; p.out = a.out; p.w = L.w32;
    p.N = a.N; p.H = a.H; p.W = a.W; p.Cin = L.cin; p.Ho = a.Ho; p.Wo = a.Wo; p.Cout = L.cout;
    p.ksz = L.k; p.stride = L.stride; p.pad = L.k / 2;
    p.M = (long long)a.N * a.Ho * a.Wo; p.K = L.k * L.k * L.cin;
    p.in_scale = a.in_scale; p.in_shift = a.in_shift;
    p.bias = nullptr; p.residual = nullptr; p.alpha = 1.f; p.act = 0; p.stats = L.stats;
    return bf16 ? run_igemm<__nv_bfloat16, __nv_bfloat16>(p, s) : run_igemm<float, float>(p, s);
}

cudaError_t launch_linear_f32(const LinearArgs &a, cudaStream_t s) {
    IGemmParams p{};
    p.in = a.A; p.out = a.out; p.w = a.W;
    p.N = a.M; p.H = 1; p.W = 1; p.Cin = a.K; p.Ho = 1; p.Wo = 1; p.Cout = a.N;
    p.ksz = 1; p.stride = 1; p.pad = 0; p.M = a.M; p.K = a.K;
    p.bias = a.bias; p.residual = a.residual; p.alpha = a.alpha; p.act = a.act; p.stats = nullptr;
    return run_igemm<float, float>(p, s);
}

cudaError_t launch_bn_finalize(const ConvLayer &L, long long count, cudaStream_t s) {
    return launch_pdl(bn_finalize_kernel, dim3(ceil_div(L.cout, 128)), dim3(128), 0, s, L.stats, L.gamma, L.beta, L.cout, 1.0 / (double)count, L.scale, L.shift);
}

cudaError_t launch_bn_fold(const float *scale, const float *shift, const ConvLayer &L, cudaStream_t s) {
    if (L.cin > 512 || L.cin % 4 != 0 || !L.w32m || !L.w16s || !L.xf) return cudaErrorInvalidValue;
    const long long total4 = (long long)L.cout * L.k * L.k * L.cin / 4;
    long long blocks = (total4 + 255) / 256;
    if (blocks > 148 * 4) blocks = 148 * 4;
    bn_fold_kernel<<<(int)blocks, 256, 0, s>>>(scale, shift, L.cin, (const float4 *)L.w32m, (uint2 *)L.w16s, L.xf, total4);
    return cudaGetLastError();
}

cudaError_t launch_bn_finalize_fold(const ConvLayer &P, long long count, const ConvLayer &L, cudaStream_t s) {
    if (P.cout != L.cin || L.cin > 512 || L.cin % 4 != 0 || !L.w32m || !L.w16s || !L.xf) return cudaErrorInvalidValue;
    const long long total4 = (long long)L.cout * L.k * L.k * L.cin / 4;
    long long blocks = (total4 + 255) / 256;
    if (blocks > 148 * 4) blocks = 148 * 4;
    return launch_pdl(bn_finalize_fold_kernel, dim3((int)blocks), dim3(256), 0, s, P.stats, P.gamma, P.beta, P.cout, 1.0 / (double)count, P.scale, P.shift,
                      (const float4 *)L.w32m, (uint2 *)L.w16s, L.xf, total4);
}

// FINAL-pass weights and the second half of the Gram-matrix statistics (kernels.h: FoldFinalArgs).  A block owns GFF_CH output
// channels: quadratic forms of its channels (thread (i, part) owns row i of G W^T over the part-th slice of the j range, G read
// through its symmetry so that consecutive threads read consecutive addresses; G is read once per GFF_CH channels), the channels'
// BN scale / shift, then their weight rows times that scale (and |scale| of the input BN), rounded to bf16 ONCE from the fp32 master.
constexpr int GFF_CH = 8;
struct GffConv {
    const float *G, *m;              // Gram partials or null
    const __nv_bfloat16 *wq;         // [Cout][C] bf16 weights the statistics refer to
    double *stats;                   // [2*Cout] in: weighted-pass contribution, out: totals (null with explicit scale / shift)
    const float *gamma, *beta;
    const float *scale_in, *shift_in;   // explicit BN (stats == null)
    const float *w32;                // fp32 master [Cout][C]
    __nv_bfloat16 *w16f;
    float *scale_out;
    int C;
};
struct GffParams {
    GffConv a, d;
    int has_d, Cout;
    const float *in_scale;
    double inv_count;
    float *shift_out;
};

// fp64 arithmetic runs at ~2 TFLOP/s on this part (measured: the all-fp64 quadratic forms of a 256 -> 1024 convolution took ~60 us), so
// the O(C^2) inner sums v_i = sum_j G_ij w_j are fp32 (four partial sums per thread and channel; G itself is an fp32 accumulation) and only
// the O(C) outer sums sum_i w_i v_i, sum_i w_i m_i and everything after them are fp64.
__device__ __forceinline__ void gff_conv(const GffConv &cv, const GffParams &p, const float *in_scale, int c0, float (*sw)[GFF_CH], double (*red)[2 * GFF_CH],
                                          float *s_sc, float *s_sh) {
    const int t = threadIdx.x, C = cv.C;
    if (cv.G) {
        for (int idx = t; idx < GFF_CH * C; idx += 256) {
            const int q = idx / C, k = idx - q * C;
            sw[k][q] = __bfloat162float(cv.wq[(size_t)(c0 + q) * C + k]);
        }
        __syncthreads();
        // thread (ig, part): columns 4 ig .. 4 ig + 3 of G over the part-th slice of the j range (C in {64, 128, 256}: 4 / 8 / 16 slices of
        // 64 / 16 / 4 rows); 128-bit loads, up to 16 in flight per thread - the loop is L2-latency bound (few blocks), not FMA bound
        const int ngrp = C / 4, parts = 256 / ngrp, ig = t % ngrp, part = t / ngrp;
        const int jlen = C / parts, j0 = part * jlen;
        float v[GFF_CH][4];
#pragma unroll
        for (int q = 0; q < GFF_CH; ++q)
#pragma unroll
            for (int e = 0; e < 4; ++e) v[q][e] = 0.f;
        const float4 *G4 = reinterpret_cast<const float4 *>(cv.G) + ig;
        auto step = [&](const float4 g, int j) {
            const float4 wa = *reinterpret_cast<const float4 *>(&sw[j][0]), wb = *reinterpret_cast<const float4 *>(&sw[j][4]);
            const float w8[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
            for (int q = 0; q < GFF_CH; ++q) {
                v[q][0] = fmaf(g.x, w8[q], v[q][0]); v[q][1] = fmaf(g.y, w8[q], v[q][1]);
                v[q][2] = fmaf(g.z, w8[q], v[q][2]); v[q][3] = fmaf(g.w, w8[q], v[q][3]);
            }
        };
        if (jlen >= 16) {
            for (int jb = j0; jb < j0 + jlen; jb += 16) {
                float4 g[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) g[u] = __ldg(G4 + (size_t)(jb + u) * ngrp);
#pragma unroll
                for (int u = 0; u < 16; ++u) step(g[u], jb + u);
            }
        } else {
            float4 g[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) g[u] = __ldg(G4 + (size_t)(j0 + u) * ngrp);
#pragma unroll
            for (int u = 0; u < 4; ++u) step(g[u], j0 + u);
        }
        double r[GFF_CH], s1[GFF_CH];
#pragma unroll
        for (int q = 0; q < GFF_CH; ++q) {
            r[q] = 0.0;
            s1[q] = 0.0;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const double wi = (double)sw[4 * ig + e][q];
                r[q] = fma((double)v[q][e], wi, r[q]);
                if (part == 0) s1[q] = fma(wi, (double)cv.m[4 * ig + e], s1[q]);
            }
        }
#pragma unroll
        for (int q = 0; q < GFF_CH; ++q)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                r[q] += __shfl_xor_sync(0xffffffffu, r[q], o);
                s1[q] += __shfl_xor_sync(0xffffffffu, s1[q], o);
            }
        if ((t & 31) == 0)
#pragma unroll
            for (int q = 0; q < GFF_CH; ++q) { red[t >> 5][q] = s1[q]; red[t >> 5][GFF_CH + q] = r[q]; }
        __syncthreads();
    }
    if (t < GFF_CH) {
        const int c = c0 + t;
        float sc, sh;
        if (cv.stats) {
            double S = cv.stats[c], Q = cv.stats[p.Cout + c];
            if (cv.G) {
                for (int wv = 0; wv < 8; ++wv) { S += red[wv][t]; Q += red[wv][GFF_CH + t]; }
                cv.stats[c] = S;
                cv.stats[p.Cout + c] = Q;
            }
            const double mean = S * p.inv_count;
            double var = Q * p.inv_count - mean * mean;
            var = var > 0.0 ? var : 0.0;
            const double a = (double)cv.gamma[c] / sqrt(var + 1e-5);
            sc = (float)a;
            sh = (float)((double)cv.beta[c] - mean * a);
        } else {
            sc = cv.scale_in[c];
            sh = cv.shift_in[c];
        }
        s_sc[t] = sc;
        s_sh[t] = sh;
        if (cv.scale_out) cv.scale_out[c] = sc;
    }
    __syncthreads();
    for (int idx = t; idx < GFF_CH * C; idx += 256) {
        const int q = idx / C, k = idx - q * C;
        float a = 1.f;
        if (in_scale) {                                         // the clamp of bn_fold_kernel: the transform parameters were derived with it
            a = fabsf(in_scale[k]);
            if (!(a >= 1e-20f)) a = 1e-20f;
        }
        const size_t o = (size_t)(c0 + q) * C + k;
        cv.w16f[o] = __float2bfloat16_rn(cv.w32[o] * a * s_sc[q]);
    }
}

__global__ void __launch_bounds__(256) gram_fold_final_kernel(const GffParams p) {
    __shared__ __align__(16) float sw[256][GFF_CH];
    __shared__ double red[8][2 * GFF_CH];
    __shared__ float s_sc[2][GFF_CH], s_sh[2][GFF_CH];
    pdl_trigger();
    pdl_wait();
    const int c0 = blockIdx.x * GFF_CH;
    gff_conv(p.a, p, p.in_scale, c0, sw, red, s_sc[0], s_sh[0]);
    if (p.has_d) {
        __syncthreads();
        gff_conv(p.d, p, nullptr, c0, sw, red, s_sc[1], s_sh[1]);
    }
    __syncthreads();
    if (threadIdx.x < GFF_CH) p.shift_out[c0 + threadIdx.x] = s_sh[0][threadIdx.x] + (p.has_d ? s_sh[1][threadIdx.x] : 0.f);
}

cudaError_t launch_bn_fold_final(const FoldFinalArgs &a, cudaStream_t s) {
    const ConvLayer *L = a.L, *D = a.ds;
    if (!L || L->k != 1 || !L->w32m || !L->w16f || !a.shift_out || L->cout % GFF_CH != 0) return cudaErrorInvalidValue;
    const bool from_stats = a.count > 0;
    if (!from_stats && (!a.scale || !a.shift)) return cudaErrorInvalidValue;
    if (D && (D->k != 1 || D->cout != L->cout || !D->w32m || !D->w16f || (!from_stats && (!a.ds_scale || !a.ds_shift)))) return cudaErrorInvalidValue;
    if (a.gram_G && (!from_stats || !a.gram_m || (L->cin != 64 && L->cin != 128 && L->cin != 256))) return cudaErrorInvalidValue;
    if (a.ds_gram_G && (!D || !from_stats || !a.ds_gram_m || (D->cin != 64 && D->cin != 128 && D->cin != 256))) return cudaErrorInvalidValue;
    GffParams p{};
    p.Cout = L->cout; p.in_scale = a.in_scale; p.inv_count = from_stats ? 1.0 / (double)a.count : 0.0; p.shift_out = a.shift_out; p.has_d = D ? 1 : 0;
    p.a = GffConv{a.gram_G, a.gram_m, (const __nv_bfloat16 *)(a.in_scale ? L->w16s : L->w16), from_stats ? L->stats : nullptr, L->gamma, L->beta, a.scale, a.shift,
                  L->w32m, (__nv_bfloat16 *)L->w16f, L->scale, L->cin};
    if (D)
        p.d = GffConv{a.ds_gram_G, a.ds_gram_m, (const __nv_bfloat16 *)D->w16, from_stats ? D->stats : nullptr, D->gamma, D->beta, a.ds_scale, a.ds_shift,
                      D->w32m, (__nv_bfloat16 *)D->w16f, D->scale, D->cin};
    return launch_pdl(gram_fold_final_kernel, dim3(L->cout / GFF_CH), dim3(256), 0, s, p);
}

static int ew_grid(long long total) {
    long long b = (total + 255) / 256;
    return (int)(b < 148 * 32 ? (b > 0 ? b : 1) : 148 * 32);
}

// Default variant (BUSCA_POOL_MONO=0 / busca_set_option("pool_mono", 0) selects the tap-by-tap kernel; compared bit for bit on a B200 by
// tests/probe_pool.py): max-pooling commutes with the per-channel BN + ReLU because x -> fma(x, scale, shift) is monotone
// (rounding included), so max_taps relu(fma(x_t)) = relu(fma(max_t x_t)) for scale >= 0 and relu(fma(min_t x_t)) for
// scale < 0.  The nine taps cost one packed bf16 max and one packed min per two channels instead of unpack + fma + max per
// channel (about 110 instead of 250 instructions per 8 channels; the kernel is latency / issue bound, not HBM bound).
__global__ void __launch_bounds__(256) bn_relu_maxpool_bf16_mono_kernel(const uint4 *__restrict__ raw, uint4 *__restrict__ out, int N, int H, int W, int C8,
                                                                         const float *__restrict__ scale, const float *__restrict__ shift) {
    const int Ho = H / 2, Wo = W / 2;
    const long long total = (long long)N * Ho * Wo * C8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % C8);
        long long t = i / C8;
        const int ox = (int)(t % Wo);
        t /= Wo;
        const int oy = (int)(t % Ho);
        const long long n = t / Ho;
        uint4 vmax, vmin;
        bool first = true;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
            const int iy = 2 * oy + dy;
            if (iy < 0 || iy >= H) continue;
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                const int ix = 2 * ox + dx;
                if (ix < 0 || ix >= W) continue;
                const uint4 v = __ldg(raw + ((n * H + iy) * W + ix) * C8 + c8);
                if (first) {
                    vmax = v; vmin = v; first = false;
                } else {
                    const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&v);
                    __nv_bfloat162 *hx = reinterpret_cast<__nv_bfloat162 *>(&vmax), *hn = reinterpret_cast<__nv_bfloat162 *>(&vmin);
#pragma unroll
                    for (int j = 0; j < 4; ++j) { hx[j] = __hmax2(hx[j], h[j]); hn[j] = __hmin2(hn[j], h[j]); }
                }
            }
        }
        float sc[8], sh[8], hi[8], lo[8], best[8];
        load8(scale + c8 * 8, sc);
        load8(shift + c8 * 8, sh);
        unpack8(vmax, hi);
        unpack8(vmin, lo);
#pragma unroll
        for (int j = 0; j < 8; ++j) best[j] = fmaxf(0.f, fmaf(sc[j] < 0.f ? lo[j] : hi[j], sc[j], sh[j]));
        out[i] = pack8(best);
    }
}

// Strip variant of the kernel above for the stem's shape (Wo * C/8 == 256): one block walks R output rows of one image downwards, a
// thread owns one (ox, 8-channel chunk) column.  Input row 2oy+1 of output row oy is input row 2(oy+1)-1 of the next one, so only the
// horizontal 3-max / 3-min of TWO new input rows are computed per output row (6 loads instead of 9), and horizontally adjacent
// windows share their edge pixel inside the block (L1).  The tensor is then read from L2 about once instead of 2.25 times.
constexpr int POOL_STRIP_ROWS = 12;
__global__ void __launch_bounds__(256) bn_relu_maxpool_bf16_strip_kernel(const uint4 *__restrict__ raw, uint4 *__restrict__ out, int N, int H, int W, int C8,
                                                                          const float *__restrict__ scale, const float *__restrict__ shift, int strips) {
    pdl_trigger();
    pdl_wait();
    const int Ho = H / 2, Wo = W / 2;
    const int n = blockIdx.x / strips, oy0 = (blockIdx.x % strips) * POOL_STRIP_ROWS;
    const int c8 = threadIdx.x % C8, ox = threadIdx.x / C8;
    float sc[8], sh[8];
    load8(scale + c8 * 8, sc);
    load8(shift + c8 * 8, sh);
    const uint4 *img = raw + (size_t)n * H * W * C8;
    auto hrow = [&](int iy, uint4 &mx, uint4 &mn) {             // horizontal max / min of input row iy over ix = 2ox-1 .. 2ox+1
        const uint4 *r = img + ((size_t)iy * W + 2 * ox) * C8 + c8;
        const uint4 b = __ldg(r), c = __ldg(r + C8);
        mx = b; mn = b;
        __nv_bfloat162 *hx = reinterpret_cast<__nv_bfloat162 *>(&mx), *hn = reinterpret_cast<__nv_bfloat162 *>(&mn);
        const __nv_bfloat162 *hc = reinterpret_cast<const __nv_bfloat162 *>(&c);
#pragma unroll
        for (int j = 0; j < 4; ++j) { hx[j] = __hmax2(hx[j], hc[j]); hn[j] = __hmin2(hn[j], hc[j]); }
        if (ox > 0) {
            const uint4 a = __ldg(r - C8);
            const __nv_bfloat162 *ha = reinterpret_cast<const __nv_bfloat162 *>(&a);
#pragma unroll
            for (int j = 0; j < 4; ++j) { hx[j] = __hmax2(hx[j], ha[j]); hn[j] = __hmin2(hn[j], ha[j]); }
        }
    };
    uint4 pmx, pmn;                                             // row 2oy-1 (the previous iteration's row 2oy+1)
    bool have_prev = false;
    if (oy0 > 0) { hrow(2 * oy0 - 1, pmx, pmn); have_prev = true; }
    const int oy1 = min(Ho, oy0 + POOL_STRIP_ROWS);
    for (int oy = oy0; oy < oy1; ++oy) {
        uint4 amx, amn, bmx, bmn;
        hrow(2 * oy, amx, amn);
        hrow(2 * oy + 1, bmx, bmn);                             // 2oy+1 <= H-1 always (H even)
        uint4 vmax = amx, vmin = amn;
        __nv_bfloat162 *hx = reinterpret_cast<__nv_bfloat162 *>(&vmax), *hn = reinterpret_cast<__nv_bfloat162 *>(&vmin);
        const __nv_bfloat162 *bx = reinterpret_cast<const __nv_bfloat162 *>(&bmx), *bn = reinterpret_cast<const __nv_bfloat162 *>(&bmn);
#pragma unroll
        for (int j = 0; j < 4; ++j) { hx[j] = __hmax2(hx[j], bx[j]); hn[j] = __hmin2(hn[j], bn[j]); }
        if (have_prev) {
            const __nv_bfloat162 *px = reinterpret_cast<const __nv_bfloat162 *>(&pmx), *pn = reinterpret_cast<const __nv_bfloat162 *>(&pmn);
#pragma unroll
            for (int j = 0; j < 4; ++j) { hx[j] = __hmax2(hx[j], px[j]); hn[j] = __hmin2(hn[j], pn[j]); }
        }
        float hi[8], lo[8], best[8];
        unpack8(vmax, hi);
        unpack8(vmin, lo);
#pragma unroll
        for (int j = 0; j < 8; ++j) best[j] = fmaxf(0.f, fmaf(sc[j] < 0.f ? lo[j] : hi[j], sc[j], sh[j]));
        out[(((size_t)n * Ho + oy) * Wo + ox) * C8 + c8] = pack8(best);
        pmx = bmx; pmn = bmn; have_prev = true;
    }
}

int g_pool_mono = -1;
void reid_set_pool_mono(int on) { g_pool_mono = on ? 1 : 0; }
static bool pool_mono_enabled() {
    if (g_pool_mono < 0) {
        const char *e = getenv("BUSCA_POOL_MONO");
        g_pool_mono = !(e && e[0] == '0');      // on by default since round 2: bit-equal on a B200 (profiles/r02a_probe_pool.log)
    }
    return g_pool_mono != 0;
}

cudaError_t launch_bn_relu_maxpool(const void *raw, void *out, int N, int H, int W, int C, const float *scale,
                                   const float *shift, int bf16, cudaStream_t s) {
    long long total = (long long)N * (H / 2) * (W / 2) * (C / 4);
    if (total == 0) return cudaSuccess;
    if (bf16 && C % 8 == 0) {
        const long long t8 = (long long)N * (H / 2) * (W / 2) * (C / 8);
        if (pool_mono_enabled() && (W / 2) * (C / 8) == 256 && (H / 2) % POOL_STRIP_ROWS == 0) {
            const int strips = (H / 2) / POOL_STRIP_ROWS;
            return launch_pdl(bn_relu_maxpool_bf16_strip_kernel, dim3(N * strips), dim3(256), 0, s, (const uint4 *)raw, (uint4 *)out, N, H, W, C / 8, scale, shift, strips);
        } else if (pool_mono_enabled())
            bn_relu_maxpool_bf16_mono_kernel<<<ew_grid(t8), 256, 0, s>>>((const uint4 *)raw, (uint4 *)out, N, H, W, C / 8, scale, shift);
        else
            bn_relu_maxpool_bf16_kernel<<<ew_grid(t8), 256, 0, s>>>((const uint4 *)raw, (uint4 *)out, N, H, W, C / 8, scale, shift);
    } else if (bf16)
        bn_relu_maxpool_kernel<__nv_bfloat16><<<ew_grid(total), 256, 0, s>>>((const __nv_bfloat16 *)raw, (__nv_bfloat16 *)out, N, H, W, C, scale, shift);
    else
        bn_relu_maxpool_kernel<float><<<ew_grid(total), 256, 0, s>>>((const float *)raw, (float *)out, N, H, W, C, scale, shift);
    return cudaGetLastError();
}

cudaError_t launch_bn_add_relu(const void *raw, const float *scale, const float *shift, const void *idt,
                               const float *idt_scale, const float *idt_shift, void *out, long long rows, int C, int bf16,
                               cudaStream_t s) {
    long long total = rows * (C / 4);
    if (total == 0) return cudaSuccess;
    if (bf16 && C % 8 == 0) {
        const long long t8 = rows * (C / 8);
        bn_add_relu_bf16_kernel<<<ew_grid(t8), 256, 0, s>>>((const uint4 *)raw, scale, shift, (const uint4 *)idt, idt_scale, idt_shift, (uint4 *)out, t8, C / 8);
    } else if (bf16)
        bn_add_relu_kernel<__nv_bfloat16><<<ew_grid(total), 256, 0, s>>>((const __nv_bfloat16 *)raw, scale, shift, (const __nv_bfloat16 *)idt, idt_scale, idt_shift, (__nv_bfloat16 *)out, rows, C);
    else
        bn_add_relu_kernel<float><<<ew_grid(total), 256, 0, s>>>((const float *)raw, scale, shift, (const float *)idt, idt_scale, idt_shift, (float *)out, rows, C);
    return cudaGetLastError();
}

cudaError_t launch_bn_relu_inplace(void *x, const float *scale, const float *shift, long long rows, int C, int bf16, cudaStream_t s) {
    long long total = rows * (C / 8);
    if (total == 0) return cudaSuccess;
    if (bf16)
        bn_relu_inplace_kernel<__nv_bfloat16><<<ew_grid(total), 256, 0, s>>>((__nv_bfloat16 *)x, scale, shift, rows, C);
    else
        bn_relu_inplace_kernel<float><<<ew_grid(total), 256, 0, s>>>((float *)x, scale, shift, rows, C);
    return cudaGetLastError();
}

cudaError_t launch_cast_bf16(const float *in, void *out, long long n, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    if (n % 4 != 0) return cudaErrorInvalidValue;
    return launch_pdl(cast_bf16_kernel, dim3(ew_grid(n / 4)), dim3(256), 0, s, (const float4 *)in, (uint2 *)out, n / 4);
    return cudaGetLastError();
}

cudaError_t launch_global_maxpool(const void *x, float *out, int N, int HW, int C, int bf16, cudaStream_t s) {
    long long total = (long long)N * C;
    if (total == 0) return cudaSuccess;
    if (bf16)
        return launch_pdl(global_maxpool_kernel<__nv_bfloat16>, dim3(ceil_div(total, 256)), dim3(256), 0, s, (const __nv_bfloat16 *)x, out, N, HW, C);
    else
        global_maxpool_kernel<float><<<ceil_div(total, 256), 256, 0, s>>>((const float *)x, out, N, HW, C);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------------
// Duplicate elimination inside one BatchNorm batch.  The candidate patches of different tracks are the SAME detections
// (T*C slots drawn from D + T crops) and every incomplete history is the same zero image, so the batch the reference
// stacks (network.py:313-316, 383-386) holds many identical images.  An image's activations depend on the rest of the
// batch only through the batch statistics, so the encoder runs once per DISTINCT bank slot and every statistic is
// weighted by the slot's multiplicity: sum_i x_i over the stacked batch == sum_u w_u x_u over the distinct images.
//   slots [n] (-1 = zero image)  ->  uniq [nu] (first occurrence order), map [n] (row of uniq), weight [nu], *nu
// table [bank_slots + 1] must be all INT_MAX-like (0x7f7f7f7f) on entry and is restored on exit.  One CTA.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) dedup_slots_kernel(const int32_t *__restrict__ slots, int n, int *__restrict__ table,
                                                            int32_t *__restrict__ uniq, int32_t *__restrict__ map,
                                                            float *__restrict__ weight, int *__restrict__ n_uniq) {
    __shared__ int part[1024];
    const int t = threadIdx.x;
    const int per = (n + 1023) / 1024, lo = t * per, hi = min(n, lo + per);
#define KEY(i) (max(slots[i], -1) + 1)                         /* every negative slot is the zero image */
    for (int i = t; i < n; i += 1024) weight[i] = 0.f;
    for (int i = lo; i < hi; ++i) atomicMin(&table[KEY(i)], i);
    __syncthreads();
    int cnt = 0;
    for (int i = lo; i < hi; ++i) cnt += table[KEY(i)] == i;
    part[t] = cnt;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {              // inclusive scan of the per-thread counts
        const int v = t >= off ? part[t - off] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    int rank = part[t] - cnt;
    if (t == 1023) *n_uniq = part[1023];
    __syncthreads();                                        // everyone has read table[] == first index
    for (int i = lo; i < hi; ++i)
        if (table[KEY(i)] == i) uniq[rank++] = slots[i];
    __syncthreads();
    rank = part[t] - cnt;
    for (int i = lo; i < hi; ++i)                           // representatives publish their rank as -(rank + 1)
        if (table[KEY(i)] == i) table[KEY(i)] = -(rank++ + 1);
    __syncthreads();
    for (int i = lo; i < hi; ++i) {
        const int r = -table[KEY(i)] - 1;
        map[i] = r;
        atomicAdd(&weight[r], 1.f);
    }
    __syncthreads();
    for (int i = lo; i < hi; ++i) table[KEY(i)] = 0x7f7f7f7f;
#undef KEY
}

__global__ void gather_rows_kernel(const float4 *__restrict__ src, const int32_t *__restrict__ map, float4 *__restrict__ dst, int rows, int cols4) {
    pdl_trigger();
    pdl_wait();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)rows * cols4) return;
    const int r = (int)(i / cols4), c = (int)(i % cols4);
    dst[i] = src[(long long)map[r] * cols4 + c];
}

// Stable partition of the distinct images of a BatchNorm batch: multiplicity 1 first, repeated images last (uniq / weight / map are
// rewritten in place through the scratch arrays; one CTA).  *n_single = number of images with multiplicity 1.
__global__ void __launch_bounds__(1024) dedup_partition_kernel(int32_t *__restrict__ uniq, float *__restrict__ weight, int32_t *__restrict__ map, int n,
                                                                const int *__restrict__ n_uniq, int32_t *__restrict__ tmp_u, float *__restrict__ tmp_w,
                                                                int32_t *__restrict__ newpos, int *__restrict__ n_single) {
    __shared__ int part[1024];
    const int t = threadIdx.x, nu = *n_uniq;
    const int per = (nu + 1023) / 1024, lo = min(nu, t * per), hi = min(nu, lo + per);
    int cnt = 0;
    for (int r = lo; r < hi; ++r) cnt += weight[r] == 1.f;
    part[t] = cnt;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        const int v = t >= off ? part[t - off] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    const int ns = part[1023];
    int rs = part[t] - cnt, rm = ns + (lo - rs);               // singles before this chunk; multiples before it = lo - singles
    for (int r = lo; r < hi; ++r) {
        const bool single = weight[r] == 1.f;
        const int pos = single ? rs++ : rm++;
        newpos[r] = pos;
        tmp_u[pos] = uniq[r];
        tmp_w[pos] = weight[r];
    }
    if (t == 0) *n_single = ns;
    __syncthreads();
    for (int r = t; r < nu; r += 1024) { uniq[r] = tmp_u[r]; weight[r] = tmp_w[r]; }
    for (int i = t; i < n; i += 1024) map[i] = newpos[map[i]];
}

cudaError_t launch_dedup_partition(int32_t *uniq, float *weight, int32_t *map, int n, const int *n_uniq, int32_t *tmp_u, float *tmp_w, int32_t *newpos,
                                   int *n_single, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    dedup_partition_kernel<<<1, 1024, 0, s>>>(uniq, weight, map, n, n_uniq, tmp_u, tmp_w, newpos, n_single);
    return cudaGetLastError();
}

cudaError_t launch_dedup_slots(const int32_t *slots, int n, int *table, int32_t *uniq, int32_t *map, float *weight, int *n_uniq, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    dedup_slots_kernel<<<1, 1024, 0, s>>>(slots, n, table, uniq, map, weight, n_uniq);
    return cudaGetLastError();
}
cudaError_t launch_gather_rows(const float *src, const int32_t *map, float *dst, int rows, int cols, cudaStream_t s) {
    if (rows <= 0) return cudaSuccess;
    const long long total = (long long)rows * (cols / 4);
    return launch_pdl(gather_rows_kernel, dim3(ceil_div(total, 256)), dim3(256), 0, s, (const float4 *)src, map, (float4 *)dst, rows, cols / 4);
    return cudaGetLastError();
}

cudaError_t launch_l2norm_rows(float *x, int rows, int cols, cudaStream_t s) {
    if (rows <= 0) return cudaSuccess;
    return launch_pdl(l2norm_rows_kernel, dim3(ceil_div(rows, 8)), dim3(256), 0, s, x, rows, cols);
    return cudaGetLastError();
}
