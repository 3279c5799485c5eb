// Decision-Transformer linears on the tensor cores with their neighbours fused into the epilogue (bf16 mode).
//
//   out = epilogue(A[M, K] * W[N, K]^T),   A and W bf16 (K-major), fp32 accumulation in TMEM
//
// Reference: busca/custom_layers.py:30-41 (post-norm encoder layer): x = LN1(x + out_proj(attention(in_proj(x)))),
// x = LN2(x + linear2(act(linear1(x)))).  Per layer the tensor-core path is five launches instead of thirteen:
//   LE_F32   in_proj:   fp32 QKV = A W^T + b                              (the attention kernel reads fp32 and writes bf16)
//   LE_LN    out_proj:  x = LayerNorm(x + A W^T + b) -> fp32 x and its bf16 copy (the next GEMM's A operand)
//   LE_BF16  linear1:   bf16 hidden = act(A W^T + b)
//   LE_LN    linear2:   x = LayerNorm(x + hidden W^T + b)
// The LayerNorm epilogue needs whole rows: a CTA then owns all 512 output columns of its 128 rows (two N = 256 accumulators side
// by side in TMEM), writes y = acc + bias + residual BACK into TMEM (tcgen05.st) and makes the mean / variance / normalise passes
// over TMEM - every row is thread-local (TMEM lane = row), the two warps of a lane quarter own one 256-column half each and exchange
// their partial sums through shared memory.  Rounding points are those of the unfused path (bf16 A operands, fp32 everything else).
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-9 = epilogue.
// Ring: 4 stages of [128 x 64] A + [256 x 64] W (48 KB); a k-iteration of an LE_LN tile is (K block, column half).
#include <cuda.h>

#include <cstdio>

#include "common.cuh"
#include "kernels.h"
#include "tc_ptx.cuh"
#include "tmap.cuh"

namespace {

constexpr int LT_THREADS = 320;
constexpr int LT_STAGES = 4;
constexpr int LT_A_BYTES = 128 * 128, LT_B_BYTES = 256 * 128, LT_STAGE_BYTES = LT_A_BYTES + LT_B_BYTES;
constexpr int LT_SMEM = 1024 + LT_STAGES * LT_STAGE_BYTES + 128 * 2 * 2 * 4 + (2 * LT_STAGES + 4) * 8 + 64;
static_assert(LT_SMEM <= 232448, "shared memory budget");

struct LinParams {
    int M, N, K, tiles_m, tiles_n, k_blocks;
    const float *bias;               // [N] or null
    float alpha;                     // LE_F32 / LE_BF16: (acc + bias) * alpha
    int act;                         // 0 none, 1 relu, 2 gelu(erf)
    const float *residual;           // [M, N] fp32 or null
    const float *gamma, *beta;       // LE_LN
    float *out_f32;                  // [M, N] or null
    __nv_bfloat16 *out_bf16;         // [M, N] or null
};

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]),
        "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]),
        "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}

__device__ __forceinline__ float act_apply(float x, int act) {
    if (act == 1) return fmaxf(x, 0.f);
    if (act == 2) return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
    return x;
}

template <int EPI>
__global__ void __launch_bounds__(LT_THREADS, 1) linear_fused_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                                      const LinParams p) {
    constexpr bool LN = EPI == LE_LN;
    constexpr int HALVES = LN ? 2 : 1;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *tiles = smem;
    float *s_red = reinterpret_cast<float *>(tiles + LT_STAGES * LT_STAGE_BYTES);      // [2 passes][2 halves][128 rows]
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_red + 128 * 2 * 2);
    uint64_t *full = bars, *empty = full + LT_STAGES, *tfull = empty + LT_STAGES, *tempty = tfull + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = p.tiles_m * p.tiles_n;
    const int k_iters = p.k_blocks * HALVES;

    pdl_trigger();
    if (threadIdx.x == 0) {
        for (int s = 0; s < LT_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        prefetch_tmap(&mapA); prefetch_tmap(&mapB);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    pdl_wait();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================================================== TMA producer
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int m0 = (tile / p.tiles_n) * 128, n0 = (tile % p.tiles_n) * 256;
            for (int it = 0; it < k_iters; ++it) {
                const int kb = it / HALVES, half = it % HALVES;
                mbar_wait<32>(&empty[stage], phase ^ 1);
                if (elect_one()) {
                    uint8_t *a_dst = tiles + stage * LT_STAGE_BYTES;
                    mbar_expect_tx(&full[stage], LT_STAGE_BYTES);
                    tma_load_2d(a_dst, &mapA, &full[stage], kb * 64, m0);
                    tma_load_2d(a_dst + LT_A_BYTES, &mapB, &full[stage], kb * 64, n0 + half * 256);
                }
                __syncwarp();
                if (++stage == LT_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        int stage = 0, tcount = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
            const int acc = LN ? 0 : (tcount & 1);
            const uint32_t acc_phase = LN ? (uint32_t)(tcount & 1) : (uint32_t)((tcount >> 1) & 1);
            mbar_wait<32>(&tempty[acc], acc_phase ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int it = 0; it < k_iters; ++it) {
                const int kb = it / HALVES, half = it % HALVES;
                mbar_wait<0>(&full[stage], phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_addr = smem_u32(tiles + stage * LT_STAGE_BYTES);
                const uint64_t da = umma_desc<128>(a_addr), db = umma_desc<128>(a_addr + LT_A_BYTES);
                const uint32_t d_tmem = tmem_base + (LN ? half * 256 : acc * 256);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, !(kb == 0 && k == 0));
                    umma_commit(&empty[stage]);
                    if (it == k_iters - 1) umma_commit(&tfull[acc]);
                }
                __syncwarp();
                if (++stage == LT_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===================================================== epilogue: TMEM lane = row; two warps per lane quarter
        const int q = warp & 3, half = (warp - 2) >> 2;
        const int r_local = q * 32 + lane;
        int tcount = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
            const int acc = LN ? 0 : (tcount & 1);
            const uint32_t acc_phase = LN ? (uint32_t)(tcount & 1) : (uint32_t)((tcount >> 1) & 1);
            const int m0 = (tile / p.tiles_n) * 128, n0 = (tile % p.tiles_n) * 256;
            const int row = m0 + r_local;
            const bool valid = row < p.M;
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
            mbar_wait<0>(&tfull[acc], acc_phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (!LN) {
                // this warp: columns [half * 128, half * 128 + 128) of the 256-column tile
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    const int cl = half * 128 + c * 32, col0 = n0 + cl;
                    uint32_t r[32];
                    tmem_ld32(t_row + acc * 256 + cl, r);
                    TMEM_LD_WAIT();
                    if (valid) {
                        const size_t o = (size_t)row * p.N + col0;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            float x[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                float v = __uint_as_float(r[j + e]);
                                if (p.bias) v += __ldg(p.bias + col0 + j + e);
                                v = act_apply(v * p.alpha, p.act);
                                x[e] = v;
                            }
                            if (p.residual) {
                                const float4 rs = *reinterpret_cast<const float4 *>(p.residual + o + j);
                                x[0] += rs.x; x[1] += rs.y; x[2] += rs.z; x[3] += rs.w;
                            }
                            if (EPI == LE_F32) {
                                *reinterpret_cast<float4 *>(p.out_f32 + o + j) = make_float4(x[0], x[1], x[2], x[3]);
                            } else {
                                *reinterpret_cast<uint2 *>(p.out_bf16 + o + j) = make_uint2(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]));
                            }
                        }
                    }
                }
            } else {
                // this warp: columns [half * 256, half * 256 + 256) of the 512-column row
                const uint32_t t_half = t_row + half * 256;
                const size_t o_row = (size_t)row * p.N + half * 256;
                float s = 0.f;
#pragma unroll 1
                for (int c = 0; c < 8; ++c) {                       // pass 1: y = acc + bias + residual -> TMEM, row sum
                    uint32_t r[32];
                    tmem_ld32(t_half + c * 32, r);
                    TMEM_LD_WAIT();
                    const int col0 = half * 256 + c * 32;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 rs = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (valid && p.residual) rs = *reinterpret_cast<const float4 *>(p.residual + o_row + c * 32 + j);
                        const float4 bs = p.bias ? __ldg(reinterpret_cast<const float4 *>(p.bias + col0 + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        const float y0 = __uint_as_float(r[j]) + bs.x + rs.x, y1 = __uint_as_float(r[j + 1]) + bs.y + rs.y;
                        const float y2 = __uint_as_float(r[j + 2]) + bs.z + rs.z, y3 = __uint_as_float(r[j + 3]) + bs.w + rs.w;
                        s += (y0 + y1) + (y2 + y3);
                        r[j] = __float_as_uint(y0); r[j + 1] = __float_as_uint(y1); r[j + 2] = __float_as_uint(y2); r[j + 3] = __float_as_uint(y3);
                    }
                    tmem_st32(t_half + c * 32, r);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                s_red[half * 128 + r_local] = s;
                asm volatile("bar.sync 1, 256;" ::: "memory");
                const float mean = (s_red[r_local] + s_red[128 + r_local]) * (1.f / 512.f);
                float qv = 0.f;
#pragma unroll 1
                for (int c = 0; c < 8; ++c) {                       // pass 2: sum of squared deviations (biased variance, as nn.LayerNorm)
                    uint32_t r[32];
                    tmem_ld32(t_half + c * 32, r);
                    TMEM_LD_WAIT();
#pragma unroll
                    for (int j = 0; j < 32; ++j) { const float d = __uint_as_float(r[j]) - mean; qv = fmaf(d, d, qv); }
                }
                s_red[256 + half * 128 + r_local] = qv;
                asm volatile("bar.sync 1, 256;" ::: "memory");
                const float rstd = rsqrtf((s_red[256 + r_local] + s_red[256 + 128 + r_local]) * (1.f / 512.f) + 1e-5f);
#pragma unroll 1
                for (int c = 0; c < 8; ++c) {                       // pass 3: normalise, fp32 row and its bf16 copy
                    uint32_t r[32];
                    tmem_ld32(t_half + c * 32, r);
                    TMEM_LD_WAIT();
                    if (valid) {
                        const int col0 = half * 256 + c * 32;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 g = __ldg(reinterpret_cast<const float4 *>(p.gamma + col0 + j)), b = __ldg(reinterpret_cast<const float4 *>(p.beta + col0 + j));
                            const float x0 = (__uint_as_float(r[j]) - mean) * rstd * g.x + b.x, x1 = (__uint_as_float(r[j + 1]) - mean) * rstd * g.y + b.y;
                            const float x2 = (__uint_as_float(r[j + 2]) - mean) * rstd * g.z + b.z, x3 = (__uint_as_float(r[j + 3]) - mean) * rstd * g.w + b.w;
                            if (p.out_f32) *reinterpret_cast<float4 *>(p.out_f32 + o_row + c * 32 + j) = make_float4(x0, x1, x2, x3);
                            if (p.out_bf16) *reinterpret_cast<uint2 *>(p.out_bf16 + o_row + c * 32 + j) = make_uint2(pack_bf16(x0, x1), pack_bf16(x2, x3));
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
        }
    }

    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

int g_lt_sms = 0;

template <int EPI>
cudaError_t launch_lt(const CUtensorMap &ma, const CUtensorMap &mb, const LinParams &p, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(linear_fused_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, LT_SMEM);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    if (!g_lt_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_lt_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int total = p.tiles_m * p.tiles_n;
    const int grid = total < g_lt_sms ? total : g_lt_sms;
    return launch_pdl(linear_fused_kernel<EPI>, dim3(grid), dim3(LT_THREADS), (size_t)LT_SMEM, s, ma, mb, p);
}

}  // namespace

cudaError_t launch_linear_fused(const void *A_bf16, const void *W_bf16, const LinearFusedArgs &a, cudaStream_t s) {
    if (a.M <= 0) return cudaSuccess;
    if (a.K % 64 != 0 || a.N % 256 != 0 || !A_bf16 || !W_bf16) return cudaErrorInvalidValue;
    if (a.epilogue == LE_LN && (a.N != 512 || !a.gamma || !a.beta || (!a.out_f32 && !a.out_bf16))) return cudaErrorInvalidValue;
    if (a.epilogue == LE_F32 && !a.out_f32) return cudaErrorInvalidValue;
    if (a.epilogue == LE_BF16 && !a.out_bf16) return cudaErrorInvalidValue;
    LinParams p{};
    p.M = a.M; p.N = a.N; p.K = a.K;
    p.tiles_m = (a.M + 127) / 128;
    p.tiles_n = a.epilogue == LE_LN ? 1 : a.N / 256;
    p.k_blocks = a.K / 64;
    p.bias = a.bias; p.alpha = a.alpha; p.act = a.act; p.residual = a.residual; p.gamma = a.gamma; p.beta = a.beta;
    p.out_f32 = a.out_f32; p.out_bf16 = reinterpret_cast<__nv_bfloat16 *>(a.out_bf16);
    CUtensorMap ma, mb;
    if (!make_map2(&ma, A_bf16, a.K, a.M, 128) || !make_map2(&mb, W_bf16, a.K, a.N, 256)) return cudaErrorInvalidValue;
    switch (a.epilogue) {
        case LE_F32: return launch_lt<LE_F32>(ma, mb, p, s);
        case LE_BF16: return launch_lt<LE_BF16>(ma, mb, p, s);
        case LE_LN: return launch_lt<LE_LN>(ma, mb, p, s);
        default: return cudaErrorInvalidValue;
    }
}
