// tcgen05 / TMEM / TMA implicit-GEMM convolution and linear kernel for sm_100a (bf16 in, fp32 accumulate).
//
//   D[M, Cout] = A[M, K] * W[Cout, K]^T,   M = N*Ho*Wo output pixels, K = taps*Cin
//
// * A is never materialised: for every filter tap the TMA engine loads a [BI images x BH rows x BW cols x 64 ch]
//   box of the NHWC activation tensor straight into a 128-byte-swizzled shared-memory tile (128 pixels x 64
//   channels, K-major) - exactly the canonical operand layout of tcgen05.mma.  Spatial zero padding is the TMA
//   out-of-bounds fill; stride-2 convolutions read one of four "parity" views of the input (a strided 4-D tensor
//   map per parity), so no gather code and no elementStrides are needed.
// * W tiles ([BN out-channels x 64] of the [Cout][K] weight matrix) arrive the same way.
// * One elected thread issues tcgen05.mma (M=128, N=BN, K=16) four times per 64-wide K block; the fp32 accumulator
//   lives in TMEM (2 x 256 columns, double buffered so the epilogue of tile i overlaps the MMAs of tile i+1).
// * Epilogue warps read the accumulator with tcgen05.ld (thread = output pixel, registers = channels), reduce the
//   per-channel sum / sum-of-squares needed by batch-statistic BatchNorm with a shuffle butterfly into a per-CTA
//   shared-memory accumulator (flushed once per CTA with fp64 atomics), and store bf16 NHWC.
// * Persistent: one CTA per SM, static round-robin over (pixel-tile, channel-tile) pairs, channel-tile fastest so
//   CTAs running concurrently share the same A boxes through L2.
//
// * The bf16 output tile is staged in shared memory (128-byte swizzle, 64 channels per group, double buffered) and
//   written with TMA stores, so HBM sees whole 128-byte lines; the per-channel statistics are reduced from the
//   staged tile (thread = channel, no shuffles).
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-9 = epilogue
// (warp % 4 = TMEM lane quarter, two warps per quarter split the 64-channel group).
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;            // bf16 elements per K block = 128 bytes = one swizzle row
constexpr int TC_THREADS = 320;
constexpr int TC_EPI_THREADS = 256;
constexpr int TC_MAX_TAPS = 9;

struct TcParams {
    int tiles_m, tiles_n, k_iters, cin_blocks, ntaps;
    int tap_map[TC_MAX_TAPS], tap_dw[TC_MAX_TAPS], tap_dh[TC_MAX_TAPS];
    int BW, BH, BI;                  // output-tile geometry, BW*BH*BI == 128, BW == Wo
    int Ho, Wo, Nimg, Cout, h_tiles; // h_tiles = Ho / BH
    void *out;                       // bf16 [M, Cout] (out_f32 == 0) or float [M, Cout]
    double *stats;                   // [2*Cout] or null
    const float *bias;               // [Cout] or null
    const float *residual;           // float [M, Cout] or null (fp32 output only)
    float alpha;
    int act;                         // 0 none, 1 relu, 2 gelu
    int out_f32;
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *smem, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(smem)),
                 "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *smem, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem)),
                 "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}
// K-major swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor), rows of KB bytes (128 or 64):
// start>>4 | LBO(1)<<16 | SBO(8 rows * KB bytes >> 4)<<32 | version 1 <<46 | layout (SWIZZLE_128B = 2, SWIZZLE_64B = 4) <<61
template <int KB>
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(8 * KB / 16) << 32) | (1ull << 46) | ((uint64_t)(KB == 128 ? 2 : 4) << 61);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

template <int BN, int KB>
struct TcCfg {
    static constexpr int BKE = KB / 2;                               // bf16 elements per K block
    static constexpr int STAGES = BN == 256 ? 3 : (BN == 128 ? 5 : 6);
    static constexpr int OUT_STAGE_BYTES = TC_BM * 128;              // one 64-channel bf16 group of the output tile
    static constexpr int A_BYTES = TC_BM * KB;                       // 16 KB (8 KB for the 64-byte stem rows)
    static constexpr int B_BYTES = BN * KB;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STATS_FLOATS = 2 * 2048;                    // per-CTA sum / sum-of-squares for up to 2048 channels
    static constexpr int SMEM = 1024 + STAGES * STAGE_BYTES + 2 * OUT_STAGE_BYTES + STATS_FLOATS * 4 + 256;
};

template <int BN, int KB>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                                                                 const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapA3,
                                                                 const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapOut,
                                                                 const TcParams p) {
    using Cfg = TcCfg<BN, KB>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);       // SWIZZLE_128B tiles need 1024 B alignment
    uint8_t *tiles = smem;
    uint8_t *ostage = smem + Cfg::STAGES * Cfg::STAGE_BYTES;                              // 2 x [128 rows][128 B], swizzled
    float *s_stats = reinterpret_cast<float *>(ostage + 2 * Cfg::OUT_STAGE_BYTES);
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_stats + Cfg::STATS_FLOATS);
    uint64_t *full = bars, *empty = bars + Cfg::STAGES, *tfull = bars + 2 * Cfg::STAGES, *tempty = tfull + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = p.tiles_m * p.tiles_n;

    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        prefetch_tmap(&mapA0); prefetch_tmap(&mapB);
    }
    if (p.stats)
        for (int i = threadIdx.x; i < 2 * p.Cout; i += TC_THREADS) s_stats[i] = 0.f;
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================================================== TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int n_tile = tile % p.tiles_n, m_tile = tile / p.tiles_n;
                const int h0 = (m_tile % p.h_tiles) * p.BH, n0 = (m_tile / p.h_tiles) * p.BI;
                for (int kt = 0; kt < p.k_iters; ++kt) {
                    const int tap = kt / p.cin_blocks, cb = kt - tap * p.cin_blocks;
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t *a_dst = tiles + stage * Cfg::STAGE_BYTES, *b_dst = a_dst + Cfg::A_BYTES;
                    mbar_expect_tx(&full[stage], Cfg::STAGE_BYTES);
                    const int m = p.tap_map[tap];
                    const CUtensorMap *mp = m == 0 ? &mapA0 : (m == 1 ? &mapA1 : (m == 2 ? &mapA2 : &mapA3));
                    tma_load_4d(a_dst, mp, &full[stage], cb * Cfg::BKE, p.tap_dw[tap], h0 + p.tap_dh[tap], n0);
                    tma_load_2d(b_dst, &mapB, &full[stage], kt * Cfg::BKE, n_tile * BN);
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        if (lane == 0) {
            // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, both K-major, N>>3, M>>4
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + acc * 256;
                for (int kt = 0; kt < p.k_iters; ++kt) {
                    mbar_wait(&full[stage], phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_addr = smem_u32(tiles + stage * Cfg::STAGE_BYTES);
                    const uint64_t da = umma_desc<KB>(a_addr), db = umma_desc<KB>(a_addr + Cfg::A_BYTES);
#pragma unroll
                    for (int k = 0; k < Cfg::BKE / 16; ++k)     // +32 bytes (2 x 16 B) per K=16 step inside the swizzle row
                        umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kt | k) != 0);
                    umma_commit(&empty[stage]);                 // frees the smem stage when these MMAs retire
                    if (kt == p.k_iters - 1) umma_commit(&tfull[acc]);
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ===================================================== epilogue (warps 2..9)
        const int e = threadIdx.x - 64;                // 0..255
        const int q = warp & 3;                        // TMEM lane quarter this warp may read
        const int half = (warp - 2) >> 2;              // which 32 of the 64 channels of a group
        const int row = q * 32 + lane;
        const int wi = row % p.BW, hi = (row / p.BW) % p.BH, ni = row / (p.BW * p.BH);
        int it = 0, gcount = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int n_tile = tile % p.tiles_n, m_tile = tile / p.tiles_n;
            const int h0 = (m_tile % p.h_tiles) * p.BH, n0 = (m_tile / p.h_tiles) * p.BI;
            mbar_wait(&tfull[acc], acc_phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (!p.out_f32) {
#pragma unroll 1
                for (int g = 0; g < BN / 64; ++g, ++gcount) {
                    uint8_t *st = ostage + (gcount & 1) * Cfg::OUT_STAGE_BYTES;
                    uint32_t r[32];
                    tmem_ld32(tmem_base + acc * 256 + g * 64 + half * 32 + ((uint32_t)(q * 32) << 16), r);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 4; ++j) {       // 4 x 16 B = this thread's 32 channels of its row, swizzled like the TMA box
                        uint32_t w[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(r[j * 8 + 2 * k]), __uint_as_float(r[j * 8 + 2 * k + 1]));
                            w[k] = *reinterpret_cast<uint32_t *>(&h2);
                        }
                        const int chunk = (half * 4 + j) ^ (row & 7);
                        *reinterpret_cast<uint4 *>(st + row * 128 + chunk * 16) = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    if (e == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // store(g-1) has released the buffer group g+1 will overwrite
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    if (e == 0) {
                        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"((uint64_t)&mapOut),
                                     "r"(smem_u32(st)), "r"(n_tile * BN + g * 64), "r"(0), "r"(h0), "r"(n0)
                                     : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    if (p.stats) {                      // thread = channel: 32 rows of one channel from the staged bf16 tile
                        const int col = e & 63, part = e >> 6;
                        float sum = 0.f, sq = 0.f;
#pragma unroll 8
                        for (int i = 0; i < 32; ++i) {
                            const int rr = part * 32 + i;
                            const __nv_bfloat16 v = *reinterpret_cast<const __nv_bfloat16 *>(st + rr * 128 + (((col >> 3) ^ (rr & 7)) << 4) + (col & 7) * 2);
                            const float x = __bfloat162float(v);
                            sum += x;
                            sq = fmaf(x, x, sq);
                        }
                        atomicAdd(&s_stats[n_tile * BN + g * 64 + col], sum);
                        atomicAdd(&s_stats[p.Cout + n_tile * BN + g * 64 + col], sq);
                    }
                }
            } else if (half == 0) {
                // fp32 output with bias / scale / activation / residual (linear layers): direct stores, 4 warps
                const int img = n0 + ni;
                const bool valid = img < p.Nimg;
                const long long m = ((long long)img * p.Ho + h0 + hi) * p.Wo + wi;
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    uint32_t r[32];
                    tmem_ld32(tmem_base + acc * 256 + c * 32 + ((uint32_t)(q * 32) << 16), r);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (valid) {
                        const int col0 = n_tile * BN + c * 32;
                        float *o = reinterpret_cast<float *>(p.out) + m * p.Cout + col0;
                        const float *res = p.residual ? p.residual + m * p.Cout + col0 : nullptr;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            float x = __uint_as_float(r[j]);
                            if (p.bias) x += __ldg(p.bias + col0 + j);
                            x *= p.alpha;
                            if (p.act == 1) x = fmaxf(x, 0.f);
                            else if (p.act == 2) x = 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
                            if (res) x += res[j];
                            o[j] = x;
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
        }
        if (e == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }

    __syncthreads();
    if (p.stats) {
        for (int i = threadIdx.x; i < 2 * p.Cout; i += TC_THREADS) {
            const float v = s_stats[i];
            if (v != 0.f) atomicAdd(p.stats + i, (double)v);
        }
    }
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---------------------------------------------------------------- host side: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 4-D bf16 map: dims {C, W, H, N} (elements), strides in elements for W, H, N; box {64, bw, bh, bi}
bool make_map4(CUtensorMap *m, const void *base, int C, int W, int H, int N, long long sW, long long sH, long long sN, int bw, int bh, int bi,
               int inner = TC_BK) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)sW * 2, (cuuint64_t)sH * 2, (cuuint64_t)sN * 2};
    cuuint32_t box[4] = {(cuuint32_t)inner, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bi};
    cuuint32_t es[4] = {1, 1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               inner == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
bool make_map2(CUtensorMap *m, const void *base, long long K, long long rows, int box_rows, int inner = TC_BK) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)inner, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               inner == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int g_num_sms = 0;

template <int BN, int KB = 128>
cudaError_t launch_tc(const CUtensorMap *maps, const CUtensorMap &mapB, const CUtensorMap &mapOut, const TcParams &p, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN, KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<BN, KB>::SMEM);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    if (!g_num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int total = p.tiles_m * p.tiles_n;
    const int grid = total < g_num_sms ? total : g_num_sms;
    conv_tc_kernel<BN, KB><<<grid, TC_THREADS, TcCfg<BN, KB>::SMEM, s>>>(maps[0], maps[1], maps[2], maps[3], mapB, mapOut, p);
    return cudaGetLastError();
}

}  // namespace

// Output-tile geometry for an Ho x Wo output: BW = Wo (<= 32... or 128 for flat rows), BH | Ho, BI = 128 / (BW*BH).
static bool tile_geometry(int Ho, int Wo, int &BW, int &BH, int &BI) {
    if (Wo > 128 || 128 % Wo != 0) return false;
    BW = Wo;
    int rows = 128 / Wo;                 // rows*BW = 128 if a single image supplies them
    BH = 1;
    for (int h = rows; h >= 1; --h)
        if (Ho % h == 0 && rows % h == 0) { BH = h; break; }
    BI = 128 / (BW * BH);
    return true;
}

cudaError_t launch_conv_tc(const ConvLayer &L, const ConvArgs &a, cudaStream_t s) {
    if (L.cin % TC_BK != 0 || L.cout % 64 != 0 || !L.w16) return cudaErrorInvalidValue;
    if (L.stride != 1 && L.stride != 2) return cudaErrorInvalidValue;
    if (L.stride == 2 && ((a.H | a.W) & 1)) return cudaErrorInvalidValue;
    TcParams p{};
    if (!tile_geometry(a.Ho, a.Wo, p.BW, p.BH, p.BI)) return cudaErrorInvalidValue;
    const int BN = L.cout >= 256 ? 256 : (L.cout >= 128 ? 128 : 64);
    p.tiles_n = L.cout / BN;
    p.h_tiles = a.Ho / p.BH;
    p.tiles_m = ((a.N + p.BI - 1) / p.BI) * p.h_tiles;
    p.cin_blocks = L.cin / TC_BK;
    p.ntaps = L.k * L.k;
    p.k_iters = p.ntaps * p.cin_blocks;
    p.Ho = a.Ho; p.Wo = a.Wo; p.Nimg = a.N; p.Cout = L.cout;
    p.out = a.out; p.stats = L.stats; p.bias = nullptr; p.residual = nullptr; p.alpha = 1.f; p.act = 0; p.out_f32 = 0;
    CUtensorMap maps[4];
    const __nv_bfloat16 *in = reinterpret_cast<const __nv_bfloat16 *>(a.in);
    const long long C = L.cin, W = a.W, H = a.H;
    bool ok = true;
    if (L.stride == 1) {
        ok = make_map4(&maps[0], in, (int)C, a.W, a.H, a.N, C, W * C, H * W * C, p.BW, p.BH, p.BI);
        maps[1] = maps[2] = maps[3] = maps[0];
        const int pad = L.k / 2;
        for (int r = 0; r < L.k; ++r)
            for (int q = 0; q < L.k; ++q) { p.tap_map[r * L.k + q] = 0; p.tap_dh[r * L.k + q] = r - pad; p.tap_dw[r * L.k + q] = q - pad; }
    } else {
        // parity views: input pixel (2*ho + r - pad, 2*wo + q - pad) = view[ph][pw] at (ho + dh, wo + dw)
        for (int ph = 0; ph < 2; ++ph)
            for (int pw = 0; pw < 2; ++pw)
                ok = ok && make_map4(&maps[ph * 2 + pw], in + (ph * W + pw) * C, (int)C, a.W / 2, a.H / 2, a.N, 2 * C, 2 * W * C, H * W * C, p.BW, p.BH, p.BI);
        const int pad = L.k / 2;
        for (int r = 0; r < L.k; ++r)
            for (int q = 0; q < L.k; ++q) {
                const int oy = r - pad, ox = q - pad;          // offset in input pixels: -1, 0, +1 (or 0 for 1x1)
                const int ph = oy & 1, pw = ox & 1;
                p.tap_map[r * L.k + q] = ph * 2 + pw;
                p.tap_dh[r * L.k + q] = (oy - ph) / 2;         // -1 -> -1, 0 -> 0, +1 -> 0
                p.tap_dw[r * L.k + q] = (ox - pw) / 2;
            }
    }
    CUtensorMap mapB, mapOut;
    ok = ok && make_map2(&mapB, L.w16, (long long)p.ntaps * L.cin, L.cout, BN);
    // output [N][Ho][Wo][Cout] bf16, stored one 64-channel group of a tile at a time; images beyond N are clipped by TMA
    ok = ok && make_map4(&mapOut, a.out, L.cout, a.Wo, a.Ho, a.N, L.cout, (long long)a.Wo * L.cout, (long long)a.Ho * a.Wo * L.cout, p.BW, p.BH, p.BI);
    if (!ok) return cudaErrorInvalidValue;
    switch (BN) {
        case 256: return launch_tc<256>(maps, mapB, mapOut, p, s);
        case 128: return launch_tc<128>(maps, mapB, mapOut, p, s);
        default: return launch_tc<64>(maps, mapB, mapOut, p, s);
    }
}

// out[M,N] (fp32) = act((A[M,K] W[N,K]^T + bias) * alpha) + residual, A and W bf16
cudaError_t launch_linear_tc(const void *A_bf16, const void *W_bf16, const LinearArgs &la, cudaStream_t s) {
    if (la.K % TC_BK != 0 || la.N % 64 != 0) return cudaErrorInvalidValue;
    TcParams p{};
    p.BW = 1; p.BH = 1; p.BI = 128;
    const int BN = la.N % 256 == 0 ? 256 : (la.N % 128 == 0 ? 128 : 64);
    p.tiles_n = la.N / BN;
    p.h_tiles = 1;
    p.tiles_m = (la.M + 127) / 128;
    p.cin_blocks = la.K / TC_BK;
    p.ntaps = 1;
    p.k_iters = p.cin_blocks;
    p.tap_map[0] = 0; p.tap_dw[0] = 0; p.tap_dh[0] = 0;
    p.Ho = 1; p.Wo = 1; p.Nimg = la.M; p.Cout = la.N;
    p.out = la.out; p.stats = nullptr; p.bias = la.bias; p.residual = la.residual; p.alpha = la.alpha; p.act = la.act; p.out_f32 = 1;
    CUtensorMap maps[4], mapB;
    bool ok = make_map4(&maps[0], A_bf16, la.K, 1, 1, la.M, la.K, la.K, la.K, 1, 1, 128);
    maps[1] = maps[2] = maps[3] = maps[0];
    ok = ok && make_map2(&mapB, W_bf16, la.K, la.N, BN);
    if (!ok) return cudaErrorInvalidValue;
    switch (BN) {
        case 256: return launch_tc<256>(maps, mapB, maps[0], p, s);      // no TMA store on the fp32-output path
        case 128: return launch_tc<128>(maps, mapB, maps[0], p, s);
        default: return launch_tc<64>(maps, mapB, maps[0], p, s);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Stem on tensor cores.  The 7x7 stride-2 convolution over 3 channels becomes a GEMM with K = 7 tap-rows x 32:
// a pre-pass writes the normalised patch as bf16 with 4 channels per pixel (B,G,R,0) and 3+5 zero pixels of horizontal
// padding per row; then, for one tap-row ky, the 7x4 (+4 zero) window of output pixel ox is 32 CONTIGUOUS elements
// starting 16 bytes after the window of ox-1.  An overlapping-stride tensor map {32 el, 64 ox (stride 16 B), rows of
// one parity (stride 2 rows), N} therefore delivers the im2col tile directly; vertical padding is TMA zero fill.
// ---------------------------------------------------------------------------------------------------------------------
namespace {
constexpr int STEM_PITCH_PX = 136;                       // 3 zero px + 128 px + 5 zero px
__global__ void __launch_bounds__(256) stem_prepass_kernel(const uint8_t *__restrict__ bank, const int32_t *__restrict__ slots,
                                                           const float *__restrict__ lut, uint2 *__restrict__ out, long long total_px) {
    __shared__ float slut[768];
    for (int i = threadIdx.x; i < 768; i += 256) slut[i] = lut[i];
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_px; i += (long long)gridDim.x * blockDim.x) {
        const int px = (int)(i % STEM_PITCH_PX);
        const long long t = i / STEM_PITCH_PX;
        const int y = (int)(t % PATCH_H);
        const int n = (int)(t / PATCH_H);
        uint2 v = make_uint2(0u, 0u);
        const int x = px - 3;
        if (x >= 0 && x < PATCH_W) {
            const int slot = slots[n];
            int b = 0, g = 0, r = 0;
            if (slot >= 0) {
                const uint8_t *p = bank + (size_t)slot * PATCH_BYTES + ((size_t)y * PATCH_W + x) * 3;
                b = p[0]; g = p[1]; r = p[2];
            }
            __nv_bfloat162 lo = __floats2bfloat162_rn(slut[b * 3 + 0], slut[g * 3 + 1]);
            __nv_bfloat162 hi = __floats2bfloat162_rn(slut[r * 3 + 2], 0.f);
            v.x = *reinterpret_cast<uint32_t *>(&lo);
            v.y = *reinterpret_cast<uint32_t *>(&hi);
        }
        out[i] = v;
    }
}
}  // namespace

size_t stem_tc_scratch_bytes(int N) { return (size_t)N * PATCH_H * STEM_PITCH_PX * 4 * 2; }

// wstem: bf16 [64][7*32], element (ky, kx*4 + c_bgr) ; scratch: stem_tc_scratch_bytes(N); out: bf16 [N,192,64,64]
cudaError_t launch_stem_tc(const uint8_t *bank, const int32_t *slots, int N, const float *lut, const void *wstem, void *scratch, void *out,
                           double *stats, cudaStream_t s) {
    if (N <= 0) return cudaSuccess;
    const long long total_px = (long long)N * PATCH_H * STEM_PITCH_PX;
    long long blocks = (total_px + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    stem_prepass_kernel<<<(int)blocks, 256, 0, s>>>(bank, slots, lut, (uint2 *)scratch, total_px);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    TcParams p{};
    p.BW = 64; p.BH = 2; p.BI = 1;
    p.tiles_n = 1; p.h_tiles = 192 / 2; p.tiles_m = N * p.h_tiles;
    p.cin_blocks = 1; p.ntaps = 7; p.k_iters = 7;
    for (int ky = 0; ky < 7; ++ky) {
        const int oy = ky - 3, ph = oy & 1;
        p.tap_map[ky] = ph; p.tap_dh[ky] = (oy - ph) / 2; p.tap_dw[ky] = 0;
    }
    p.Ho = 192; p.Wo = 64; p.Nimg = N; p.Cout = 64;
    p.out = out; p.stats = stats; p.alpha = 1.f;
    const __nv_bfloat16 *in = reinterpret_cast<const __nv_bfloat16 *>(scratch);
    const long long pitch = (long long)STEM_PITCH_PX * 4;             // elements per padded row
    CUtensorMap maps[4], mapB, mapOut;
    bool ok = true;
    for (int ph = 0; ph < 2; ++ph)      // dims {32 window elements, 64 ox (stride 8 el = 16 B), 192 rows of this parity, N}
        ok = ok && make_map4(&maps[ph], in + ph * pitch, 32, 64, 192, N, 8, 2 * pitch, (long long)PATCH_H * pitch, 64, 2, 1, 32);
    maps[2] = maps[3] = maps[0];
    ok = ok && make_map2(&mapB, wstem, 7 * 32, 64, 64, 32);
    ok = ok && make_map4(&mapOut, out, 64, 64, 192, N, 64, 64 * 64, 192LL * 64 * 64, 64, 2, 1);
    if (!ok) return cudaErrorInvalidValue;
    return launch_tc<64, 64>(maps, mapB, mapOut, p, s);
}
