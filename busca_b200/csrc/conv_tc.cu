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
// * Batch-statistic BatchNorm makes every conv output wait for a grid-wide reduction, so the BN + ReLU of the
//   PRODUCING convolution is applied here, to the A tile, in shared memory ("transform" warps: TMA -> smem ->
//   transform -> tcgen05.mma), as ONE exact packed-bf16 instruction per two elements:
//       relu(s*x + t) = |s| * max(sgn(s)*x, theta) + t,      theta = -t/|s|
//   The transform writes a = max(x ^ signmask, theta) (max.bf16x2: no rounding at all), |s| is folded into this
//   conv's weights once per call (bn_fold: bf16(W*|s|), a single rounding), and the "+ t" term only adds a per-
//   output-channel constant sum(W*t) to the raw output - which the batch-statistic BN that follows EVERY conv
//   removes exactly.  Spatial padding (and rows of images beyond N) must read theta, not 0, for that constant to be
//   uniform: |s|*theta + t = 0 is the zero padding of the reference.  The raw conv output therefore makes exactly
//   one HBM round trip, no separate BN-apply pass exists and the activation is never re-rounded to bf16.
// * One elected thread issues tcgen05.mma (M=128, N=BN, K=16) four times per 64-wide K block; the fp32 accumulator
//   lives in TMEM (2 x 256 columns, double buffered so the epilogue of tile i overlaps the MMAs of tile i+1).
// * Epilogue modes:
//     RAW    tcgen05.ld -> bf16 -> 128B-swizzled staging tile -> TMA store; per-channel sum / sum of squares of the
//            (bf16-rounded) output are read back from the staging tile with 64-bit shared loads into per-thread
//            register accumulators that live across tiles (no shuffles, no atomics in the loop).
//     STATS  the same without the store: the statistics-only first pass of a bottleneck's last 1x1 convolution.
//     FINAL  second pass of that convolution, with the scale of ITS OWN BatchNorm folded into the weights as well (bn_fold_final,
//            reid.cu): out = relu(acc + t3 + identity), the identity tile TMA-loaded into the staging buffer the result is then
//            stored from; with DUAL the downsample 1x1 conv (weights times ITS BN scale) accumulates into the SAME TMEM tile
//            and the shift is t3 + t_ds.  The raw output of conv3 / downsample is therefore NEVER written to HBM.
//     F32    fp32 output with bias / scale / activation / residual (Transformer linears).
// * Persistent: one CTA per SM, static round-robin over (pixel-tile, channel-tile) pairs, channel-tile fastest so
//   CTAs running concurrently share the same A boxes through L2.
//
// Warp roles (608 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-9 = A-tile
// transform (two groups of four warps take alternate k-iterations, so the fixed wait / proxy-fence / arrive latency
// of one stage overlaps the other group's work), warps 10-17 = epilogue (warp % 4 = TMEM lane quarter, two warps per
// quarter split a 64-channel group), warp 18 = identity-tile loader (FINAL mode).
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "kernels.h"
#include "tc_ptx.cuh"
#include "tmap.cuh"

namespace {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;            // bf16 elements per K block = 128 bytes = one swizzle row
constexpr int TC_THREADS = 608;
constexpr int TC_XF_WARP0 = 2;       // transform warps 2..9: two groups of four, alternating k-iterations
constexpr int TC_EPI_WARP0 = 10;     // epilogue warps 10..17
constexpr int TC_IDT_WARP = 18;
constexpr int TC_MAX_TAPS = 9;
constexpr int TC_XBUFS = 3;          // staging tiles (128 rows x 128 B)
constexpr int TC_XBUFS_MAX = 5;      // ... of the resident-weight FINAL launches (identity tiles in flight: see XB in conv_tc_kernel)

enum { MODE_RAW = 0, MODE_STATS = 1, MODE_FINAL = 2, MODE_F32 = 3 };

struct TcParams {
    int tiles_m, tiles_n, k_iters, cin_blocks, ntaps;
    int k1_iters;                    // k-iterations of the main operand (== k_iters unless DUAL)
    int tap_map[TC_MAX_TAPS], tap_dw[TC_MAX_TAPS], tap_dh[TC_MAX_TAPS];
    int tap_bit[TC_MAX_TAPS];        // (dh+1)*3 + (dw+1): bit of the per-row 3x3 in-bounds mask
    int BW, BH, BI;                  // output-tile geometry, BW*BH*BI == 128, BW == Wo
    int Ho, Wo, Nimg, Cout, h_tiles; // h_tiles = Ho / BH
    int Hv, Wv;                      // extent of the (parity view of the) input the taps index: padding mask of the transform
    int mode;
    const uint16_t *a_xf;            // [2*Cin] theta (bf16) then sign masks: transform of the A tile in shared memory, or null
    const float *e_shift;            // FINAL: shift of this conv's BN (+ shift of the downsample BN) [Cout]; the scales are in the weights
    void *out;                       // F32: float [M, Cout]
    double *stats;                   // [2*Cout] or null
    const float *bias;               // F32: [Cout] or null
    const float *residual;           // F32: float [M, Cout] or null
    float alpha;
    int act;                         // F32: 0 none, 1 relu, 2 gelu
    const float *img_w;              // [Nimg] multiplicity of every image in the batch statistics, or null (= 1): see dedup_slots_kernel
    int img_shift;                   // log2(BW*BH): tile row >> img_shift = image of the row within the tile
};

// Coordinates of the tiles a persistent CTA visits (tile = blockIdx.x + j * gridDim.x; tile = m_tile * tiles_n + n_tile;
// m_tile = image-group * h_tiles + row-group), stepped without the three integer divisions per tile.
// Tile sequence of a CTA PAIR (MC kernels): the two CTAs of a cluster take the pixel tiles 2 mp and 2 mp + 1 of the same channel tile in
// lockstep (the weight tile is multicast to both).  Plain divisions: these kernels have >= 4 k-iterations per tile.
struct PairWalk {
    int tile, n_tile, hi, ni;        // tile = pair-tile index; hi / ni = row group / image group of THIS CTA's pixel tile
    int step, cr;
    __device__ __forceinline__ PairWalk(const TcParams &p, int rank) : step((int)gridDim.x >> 1), cr(rank) {
        tile = (int)blockIdx.x >> 1;
        set(p);
    }
    __device__ __forceinline__ void set(const TcParams &p) {
        n_tile = tile % p.tiles_n;
        const int m_tile = 2 * (tile / p.tiles_n) + cr;
        hi = m_tile % p.h_tiles;
        ni = m_tile / p.h_tiles;
    }
    __device__ __forceinline__ void next(const TcParams &p) { tile += step; set(p); }
};

struct TileWalk {
    int tile, n_tile, hi, ni;
    int rn, qh, qi;
    __device__ __forceinline__ explicit TileWalk(const TcParams &p) {
        tile = blockIdx.x;
        n_tile = tile % p.tiles_n;
        const int m_tile = tile / p.tiles_n;
        hi = m_tile % p.h_tiles;
        ni = m_tile / p.h_tiles;
        const int qn = (int)gridDim.x / p.tiles_n;
        rn = (int)gridDim.x % p.tiles_n;
        qh = qn % p.h_tiles;
        qi = qn / p.h_tiles;
    }
    __device__ __forceinline__ void next(const TcParams &p) {
        tile += gridDim.x;
        n_tile += rn;
        int e = 0;
        if (n_tile >= p.tiles_n) { n_tile -= p.tiles_n; e = 1; }
        hi += qh + e;                                   // < 2 * h_tiles: one carry is enough
        ni += qi;
        if (hi >= p.h_tiles) { hi -= p.h_tiles; ++ni; }
    }
};

template <bool MC>
__device__ __forceinline__ typename std::conditional<MC, PairWalk, TileWalk>::type make_walk(const TcParams &p, int cr) {
    if constexpr (MC) return PairWalk(p, cr);
    else return TileWalk(p);
}

// RESB: the CTA's whole weight slab [BN x K] stays resident in shared memory (loaded once per launch; the host keeps a
// CTA on one channel tile), so the ring only streams A boxes.  L2 -> SM delivery (~6300 B/clk chip-wide) is the bound of
// every small-K convolution: with BN = 256 the weight tile is 2/3 of the bytes a k-iteration pulls.
template <int BN, int KB, bool DUAL, bool RESB>
struct TcCfg {
    static constexpr int BKE = KB / 2;                               // bf16 elements per K block
    static constexpr int STAGES = RESB ? 4 : (BN == 256 ? 3 : (BN == 128 ? 4 : 6));
    static constexpr int XBUF_BYTES = TC_BM * 128;                   // one 64-channel bf16 group of an output / identity tile
    static constexpr int A_BYTES = TC_BM * KB;                       // 16 KB (8 KB for the 64-byte stem rows)
    static constexpr int B_BYTES = BN * KB;
    static constexpr int STAGE_BYTES = RESB ? A_BYTES : A_BYTES + B_BYTES;
    static constexpr int RES_BYTES = RESB ? 65536 : 0;              // resident weight slab: k_iters * B_BYTES must fit
    static constexpr int PAR_FLOATS = 2 * 2048;                      // statistics (RAW/STATS: 2 Cout) or the epilogue shift (FINAL: Cout)
    static constexpr int APAR_FLOATS = 512;                          // transform parameters: theta bf16 [512], sign mask u16 [512]
    static constexpr int SMEM = 1024 + STAGES * STAGE_BYTES + RES_BYTES + TC_XBUFS * XBUF_BYTES + (PAR_FLOATS + APAR_FLOATS) * 4 + 512;
    static_assert(SMEM <= 232448, "shared memory budget (227 KB per CTA)");
};

// MC: launched as clusters of two CTAs that walk pixel tiles 2 mp / 2 mp + 1 of the same channel tile in lockstep; each CTA loads HALF
// of every weight tile and multicasts it to both (measured, profiles/r03i: the streamed-weight kernels move 48 KB per k-iteration and
// SM against ~39 B/clk of L2 delivery - 1220 cycles where the MMAs need 770 - and two thirds of those bytes are the weight tile every
// CTA fetches for itself).  A shared-memory stage is then free when BOTH CTAs' MMAs have retired (commit multicast to both).
// S4 (statistics-only passes, BN = 256, streamed weights): the three staging tiles - unused in that mode - are a FOURTH ring stage.
template <int BN, int KB, bool DUAL, bool RESB, bool MC, bool S4 = false, bool CG2 = false, bool RAW4 = false>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                                                                 const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapA3,
                                                                 const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapB2,
                                                                 const __grid_constant__ CUtensorMap mapOut, const __grid_constant__ CUtensorMap mapIdt,
                                                                 const TcParams p) {
    using Cfg = TcCfg<BN, KB, DUAL, RESB>;
    // CG2 (always together with MC's lockstep pair walk): ONE tcgen05.mma.cta_group::2 per K step covers both pixel tiles of the pair
    // (M = 256); each CTA keeps its own A tile and only ITS HALF of the weight tile (32 KB per stage instead of 48 KB: four stages, a
    // third fewer bytes into every SM per k-iteration, and the tensor core fetches half of B from each SM's shared memory).  The leader
    // CTA issues; the peer's MMA warp relays "my stage is full / transformed" and "my accumulator is drained" to the leader's barriers.
    static_assert(!CG2 || (MC && !RESB && !S4 && BN == 256), "cta_group::2 variant: paired, streamed weights, BN = 256");
    // RAW4 (RAW epilogue, BN = 256, streamed weights, long k-loops: the 3x3 convolutions of the tap loop): these launches are bound by the
    // round trip of a ring stage (load -> transform -> MMA -> commit: ~2 us over three 48 KB stages), not by the MMAs (0.4 us per
    // k-iteration).  A FOURTH stage in place of two of the three staging tiles: the epilogue then packs, stores and waits on ONE tile (a
    // fraction of a microsecond per 36-k-iteration tile), the statistics block shrinks to the 2 x 512 channels RAW launches can have.
    // CG2 + RAW4: the same trade for the pair kernel - SIX 32 KB stages.
    // BN = 128 (32 KB stages): two more stages, six in all.
    // DUAL (FINAL epilogue with the downsample conv in the same accumulator: no identity loader, the staging tiles only rotate between the
    // stores): the same trade, with the 2048-float shift block FINAL needs.
    static_assert(!RAW4 || (!RESB && MC == CG2 && !S4 && (!DUAL || (BN == 256 && !CG2)) && (BN == 256 || (BN == 128 && !CG2))), "deeper-ring variant");
    constexpr int STAGES = CG2 ? (RAW4 ? 6 : 4) : (RAW4 ? Cfg::STAGES + (BN == 128 ? 2 : 1) : (S4 ? Cfg::STAGES + 1 : Cfg::STAGES));
    // Resident weights, FINAL with an identity tensor (layers 1-2: at the HBM roofline, 64 KB of identity and 64 KB of output per 16-32 KB
    // of A): FIVE staging tiles, i.e. five identity tiles in flight per SM instead of three (the 32 KB were unused).
    constexpr int XB = RAW4 ? 1 : ((RESB && !DUAL && BN == 256) ? TC_XBUFS_MAX : TC_XBUFS);   // staging tiles of the epilogue
    constexpr int PAR_F = RAW4 ? (DUAL ? 2048 : 1024) : ((RESB && !DUAL && BN == 256) ? 2048 : Cfg::PAR_FLOATS);   // five staging tiles: a smaller statistics / shift block
    constexpr int STAGE_BYTES = CG2 ? Cfg::A_BYTES + Cfg::B_BYTES / 2 : Cfg::STAGE_BYTES;
    static_assert(!S4 || (!RESB && Cfg::STAGE_BYTES == TC_XBUFS * Cfg::XBUF_BYTES), "the extra stage aliases the staging tiles");
    constexpr int G = BN / 64;                                                            // 64-channel groups per tile
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);       // SWIZZLE_128B tiles need 1024 B alignment
    uint8_t *tiles = smem;
    uint8_t *resb = smem + Cfg::STAGES * Cfg::STAGE_BYTES;                                // RESB: k_iters x [BN rows][KB bytes], swizzled
    uint8_t *xbuf = RAW4 ? smem + STAGES * STAGE_BYTES : resb + Cfg::RES_BYTES;     // XB x [128 rows][128 B], swizzled
    float *s_par = reinterpret_cast<float *>(xbuf + XB * Cfg::XBUF_BYTES);
    float *s_apar = s_par + PAR_F;
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_apar + Cfg::APAR_FLOATS);
    uint64_t *full = bars, *empty = full + STAGES, *ready = empty + STAGES, *tfull = ready + STAGES, *tempty = tfull + 2;
    uint64_t *xfull = tempty + 2, *xfree = xfull + TC_XBUFS_MAX, *bfull = xfree + TC_XBUFS_MAX;
    uint64_t *pfull = bfull + 1, *ptempty = pfull + STAGES;          // CG2, leader CTA: the peer's stage is ready / the peer's accumulator is drained
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(ptempty + 2);
    uint32_t *xcnt = tmem_slot + 1;                                  // CG2, peer CTA: transform threads done with a stage (the 128th reports to the leader)

    static_assert(!(MC && RESB), "the paired kernel streams its weights");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cr = MC ? (int)cluster_ctarank() : 0;                  // rank in the CTA pair
    const int tile0 = MC ? (int)blockIdx.x >> 1 : (int)blockIdx.x, tile_step = MC ? (int)gridDim.x >> 1 : (int)gridDim.x;
    const int total_tiles = MC ? ((p.tiles_m + 1) >> 1) * p.tiles_n : p.tiles_m * p.tiles_n;      // MC: pair tiles
    const bool xform = p.a_xf != nullptr;
    const bool want_stats = p.stats != nullptr && (p.mode == MODE_RAW || p.mode == MODE_STATS);

    pdl_trigger();
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], (MC && !CG2) ? 2 : 1); mbar_init(&ready[s], CG2 ? 129 : ((STAGES % 2) == 0 ? 128 : 256)); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 8); }
        for (int b = 0; b < TC_XBUFS_MAX; ++b) { mbar_init(&xfull[b], 1); mbar_init(&xfree[b], 1); }
        mbar_init(bfull, 1);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&pfull[s], 1); xcnt[s] = 0; }
        for (int a = 0; a < 2; ++a) mbar_init(&ptempty[a], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        prefetch_tmap(&mapA0); prefetch_tmap(&mapB);
    }
    if (want_stats)
        for (int i = threadIdx.x; i < 2 * p.Cout; i += TC_THREADS) s_par[i] = 0.f;
    if (warp == 1) {
        if (CG2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    pdl_wait();              // everything above is private to this CTA; from here on the predecessor's results are read
    if (p.mode == MODE_FINAL)
        for (int i = threadIdx.x; i < p.Cout; i += TC_THREADS) s_par[i] = p.e_shift[i];
    if (xform) {
        const int cin = p.cin_blocks * Cfg::BKE;
        uint16_t *sp = reinterpret_cast<uint16_t *>(s_apar);
        for (int i = threadIdx.x; i < cin; i += TC_THREADS) { sp[i] = p.a_xf[i]; sp[512 + i] = p.a_xf[cin + i]; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (MC) cluster_sync_all();          // the peer's barriers are initialised before anything is multicast to them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================================================== TMA producer (the warp walks the loops, one elected lane issues)
        {
            int stage = 0;
            uint32_t phase = 0;
            if (RESB) {
                // the weight slab of this CTA's channel tile (gridDim.x % tiles_n == 0: every tile of the CTA has the same n_tile)
                const int n_tile = blockIdx.x % p.tiles_n;
                if (elect_one()) {
                    mbar_expect_tx(bfull, (uint32_t)p.k_iters * Cfg::B_BYTES);
                    for (int kt = 0; kt < p.k_iters; ++kt) {
                        if (!DUAL || kt < p.k1_iters) tma_load_2d(resb + kt * Cfg::B_BYTES, &mapB, bfull, kt * Cfg::BKE, n_tile * BN);
                        else tma_load_2d(resb + kt * Cfg::B_BYTES, &mapB2, bfull, (kt - p.k1_iters) * Cfg::BKE, n_tile * BN);
                    }
                }
                __syncwarp();
            }
            for (int tile = tile0; tile < total_tiles; tile += tile_step) {
                const int n_tile = tile % p.tiles_n, m_tile = MC ? 2 * (tile / p.tiles_n) + cr : tile / p.tiles_n;
                const int h0 = (m_tile % p.h_tiles) * p.BH, n0 = (m_tile / p.h_tiles) * p.BI;
                int tap = 0, cb = 0;
                for (int kt = 0; kt < p.k_iters; ++kt) {
                    mbar_wait<32>(&empty[stage], phase ^ 1);      // MC: both CTAs of the pair have retired the MMAs that read this stage
                    uint8_t *a_dst = tiles + stage * STAGE_BYTES, *b_dst = a_dst + Cfg::A_BYTES;
                    const bool main_op = !DUAL || kt < p.k1_iters;
                    if (CG2 && !xform) {
                        // no transform role in the way: both CTAs' loads complete on the LEADER's barrier (twice the bytes), no relay hop
                        if (elect_one()) {
                            if (cr == 0) mbar_expect_tx(&full[stage], 2 * STAGE_BYTES);
                            const uint32_t lf = cluster_addr_of(&full[stage], 0);
                            if (main_op) {
                                const int m = p.tap_map[tap];
                                const CUtensorMap *mp = m == 0 ? &mapA0 : (m == 1 ? &mapA1 : (m == 2 ? &mapA2 : &mapA3));
                                tma_load_4d_cg2(a_dst, mp, lf, cb * Cfg::BKE, p.tap_dw[tap], h0 + p.tap_dh[tap], n0);
                                tma_load_2d_cg2(b_dst, &mapB, lf, kt * Cfg::BKE, n_tile * BN + cr * (BN / 2));
                            } else {
                                const int cb2 = kt - p.k1_iters;
                                tma_load_4d_cg2(a_dst, &mapA1, lf, cb2 * Cfg::BKE, 0, h0, n0);
                                tma_load_2d_cg2(b_dst, &mapB2, lf, cb2 * Cfg::BKE, n_tile * BN + cr * (BN / 2));
                            }
                        }
                    } else if (elect_one()) {
                        mbar_expect_tx(&full[stage], STAGE_BYTES);
                        if (main_op) {
                            const int m = p.tap_map[tap];
                            const CUtensorMap *mp = m == 0 ? &mapA0 : (m == 1 ? &mapA1 : (m == 2 ? &mapA2 : &mapA3));
                            tma_load_4d(a_dst, mp, &full[stage], cb * Cfg::BKE, p.tap_dw[tap], h0 + p.tap_dh[tap], n0);
                            if (CG2) tma_load_2d(b_dst, &mapB, &full[stage], kt * Cfg::BKE, n_tile * BN + cr * (BN / 2));     // this CTA's half only
                            else if (MC) tma_load_2d_mc(b_dst + cr * (Cfg::B_BYTES / 2), &mapB, &full[stage], kt * Cfg::BKE, n_tile * BN + cr * (BN / 2), (uint16_t)3);
                            else if (!RESB) tma_load_2d(b_dst, &mapB, &full[stage], kt * Cfg::BKE, n_tile * BN);
                        } else {                                     // downsample branch: 1x1 (strided view) on the block input
                            const int cb2 = kt - p.k1_iters;
                            tma_load_4d(a_dst, &mapA1, &full[stage], cb2 * Cfg::BKE, 0, h0, n0);
                            if (CG2) tma_load_2d(b_dst, &mapB2, &full[stage], cb2 * Cfg::BKE, n_tile * BN + cr * (BN / 2));
                            else if (MC) tma_load_2d_mc(b_dst + cr * (Cfg::B_BYTES / 2), &mapB2, &full[stage], cb2 * Cfg::BKE, n_tile * BN + cr * (BN / 2), (uint16_t)3);
                            else if (!RESB) tma_load_2d(b_dst, &mapB2, &full[stage], cb2 * Cfg::BKE, n_tile * BN);
                        }
                    }
                    __syncwarp();
                    if (main_op) {
                        if (++cb == p.cin_blocks) { cb = 0; ++tap; }
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1 && CG2 && cr == 1) {
        // ===================================================== CG2, peer CTA: no MMA issue - relay this CTA's pipeline state to the leader
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int tile = tile0; tile < total_tiles; tile += tile_step, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            mbar_wait<32>(&tempty[acc], acc_phase ^ 1);                 // this CTA's epilogue has drained its half of the accumulator
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            if (lane == 0) mbar_arrive_remote(&ptempty[acc], 0);
            __syncwarp();
            // (the stages report to the leader directly: the loads complete on its full barrier when there is no transform, else this
            // CTA's transform threads arrive on its ready barrier)
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer (the warp walks the loops, one elected lane issues: r02d)
        {
            // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, both K-major, N>>3, M>>4
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            if (RESB) mbar_wait<32>(bfull, 0);                   // the resident weight slab has landed
            for (int tile = tile0; tile < total_tiles; tile += tile_step, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait<32>(&tempty[acc], acc_phase ^ 1);
                if (CG2) mbar_wait_cluster(&ptempty[acc], acc_phase);       // ... and the peer's half (its relay arrives once per tile)
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int kt = 0; kt < p.k_iters; ++kt) {
                    mbar_wait<0>(&full[stage], phase);
                    if (xform) {                                        // the transform warps have rewritten the A tile (CG2: of both CTAs)
                        if (CG2) mbar_wait_cluster(&ready[stage], phase);
                        else mbar_wait<0>(&ready[stage], phase);
                    }
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t d_tmem = tmem_base + acc * 256;        // DUAL: the downsample conv accumulates into the same tile
                    const bool first = kt == 0;
                    const uint32_t a_addr = smem_u32(tiles + stage * STAGE_BYTES);
                    const uint64_t da = umma_desc<KB>(a_addr);
                    const uint64_t db = umma_desc<KB>(RESB ? smem_u32(resb) + (uint32_t)kt * Cfg::B_BYTES : a_addr + Cfg::A_BYTES);
                    if (elect_one()) {
                        if (BN == 256 && !DUAL && p.mode == MODE_STATS) {
                            // statistics-only pass: D^T = W * A^T (operands swapped: M = 128 output channels, N = 128 pixels, two channel
                            // halves), so that TMEM lanes are CHANNELS and the per-channel sums over pixels are thread-local in the epilogue
                            const uint32_t idesc_t = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_BM >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
#pragma unroll
                            for (int h = 0; h < 2; ++h)
#pragma unroll
                                for (int k = 0; k < Cfg::BKE / 16; ++k)
                                    umma_bf16(tmem_base + acc * 256 + h * 128, db + (uint64_t)(h * (128 * KB / 16)) + 2 * k, da + 2 * k, idesc_t, !(first && k == 0));
                        } else if (CG2) {
                            // M = 256 over the pair, N = 256: A descriptor = this stage's pixel tile in EACH CTA, B descriptor = each CTA's weight half
                            const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
#pragma unroll
                            for (int k = 0; k < Cfg::BKE / 16; ++k)
                                umma_bf16_cg2(d_tmem, da + 2 * k, db + 2 * k, idesc2, !(first && k == 0));
                        } else {
#pragma unroll
                            for (int k = 0; k < Cfg::BKE / 16; ++k)     // +32 bytes (2 x 16 B) per K=16 step inside the swizzle row
                                umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, !(first && k == 0));
                        }
                        if (CG2) umma_commit_cg2(&empty[stage], (uint16_t)3);  // both CTAs' stages are free when the pair's MMAs retire
                        else if (MC) umma_commit_mc(&empty[stage], (uint16_t)3);    // the stage holds halves written by both CTAs: it is free when both have read it
                        else umma_commit(&empty[stage]);                        // frees the smem stage when these MMAs retire
                        if (kt == p.k_iters - 1) {
                            if (CG2) umma_commit_cg2(&tfull[acc], (uint16_t)3);     // each CTA's epilogue reads its own 128 accumulator rows
                            else umma_commit(&tfull[acc]);
                        }
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp < TC_EPI_WARP0) {
        // ===================================================== A-tile transform: a = max(x ^ signmask, theta); theta in the padding
        // This role is a serial chain per stage (barrier wait -> 8 x 16 B per thread -> proxy fence -> arrive) and, for the
        // convolutions with few k-iterations per tile, THE critical path of the kernel (profiles/r01f): nothing that can be
        // hoisted out of the per-tile / per-k-iteration path is computed in it.
        if (xform) {
            // With an even number of stages a group owns the stages of its own parity: every phase of those barriers is its own,
            // so it never has to look at the other group's k-iterations.  With an odd number (3 stages of 48 KB, BN = 256) a
            // group meets a stage on every second phase only; it then OBSERVES the other group's k-iterations too (see below).
            constexpr bool EVEN = (STAGES % 2) == 0;
            const int grp = (warp - TC_XF_WARP0) >> 2;        // this group handles k-iterations with (global index & 1) == grp
            const int tt = (threadIdx.x - TC_XF_WARP0 * 32) & 127;
            const int c = tt & 7, rb = tt >> 3;               // logical 16-byte chunk (8 channels), first row
            const uint32_t col_off = (uint32_t)((c ^ (rb & 7)) << 4);
            int r_hi[8], r_ni[8];
            uint32_t r_wb[8];                                 // bit (dw+1): column r_wi+dw is inside the (view of the) input
            uint32_t mi9[8], alli = 0x1ffu;                   // masks of a tile that lies inside the image vertically, all images < N
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = rb + 16 * i;
                const int wi = r % p.BW;
                r_hi[i] = (r / p.BW) % p.BH;
                r_ni[i] = r / (p.BW * p.BH);
                r_wb[i] = ((unsigned)(wi - 1) < (unsigned)p.Wv ? 1u : 0u) | ((unsigned)wi < (unsigned)p.Wv ? 2u : 0u) | ((unsigned)(wi + 1) < (unsigned)p.Wv ? 4u : 0u);
                mi9[i] = r_wb[i] | (r_wb[i] << 3) | (r_wb[i] << 6);
                alli &= mi9[i];
            }
            const uint32_t par0 = smem_u32(s_apar) + (uint32_t)c * 16;
            const uint32_t tiles0 = smem_u32(tiles) + col_off + (uint32_t)rb * 128;
            uint32_t ki = 0;                                  // global k-iteration counter (stage ring position) at the start of the tile
            typename std::conditional<MC, PairWalk, TileWalk>::type tw = make_walk<MC>(p, cr);
            for (; tw.tile < total_tiles; tw.next(p), ki += (uint32_t)p.k_iters) {
                const int kt0 = EVEN ? (int)((ki ^ (uint32_t)grp) & 1u) : 0;    // first k-iteration of this tile the group owns
                if (EVEN && kt0 >= p.k_iters) continue;                         // single-k-iteration tile of the other group
                const int h0 = tw.hi * p.BH, n0 = tw.ni * p.BI;
                // per row: bit (dh+1)*3+(dw+1) = the tap reads inside the input; bit 9 = row of an image beyond N (stays all-zero, so
                // its output is exactly 0 and drops out of the batch statistics, as with the TMA zero fill of the untransformed path)
                uint32_t m9[8], all9;
                if (n0 + p.BI <= p.Nimg && (p.ntaps == 1 || (h0 >= 1 && h0 + p.BH < p.Hv))) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) m9[i] = mi9[i];
                    all9 = alli;
                } else {
                    all9 = 0x1ffu;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int hh = h0 + r_hi[i];
                        uint32_t m = 0x200u;
                        if (n0 + r_ni[i] < p.Nimg) {
                            m = 0;
                            if ((unsigned)(hh - 1) < (unsigned)p.Hv) m |= r_wb[i];
                            if ((unsigned)hh < (unsigned)p.Hv) m |= r_wb[i] << 3;
                            if ((unsigned)(hh + 1) < (unsigned)p.Hv) m |= r_wb[i] << 6;
                        }
                        m9[i] = m;
                        all9 &= m;
                    }
                }
                int tap = 0, cb = kt0;
                while (cb >= p.cin_blocks) { cb -= p.cin_blocks; ++tap; }
                for (int kt = kt0; kt < p.k_iters; kt += EVEN ? 2 : 1) {
                    const bool main_op = !DUAL || kt < p.k1_iters;
                    const uint32_t kig = ki + (uint32_t)kt;
                    const uint32_t stage = kig % STAGES, phase = (kig / STAGES) & 1;
                    // odd number of stages: BOTH groups observe every phase of every stage (a parity wait can only tell phases
                    // apart that are at most one apart) and BOTH arrive on ready[stage] (256 arrivals per phase): the observing
                    // group is then part of the MMA's dependency chain, so it can never fall two phases of a stage behind the ring
                    // (a late observer could otherwise miss a phase and wait for one that needs its own next arrival: deadlock).
                    if (!EVEN && (int)(kig & 1) != grp) {
                        mbar_wait<0>(&full[stage], phase);
                        mbar_arrive(&ready[stage]);
                    } else {
                        if (main_op) {
                            const int bit = p.tap_bit[tap];
                            const uint4 th = lds128(par0 + (uint32_t)cb * 128), sg = lds128(par0 + (uint32_t)cb * 128 + 1024);
                            mbar_wait<0>(&full[stage], phase);
                            const uint32_t base = tiles0 + stage * STAGE_BYTES;
                            if ((all9 >> bit) & 1) {
                                uint4 v[8];
#pragma unroll
                                for (int i = 0; i < 8; ++i) v[i] = lds128(base + (uint32_t)i * 2048);
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    v[i].x = max_bf16x2(v[i].x ^ sg.x, th.x); v[i].y = max_bf16x2(v[i].y ^ sg.y, th.y);
                                    v[i].z = max_bf16x2(v[i].z ^ sg.z, th.z); v[i].w = max_bf16x2(v[i].w ^ sg.w, th.w);
                                    sts128(base + (uint32_t)i * 2048, v[i]);
                                }
                            } else {
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    const uint32_t addr = base + (uint32_t)i * 2048;
                                    uint4 v = lds128(addr);
                                    if ((m9[i] >> bit) & 1) {
                                        v.x = max_bf16x2(v.x ^ sg.x, th.x); v.y = max_bf16x2(v.y ^ sg.y, th.y);
                                        v.z = max_bf16x2(v.z ^ sg.z, th.z); v.w = max_bf16x2(v.w ^ sg.w, th.w);
                                    } else {
                                        v = (m9[i] & 0x200u) ? make_uint4(0u, 0u, 0u, 0u) : th;
                                    }
                                    sts128(addr, v);
                                }
                            }
                            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the MMA (async proxy)
                        } else {
                            mbar_wait<0>(&full[stage], phase);
                        }
                        if (CG2 && cr == 1) {
                            // the peer's transform reports to the LEADER (the only MMA issuer) with ONE remote arrival per stage: its 128 threads
                            // count themselves in shared memory (acq_rel: the last one has observed everybody's writes and proxy fences)
                            uint32_t old;
                            asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(&xcnt[stage])) : "memory");
                            if (old == 127u) {
                                xcnt[stage] = 0;                 // next use of this stage's counter is ordered behind the stage's empty -> full cycle
                                mbar_arrive_remote(&ready[stage], 0);
                            }
                        } else {
                            mbar_arrive(&ready[stage]);     // every k-iteration, so the barrier phase tracks the stage ring (CG2 leader: 128 + 1 arrivals)
                        }
                    }
                    if (main_op) {
                        cb += EVEN ? 2 : 1;
                        while (cb >= p.cin_blocks) { cb -= p.cin_blocks; ++tap; }
                    }
                }
            }
        }
    } else if (warp < TC_IDT_WARP) {
        // ===================================================== epilogue (warps 6..13)
        const int e = threadIdx.x - TC_EPI_WARP0 * 32;   // 0..255
        const int q = warp & 3;                          // TMEM lane quarter this warp may read
        const int half = (warp - TC_EPI_WARP0) >> 2;     // which 32 of the 64 channels of a group
        const int row = q * 32 + lane;
        const uint32_t xb0 = smem_u32(xbuf);
        const uint32_t row_off = (uint32_t)row * 128;
        // statistics mapping: 4 channels (8 bytes) x 8 rows per thread and group
        const int cq = e & 15, rsub = e >> 4;
        const uint32_t st_off = (uint32_t)rsub * 128 + (uint32_t)((((cq >> 1) ^ (rsub & 7)) << 4) + (cq & 1) * 8);
        float acc_s[G][4], acc_q[G][4];
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int j = 0; j < 4; ++j) { acc_s[g][j] = 0.f; acc_q[g][j] = 0.f; }
        int stats_ntile = -1;
        auto flush_stats = [&]() {
            if (stats_ntile < 0) return;
#pragma unroll
            for (int g = 0; g < G; ++g)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int ch = stats_ntile * BN + g * 64 + cq * 4 + j;
                    atomicAdd(&s_par[ch], acc_s[g][j]);
                    atomicAdd(&s_par[p.Cout + ch], acc_q[g][j]);
                    acc_s[g][j] = 0.f;
                    acc_q[g][j] = 0.f;
                }
        };
        float ts[2] = {0.f, 0.f}, tq[2] = {0.f, 0.f};   // transposed statistics pass: channel = n_tile*256 + h*128 + q*32 + lane
        auto flush_stats_t = [&]() {
            if (stats_ntile < 0) return;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int ch = stats_ntile * BN + h * 128 + q * 32 + lane;
                atomicAdd(&s_par[ch], ts[h]);
                atomicAdd(&s_par[p.Cout + ch], tq[h]);
                ts[h] = 0.f;
                tq[h] = 0.f;
            }
        };
        const bool stats_t = BN == 256 && !DUAL && p.mode == MODE_STATS;
        const bool weighted = want_stats && p.img_w != nullptr;
        int it = 0, gcount = 0;
        typename std::conditional<MC, PairWalk, TileWalk>::type tw = make_walk<MC>(p, cr);
        for (; tw.tile < total_tiles; tw.next(p), ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int n_tile = tw.n_tile;
            const int h0 = tw.hi * p.BH, n0 = tw.ni * p.BI;
            const uint32_t t_acc = tmem_base + acc * 256 + ((uint32_t)(q * 32) << 16);
            if (BN == 256 && !DUAL && p.mode == MODE_STATS) {
                // transposed accumulator (see the MMA issuer): lane = channel, columns = pixels; this warp sums 64 of the 128 pixels
                if (n_tile != stats_ntile) { flush_stats_t(); stats_ntile = n_tile; }
                // multiplicity of the image each 16-column run of this warp's 64 pixels belongs to (an image is >= 16 pixels)
                float wc[4] = {1.f, 1.f, 1.f, 1.f};
                if (weighted) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) wc[u] = __ldg(p.img_w + min(n0 + ((half * 64 + u * 16) >> p.img_shift), p.Nimg - 1));
                }
                mbar_wait<0>(&tfull[acc], acc_phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int c2 = 0; c2 < 2; ++c2) {
                        uint32_t r[32];
                        tmem_ld32(t_acc + h * 128 + half * 64 + c2 * 32, r);
                        TMEM_LD_WAIT();
                        // two 16-column runs, each inside one image (constant multiplicity): two partial sums per run keep the
                        // dependency chains short, the multiplicity is applied once per run
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            float ps0 = 0.f, ps1 = 0.f, pq0 = 0.f, pq1 = 0.f;
#pragma unroll
                            for (int j = 0; j < 16; j += 2) {
                                const float x0 = __uint_as_float(r[u * 16 + j]), x1 = __uint_as_float(r[u * 16 + j + 1]);
                                ps0 += x0; pq0 = fmaf(x0, x0, pq0);
                                ps1 += x1; pq1 = fmaf(x1, x1, pq1);
                            }
                            ts[h] = fmaf(ps0 + ps1, wc[c2 * 2 + u], ts[h]);
                            tq[h] = fmaf(pq0 + pq1, wc[c2 * 2 + u], tq[h]);
                        }
                    }
            } else if (p.mode == MODE_RAW || p.mode == MODE_STATS) {
                if (want_stats && n_tile != stats_ntile) { flush_stats(); stats_ntile = n_tile; }
                float wr[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};      // multiplicity of the image of each of this thread's 8 statistics rows
                if (weighted) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) wr[i] = __ldg(p.img_w + min(n0 + ((rsub + 16 * i) >> p.img_shift), p.Nimg - 1));
                }
                mbar_wait<0>(&tfull[acc], acc_phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int g = 0; g < G; ++g, ++gcount) {
                    // three staging tiles in rotation: the store of group g-1 may still be reading its tile while group g is packed and
                    // handed to the TMA engine (wait_group.read 1 only retires the stores up to g-2, whose tile group g+1 overwrites)
                    const int xb = gcount % XB;
                    const uint32_t st = xb0 + xb * Cfg::XBUF_BYTES;
                    uint32_t r[32];
                    tmem_ld32(t_acc + g * 64 + half * 32, r);
                    TMEM_LD_WAIT();
                    if (RAW4) {                                     // one staging tile: the previous store must have read it before it is packed again
                        if (p.mode == MODE_RAW && e == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        EPI_BAR();
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {       // 4 x 16 B = this thread's 32 channels of its row, swizzled like the TMA box
                        uint4 w;
                        w.x = pack_bf16(__uint_as_float(r[j * 8 + 0]), __uint_as_float(r[j * 8 + 1]));
                        w.y = pack_bf16(__uint_as_float(r[j * 8 + 2]), __uint_as_float(r[j * 8 + 3]));
                        w.z = pack_bf16(__uint_as_float(r[j * 8 + 4]), __uint_as_float(r[j * 8 + 5]));
                        w.w = pack_bf16(__uint_as_float(r[j * 8 + 6]), __uint_as_float(r[j * 8 + 7]));
                        sts128(st + row_off + (uint32_t)(((half * 4 + j) ^ (row & 7)) << 4), w);
                    }
                    if (p.mode == MODE_RAW) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        if (!RAW4 && e == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // store(g-2) has released the tile group g+1 will overwrite
                    }
                    EPI_BAR();
                    if (p.mode == MODE_RAW && e == 0) tma_store_4d(xbuf + xb * Cfg::XBUF_BYTES, &mapOut, n_tile * BN + g * 64, 0, h0, n0);
                    if (want_stats) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const uint2 v = lds64(st + st_off + (uint32_t)i * 2048);
                            const float x0 = bf_lo(v.x), x1 = bf_hi(v.x), x2 = bf_lo(v.y), x3 = bf_hi(v.y);
                            const float w0 = x0 * wr[i], w1 = x1 * wr[i], w2 = x2 * wr[i], w3 = x3 * wr[i];
                            acc_s[g][0] += w0; acc_q[g][0] = fmaf(w0, x0, acc_q[g][0]);
                            acc_s[g][1] += w1; acc_q[g][1] = fmaf(w1, x1, acc_q[g][1]);
                            acc_s[g][2] += w2; acc_q[g][2] = fmaf(w2, x2, acc_q[g][2]);
                            acc_s[g][3] += w3; acc_q[g][3] = fmaf(w3, x3, acc_q[g][3]);
                        }
                    }
                }
            } else if (p.mode == MODE_FINAL) {
                mbar_wait<0>(&tfull[acc], acc_phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
                for (int g = 0; g < G; ++g, ++gcount) {
                    const int b = gcount % XB;
                    const uint32_t xph = (uint32_t)(gcount / XB) & 1;
                    const uint32_t st = xb0 + b * Cfg::XBUF_BYTES;
                    const int col0 = n_tile * BN + g * 64 + half * 32;
                    // shift of this thread's 32 channels first (shared-memory latency overlaps the TMEM load), then the accumulator
                    const uint32_t pt = smem_u32(s_par) + (uint32_t)col0 * 4;
                    float4 t4[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) t4[j] = lds128f(pt + j * 16);
                    uint32_t r[32];                             // accumulator, then (in place) acc + shift as float bits
                    tmem_ld32(t_acc + g * 64 + half * 32, r);
                    TMEM_LD_WAIT();
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        r[4 * j + 0] = __float_as_uint(__uint_as_float(r[4 * j + 0]) + t4[j].x);
                        r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) + t4[j].y);
                        r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) + t4[j].z);
                        r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) + t4[j].w);
                    }
                    if (!DUAL) mbar_wait<0>(&xfull[b], xph);    // identity tile of this group has landed in the staging buffer
                    if (DUAL && RAW4) {                         // one staging tile: the previous group's store must have read it
                        if (e == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        EPI_BAR();
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t addr = st + row_off + (uint32_t)(((half * 4 + j) ^ (row & 7)) << 4);
                        const float *v = reinterpret_cast<const float *>(r) + 8 * j;
                        uint4 w;
                        if (!DUAL) {
                            const uint4 d = lds128(addr);
                            // ReLU after the rounding (one packed max per two channels): rounding is monotone and keeps 0, so it commutes
                            w.x = max_bf16x2(pack_bf16(v[0] + bf_lo(d.x), v[1] + bf_hi(d.x)), 0u);
                            w.y = max_bf16x2(pack_bf16(v[2] + bf_lo(d.y), v[3] + bf_hi(d.y)), 0u);
                            w.z = max_bf16x2(pack_bf16(v[4] + bf_lo(d.z), v[5] + bf_hi(d.z)), 0u);
                            w.w = max_bf16x2(pack_bf16(v[6] + bf_lo(d.w), v[7] + bf_hi(d.w)), 0u);
                        } else {
                            w.x = max_bf16x2(pack_bf16(v[0], v[1]), 0u);
                            w.y = max_bf16x2(pack_bf16(v[2], v[3]), 0u);
                            w.z = max_bf16x2(pack_bf16(v[4], v[5]), 0u);
                            w.w = max_bf16x2(pack_bf16(v[6], v[7]), 0u);
                        }
                        sts128(addr, w);
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    if (e == 0) {
                        if (DUAL) {
                            // no identity loader: the three staging tiles only rotate between the stores (as in the RAW epilogue)
                            if (!RAW4) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        } else {
                            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // every earlier store has finished reading its buffer
                            if (gcount > 0) mbar_arrive(&xfree[(gcount - 1) % XB]);
                        }
                    }
                    EPI_BAR();
                    if (e == 0) tma_store_4d(xbuf + b * Cfg::XBUF_BYTES, &mapOut, n_tile * BN + g * 64, 0, h0, n0);
                }
            } else if (half == 0) {
                // fp32 output with bias / scale / activation / residual (linear layers): direct stores, 4 warps
                mbar_wait<0>(&tfull[acc], acc_phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int wi = row % p.BW, hi = (row / p.BW) % p.BH, ni = row / (p.BW * p.BH);
                const int img = n0 + ni;
                const bool valid = img < p.Nimg;
                const long long m = ((long long)img * p.Ho + h0 + hi) * p.Wo + wi;
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    uint32_t r[32];
                    tmem_ld32(t_acc + c * 32, r);
                    TMEM_LD_WAIT();
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
            } else {
                mbar_wait<0>(&tfull[acc], acc_phase);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
        }
        if (want_stats) { if (stats_t) flush_stats_t(); else flush_stats(); }
        if (e == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    } else {
        // ===================================================== identity-tile loader (FINAL, single accumulator)
        if (p.mode == MODE_FINAL && !DUAL) {
            int gcount = 0;
            for (int tile = tile0; tile < total_tiles; tile += tile_step) {
                const int n_tile = tile % p.tiles_n, m_tile = MC ? 2 * (tile / p.tiles_n) + cr : tile / p.tiles_n;
                const int h0 = (m_tile % p.h_tiles) * p.BH, n0 = (m_tile / p.h_tiles) * p.BI;
                for (int g = 0; g < G; ++g, ++gcount) {
                    const int b = gcount % XB;
                    const uint32_t xph = (uint32_t)(gcount / XB) & 1;
                    mbar_wait<64>(&xfree[b], xph ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx(&xfull[b], Cfg::XBUF_BYTES);
                        tma_load_4d(xbuf + b * Cfg::XBUF_BYTES, &mapIdt, &xfull[b], n_tile * BN + g * 64, 0, h0, n0);
                    }
                    __syncwarp();
                }
            }
        }
    }

    __syncthreads();
    if (MC) cluster_sync_all();          // neither CTA leaves while the other may still signal its barriers
    if (want_stats) {
        for (int i = threadIdx.x; i < 2 * p.Cout; i += TC_THREADS) {
            const float v = s_par[i];
            if (v != 0.f) atomicAdd(p.stats + i, (double)v);
        }
    }
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (CG2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// 3x3 stride-1 convolution, HALO-BOX formulation (BUSCA_HALO=1).  EXPERIMENTAL: written after the GPU budget of round 1
// was spent, compiled but not yet run on a B200 - off by default until tests/test_gpu_conv_tc.py has passed with it.
//
// Why (profiles/r01f_conv_ncu_source_summary.md): the tap-by-tap kernel above fetches and transforms nine shifted copies
// of the same pixels per 64-channel block and is bound by shared-memory bandwidth.  Here ONE box per channel block is
// loaded - the (R+2) x (W+2) halo of R output rows of one image - transformed ONCE, and the nine taps are nine tcgen05
// A descriptors whose start address is shifted by (r*(W+2) + q) rows of 128 bytes inside that box: with the M index of
// the MMA running over FLAT positions f = ho*(W+2) + wo of the halo grid, input pixel (ho+r-1, wo+q-1) of every output
// pixel sits exactly (r*(W+2) + q) rows further.  Positions with wo >= W or ho >= R are junk rows of the accumulator
// (75-87 % of the 128 rows are real outputs); they are neither stored nor counted in the statistics.  The row shift is
// not a multiple of the 8-row swizzle atom; the hardware de-swizzles by address bits, so the plain descriptor is right
// (measured: tests/probe_umma.py, profiles/r02a_probe_umma.log - base_offset must stay 0).
// Roles: warp 0 = weight-tile producer, warp 18 = halo producer, warp 1 = MMA, warps 2-9 = transform (one group of 256:
// the transform runs once per nine taps and is off the critical path), warps 10-17 = epilogue (RAW mode only).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int HALO_STAGE_BYTES = 26624;          // probe kernel: 208 rows
constexpr int HALO_MAX_STAGES = 8;

// Round-2 measurements (profiles/r02b_probe_halo.log) showed what really bounds the 3x3 kernels: not shared-memory bandwidth but
// the WEIGHT tiles - re-fetched from L2 for every 128-pixel tile with only SB x B_BYTES in flight per SM against a ~1.6 us L2
// round trip (64 KB in flight -> one 16 KB tile per ~750 cycles, while its MMAs need 256).  Hence
//   RESW : the whole [Cout x 9 Cin] weight matrix stays resident in shared memory (64 -> 64: 72 KB), the ring streams halo boxes only;
//   MT=2 : every weight tile is used for TWO pixel tiles (two halo boxes, two TMEM accumulators), halving the weight bytes per FLOP.
struct HaloParams {
    int W, H, R, P;                  // output (= input) width / height, output rows per tile, halo pitch W + 2
    int Nimg, cin_blocks, h_tiles;   // h_tiles = H / R
    int total_tiles;                 // Nimg * h_tiles (Cout == BN: one channel tile)
    int n_groups;                    // ceil(total_tiles / MT): a group = MT pixel tiles sharing every weight tile
    int halo_rows;                   // (R + 2) * P rows of 128 B per box
    int valid_rows;                  // R * W dense output rows per tile
    int a_stage_bytes;               // bytes per halo stage: >= (2 P + 2 + 128) rows of 128 B (the taps address beyond the box), 1024-aligned
    int SA, SB;                      // ring depths (halo boxes, weight tiles)
    const uint16_t *a_xf;            // [2*Cin] theta (bf16), sign masks
    double *stats;                   // [2*Cout]
    const float *img_w;              // [Nimg] multiplicities or null
};

// K-major SWIZZLE_128B descriptor whose start is a whole number of 128-byte rows into a 1024-byte-aligned tile
// (measured on a B200, profiles/r02a_probe_umma.log: the 16-byte-chunk XOR of SWIZZLE_128B is taken from the ADDRESS bits of
// every row the tensor core reads, so a row-shifted start needs NO base_offset; setting bits 49-51 de-swizzles wrongly)
__device__ __forceinline__ uint64_t umma_desc_rowshift(uint32_t saddr) { return umma_desc<128>(saddr); }

template <int BN, int MT, bool RESW>
__global__ void __launch_bounds__(TC_THREADS, 1) conv3x3_halo_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                                      const __grid_constant__ CUtensorMap mapOut, const HaloParams p) {
    constexpr int G = BN / 64;
    constexpr int B_BYTES = BN * 128;
    constexpr int XBUF_BYTES = TC_BM * 128;
    constexpr int NACC = (2 * MT * BN <= 512) ? 2 : 1;      // TMEM accumulator buffers (each MT x BN columns)
    static_assert(MT * BN <= 512, "TMEM columns");
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *a_tiles = smem;
    uint8_t *b_tiles = a_tiles + p.SA * p.a_stage_bytes;                        // RESW: 9 * cin_blocks tiles, loaded once
    uint8_t *xbuf = b_tiles + (RESW ? 9 * p.cin_blocks : p.SB) * B_BYTES;
    float *s_par = reinterpret_cast<float *>(xbuf + TC_XBUFS * XBUF_BYTES);     // [2*BN] statistics
    float *s_apar = s_par + 2 * 256;
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_apar + 512);
    uint64_t *a_full = bars, *a_ready = a_full + HALO_MAX_STAGES, *a_empty = a_ready + HALO_MAX_STAGES, *b_full = a_empty + HALO_MAX_STAGES,
             *b_empty = b_full + HALO_MAX_STAGES;
    uint64_t *tfull = b_empty + HALO_MAX_STAGES, *tempty = tfull + 2, *bres = tempty + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bres + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_trigger();
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.SA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_ready[s], 256); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < p.SB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 8); }
        mbar_init(bres, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        prefetch_tmap(&mapA); prefetch_tmap(&mapB); prefetch_tmap(&mapOut);
    }
    for (int i = threadIdx.x; i < 2 * BN; i += TC_THREADS) s_par[i] = 0.f;
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    pdl_wait();
    {
        const int cin = p.cin_blocks * 64;
        uint16_t *sp = reinterpret_cast<uint16_t *>(s_apar);
        for (int i = threadIdx.x; i < cin; i += TC_THREADS) { sp[i] = p.a_xf[i]; sp[512 + i] = p.a_xf[cin + i]; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================================================== weight tiles: resident, or (tap, channel block) in the order the MMA consumes them
        if (RESW) {
            if (elect_one()) {
                const int nb = 9 * p.cin_blocks;
                mbar_expect_tx(bres, (uint32_t)nb * B_BYTES);
                for (int i = 0; i < nb; ++i) tma_load_2d(b_tiles + i * B_BYTES, &mapB, bres, i * 64, 0);
            }
            __syncwarp();
        } else {
            int sb = 0;
            uint32_t ph = 0;
            for (int grp = blockIdx.x; grp < p.n_groups; grp += gridDim.x)
                for (int cb = 0; cb < p.cin_blocks; ++cb)
                    for (int tap = 0; tap < 9; ++tap) {
                        mbar_wait<32>(&b_empty[sb], ph ^ 1);
                        if (elect_one()) {
                            mbar_expect_tx(&b_full[sb], B_BYTES);
                            tma_load_2d(b_tiles + sb * B_BYTES, &mapB, &b_full[sb], (tap * p.cin_blocks + cb) * 64, 0);
                        }
                        __syncwarp();
                        if (++sb == p.SB) { sb = 0; ph ^= 1; }
                    }
        }
    } else if (warp == TC_IDT_WARP) {
        // ===================================================== halo boxes: rows h0-1 .. h0+R, columns -1 .. W (zero fill outside)
        {
            int sa = 0;
            uint32_t ph = 0;
            for (int grp = blockIdx.x; grp < p.n_groups; grp += gridDim.x) {
                const int nt = min(MT, p.total_tiles - grp * MT);
                for (int cb = 0; cb < p.cin_blocks; ++cb)
                    for (int m = 0; m < nt; ++m) {
                        const int tile = grp * MT + m;
                        const int n = tile / p.h_tiles, h0 = (tile % p.h_tiles) * p.R;
                        mbar_wait<32>(&a_empty[sa], ph ^ 1);
                        if (elect_one()) {
                            mbar_expect_tx(&a_full[sa], (uint32_t)p.halo_rows * 128u);
                            tma_load_4d(a_tiles + sa * p.a_stage_bytes, &mapA, &a_full[sa], cb * 64, -1, h0 - 1, n);
                        }
                        __syncwarp();
                        if (++sa == p.SA) { sa = 0; ph ^= 1; }
                    }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer: nine shifted views of each halo box per channel block
        // The whole warp walks the loops (uniform values -> uniform registers, no per-instruction lane election code); one
        // elected lane issues the tcgen05 instructions.  Measured (r02d): with an `if (lane == 0)` body the issue loop cost
        // ~18 SASS instructions per MMA - as long as a 128x64 / 128x128 MMA itself.
        {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            int sa = 0, sb = 0, it = 0;
            uint32_t pa = 0, pb = 0;
            if (RESW) mbar_wait<32>(bres, 0);
            for (int grp = blockIdx.x; grp < p.n_groups; grp += gridDim.x, ++it) {
                const int acc = NACC == 2 ? (it & 1) : 0;
                const uint32_t acc_phase = NACC == 2 ? ((it >> 1) & 1) : (it & 1);
                const int nt = min(MT, p.total_tiles - grp * MT);
                mbar_wait<32>(&tempty[acc], acc_phase ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + acc * (MT * BN);
                for (int cb = 0; cb < p.cin_blocks; ++cb) {
                    uint32_t a_base[MT];
                    int st[MT];
#pragma unroll
                    for (int m = 0; m < MT; ++m) {
                        a_base[m] = 0; st[m] = 0;
                        if (m < nt) {
                            st[m] = sa;
                            mbar_wait<0>(&a_ready[sa], pa);           // landed (the transform waited for a_full) and transformed
                            a_base[m] = smem_u32(a_tiles + sa * p.a_stage_bytes);
                            if (++sa == p.SA) { sa = 0; pa ^= 1; }
                        }
                    }
                    int r = 0, q = 0;
                    for (int tap = 0; tap < 9; ++tap) {
                        if (!RESW) mbar_wait<0>(&b_full[sb], pb);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint64_t db = umma_desc<128>(smem_u32(b_tiles + (RESW ? (tap * p.cin_blocks + cb) : sb) * B_BYTES));
                        const uint32_t shift = (uint32_t)(r * p.P + q) * 128u;
                        if (elect_one()) {
#pragma unroll
                            for (int m = 0; m < MT; ++m) {
                                if (m < nt) {
                                    const uint64_t da = umma_desc_rowshift(a_base[m] + shift);
#pragma unroll
                                    for (int k = 0; k < 4; ++k) umma_bf16(d_tmem + m * BN, da + 2 * k, db + 2 * k, idesc, !(cb == 0 && tap == 0 && k == 0));
                                }
                            }
                            if (!RESW) umma_commit(&b_empty[sb]);
                        }
                        __syncwarp();
                        if (!RESW) {
                            if (++sb == p.SB) { sb = 0; pb ^= 1; }
                        }
                        if (++q == 3) { q = 0; ++r; }
                    }
                    if (elect_one()) {
#pragma unroll
                        for (int m = 0; m < MT; ++m)
                            if (m < nt) umma_commit(&a_empty[st[m]]);     // the nine taps have read the box
                    }
                    __syncwarp();
                }
                if (elect_one()) umma_commit(&tfull[acc]);
                __syncwarp();
            }
        }
    } else if (warp < TC_EPI_WARP0) {
        // ===================================================== transform of the halo box, once per channel block
        const int tt = threadIdx.x - TC_XF_WARP0 * 32;      // 0..255
        const int c = tt & 7, rb = tt >> 3;                 // 16-byte chunk, first row; rows rb + 32*i (same swizzle phase)
        const uint32_t col_off = (uint32_t)((c ^ (rb & 7)) << 4);
        constexpr int NI = 6;                               // 6 * 32 = 192 >= 170 rows
        int hr[NI];
        bool in_box[NI], col_in[NI];
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const int row = rb + 32 * i;
            in_box[i] = row < p.halo_rows;
            hr[i] = row / p.P;
            const int wr = row % p.P;
            col_in[i] = wr >= 1 && wr <= p.W;
        }
        const uint32_t par0 = smem_u32(s_apar) + (uint32_t)c * 16;
        const uint32_t tiles0 = smem_u32(a_tiles) + col_off + (uint32_t)rb * 128;
        int sa = 0;
        uint32_t ph = 0;
        for (int grp = blockIdx.x; grp < p.n_groups; grp += gridDim.x) {
            const int nt = min(MT, p.total_tiles - grp * MT);
            for (int cb = 0; cb < p.cin_blocks; ++cb) {
                const uint4 th = lds128(par0 + (uint32_t)cb * 128), sg = lds128(par0 + (uint32_t)cb * 128 + 1024);
                for (int m = 0; m < nt; ++m) {
                    const int h0 = ((grp * MT + m) % p.h_tiles) * p.R;
                    const bool top_out = h0 == 0, bot_out = h0 + p.R == p.H;       // halo row 0 / R+1 lies outside the image
                    mbar_wait<0>(&a_full[sa], ph);
                    const uint32_t base = tiles0 + sa * p.a_stage_bytes;
#pragma unroll
                    for (int i = 0; i < NI; ++i) {
                        if (!in_box[i]) continue;
                        const uint32_t addr = base + (uint32_t)i * 4096;
                        uint4 v = lds128(addr);
                        const bool inside = col_in[i] && !(top_out && hr[i] == 0) && !(bot_out && hr[i] == p.R + 1);
                        if (inside) {
                            v.x = max_bf16x2(v.x ^ sg.x, th.x); v.y = max_bf16x2(v.y ^ sg.y, th.y);
                            v.z = max_bf16x2(v.z ^ sg.z, th.z); v.w = max_bf16x2(v.w ^ sg.w, th.w);
                        } else {
                            v = th;                                  // padding reads theta: |s|*theta + t = 0 (see the header)
                        }
                        sts128(addr, v);
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_arrive(&a_ready[sa]);
                    if (++sa == p.SA) { sa = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp < TC_IDT_WARP) {
        // ===================================================== epilogue: real rows -> dense staging tile -> TMA store; statistics
        const int e = threadIdx.x - TC_EPI_WARP0 * 32;   // 0..255
        const int q = warp & 3;
        const int half = (warp - TC_EPI_WARP0) >> 2;
        const int f = q * 32 + lane;                     // accumulator row = flat halo position
        const int ho = f / p.P, wo = f % p.P;
        const bool real = ho < p.R && wo < p.W;
        const int d = ho * p.W + wo;                     // dense row of the staging tile ([R][W] pixels)
        const uint32_t xb0 = smem_u32(xbuf);
        const int cq = e & 15, rsub = e >> 4;            // statistics: 4 channels x rows rsub + 16*i
        const uint32_t st_off = (uint32_t)rsub * 128 + (uint32_t)((((cq >> 1) ^ (rsub & 7)) << 4) + (cq & 1) * 8);
        float acc_s[G][4], acc_q[G][4];
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int j = 0; j < 4; ++j) { acc_s[g][j] = 0.f; acc_q[g][j] = 0.f; }
        int it = 0, gcount = 0;
        for (int grp = blockIdx.x; grp < p.n_groups; grp += gridDim.x, ++it) {
            const int acc = NACC == 2 ? (it & 1) : 0;
            const uint32_t acc_phase = NACC == 2 ? ((it >> 1) & 1) : (it & 1);
            const int nt = min(MT, p.total_tiles - grp * MT);
            mbar_wait<0>(&tfull[acc], acc_phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int m = 0; m < nt; ++m) {
                const int tile = grp * MT + m;
                const int n = tile / p.h_tiles, h0 = (tile % p.h_tiles) * p.R;
                const float wimg = p.img_w ? __ldg(p.img_w + n) : 1.f;
                const uint32_t t_acc = tmem_base + acc * (MT * BN) + m * BN + ((uint32_t)(q * 32) << 16);
#pragma unroll
                for (int g = 0; g < G; ++g, ++gcount) {
                    const int xb = gcount % TC_XBUFS;          // three staging tiles in rotation (see conv_tc_kernel)
                    const uint32_t st = xb0 + xb * XBUF_BYTES;
                    uint32_t r[32];
                    tmem_ld32(t_acc + g * 64 + half * 32, r);
                    TMEM_LD_WAIT();
                    if (real) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint4 w;
                            w.x = pack_bf16(__uint_as_float(r[j * 8 + 0]), __uint_as_float(r[j * 8 + 1]));
                            w.y = pack_bf16(__uint_as_float(r[j * 8 + 2]), __uint_as_float(r[j * 8 + 3]));
                            w.z = pack_bf16(__uint_as_float(r[j * 8 + 4]), __uint_as_float(r[j * 8 + 5]));
                            w.w = pack_bf16(__uint_as_float(r[j * 8 + 6]), __uint_as_float(r[j * 8 + 7]));
                            sts128(st + (uint32_t)d * 128 + (uint32_t)(((half * 4 + j) ^ (d & 7)) << 4), w);
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    if (e == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    EPI_BAR();
                    if (e == 0) tma_store_4d(xbuf + xb * XBUF_BYTES, &mapOut, g * 64, 0, h0, n);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (rsub + 16 * i < p.valid_rows) {
                            const uint2 v = lds64(st + st_off + (uint32_t)i * 2048);
                            const float x0 = bf_lo(v.x), x1 = bf_hi(v.x), x2 = bf_lo(v.y), x3 = bf_hi(v.y);
                            const float w0 = x0 * wimg, w1 = x1 * wimg, w2 = x2 * wimg, w3 = x3 * wimg;
                            acc_s[g][0] += w0; acc_q[g][0] = fmaf(w0, x0, acc_q[g][0]);
                            acc_s[g][1] += w1; acc_q[g][1] = fmaf(w1, x1, acc_q[g][1]);
                            acc_s[g][2] += w2; acc_q[g][2] = fmaf(w2, x2, acc_q[g][2]);
                            acc_s[g][3] += w3; acc_q[g][3] = fmaf(w3, x3, acc_q[g][3]);
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
        }
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                atomicAdd(&s_par[g * 64 + cq * 4 + j], acc_s[g][j]);
                atomicAdd(&s_par[BN + g * 64 + cq * 4 + j], acc_q[g][j]);
            }
        if (e == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }

    __syncthreads();
    if (p.stats)
        for (int i = threadIdx.x; i < 2 * BN; i += TC_THREADS) {
            const float v = s_par[i];
            if (v != 0.f) atomicAdd(p.stats + i, (double)v);
        }
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Probe of the one hardware assumption conv3x3_halo_kernel rests on (busca_debug_umma_rowshift, tests/probe_umma.py): an
// A descriptor whose start address lies `shift` rows of 128 bytes into a 1024-byte-aligned SWIZZLE_128B tile.  One CTA,
// A = 208 rows x 64 bf16 written with the TMA swizzle by ordinary stores, B = the 64 x 64 identity, so D[m][n] must be
// A[m + shift][n]: fill 0 -> A[row][k] = row (which ROW each accumulator row read), fill 1 -> A[row][k] = k (whether the
// 16-byte chunks are de-swizzled with the right phase).  use_base_offset selects descriptor bits 49-51 = (start >> 7) & 7
// (PTX ISA) or 0.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) umma_rowshift_probe_kernel(int shift, int fill, int use_base_offset, float *__restrict__ out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *a_tile = smem;                              // 208 rows x 128 B
    uint8_t *b_tile = smem + HALO_STAGE_BYTES;           // 64 rows x 128 B
    uint64_t *bar = reinterpret_cast<uint64_t *>(b_tile + 64 * 128);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 208 * 64; i += 128) {
        const int row = i >> 6, k = i & 63;
        const float v = fill == 0 ? (float)row : (float)k;
        *reinterpret_cast<__nv_bfloat16 *>(a_tile + row * 128 + (((k >> 3) ^ (row & 7)) << 4) + (k & 7) * 2) = __float2bfloat16(v);
    }
    for (int i = threadIdx.x; i < 64 * 64; i += 128) {
        const int n = i >> 6, k = i & 63;
        *reinterpret_cast<__nv_bfloat16 *>(b_tile + n * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2) = __float2bfloat16(n == k ? 1.f : 0.f);
    }
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a_addr = smem_u32(a_tile) + (uint32_t)shift * 128u;
        const uint64_t da = umma_desc<128>(a_addr) | (use_base_offset ? ((uint64_t)((a_addr >> 7) & 7u) << 49) : 0ull);
        const uint64_t db = umma_desc<128>(smem_u32(b_tile));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, da + 2 * k, db + 2 * k, idesc, k != 0);
        umma_commit(bar);
    }
    mbar_wait<32>(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int m = warp * 32 + lane;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + h * 32, r);
        TMEM_LD_WAIT();
#pragma unroll
        for (int j = 0; j < 32; ++j) out[m * 64 + h * 32 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Batch statistics of a 1x1 convolution WITHOUT computing it: the Gram matrix of its input.
//   y[p][c] = sum_k W[c][k] a[p][k]   =>   sum_p y[p][c] = W[c] . m,   sum_p y[p][c]^2 = W[c]^T G W[c],   m = sum_p a[p], G = sum_p a[p] a[p]^T
// A bottleneck's last convolution has Cout = 4 Cin, so G (Cin x Cin) costs a quarter of the FLOPs of the statistics-only pass it
// replaces and needs no per-tile epilogue at all: the accumulator stays in TMEM for the whole launch.  The pixel tiles are the same
// K-major, 128-byte-swizzled [128 pixels x 64 channels] TMA boxes conv_tc_kernel consumes (and the same in-place max-transform of the
// producing BN + ReLU); they are read MN-major by BOTH tcgen05 operands (probe: umma_gram_probe_kernel), the column sums m come from
// a second MMA against an all-ones B tile.  NB = Cin / 64 in {1, 2, 4}: one stage holds all NB channel tiles of a pixel tile, 16 KB
// apart (= the LBO of the MN-major descriptors).  TMEM: NB=1 G[64x64] | m ; NB=2 G[128x128] | m ; NB=4 rows 0-127 x 256 | rows 128-255 x
// columns 128-255 | m, m (the missing quadrant is the transpose of a computed one).  At the end every CTA adds its partial to G / m in
// global memory (fp32 reductions); the quadratic forms are gram_quadform_kernel (reid.cu), in fp64.
// Images are not weighted here: the host sends images with multiplicity 1 (the vast majority) through this path and the few repeated
// ones through the weighted statistics-only pass of conv_tc_kernel.
// ---------------------------------------------------------------------------------------------------------------------
// four consecutive floats added with one reduction (REDG.E.ADD.F32x4): a quarter of the atomic operations of a TMEM dump
__device__ __forceinline__ void red_add_v4(float *p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(__uint_as_float(a)), "f"(__uint_as_float(b)), "f"(__uint_as_float(c)),
                 "f"(__uint_as_float(d))
                 : "memory");
}
// MN-major SWIZZLE_128B operand: 64 channels per 128-byte row, LBO between 64-channel groups, SBO = 1 KB between 8-pixel groups
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
constexpr int GRAM_THREADS = 448;          // warp 0 TMA, warp 1 MMA, warps 2-9 transform, warps 10-13 final TMEM dump
constexpr int GRAM_DUMP_WARP0 = 10;

struct GramParams {
    int tiles_m, h_tiles, BH, BI, Nimg, img_shift;
    const uint16_t *a_xf;            // [2*Cin] theta (bf16), sign masks - or null (input already activated)
    float *gpart;                    // [C*C] fp32, zero on entry: every CTA adds its partial G (all four quadrants)
    float *spart;                    // [C]   fp32, zero on entry: column sums m
};

template <int NB>
__global__ void __launch_bounds__(GRAM_THREADS, 1) gram_stats_kernel(const __grid_constant__ CUtensorMap mapA, const GramParams p) {
    constexpr int STAGES = NB == 1 ? 8 : (NB == 2 ? 5 : 3);
    constexpr int STAGE_BYTES = NB * 16384;
    constexpr int C = NB * 64;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *tiles = smem;
    uint8_t *ones = tiles + STAGES * STAGE_BYTES;                 // 2 KB of bf16 1.0
    uint16_t *s_apar = reinterpret_cast<uint16_t *>(ones + 2048); // theta [512], sign [512]
    uint64_t *bars = reinterpret_cast<uint64_t *>(ones + 2048 + 2048);
    uint64_t *full = bars, *ready = full + STAGES, *empty = ready + STAGES, *done = empty + STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool xform = p.a_xf != nullptr;

    pdl_trigger();
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&ready[s], 256); mbar_init(&empty[s], 1); }
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        prefetch_tmap(&mapA);
    }
    for (int i = threadIdx.x; i < 1024; i += GRAM_THREADS) reinterpret_cast<uint16_t *>(ones)[i] = 0x3F80;     // bf16 1.0
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    pdl_wait();
    if (xform)
        for (int i = threadIdx.x; i < C; i += GRAM_THREADS) { s_apar[i] = p.a_xf[i]; s_apar[512 + i] = p.a_xf[C + i]; }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // the ones tile is read by the tensor core (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    constexpr uint32_t COL_S0 = NB == 1 ? 64 : (NB == 2 ? 128 : 384), COL_S1 = 400, COL_G1 = 256;

    if (warp == 0) {
        // ===================================================== TMA producer: the NB channel tiles of one pixel tile per stage
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < p.tiles_m; tile += gridDim.x) {
            const int h0 = (tile % p.h_tiles) * p.BH, n0 = (tile / p.h_tiles) * p.BI;
            mbar_wait<32>(&empty[stage], phase ^ 1);
            if (elect_one()) {
                mbar_expect_tx(&full[stage], STAGE_BYTES);
#pragma unroll
                for (int cb = 0; cb < NB; ++cb) tma_load_4d(tiles + stage * STAGE_BYTES + cb * 16384, &mapA, &full[stage], cb * 64, 0, h0, n0);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer: G += T^T T, m += T^T 1 for every pixel tile T; one commit at the end
        constexpr int M = NB == 1 ? 64 : 128;
        constexpr uint32_t ID_MN = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(M >> 4) << 24);   // A, B MN-major
        constexpr uint32_t ID_ONES = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(M >> 4) << 24);   // B K-major
        const uint64_t d_ones = umma_desc<128>(smem_u32(ones));
        int stage = 0;
        uint32_t phase = 0;
        bool first = true;
        for (int tile = blockIdx.x; tile < p.tiles_m; tile += gridDim.x) {
            mbar_wait<0>(xform ? &ready[stage] : &full[stage], phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t base = smem_u32(tiles + stage * STAGE_BYTES);
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {                        // 16 pixels per step = two 8-pixel groups = 2 KB
                    const uint32_t acc = !(first && ks == 0);
                    const uint64_t d0 = umma_desc_mn(base + ks * 2048, 16384);
                    if (NB == 4) {
                        const uint64_t d1 = umma_desc_mn(base + 2 * 16384 + ks * 2048, 16384);
                        umma_bf16(tmem_base, d0, d0, ID_MN | ((uint32_t)(256 >> 3) << 17), acc);
                        umma_bf16(tmem_base + COL_G1, d1, d1, ID_MN | ((uint32_t)(128 >> 3) << 17), acc);
                        umma_bf16(tmem_base + COL_S0, d0, d_ones, ID_ONES, acc);
                        umma_bf16(tmem_base + COL_S1, d1, d_ones, ID_ONES, acc);
                    } else {
                        umma_bf16(tmem_base, d0, d0, ID_MN | ((uint32_t)(C >> 3) << 17), acc);
                        umma_bf16(tmem_base + COL_S0, d0, d_ones, ID_ONES, acc);
                    }
                }
                umma_commit(&empty[stage]);
            }
            __syncwarp();
            first = false;
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit(done);
        __syncwarp();
    } else if (warp < GRAM_DUMP_WARP0) {
        // ===================================================== transform: a = max(x ^ signmask, theta) in place; rows of images beyond N stay 0
        if (xform) {
            const int tt = threadIdx.x - 64;                    // 0..255
            const int c = tt & 7, rb = tt >> 3;                 // 16-byte chunk, rows rb + 32 i
            const uint32_t col_off = (uint32_t)((c ^ (rb & 7)) << 4);
            const uint32_t par0 = smem_u32(s_apar) + (uint32_t)c * 16;
            const uint32_t tiles0 = smem_u32(tiles) + col_off + (uint32_t)rb * 128;
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.tiles_m; tile += gridDim.x) {
                const int n0 = (tile / p.h_tiles) * p.BI;
                const bool all_in = n0 + p.BI <= p.Nimg;
                mbar_wait<0>(&full[stage], phase);
                const uint32_t base = tiles0 + stage * STAGE_BYTES;
#pragma unroll
                for (int cb = 0; cb < NB; ++cb) {
                    const uint4 th = lds128(par0 + (uint32_t)cb * 128), sg = lds128(par0 + (uint32_t)cb * 128 + 1024);
                    uint4 v[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) v[i] = lds128(base + cb * 16384 + (uint32_t)i * 4096);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (all_in || n0 + ((rb + 32 * i) >> p.img_shift) < p.Nimg) {
                            v[i].x = max_bf16x2(v[i].x ^ sg.x, th.x); v[i].y = max_bf16x2(v[i].y ^ sg.y, th.y);
                            v[i].z = max_bf16x2(v[i].z ^ sg.z, th.z); v[i].w = max_bf16x2(v[i].w ^ sg.w, th.w);
                        } else {
                            v[i] = make_uint4(0u, 0u, 0u, 0u);
                        }
                        sts128(base + cb * 16384 + (uint32_t)i * 4096, v[i]);
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(&ready[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===================================================== final dump: TMEM -> G and m in global memory (fp32 reductions, no return value)
        const int qq = warp & 3;                                // TMEM lane quarter this warp may read
        mbar_wait<64>(done, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float *g = p.gpart, *sv = p.spart;
        const uint32_t t_lane = tmem_base + ((uint32_t)(qq * 32) << 16);
        if (NB == 1) {
            const int row = qq * 16 + lane;                     // M = 64: rows in lanes 0..15 of every quarter
            for (int c0 = 0; c0 < 64; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(t_lane + c0, r);
                TMEM_LD_WAIT();
                if (lane < 16)
#pragma unroll
                    for (int j = 0; j < 32; j += 4) red_add_v4(&g[row * C + c0 + j], r[j], r[j + 1], r[j + 2], r[j + 3]);
            }
            uint32_t r16[16];
            tmem_ld16(t_lane + COL_S0, r16);
            TMEM_LD_WAIT();
            if (lane < 16) atomicAdd(&sv[row], __uint_as_float(r16[0]));
        } else {
            const int row = qq * 32 + lane;
            for (int c0 = 0; c0 < (NB == 4 ? 256 : C); c0 += 32) {
                uint32_t r[32];
                tmem_ld32(t_lane + c0, r);
                TMEM_LD_WAIT();
#pragma unroll
                for (int j = 0; j < 32; j += 4) red_add_v4(&g[row * C + c0 + j], r[j], r[j + 1], r[j + 2], r[j + 3]);
                if (NB == 4 && c0 >= 128)                       // the quadrant that is not computed: its transpose (consecutive lanes = consecutive addresses)
#pragma unroll
                    for (int j = 0; j < 32; ++j) atomicAdd(&g[(c0 + j) * C + row], __uint_as_float(r[j]));
            }
            uint32_t r16[16];
            tmem_ld16(t_lane + COL_S0, r16);
            TMEM_LD_WAIT();
            atomicAdd(&sv[row], __uint_as_float(r16[0]));
            if (NB == 4) {
                for (int c0 = 0; c0 < 128; c0 += 32) {          // rows 128..255, columns 128..255
                    uint32_t r[32];
                    tmem_ld32(t_lane + COL_G1 + c0, r);
                    TMEM_LD_WAIT();
#pragma unroll
                    for (int j = 0; j < 32; j += 4) red_add_v4(&g[(128 + row) * C + 128 + c0 + j], r[j], r[j + 1], r[j + 2], r[j + 3]);
                }
                tmem_ld16(t_lane + COL_S1, r16);
                TMEM_LD_WAIT();
                atomicAdd(&sv[128 + row], __uint_as_float(r16[0]));
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Probe of the operand form the Gram-matrix statistics rest on (busca_debug_gram, tests/probe_gram.py): D = A^T A with BOTH operands
// read MN-major from the same K-major-stored pixel tiles.  nb tiles of [128 pixels x 64 channels] bf16 (128-byte rows, SWIZZLE_128B,
// 16 KB apart) are exactly the canonical MN-major SW128 layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units with the channel
// dimension as M/N (64 channels = one 128-byte row), LBO = 16 KB between 64-channel groups and SBO = 1 KB between 8-pixel groups.
//   nb = 1: M = 64  (TMEM rows in lanes (r % 16) + 32 * (r / 16)), N = 64
//   nb = 2: M = 128, N = 128;   nb = 4: two M halves of 128 channels, N = 256
// ---------------------------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(128, 1) umma_gram_probe_kernel(const uint16_t *__restrict__ a /* [128][64*nb] */, int nb, float *__restrict__ out /* [64nb][64nb] */) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 4 * 16384);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int C = 64 * nb;
    for (int i = threadIdx.x; i < 128 * C; i += 128) {
        const int row = i / C, ch = i % C, t = ch >> 6, k = ch & 63;
        *reinterpret_cast<uint16_t *>(smem + t * 16384 + row * 128 + (((k >> 3) ^ (row & 7)) << 4) + (k & 7) * 2) = a[i];
    }
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const int M = nb == 1 ? 64 : 128, N = C, halves = nb == 4 ? 2 : 1;
    if (threadIdx.x == 0) {
        // both operands MN-major (bits 15, 16)
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t base = smem_u32(smem);
        for (int h = 0; h < halves; ++h)
            for (int ks = 0; ks < 8; ++ks) {                       // 16 pixels per step = two 8-pixel groups = 2 KB
                const uint64_t da = umma_desc_mn(base + h * 2 * 16384 + ks * 2048, 16384);
                const uint64_t db = umma_desc_mn(base + ks * 2048, 16384);
                umma_bf16(tmem_base + h * 256, da, db, idesc, ks != 0);
            }
        umma_commit(bar);
    }
    mbar_wait<32>(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int h = 0; h < halves; ++h)
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + h * 256 + c0, r);
            TMEM_LD_WAIT();
            int row;
            bool valid = true;
            if (M == 64) { row = warp * 16 + lane; valid = lane < 16; }
            else row = h * 128 + warp * 32 + lane;
            if (valid)
                for (int j = 0; j < 32; ++j) out[(size_t)row * C + c0 + j] = __uint_as_float(r[j]);
        }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

int g_num_sms = 0;
char g_last_kernel[64] = "";      // template instantiation of the last launch, spelled as ncu prints it

struct TcMaps {
    CUtensorMap a[4], b, b2, out, idt;
};

// BUSCA_RESB=0 disables the resident-weight variant (A/B comparison on the GPU box)
bool resb_enabled() {
    static int on = -1;
    if (on < 0) {
        const char *e = getenv("BUSCA_RESB");
        on = !(e && e[0] == '0');
    }
    return on != 0;
}

// The paired (weight-multicast) variant is OFF unless BUSCA_MC=1: measured on a B200 (profiles/r03j) it changes nothing - 43.0 vs 42.9
// ms/frame - i.e. the streamed-weight kernels are not bound by L2 delivery of the weight tiles but by the bytes the ring keeps in flight
// per SM (see S4).  It stays because the tests hold it bit-identical and a cta_group::2 kernel would start from this lockstep pair.
bool mc_enabled() {
    static int on = -1;
    if (on < 0) {
        const char *e = getenv("BUSCA_MC");
        on = (e && e[0] == '1');
    }
    return on != 0;
}

template <int BN, int KB, bool DUAL, bool RESB, bool MC, bool S4 = false, bool CG2 = false, bool RAW4 = false>
cudaError_t launch_tc_v(const TcMaps &m, const TcParams &p, cudaStream_t s) {
    using Cfg = TcCfg<BN, KB, DUAL, RESB>;
    // RAW4: four 48 KB stages + one staging tile + 2 x 512 statistics floats + transform parameters + barriers
    constexpr int SMEM_BYTES = RAW4 ? 1024 + 196608 + Cfg::XBUF_BYTES + ((DUAL ? 2048 : 1024) + Cfg::APAR_FLOATS) * 4 + 512     // 4 x 48 KB or 6 x 32 KB of ring
                               : Cfg::SMEM + ((RESB && !DUAL && BN == 256) ? (TC_XBUFS_MAX - TC_XBUFS) * Cfg::XBUF_BYTES - (Cfg::PAR_FLOATS - 2048) * 4 : 0);
    static_assert(SMEM_BYTES <= 232448, "shared memory budget (227 KB per CTA)");
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN, KB, DUAL, RESB, MC, S4, CG2, RAW4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    if (!g_num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    snprintf(g_last_kernel, sizeof(g_last_kernel), "conv_tc_kernel<%d, %d, %d, %d, %d, %d, %d, %d>", BN, KB, (int)DUAL, (int)RESB, (int)MC, (int)S4, (int)CG2, (int)RAW4);
    if (MC) {
        // clusters of two CTAs: pair tiles = ceil(tiles_m / 2) x tiles_n, one pair per two SMs, a pair stays on one channel tile when cheap
        const int pair_tiles = ((p.tiles_m + 1) / 2) * p.tiles_n;
        int pairs = pair_tiles < g_num_sms / 2 ? pair_tiles : g_num_sms / 2;
        if (p.tiles_n > 1 && pairs > p.tiles_n && pairs % p.tiles_n != 0 && (pairs % p.tiles_n) * 32 < pairs) pairs -= pairs % p.tiles_n;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = s;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
        cfg.attrs = at; cfg.numAttrs = 2;
        return cudaLaunchKernelEx(&cfg, conv_tc_kernel<BN, KB, DUAL, RESB, MC, S4, CG2, RAW4>, m.a[0], m.a[1], m.a[2], m.a[3], m.b, m.b2, m.out, m.idt, p);
    }
    const int total = p.tiles_m * p.tiles_n;
    int grid = total < g_num_sms ? total : g_num_sms;
    // keep a CTA on one channel tile (its statistics accumulators stay in registers) when that costs < 3 % of the SMs;
    // with resident weights it is a requirement (total is a multiple of tiles_n, so grid >= tiles_n stays one)
    if (p.tiles_n > 1 && grid > p.tiles_n && grid % p.tiles_n != 0 && (RESB || (grid % p.tiles_n) * 32 < grid)) grid -= grid % p.tiles_n;
    return launch_pdl(conv_tc_kernel<BN, KB, DUAL, RESB, MC, S4, CG2, RAW4>, dim3(grid), dim3(TC_THREADS), SMEM_BYTES, s, m.a[0], m.a[1], m.a[2], m.a[3], m.b, m.b2, m.out, m.idt, p);
}

// whether a launch takes the paired variant: streamed weights, BN = 256, enough pixel tiles to keep every pair busy
// (busca_set_option("mc_min_tiles", n): the tests lower the threshold to run small shapes through it)
int g_mc_min_tiles = -1;
bool tc_use_mc(int BN, const TcParams &p) {
    const int min_tiles = g_mc_min_tiles >= 0 ? g_mc_min_tiles : 4 * (g_num_sms ? g_num_sms : 148);
    return (mc_enabled() || g_mc_min_tiles >= 0) && BN == 256 && p.mode != MODE_F32 && p.k_iters > 2 && p.tiles_m >= min_tiles;
}

// cta_group::2 variant: BN = 256, streamed weights, RAW or FINAL epilogues.  Measured per launch on a B200 (profiles/r04x_cg2_per_launch.txt,
// bit-identical outputs): convolutions WITHOUT an A-tile transform (the 1x1 conv1 of layers 3-4, whose input is a finished block output)
// run 10-13 % faster - a third fewer bytes into every SM per k-iteration, four ring stages instead of three, half of B fetched per SM;
// launches WITH the transform run 7-26 % slower - their critical path is the transform role, which the pair only adds a cross-CTA
// dependency to (the MMA waits for the slower of two CTAs every stage).  Default: on for the first kind only.
// BUSCA_CG2=0: never; BUSCA_CG2=2 or busca_set_option("cg2_min_tiles", n): every eligible launch (tests, A/B runs).
int g_cg2_min_tiles = -1;
int cg2_mode() {
    static int mode = -1;
    if (mode < 0) {
        const char *e = getenv("BUSCA_CG2");
        mode = e ? (e[0] == '0' ? 0 : (e[0] == '2' ? 2 : 1)) : 1;
    }
    return mode;
}
bool tc_use_cg2(int BN, const TcParams &p) {
    if (BN != 256 || !(p.mode == MODE_RAW || p.mode == MODE_FINAL) || p.k_iters <= 2) return false;
    if (g_cg2_min_tiles >= 0) return p.tiles_m >= g_cg2_min_tiles;
    const int min_tiles = 2 * (g_num_sms ? g_num_sms : 148);
    if (cg2_mode() == 0 || p.tiles_m < min_tiles) return false;
    return cg2_mode() == 2 || p.a_xf == nullptr;
}

template <int BN, int KB = 128, bool DUAL = false>
cudaError_t launch_tc(const TcMaps &m, const TcParams &p, cudaStream_t s) {
    // resident weights whenever the CTA's slab [BN x K] fits and every CTA can stay on one channel tile
    const int total = p.tiles_m * p.tiles_n;
    // (only for tiles of one or two k-iterations: there the weight tile would otherwise be re-fetched for every 16 KB of A;
    // longer k-loops keep the deeper ring, which hides the TMA + transform latency of a stage)
    if (resb_enabled() && p.k_iters <= 2 && (long long)p.k_iters * TcCfg<BN, KB, DUAL, true>::B_BYTES <= TcCfg<BN, KB, DUAL, true>::RES_BYTES &&
        total >= p.tiles_n && (BN != 256 || DUAL || (p.mode == MODE_FINAL ? p.Cout <= 2048 : 2 * p.Cout <= 2048)))
        return launch_tc_v<BN, KB, DUAL, true, false>(m, p, s);
    if constexpr (BN == 128 && !DUAL) {
        static const bool raw4 = !(getenv("BUSCA_RAW4") && getenv("BUSCA_RAW4")[0] == '0');
        if (raw4 && p.mode == MODE_RAW && p.k_iters >= 8 && p.Cout <= 512 && p.a_xf != nullptr)
            return launch_tc_v<BN, KB, DUAL, false, false, false, false, true>(m, p, s);         // six ring stages, one staging tile
    }
    if constexpr (BN == 256) {
        if (tc_use_cg2(BN, p)) {
            if constexpr (!DUAL) {
                static const bool raw4 = !(getenv("BUSCA_RAW4") && getenv("BUSCA_RAW4")[0] == '0');
                if (raw4 && p.mode == MODE_RAW && p.Cout <= 512) return launch_tc_v<BN, KB, DUAL, false, true, false, true, true>(m, p, s);   // six ring stages
            }
            return launch_tc_v<BN, KB, DUAL, false, true, false, true>(m, p, s);
        }
        if (tc_use_mc(BN, p)) return launch_tc_v<BN, KB, DUAL, false, true>(m, p, s);
        if constexpr (DUAL) {
            static const bool raw4d = !(getenv("BUSCA_RAW4") && getenv("BUSCA_RAW4")[0] == '0');
            if (raw4d && p.mode == MODE_FINAL && p.k_iters >= 8 && p.Cout <= 2048)
                return launch_tc_v<BN, KB, DUAL, false, false, false, false, true>(m, p, s);     // four ring stages, one staging tile
        }
        if constexpr (!DUAL) {
            static const bool raw4 = !(getenv("BUSCA_RAW4") && getenv("BUSCA_RAW4")[0] == '0');
            if (raw4 && p.mode == MODE_RAW && p.k_iters >= 8 && p.Cout <= 512 && p.a_xf != nullptr)
                return launch_tc_v<BN, KB, DUAL, false, false, false, false, true>(m, p, s);     // four ring stages, one staging tile
            static const bool s4 = !(getenv("BUSCA_S4") && getenv("BUSCA_S4")[0] == '0');
            if (s4 && p.mode == MODE_STATS) return launch_tc_v<BN, KB, DUAL, false, false, true>(m, p, s);    // no staging tiles in that mode: a fourth ring stage
        }
    }
    return launch_tc_v<BN, KB, DUAL, false, false>(m, p, s);
}

}  // namespace

namespace {
// BUSCA_HALO=1 (or busca_set_option("halo", 1)) routes the stride-1 3x3 convolutions through conv3x3_halo_kernel
// (experimental, see its header)
int g_halo = -1;
bool halo_enabled() {
    if (g_halo < 0) {
        const char *e = getenv("BUSCA_HALO");
        g_halo = !(e && e[0] == '0');            // on by default since round 2 (BUSCA_HALO=0: the tap-by-tap kernel)
    }
    return g_halo != 0;
}

template <int BN, int MT, bool RESW>
cudaError_t launch_halo_v(const CUtensorMap &ma, const CUtensorMap &mb, const CUtensorMap &mo, HaloParams p, cudaStream_t s) {
    constexpr int B_BYTES = BN * 128;
    const int fixed = 1024 + TC_XBUFS * TC_BM * 128 + (2 * 256 + 512) * 4 + (5 * HALO_MAX_STAGES + 5) * 8 + 64;
    const int budget = 232448 - fixed;
    p.n_groups = (p.total_tiles + MT - 1) / MT;
    if (RESW) {
        const int wbytes = 9 * p.cin_blocks * B_BYTES;
        p.SB = 1;
        p.SA = (budget - wbytes) / p.a_stage_bytes;
    } else {
        // halo ring: the MT boxes the MMA is reading plus MT being filled / transformed; the rest of the budget goes to weight tiles in flight
        p.SA = 2 * MT;
        p.SB = (budget - p.SA * p.a_stage_bytes) / B_BYTES;
    }
    if (p.SA > HALO_MAX_STAGES) p.SA = HALO_MAX_STAGES;
    if (p.SB > HALO_MAX_STAGES) p.SB = HALO_MAX_STAGES;
    if (p.SA < 2 * MT || p.SB < 1 || (!RESW && p.SB < 2)) return cudaErrorInvalidValue;
    const int smem = fixed + p.SA * p.a_stage_bytes + (RESW ? 9 * p.cin_blocks : p.SB) * B_BYTES;
    static int attr_smem = 0;
    if (smem > attr_smem) {
        cudaError_t e = cudaFuncSetAttribute(conv3x3_halo_kernel<BN, MT, RESW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        attr_smem = smem;
    }
    if (!g_num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int grid = p.n_groups < g_num_sms ? p.n_groups : g_num_sms;
    cudaError_t le = launch_pdl(conv3x3_halo_kernel<BN, MT, RESW>, dim3(grid), dim3(TC_THREADS), (size_t)smem, s, ma, mb, mo, p);
    if (le != cudaSuccess) return le;
    snprintf(g_last_kernel, sizeof(g_last_kernel), "conv3x3_halo_kernel<%d, %d, %d>", BN, MT, (int)RESW);
    return cudaGetLastError();
}

// true if the shape is one the halo kernel covers (and it is enabled)
bool halo_applies(const ConvLayer &L, const ConvArgs &a, const ConvTcOpts &o) {
    if (!halo_enabled() || L.k != 3 || L.stride != 1 || o.mode != TC_MODE_RAW || !a.in_xf || !L.w16s) return false;
    if (L.cin % TC_BK != 0 || L.cin > 512 || (L.cout != 64 && L.cout != 128 && L.cout != 256)) return false;
    if (a.Wo != a.W || a.Ho != a.H || (a.W != 32 && a.W != 16 && a.W != 8)) return false;
    return true;
}

// BUSCA_HALO_MT=1 forces one pixel tile per weight tile (A/B measurements)
int halo_mt() {
    static int mt = 0;
    if (!mt) {
        const char *e = getenv("BUSCA_HALO_MT");
        mt = (e && e[0] == '1') ? 1 : 2;
    }
    return mt;
}

cudaError_t launch_conv3x3_halo(const ConvLayer &L, const ConvArgs &a, cudaStream_t s) {
    HaloParams p{};
    p.W = a.W; p.H = a.H; p.P = a.W + 2;
    p.R = 0;
    for (int r = 128 / p.P; r >= 1; --r)
        if (a.H % r == 0) { p.R = r; break; }
    p.halo_rows = (p.R + 2) * p.P;
    p.valid_rows = p.R * p.W;
    p.a_stage_bytes = (((2 * p.P + 2 + 128) * 128) + 1023) & ~1023;
    if (p.R < 1 || p.halo_rows > 192 || p.halo_rows * 128 > p.a_stage_bytes || p.valid_rows > 128) return cudaErrorInvalidValue;
    p.Nimg = a.N; p.cin_blocks = L.cin / TC_BK; p.h_tiles = a.H / p.R; p.total_tiles = a.N * p.h_tiles;
    p.a_xf = a.in_xf; p.stats = L.stats; p.img_w = a.img_w;
    const long long C = L.cin, W = a.W, H = a.H;
    CUtensorMap ma, mb, mo;
    bool ok = make_map4(&ma, a.in, (int)C, a.W, a.H, a.N, C, W * C, H * W * C, p.P, p.R + 2, 1);
    ok = ok && make_map2(&mb, L.w16s, 9LL * L.cin, L.cout, L.cout);
    ok = ok && make_map4(&mo, a.out, L.cout, a.W, a.H, a.N, L.cout, W * L.cout, H * W * L.cout, a.W, p.R, 1);
    if (!ok) return cudaErrorInvalidValue;
    const bool mt2 = halo_mt() == 2;
    switch (L.cout) {
        case 256: return launch_halo_v<256, 1, false>(ma, mb, mo, p, s);      // measured (r02e): 0.112 ms vs 0.138 ms with MT = 2 (single TMEM buffer)
        case 128: return mt2 ? launch_halo_v<128, 2, false>(ma, mb, mo, p, s) : launch_halo_v<128, 1, false>(ma, mb, mo, p, s);
        default:
            if (L.cin == 64) return launch_halo_v<64, 1, true>(ma, mb, mo, p, s);         // 72 KB of weights: resident
            return launch_halo_v<64, 1, false>(ma, mb, mo, p, s);
    }
}
}  // namespace

const char *conv_tc_last_kernel() { return g_last_kernel; }
void conv_tc_set_mc_min_tiles(int n) { g_mc_min_tiles = n; }
void conv_tc_set_cg2_min_tiles(int n) { g_cg2_min_tiles = n; }
void conv_tc_set_halo(int on) { g_halo = on ? 1 : 0; }
static bool tile_geometry(int Ho, int Wo, int &BW, int &BH, int &BI);
namespace {
template <int NB>
cudaError_t launch_gram_v(const CUtensorMap &ma, const GramParams &p, int grid, cudaStream_t s) {
    constexpr int STAGES = NB == 1 ? 8 : (NB == 2 ? 5 : 3);
    const int smem = 1024 + STAGES * NB * 16384 + 4096 + (3 * STAGES + 2) * 8 + 64;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gram_stats_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    cudaError_t le = launch_pdl(gram_stats_kernel<NB>, dim3(grid), dim3(GRAM_THREADS), (size_t)smem, s, ma, p);
    if (le != cudaSuccess) return le;
    snprintf(g_last_kernel, sizeof(g_last_kernel), "gram_stats_kernel<%d>", NB);
    return cudaGetLastError();
}
}  // namespace

int gram_max_ctas() {
    if (!g_num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return g_num_sms;
}

// Gram matrix / column sums of the (transformed) input of the 1x1 convolution L over images [0, a.N), ADDED to gpart [C*C], spart [C]
// (zero them first); returns the grid size used in *grid_out.
cudaError_t launch_gram_stats(const ConvLayer &L, const ConvArgs &a, float *gpart, float *spart, int *grid_out, cudaStream_t s) {
    if (L.k != 1 || (L.cin != 64 && L.cin != 128 && L.cin != 256) || (L.stride != 1 && L.stride != 2)) return cudaErrorInvalidValue;
    GramParams p{};
    int BW;
    if (!tile_geometry(a.Ho, a.Wo, BW, p.BH, p.BI)) return cudaErrorInvalidValue;
    p.h_tiles = a.Ho / p.BH;
    p.tiles_m = ((a.N + p.BI - 1) / p.BI) * p.h_tiles;
    p.Nimg = a.N;
    p.img_shift = 0;
    while ((1 << p.img_shift) < BW * p.BH) ++p.img_shift;
    if ((1 << p.img_shift) != BW * p.BH) return cudaErrorInvalidValue;
    p.a_xf = a.in_xf; p.gpart = gpart; p.spart = spart;
    const long long C = L.cin, W = a.W, H = a.H, st = L.stride;
    CUtensorMap ma;
    if (!make_map4(&ma, a.in, (int)C, (int)(W / st), (int)(H / st), a.N, st * C, st * W * C, H * W * C, BW, p.BH, p.BI)) return cudaErrorInvalidValue;
    const int grid = p.tiles_m < gram_max_ctas() ? p.tiles_m : gram_max_ctas();
    *grid_out = grid;
    switch (L.cin) {
        case 64: return launch_gram_v<1>(ma, p, grid, s);
        case 128: return launch_gram_v<2>(ma, p, grid, s);
        default: return launch_gram_v<4>(ma, p, grid, s);
    }
}

cudaError_t launch_umma_gram_probe(const uint16_t *a_dev, int nb, float *out_dev, cudaStream_t s) {
    if (nb != 1 && nb != 2 && nb != 4) return cudaErrorInvalidValue;
    const int smem = 1024 + 4 * 16384 + 64;
    cudaError_t e = cudaFuncSetAttribute(umma_gram_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    umma_gram_probe_kernel<<<1, 128, smem, s>>>(a_dev, nb, out_dev);
    return cudaGetLastError();
}
cudaError_t launch_umma_rowshift_probe(int shift, int fill, int use_base_offset, float *out_dev, cudaStream_t s) {
    if (shift < 0 || shift + 128 > 208) return cudaErrorInvalidValue;
    const int smem = 1024 + HALO_STAGE_BYTES + 64 * 128 + 64;
    cudaError_t e = cudaFuncSetAttribute(umma_rowshift_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    umma_rowshift_probe_kernel<<<1, 128, smem, s>>>(shift, fill, use_base_offset, out_dev);
    return cudaGetLastError();
}

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

cudaError_t launch_conv_tc(const ConvLayer &L, const ConvArgs &a, const ConvTcOpts &o, cudaStream_t s) {
    if (L.cin % TC_BK != 0 || L.cout % 64 != 0 || !L.w16) return cudaErrorInvalidValue;
    if (L.stride != 1 && L.stride != 2) return cudaErrorInvalidValue;
    if (L.stride == 2 && ((a.H | a.W) & 1)) return cudaErrorInvalidValue;
    if (a.in_xf && (L.cin > 512 || !L.w16s)) return cudaErrorInvalidValue;
    if (halo_applies(L, a, o)) return launch_conv3x3_halo(L, a, s);
    const bool dual = o.mode == TC_MODE_FINAL && o.ds != nullptr;
    if (o.mode == TC_MODE_FINAL && (L.k != 1 || L.stride != 1 || !o.e_shift || !L.w16f || (!dual && !o.idt))) return cudaErrorInvalidValue;
    if (dual && (o.ds->k != 1 || o.ds->cout != L.cout || o.ds->cin % TC_BK != 0 || !o.ds->w16f || !o.ds_in)) return cudaErrorInvalidValue;
    TcParams p{};
    if (!tile_geometry(a.Ho, a.Wo, p.BW, p.BH, p.BI)) return cudaErrorInvalidValue;
    const int BN = L.cout >= 256 ? 256 : (L.cout >= 128 ? 128 : 64);
    if (dual && BN != 256) return cudaErrorInvalidValue;
    if (L.cout % BN != 0) return cudaErrorInvalidValue;
    p.tiles_n = L.cout / BN;
    p.h_tiles = a.Ho / p.BH;
    p.tiles_m = ((a.N + p.BI - 1) / p.BI) * p.h_tiles;
    p.cin_blocks = L.cin / TC_BK;
    p.ntaps = L.k * L.k;
    p.k1_iters = p.ntaps * p.cin_blocks;
    p.k_iters = p.k1_iters + (dual ? o.ds->cin / TC_BK : 0);
    p.Ho = a.Ho; p.Wo = a.Wo; p.Nimg = a.N; p.Cout = L.cout;
    p.Hv = a.H / L.stride; p.Wv = a.W / L.stride;
    p.mode = o.mode == TC_MODE_FINAL ? MODE_FINAL : (o.mode == TC_MODE_STATS ? MODE_STATS : MODE_RAW);
    p.a_xf = a.in_xf;
    p.e_shift = o.e_shift;
    p.out = a.out; p.stats = p.mode == MODE_FINAL ? nullptr : L.stats; p.bias = nullptr; p.residual = nullptr; p.alpha = 1.f; p.act = 0;
    p.img_w = a.img_w;
    p.img_shift = 0;
    while ((1 << p.img_shift) < p.BW * p.BH) ++p.img_shift;
    if ((1 << p.img_shift) != p.BW * p.BH || (a.img_w && p.BW * p.BH < 16)) return cudaErrorInvalidValue;
    TcMaps m;
    const __nv_bfloat16 *in = reinterpret_cast<const __nv_bfloat16 *>(a.in);
    const long long C = L.cin, W = a.W, H = a.H;
    bool ok = true;
    if (L.stride == 1) {
        ok = make_map4(&m.a[0], in, (int)C, a.W, a.H, a.N, C, W * C, H * W * C, p.BW, p.BH, p.BI);
        m.a[1] = m.a[2] = m.a[3] = m.a[0];
        const int pad = L.k / 2;
        for (int r = 0; r < L.k; ++r)
            for (int q = 0; q < L.k; ++q) {
                p.tap_map[r * L.k + q] = 0; p.tap_dh[r * L.k + q] = r - pad; p.tap_dw[r * L.k + q] = q - pad;
                p.tap_bit[r * L.k + q] = (r - pad + 1) * 3 + (q - pad + 1);
            }
    } else {
        // parity views: input pixel (2*ho + r - pad, 2*wo + q - pad) = view[ph][pw] at (ho + dh, wo + dw)
        for (int ph = 0; ph < 2; ++ph)
            for (int pw = 0; pw < 2; ++pw)
                ok = ok && make_map4(&m.a[ph * 2 + pw], in + (ph * W + pw) * C, (int)C, a.W / 2, a.H / 2, a.N, 2 * C, 2 * W * C, H * W * C, p.BW, p.BH, p.BI);
        const int pad = L.k / 2;
        for (int r = 0; r < L.k; ++r)
            for (int q = 0; q < L.k; ++q) {
                const int oy = r - pad, ox = q - pad;          // offset in input pixels: -1, 0, +1 (or 0 for 1x1)
                const int ph = oy & 1, pw = ox & 1;
                p.tap_map[r * L.k + q] = ph * 2 + pw;
                p.tap_dh[r * L.k + q] = (oy - ph) / 2;         // -1 -> -1, 0 -> 0, +1 -> 0
                p.tap_dw[r * L.k + q] = (ox - pw) / 2;
                p.tap_bit[r * L.k + q] = (p.tap_dh[r * L.k + q] + 1) * 3 + (p.tap_dw[r * L.k + q] + 1);
            }
    }
    // w16s = bf16(W * |scale_in|); FINAL: w16f = bf16(W * |scale_in| * scale of this conv's BN)
    // (paired variant: each CTA of a pair loads and multicasts half of a weight tile, so the box is BN / 2 rows)
    const bool resb_fit = resb_enabled() && p.k_iters <= 2 && (long long)p.k_iters * BN * 128 <= 65536;
    const int b_rows = (!resb_fit && (tc_use_mc(BN, p) || tc_use_cg2(BN, p))) ? BN / 2 : BN;
    ok = ok && make_map2(&m.b, p.mode == MODE_FINAL ? L.w16f : (a.in_xf ? L.w16s : L.w16), (long long)p.ntaps * L.cin, L.cout, b_rows);
    m.b2 = m.b;
    if (dual) {
        // downsample branch: 1x1 conv (stride ds->stride) on the block input [N, ds_H, ds_W, ds->cin]; view (0,0) for stride 2
        const long long Cd = o.ds->cin, Wd = o.ds_W, Hd = o.ds_H;
        const int sd = o.ds->stride;
        if (Hd / sd != a.Ho || Wd / sd != a.Wo) return cudaErrorInvalidValue;
        ok = ok && make_map4(&m.a[1], o.ds_in, (int)Cd, (int)(Wd / sd), (int)(Hd / sd), a.N, sd * Cd, sd * Wd * Cd, Hd * Wd * Cd, p.BW, p.BH, p.BI);
        ok = ok && make_map2(&m.b2, o.ds->w16f, Cd, L.cout, b_rows);
    }
    // output [N][Ho][Wo][Cout] bf16, stored one 64-channel group of a tile at a time; images beyond N are clipped by TMA
    ok = ok && make_map4(&m.out, a.out, L.cout, a.Wo, a.Ho, a.N, L.cout, (long long)a.Wo * L.cout, (long long)a.Ho * a.Wo * L.cout, p.BW, p.BH, p.BI);
    m.idt = m.out;
    if (p.mode == MODE_FINAL && !dual)
        ok = ok && make_map4(&m.idt, o.idt, L.cout, a.Wo, a.Ho, a.N, L.cout, (long long)a.Wo * L.cout, (long long)a.Ho * a.Wo * L.cout, p.BW, p.BH, p.BI);
    if (!ok) return cudaErrorInvalidValue;
    if (dual) return launch_tc<256, 128, true>(m, p, s);
    switch (BN) {
        case 256: return launch_tc<256>(m, p, s);
        case 128: return launch_tc<128>(m, p, s);
        default: return launch_tc<64>(m, p, s);
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
    p.k_iters = p.k1_iters = p.cin_blocks;
    p.tap_map[0] = 0; p.tap_dw[0] = 0; p.tap_dh[0] = 0;
    p.Ho = 1; p.Wo = 1; p.Nimg = la.M; p.Cout = la.N; p.Hv = 1; p.Wv = 1;
    p.mode = MODE_F32;
    p.out = la.out; p.stats = nullptr; p.bias = la.bias; p.residual = la.residual; p.alpha = la.alpha; p.act = la.act;
    TcMaps m;
    bool ok = make_map4(&m.a[0], A_bf16, la.K, 1, 1, la.M, la.K, la.K, la.K, 1, 1, 128);
    m.a[1] = m.a[2] = m.a[3] = m.a[0];
    ok = ok && make_map2(&m.b, W_bf16, la.K, la.N, BN);
    if (!ok) return cudaErrorInvalidValue;
    m.b2 = m.b;
    m.out = m.idt = m.a[0];                                      // no TMA store on the fp32-output path
    switch (BN) {
        case 256: return launch_tc<256>(m, p, s);
        case 128: return launch_tc<128>(m, p, s);
        default: return launch_tc<64>(m, p, s);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Stem on tensor cores.  The 7x7 stride-2 convolution over 3 channels becomes a GEMM with K = 4 row-PAIRS x 64:
// a pre-pass writes the normalised patch as bf16 "entries": entry e of an image holds, per pixel, the 4 channels (B,G,R,0) of input
// row 2e-3 followed by the 4 channels of row 2e-2 (16 bytes), with 3+5 zero pixels of horizontal padding per row (rows outside the
// image are zero).  For output row oy and row pair p (filter rows 2p, 2p+1) the 7x8 (+8 zero) window of output pixel ox is then 64
// CONTIGUOUS elements (128 bytes) of entry oy + p, starting 32 bytes after the window of ox-1: an overlapping-stride tensor map
// {64 el, 64 ox (stride 32 B), 196 entries, N} delivers the im2col tile directly, as ordinary 128-byte-swizzled rows.
// ---------------------------------------------------------------------------------------------------------------------
namespace {
constexpr int STEM_PITCH_PX = 136;                       // 3 zero px + 128 px + 5 zero px
constexpr int STEM_ENTRIES = PATCH_H / 2 + 4;            // 196: entry e = rows (2e-3, 2e-2), e = 0 .. 195 covers rows -3 .. 388
constexpr int STEM_EPB = 4;                              // entries per block of the pre-pass
__global__ void __launch_bounds__(256) stem_prepass_kernel(const uint8_t *__restrict__ bank, const int32_t *__restrict__ slots,
                                                           const float *__restrict__ lut, uint4 *__restrict__ out) {
    __shared__ float slut[768];
    pdl_trigger();
    pdl_wait();
    for (int i = threadIdx.x; i < 768; i += 256) slut[i] = lut[i];
    __syncthreads();
    const int n = blockIdx.y, e0 = blockIdx.x * STEM_EPB;
    const int slot = slots[n];
    const uint8_t *img = bank + (size_t)(slot < 0 ? 0 : slot) * PATCH_BYTES;
    uint4 *dst = out + ((size_t)n * STEM_ENTRIES + e0) * STEM_PITCH_PX;
    for (int i = threadIdx.x; i < STEM_EPB * STEM_PITCH_PX; i += 256) {
        const int el = i / STEM_PITCH_PX, px = i - el * STEM_PITCH_PX;
        const int x = px - 3, y0 = 2 * (e0 + el) - 3;
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        if (x >= 0 && x < PATCH_W) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int y = y0 + r;
                if (y >= 0 && y < PATCH_H) {
                    int b = 0, g = 0, rr = 0;
                    if (slot >= 0) {
                        const uint8_t *q = img + ((size_t)y * PATCH_W + x) * 3;
                        b = q[0]; g = q[1]; rr = q[2];
                    }
                    __nv_bfloat162 lo = __floats2bfloat162_rn(slut[b * 3 + 0], slut[g * 3 + 1]);
                    __nv_bfloat162 hi = __floats2bfloat162_rn(slut[rr * 3 + 2], 0.f);
                    w[2 * r] = *reinterpret_cast<uint32_t *>(&lo);
                    w[2 * r + 1] = *reinterpret_cast<uint32_t *>(&hi);
                }
            }
        }
        dst[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}
}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// Stem GEMM, row-marching formulation.
// Measured (profiles/r02u): a tap-by-tap stem (one im2col box per row pair and tile through conv_tc_kernel) is bound by the TMA REQUEST rate - every 128-pixel tile pulls four im2col boxes whose
// 128-byte rows start 32 bytes apart (each row its own, line-straddling L2 request: ~1000 requests per tile, ~3 cycles each) - and
// three of a tile's four boxes are fetched again by the tile of the next output row.  Here a CTA marches DOWN a column strip instead:
//   * M tile = ONE output row of an image PAIR (2 x 64 pixels), so the box of padded-row entry 2j (128 rows x 128 B) is the A operand
//     of output row j - p for each of the four row pairs p = 0..3: it is loaded ONCE and multiplied by the four weight slices
//     (resident in shared memory) into the four accumulators that are open at that moment.  A quarter of the requests and bytes.
//   * accumulators: a ring of eight 64-column TMEM tiles; the tile of output row oy opens at step j = oy and is complete after step
//     oy + 3.  The (up to) four tiles open at a step are ADJACENT in TMEM and the weight slices are stored in the order p = 3, 2, 1, 0,
//     so ONE tcgen05.mma with N = 256 updates all four (two instructions where the ring wraps).  Measured (profiles/r02x): an M = 128
//     MMA costs about 64 + N/2 cycles per K = 16 step - operand fetch, the A tile is re-read by every instruction - i.e. 96 cycles at
//     N = 64 but 192 at N = 256: half the tensor-pipe time of four N = 64 instructions.  Since every instruction accumulates, a tile
//     must be ZERO when it opens: the epilogue warps clear a tile (tcgen05.st) right after reading it, and all of TMEM at the start.
//   * two epilogue teams of eight warps take alternate output rows (TMEM -> bf16 -> swizzled staging -> TMA store, statistics from
//     the staging tile as in conv_tc_kernel's RAW mode), so a row's epilogue has two row-times to finish.
// Work units: (image pair, segment of 48 output rows); a segment re-reads 3 entries of its predecessor (6 %).
// ---------------------------------------------------------------------------------------------------------------------
namespace {
constexpr int SM_STAGES = 5;
constexpr int SM_SLOTS = 8;
constexpr int SM_TEAM_WARPS = 8;
constexpr int SM_THREADS = 64 + 2 * SM_TEAM_WARPS * 32;      // warp 0 TMA, warp 1 MMA, warps 2-9 team 0, warps 10-17 team 1
constexpr int SM_A_BYTES = 16384, SM_W_BYTES = 4 * 8192, SM_XBUF = 16384;
constexpr int SM_MAX_STEPS = 64;                             // seg_rows + 3 steps per unit
constexpr int SM_SMEM = 1024 + SM_STAGES * SM_A_BYTES + SM_W_BYTES + 2 * 3 * SM_XBUF + 128 * 4 + 2 * SM_MAX_STEPS * 16 + (2 * SM_STAGES + 1 + 2 * SM_SLOTS) * 8 + 64;
static_assert(SM_SMEM <= 232448, "shared memory budget");
struct StemParams {
    int N, segs, seg_rows, total_units;
    double *stats;                   // [128] sum, sum of squares
    const float *img_w;              // [N] multiplicities or null
};

__global__ void __launch_bounds__(SM_THREADS, 1) stem_march_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                                    const __grid_constant__ CUtensorMap mapOut, const StemParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *a_tiles = smem;
    uint8_t *w_tiles = a_tiles + SM_STAGES * SM_A_BYTES;                 // four [64 cout x 64 k] slices, one per row pair
    uint8_t *xbuf = w_tiles + SM_W_BYTES;                                // 2 teams x 3 staging tiles
    float *s_par = reinterpret_cast<float *>(xbuf + 6 * SM_XBUF);        // [128] statistics
    uint4 *s_tab = reinterpret_cast<uint4 *>(s_par + 128);               // [2 * steps] the MMA issuer's step table
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_tab + 2 * SM_MAX_STEPS);
    uint64_t *a_full = bars, *a_empty = a_full + SM_STAGES, *wfull = a_empty + SM_STAGES, *tfull = wfull + 1, *tempty = tfull + SM_SLOTS;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + SM_SLOTS);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    pdl_trigger();
    if (threadIdx.x == 0) {
        for (int s = 0; s < SM_STAGES; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        mbar_init(wfull, 1);
        for (int t = 0; t < SM_SLOTS; ++t) { mbar_init(&tfull[t], 1); mbar_init(&tempty[t], SM_TEAM_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        prefetch_tmap(&mapA); prefetch_tmap(&mapB); prefetch_tmap(&mapOut);
    }
    for (int i = threadIdx.x; i < 128; i += SM_THREADS) s_par[i] = 0.f;
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    pdl_wait();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    if (warp >= 2) {
        // clear all 512 columns: warp % 4 = lane quarter, the four warps of a quarter take 128 columns each
        const int part = (warp - 2) >> 2;
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_st32_zero(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(part * 128 + c * 32));
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    if (warp == 0) {
        // ===================================================== TMA producer: the weight slices once, then one entry box per step
        if (elect_one()) {
            mbar_expect_tx(wfull, SM_W_BYTES);
            for (int pr = 0; pr < 4; ++pr) tma_load_2d(w_tiles + (3 - pr) * 8192, &mapB, wfull, pr * 64, 0);      // stored in the order p = 3, 2, 1, 0
        }
        __syncwarp();
        int stage = 0;
        uint32_t phase = 0;
        for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
            const int n0 = (u / p.segs) * 2, oy0 = (u % p.segs) * p.seg_rows;
            for (int j = oy0; j < oy0 + p.seg_rows + 3; ++j) {
                mbar_wait<32>(&a_empty[stage], phase ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&a_full[stage], SM_A_BYTES);
                    tma_load_4d(a_tiles + stage * SM_A_BYTES, &mapA, &a_full[stage], 0, 0, j, n0);
                }
                __syncwarp();
                if (++stage == SM_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        // Everything a step needs is tabulated once (the issuing warp is a single instruction stream: measured ~110 dependent SASS
        // instructions and ~1200 cycles per step when p_lo / p_hi / slots / descriptors were derived per step).  seg_rows is a multiple of
        // the ring size, so the tile index of a unit's first row is = 0 (mod 8) and the table is the same for every unit: step s of a unit
        // (entry oy0 + s) updates tiles s - p_hi .. s - p_lo.
        const uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 4) << 24);      // N is filled in per instruction
        const int nsteps = p.seg_rows + 3;
        for (int sidx = lane; sidx < nsteps; sidx += 32) {
            const int p_lo = max(0, sidx - (p.seg_rows - 1)), p_hi = min(3, sidx);
            const int n = p_hi - p_lo + 1, slot_lo = (sidx - p_hi) & (SM_SLOTS - 1);
            const int n1 = min(n, SM_SLOTS - slot_lo);
            uint4 e0, e1;
            e0.x = (uint32_t)slot_lo * 64u;                                           // TMEM column of the first instruction
            e0.y = ((uint32_t)(3 - p_hi) * 8192u) >> 4;                               // weight-slice offset in descriptor units
            e0.z = idesc0 | ((uint32_t)(n1 * 64 >> 3) << 17);
            e0.w = n1 < n ? 1u : 0u;                                                  // the ring wraps: a second instruction at column 0
            e1.x = ((uint32_t)(3 - p_hi + n1) * 8192u) >> 4;
            e1.y = idesc0 | ((uint32_t)((n - n1) * 64 >> 3) << 17);
            e1.z = sidx < p.seg_rows ? (uint32_t)(sidx & (SM_SLOTS - 1)) : 0xffffffffu;    // tile that opens at this step (wait for its slot)
            e1.w = sidx >= 3 ? (uint32_t)((sidx - 3) & (SM_SLOTS - 1)) : 0xffffffffu;      // tile that is complete after this step
            s_tab[2 * sidx] = e0;
            s_tab[2 * sidx + 1] = e1;
        }
        __syncwarp();
        const uint64_t da0 = umma_desc<128>(smem_u32(a_tiles)), dw0 = umma_desc<128>(smem_u32(w_tiles));
        int stage = 0;
        uint32_t phase = 0;
        uint32_t use = 0;                                // (tile index of the unit's first row) >> 3: parity source of the slot barriers
        mbar_wait<32>(wfull, 0);
        for (int u = blockIdx.x; u < p.total_units; u += gridDim.x, use += (uint32_t)(p.seg_rows >> 3)) {
            for (int sidx = 0; sidx < nsteps; ++sidx) {
                const uint4 e0 = s_tab[2 * sidx], e1 = s_tab[2 * sidx + 1];
                if (e1.z != 0xffffffffu) mbar_wait<32>(&tempty[e1.z], ((use + (uint32_t)(sidx >> 3)) & 1u) ^ 1u);
                mbar_wait<0>(&a_full[stage], phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t da = da0 + (uint64_t)((uint32_t)stage * (SM_A_BYTES >> 4));
                if (elect_one()) {
                    const uint64_t db1 = dw0 + e0.y;
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + e0.x, da + 2 * k, db1 + 2 * k, e0.z, 1u);
                    if (e0.w) {
                        const uint64_t db2 = dw0 + e1.x;
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, da + 2 * k, db2 + 2 * k, e1.y, 1u);
                    }
                    umma_commit(&a_empty[stage]);
                    if (e1.w != 0xffffffffu) umma_commit(&tfull[e1.w]);
                }
                __syncwarp();
                if (++stage == SM_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===================================================== epilogue teams: team = parity of the CTA's running tile index
        const int team = (warp - 2) / SM_TEAM_WARPS;
        const int e = threadIdx.x - 64 - team * (SM_TEAM_WARPS * 32);      // 0..255 within the team
        const int q = warp & 3;                                            // TMEM lane quarter this warp may read
        const int half = ((warp - 2) % SM_TEAM_WARPS) >> 2;                // which 32 of the 64 channels
        const int row = q * 32 + lane;
        const uint32_t xb0 = smem_u32(xbuf) + team * 3 * SM_XBUF;
        const uint32_t row_off = (uint32_t)row * 128;
        const int cq = e & 15, rsub = e >> 4;                              // statistics: 4 channels x rows rsub + 16 i (i < 4: first image)
        const uint32_t st_off = (uint32_t)rsub * 128 + (uint32_t)((((cq >> 1) ^ (rsub & 7)) << 4) + (cq & 1) * 8);
        float acc_s[4] = {0.f, 0.f, 0.f, 0.f}, acc_q[4] = {0.f, 0.f, 0.f, 0.f};
        int t_base = 0, cnt = 0;
        for (int u = blockIdx.x; u < p.total_units; u += gridDim.x, t_base += p.seg_rows) {
            const int n0 = (u / p.segs) * 2, oy0 = (u % p.segs) * p.seg_rows;
            float w0 = 1.f, w1 = 1.f;
            if (p.img_w) { w0 = __ldg(p.img_w + n0); w1 = __ldg(p.img_w + min(n0 + 1, p.N - 1)); }
            for (int r = (t_base & 1) == team ? 0 : 1; r < p.seg_rows; r += 2, ++cnt) {
                const int t = t_base + r, slot = t & (SM_SLOTS - 1);
                const uint32_t st = xb0 + (uint32_t)(cnt % 3) * SM_XBUF;
                mbar_wait<0>(&tfull[slot], (uint32_t)(t >> 3) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t rr[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)slot * 64u + half * 32, rr);
                TMEM_LD_WAIT();
                tmem_st32_zero(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)slot * 64u + half * 32);     // every MMA accumulates: leave the tile cleared
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[slot]);                  // the accumulator is in registers: the slot may be reopened
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint4 w;
                    w.x = pack_bf16(__uint_as_float(rr[j * 8 + 0]), __uint_as_float(rr[j * 8 + 1]));
                    w.y = pack_bf16(__uint_as_float(rr[j * 8 + 2]), __uint_as_float(rr[j * 8 + 3]));
                    w.z = pack_bf16(__uint_as_float(rr[j * 8 + 4]), __uint_as_float(rr[j * 8 + 5]));
                    w.w = pack_bf16(__uint_as_float(rr[j * 8 + 6]), __uint_as_float(rr[j * 8 + 7]));
                    sts128(st + row_off + (uint32_t)(((half * 4 + j) ^ (row & 7)) << 4), w);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                if (e == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the store two tiles back has released the tile the next one overwrites
                asm volatile("bar.sync %0, 256;" ::"r"(1 + team) : "memory");
                if (e == 0) tma_store_4d(xbuf + (team * 3 + cnt % 3) * SM_XBUF, &mapOut, 0, 0, oy0 + r, n0);
                if (p.stats) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const uint2 v = lds64(st + st_off + (uint32_t)i * 2048);
                        const float wi = i < 4 ? w0 : w1;
                        const float x0 = bf_lo(v.x), x1 = bf_hi(v.x), x2 = bf_lo(v.y), x3 = bf_hi(v.y);
                        const float y0 = x0 * wi, y1 = x1 * wi, y2 = x2 * wi, y3 = x3 * wi;
                        acc_s[0] += y0; acc_q[0] = fmaf(y0, x0, acc_q[0]);
                        acc_s[1] += y1; acc_q[1] = fmaf(y1, x1, acc_q[1]);
                        acc_s[2] += y2; acc_q[2] = fmaf(y2, x2, acc_q[2]);
                        acc_s[3] += y3; acc_q[3] = fmaf(y3, x3, acc_q[3]);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            atomicAdd(&s_par[cq * 4 + j], acc_s[j]);
            atomicAdd(&s_par[64 + cq * 4 + j], acc_q[j]);
        }
        if (e == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }

    __syncthreads();
    if (p.stats)
        for (int i = threadIdx.x; i < 128; i += SM_THREADS) {
            const float v = s_par[i];
            if (v != 0.f) atomicAdd(p.stats + i, (double)v);
        }
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}
}  // namespace

size_t stem_tc_scratch_bytes(int N) { return (size_t)N * STEM_ENTRIES * STEM_PITCH_PX * 16; }

// wstem: bf16 [64][4 row pairs][64], element (p, kx*8 + r*4 + c_bgr) = W[o][c][2p+r][kx]; scratch: stem_tc_scratch_bytes(N); out: bf16 [N,192,64,64]
cudaError_t launch_stem_tc(const uint8_t *bank, const int32_t *slots, int N, const float *lut, const void *wstem, void *scratch, void *out,
                           double *stats, const float *img_w, cudaStream_t s) {
    if (N <= 0) return cudaSuccess;
    static_assert(STEM_ENTRIES % STEM_EPB == 0, "pre-pass blocks");
    cudaError_t e = launch_pdl(stem_prepass_kernel, dim3(STEM_ENTRIES / STEM_EPB, N), dim3(256), 0, s, bank, slots, lut, (uint4 *)scratch);
    if (e != cudaSuccess) return e;
    {
        const __nv_bfloat16 *in = reinterpret_cast<const __nv_bfloat16 *>(scratch);
        const long long pitch = (long long)STEM_PITCH_PX * 8;
        CUtensorMap ma, mb, mo;
        // dims {64 window elements, 64 ox (stride 16 el = 32 B), 196 entries, N}; box = one entry of an image pair
        bool ok = make_map4(&ma, in, 64, 64, STEM_ENTRIES, N, 16, pitch, (long long)STEM_ENTRIES * pitch, 64, 1, 2);
        ok = ok && make_map2(&mb, wstem, 4 * 64, 64, 64);
        ok = ok && make_map4(&mo, out, 64, 64, 192, N, 64, 64 * 64, 192LL * 64 * 64, 64, 1, 2);
        if (!ok) return cudaErrorInvalidValue;
        StemParams sp{};
        sp.N = N; sp.segs = 4; sp.seg_rows = 192 / sp.segs; sp.total_units = ((N + 1) / 2) * sp.segs;
        static_assert(192 / 4 % SM_SLOTS == 0 && 192 / 4 + 3 <= SM_MAX_STEPS, "the step table assumes segments of a multiple of the ring size");
        sp.stats = stats; sp.img_w = img_w;
        static bool attr_set = false;
        if (!attr_set) {
            e = cudaFuncSetAttribute(stem_march_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_SMEM);
            if (e != cudaSuccess) return e;
            attr_set = true;
        }
        if (!g_num_sms) {
            int dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        }
        const int grid = sp.total_units < g_num_sms ? sp.total_units : g_num_sms;
        e = launch_pdl(stem_march_kernel, dim3(grid), dim3(SM_THREADS), (size_t)SM_SMEM, s, ma, mb, mo, sp);
        snprintf(g_last_kernel, sizeof(g_last_kernel), "stem_march_kernel");
        return e;
    }
}
