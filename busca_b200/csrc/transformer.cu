// Decision Transformer side kernels: candidate assembly, spatio-temporal index triples, token assembly + PE add,
// short-sequence attention (warp-shuffle softmax), LayerNorm, decoder + softmax, decision.
// Reference: busca/network.py:103-165, 203-232, 324-380, 403; busca/encodings.py:43-272;
// busca/custom_layers.py:30-41; byte_tracker.py:504-526.  The dense GEMMs are in reid.cu (fp32) / conv_tc.cu (bf16).
#include "common.cuh"
#include "kernels.h"

namespace {

// ---------------------------------------------------------------------------------------------
// candidate boxes / patch slots from the index table        network.py:342-380
// ---------------------------------------------------------------------------------------------
__global__ void assemble_candidates_kernel(const int *__restrict__ cand, int T, int D, int C, const double *__restrict__ det_ltwh,
                                           const int32_t *__restrict__ det_slots, const double *__restrict__ kal_ltwh,
                                           const int32_t *__restrict__ kal_slots, double *__restrict__ can_ltwh,
                                           int32_t *__restrict__ can_slots, int sentinel_fp64) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T * C) return;
    const int t = i / C;
    const int j = cand[i];
    double b[4];
    int slot = -1;
    if (j < 0) {                                   // missing_candidate_bbox('ltwh'), tracking.py:11-12
        const double m = -3.4028234663852886e38;
        const double q = sentinel_fp64 ? (3.4028234663852886e38 / 100.0) : (double)__fdiv_rn(3.4028234663852886e38f, 100.f);
        b[0] = m; b[1] = m; b[2] = q; b[3] = q;
    } else if (j < D) {
        for (int k = 0; k < 4; ++k) b[k] = det_ltwh[4 * j + k];
        slot = det_slots ? det_slots[j] : -1;
    } else {
        for (int k = 0; k < 4; ++k) b[k] = kal_ltwh[4 * t + k];
        slot = kal_slots ? kal_slots[t] : -1;
    }
    for (int k = 0; k < 4; ++k) can_ltwh[4 * i + k] = b[k];
    can_slots[i] = slot;
}

// ---------------------------------------------------------------------------------------------
// index triples                                              encodings.py:150-180, 183-235, 239-272
// ---------------------------------------------------------------------------------------------
template <typename R> struct Rn;
template <> struct Rn<float> {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
    // correctly rounded in all but ~1e-9 of the cases; torch's CPU log (SLEEF u10) is within 1 ulp of it
    static __device__ __forceinline__ float log(float a) { return (float)::log((double)a); }
};
template <> struct Rn<double> {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double sqrt(double a) { return __dsqrt_rn(a); }
    static __device__ __forceinline__ double log(double a) { return ::log(a); }
};

template <typename R>
__device__ void spatial_bins(const R *box, const R *ref, int &xy_bin, int &size_bin) {
    using M = Rn<R>;
    const R one = (R)1, half = (R)0.5, eps = (R)1e-3;
    const R wr = M::add(M::sub(ref[2], ref[0]), one), hr = M::add(M::sub(ref[3], ref[1]), one);
    const R cxr = M::mul(half, M::add(ref[0], ref[2])), cyr = M::mul(half, M::add(ref[1], ref[3]));
    const R w = M::add(M::sub(box[2], box[0]), one), h = M::add(M::sub(box[3], box[1]), one);
    const R cx = M::mul(half, M::add(box[0], box[2])), cy = M::mul(half, M::add(box[1], box[3]));
    R dx = M::div(M::sub(cx, cxr), w), dy = M::div(M::sub(cy, cyr), h);
    dx = M::mul(dx, dx);
    dy = M::mul(dy, dy);
    const R xy = M::log(M::add(M::sqrt(M::add(dx, dy)), eps));
    const R lw = M::log(M::add(M::div(w, wr), eps)), lh = M::log(M::add(M::div(h, hr), eps));
    const R size = M::add(lw, lh);
    R a = M::mul(xy, (R)15), b = M::mul(size, (R)15);
    a = a < (R)-PE_MAX_XY ? (R)-PE_MAX_XY : (a > (R)PE_MAX_XY ? (R)PE_MAX_XY : a);
    b = b < (R)-PE_MAX_SIZE ? (R)-PE_MAX_SIZE : (b > (R)PE_MAX_SIZE ? (R)PE_MAX_SIZE : b);
    xy_bin = (int)a + PE_MAX_XY;                   // .to(torch.long): truncation toward zero
    size_bin = (int)b + PE_MAX_SIZE;
}

__device__ __forceinline__ void ltwh64_to_ltrb32(const double *b, float *o) {
    // .float() then ltwh_to_ltrb in fp32 (network.py:319, 389, 393-394, 483-489)
    const float l = (float)b[0], t = (float)b[1], w = (float)b[2], h = (float)b[3];
    o[0] = l; o[1] = t; o[2] = __fadd_rn(w, l); o[3] = __fadd_rn(h, t);
}

__global__ void pe_index_kernel(const double *__restrict__ mem_ltwh, const double *__restrict__ can_ltwh, int T, int L, int C,
                                int sentinel_fp64, int32_t *__restrict__ idx) {
    const int S = L + 2 * (C + 2);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T * S) return;
    const int t = i / S, s = i - t * S;
    float ref[4];
    ltwh64_to_ltrb32(mem_ltwh + ((size_t)t * L + (L - 1)) * 4, ref);
    int xy, sz, tb;
    if (s < L) {
        float box[4];
        ltwh64_to_ltrb32(mem_ltwh + ((size_t)t * L + s) * 4, box);
        spatial_bins<float>(box, ref, xy, sz);
        int v = 2 * (s - (L - 1));
        v = v < -PE_MAX_T ? -PE_MAX_T : (v > PE_MAX_T ? PE_MAX_T : v);
        tb = v + PE_MAX_T;
    } else {
        const int j = s - L, pair = j >> 1, second = j & 1;
        tb = (second ? 4 : 2) + PE_MAX_T;          // clamp(2*{1,2}) + 30
        // 0: reference box, 1: candidate box, 2: the distant fake box (BAD token and its SEP)
        const int kind = pair == C + 1 ? 2 : ((second && pair < C) ? 1 : 0);
        float box32[4];
        if (kind == 1) ltwh64_to_ltrb32(can_ltwh + ((size_t)t * C + pair) * 4, box32);
        else { box32[0] = ref[0]; box32[1] = ref[1]; box32[2] = ref[2]; box32[3] = ref[3]; }
        if (sentinel_fp64) {
            // numpy 1.23.5: the float64 sentinel promotes the whole candidate-side concatenation
            double box[4], r64[4] = {ref[0], ref[1], ref[2], ref[3]};
            if (kind == 2) {
                const double m = -3.4028234663852886e38;
                box[0] = m; box[1] = m; box[2] = -m / 100.0; box[3] = -m / 100.0;       // used AS IF ltrb (encodings.py:21,124)
            } else { box[0] = box32[0]; box[1] = box32[1]; box[2] = box32[2]; box[3] = box32[3]; }
            spatial_bins<double>(box, r64, xy, sz);
        } else {
            if (kind == 2) {
                const float m = -3.4028234663852886e38f, q = __fdiv_rn(3.4028234663852886e38f, 100.f);
                box32[0] = m; box32[1] = m; box32[2] = q; box32[3] = q;
            }
            spatial_bins<float>(box32, ref, xy, sz);
        }
    }
    idx[3 * (size_t)i + 0] = xy;
    idx[3 * (size_t)i + 1] = sz;
    idx[3 * (size_t)i + 2] = tb;
}

// ---------------------------------------------------------------------------------------------
// token assembly + positional encoding add               network.py:103-165; encodings.py:65-94
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) build_tokens_kernel(const float *__restrict__ mem_enc, const float *__restrict__ can_enc,
                                                           const float *__restrict__ sep, const float *__restrict__ non,
                                                           const float *__restrict__ bad, const int32_t *__restrict__ idx,
                                                           PeTables pe, int T, int L, int C, float *__restrict__ x) {
    const int S = L + 2 * (C + 2);
    const int s = blockIdx.x, t = blockIdx.y;
    const float *base;
    if (s < L) base = mem_enc + ((size_t)t * L + s) * EMB_DIM;
    else {
        const int j = s - L, pair = j >> 1;
        if (!(j & 1)) base = sep;
        else if (pair < C) base = can_enc + ((size_t)t * C + pair) * EMB_DIM;
        else base = pair == C ? non : bad;
    }
    const int32_t *tri = idx + 3 * ((size_t)t * S + s);
    const __half *px = pe.xy + (size_t)tri[0] * PE_CH, *ps = pe.size + (size_t)tri[1] * PE_CH, *pt = pe.t + (size_t)tri[2] * PE_CH_T;
    float *o = x + ((size_t)t * S + s) * EMB_DIM;
    for (int c = threadIdx.x; c < EMB_DIM; c += 128) {
        const __half e = c < PE_CH ? px[c] : (c < 2 * PE_CH ? ps[c - PE_CH] : pt[c - 2 * PE_CH]);
        o[c] = base[c] + __half2float(e);
    }
}

// ---------------------------------------------------------------------------------------------
// attention for S <= 64 tokens, head dim 128: one CTA per (head, track), one warp per query row,
// scores and softmax in registers with warp shuffles, K/V staged in shared memory.
// ---------------------------------------------------------------------------------------------
constexpr int ATT_WARPS = 4;
constexpr int DH = 128;

template <typename TOut>
__global__ void __launch_bounds__(ATT_WARPS * 32) attention_kernel(const float *__restrict__ qkv, TOut *__restrict__ out, int S,
                                                                  int nhead, float scale) {
    extern __shared__ float sm[];
    float *Ks = sm;                                // [S][DH+1]
    float *Vs = Ks + S * (DH + 1);                 // [S][DH]
    float *Qs = Vs + S * DH;                       // [ATT_WARPS][DH]
    const int h = blockIdx.x, t = blockIdx.y;
    const int D3 = 3 * nhead * DH, Dm = nhead * DH;
    const float *base = qkv + (size_t)t * S * D3;
    for (int i = threadIdx.x; i < S * DH; i += blockDim.x) {
        const int j = i / DH, d = i - j * DH;
        Ks[j * (DH + 1) + d] = base[(size_t)j * D3 + Dm + h * DH + d];
        Vs[j * DH + d] = base[(size_t)j * D3 + 2 * Dm + h * DH + d];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *q = Qs + warp * DH;
    for (int i = warp; i < S; i += ATT_WARPS) {
        for (int d = lane; d < DH; d += 32) q[d] = base[(size_t)i * D3 + h * DH + d];
        __syncwarp();
        float sc[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int j = lane + 32 * r;
            float a = -CUDART_INF_F;
            if (j < S) {
                a = 0.f;
                const float *kr = Ks + j * (DH + 1);
#pragma unroll 8
                for (int d = 0; d < DH; ++d) a = fmaf(q[d], kr[d], a);
                a *= scale;
            }
            sc[r] = a;
        }
        const float m = warp_max(fmaxf(sc[0], sc[1]));
        float e0 = expf(sc[0] - m), e1 = (lane + 32 < S) ? expf(sc[1] - m) : 0.f;
        if (lane >= S) e0 = 0.f;
        const float inv = 1.f / warp_sum(e0 + e1);
        e0 *= inv;
        e1 *= inv;
        float o[4] = {0.f, 0.f, 0.f, 0.f};
        for (int j = 0; j < S; ++j) {
            const float p = __shfl_sync(0xffffffffu, j < 32 ? e0 : e1, j & 31);
            const float *vr = Vs + j * DH;
#pragma unroll
            for (int r = 0; r < 4; ++r) o[r] = fmaf(p, vr[lane + 32 * r], o[r]);
        }
        TOut *dst = out + ((size_t)t * S + i) * Dm + h * DH;
#pragma unroll
        for (int r = 0; r < 4; ++r) ActIO<TOut>::st(dst + lane + 32 * r, o[r]);
        __syncwarp();
    }
}

// LayerNorm over 512 columns, one warp per row (biased variance, eps 1e-5); out may alias x.
__global__ void layernorm_kernel(const float *x, const float *__restrict__ g, const float *__restrict__ b, float *out, int rows) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const float *p = x + (size_t)row * EMB_DIM;
    float v[16], s = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) { v[k] = p[lane + 32 * k]; s += v[k]; }
    const float mean = warp_sum(s) * (1.f / EMB_DIM);
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) { const float d = v[k] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / EMB_DIM) + 1e-5f);
    float *o = out + (size_t)row * EMB_DIM;
#pragma unroll
    for (int k = 0; k < 16; ++k) { const int c = lane + 32 * k; o[c] = (v[k] - mean) * rstd * g[c] + b[c]; }
}

// decoder = LayerNorm(512) + Linear(512 -> 1) on the C+2 candidate rows, then softmax over them
// (network.py:93-96, 222-232, 403).  One CTA per track, one warp per candidate row.
__global__ void decoder_kernel(const float *__restrict__ x, int S, int L, int C, const float *__restrict__ g, const float *__restrict__ b,
                               const float *__restrict__ w, const float *__restrict__ bias, float *__restrict__ logits,
                               float *__restrict__ probs, float *__restrict__ cand_rows, float *__restrict__ mem_logits) {
    __shared__ float slog[32];
    const int t = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nc = C + 2;
    if (warp < nc) {
        const float *p = x + ((size_t)t * S + L + 1 + 2 * warp) * EMB_DIM;
        float v[16], s = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) { v[k] = p[lane + 32 * k]; s += v[k]; }
        if (cand_rows) {
            float *cr = cand_rows + ((size_t)t * nc + warp) * EMB_DIM;
#pragma unroll
            for (int k = 0; k < 16; ++k) cr[lane + 32 * k] = v[k];
        }
        const float mean = warp_sum(s) * (1.f / EMB_DIM);
        float q = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) { const float d = v[k] - mean; q = fmaf(d, d, q); }
        const float rstd = rsqrtf(warp_sum(q) * (1.f / EMB_DIM) + 1e-5f);
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) { const int c = lane + 32 * k; dot = fmaf((v[k] - mean) * rstd * g[c] + b[c], w[c], dot); }
        dot = warp_sum(dot) + bias[0];
        if (lane == 0) slog[warp] = dot;
    }
    if (mem_logits) {
        for (int c = threadIdx.x; c < EMB_DIM; c += blockDim.x) {
            float s = 0.f;
            for (int i = 0; i < L; ++i) s += x[((size_t)t * S + i) * EMB_DIM + c];
            mem_logits[(size_t)t * EMB_DIM + c] = s / (float)L;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = -CUDART_INF_F;
        for (int k = 0; k < nc; ++k) m = fmaxf(m, slog[k]);
        float e[32], sum = 0.f;
        for (int k = 0; k < nc; ++k) { e[k] = expf(slog[k] - m); sum += e[k]; }
        for (int k = 0; k < nc; ++k) {
            if (logits) logits[(size_t)t * nc + k] = slog[k];
            if (probs) probs[(size_t)t * nc + k] = e[k] / sum;
        }
    }
}

// decision: reliable[t] and p'[t, kalman slot] > busca_thresh        byte_tracker.py:504-526
// p' = p, or with select_highest_candidate (network.py:415-424) the one-hot of the FIRST maximum over all C+2 outputs (1.0, or the
// maximum itself with keep_highest_value), all-zero when highest_candidate_minimum_thresh > 0 and the maximum is below it
__global__ void decide_kernel(const float *__restrict__ probs, const uint8_t *__restrict__ reliable, int T, int D, int C, float thresh,
                              int select_highest, float min_thresh, int keep_value, uint8_t *__restrict__ keep) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const int kslot = min(D, C - 1);
    const float *p = probs + (size_t)t * (C + 2);
    float pk = p[kslot];
    if (select_highest) {
        int best = 0;
        for (int k = 1; k < C + 2; ++k)
            if (p[k] > p[best]) best = k;
        const bool on = min_thresh == 0.f || (min_thresh > 0.f && p[best] >= min_thresh);
        pk = (on && best == kslot) ? (keep_value ? p[best] : 1.f) : 0.f;
    }
    keep[t] = (reliable ? reliable[t] != 0 : true) && pk > thresh;
}

}  // namespace

cudaError_t launch_assemble_candidates(const int *cand, int T, int D, int C, const double *det_ltwh, const int32_t *det_slots,
                                       const double *kal_ltwh, const int32_t *kal_slots, double *can_ltwh, int32_t *can_slots,
                                       int sentinel_fp64, cudaStream_t s) {
    if (T * C <= 0) return cudaSuccess;
    assemble_candidates_kernel<<<ceil_div(T * C, 128), 128, 0, s>>>(cand, T, D, C, det_ltwh, det_slots, kal_ltwh, kal_slots, can_ltwh,
                                                                 can_slots, sentinel_fp64);
    return cudaGetLastError();
}

cudaError_t launch_pe_index(const double *mem_ltwh, const double *can_ltwh, int T, int L, int C, int sentinel_fp64, int32_t *idx,
                            cudaStream_t s) {
    const int n = T * (L + 2 * (C + 2));
    if (n <= 0) return cudaSuccess;
    pe_index_kernel<<<ceil_div(n, 128), 128, 0, s>>>(mem_ltwh, can_ltwh, T, L, C, sentinel_fp64, idx);
    return cudaGetLastError();
}

cudaError_t launch_build_tokens(const float *mem_enc, const float *can_enc, const float *sep, const float *non, const float *bad,
                                const int32_t *idx, PeTables pe, int T, int L, int C, float *x, cudaStream_t s) {
    if (T <= 0) return cudaSuccess;
    dim3 grid(L + 2 * (C + 2), T);
    build_tokens_kernel<<<grid, 128, 0, s>>>(mem_enc, can_enc, sep, non, bad, idx, pe, T, L, C, x);
    return cudaGetLastError();
}

cudaError_t launch_attention(const float *qkv, void *out, int out_is_bf16, int T, int S, int nhead, int dh, cudaStream_t s) {
    if (T <= 0) return cudaSuccess;
    if (dh != DH || S > 64) return cudaErrorInvalidValue;
    size_t smem = ((size_t)S * (DH + 1) + (size_t)S * DH + ATT_WARPS * DH) * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = out_is_bf16 ? cudaFuncSetAttribute(attention_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                                    : cudaFuncSetAttribute(attention_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    dim3 grid(nhead, T);
    if (out_is_bf16) attention_kernel<__nv_bfloat16><<<grid, ATT_WARPS * 32, smem, s>>>(qkv, (__nv_bfloat16 *)out, S, nhead, 1.f / sqrtf((float)dh));
    else attention_kernel<float><<<grid, ATT_WARPS * 32, smem, s>>>(qkv, (float *)out, S, nhead, 1.f / sqrtf((float)dh));
    return cudaGetLastError();
}

cudaError_t launch_layernorm(const float *x, const float *gamma, const float *beta, float *out, int rows, int cols, cudaStream_t s) {
    if (rows <= 0) return cudaSuccess;
    if (cols != EMB_DIM) return cudaErrorInvalidValue;
    layernorm_kernel<<<ceil_div(rows, 8), 256, 0, s>>>(x, gamma, beta, out, rows);
    return cudaGetLastError();
}

cudaError_t launch_decoder(const float *x, int T, int S, int L, int C, const float *ln_g, const float *ln_b, const float *w,
                           const float *b, float *logits, float *probs, float *cand_rows, float *mem_logits, cudaStream_t s) {
    if (T <= 0) return cudaSuccess;
    if (C + 2 > 30) return cudaErrorInvalidValue;
    decoder_kernel<<<T, 32 * (C + 2), 0, s>>>(x, S, L, C, ln_g, ln_b, w, b, logits, probs, cand_rows, mem_logits);
    return cudaGetLastError();
}

cudaError_t launch_decide(const float *probs, const int *cand, const uint8_t *reliable, int T, int D, int C, float thresh,
                          int select_highest, float min_thresh, int keep_value, uint8_t *keep, cudaStream_t s) {
    if (T <= 0) return cudaSuccess;
    (void)cand;
    decide_kernel<<<ceil_div(T, 128), 128, 0, s>>>(probs, reliable, T, D, C, thresh, select_highest, min_thresh, keep_value, keep);
    return cudaGetLastError();
}
