// Per-frame geometry: motion proposals, centre-distance and IoU matrices, top-C candidate selection.
// One launch per frame (grid = tracks x batch).  All arithmetic is IEEE fp64 with the *_rn intrinsics so
// that nothing is contracted into FMAs: results are bit-identical to numpy / scipy.cdist / cython_bbox.
//
// Reference: byte_tracker.py:50-61,140-161 (multi_predict, tlwh, tlbr); busca/tracking.py:23-60
// (center_distance); matching.py:53-70 -> cython_bbox.bbox_overlaps; busca/network.py:324-380 (selection).
#include "common.cuh"
#include "kernels.h"

namespace {

struct Box { double x1, y1, x2, y2; };

__device__ __forceinline__ double center_dist(const Box &a, const Box &b) {
    // (tlbr[:2] + tlbr[2:]) / 2.0 ; cdist 'euclidean': s = dx*dx; s += dy*dy; sqrt(s)
    double acx = __ddiv_rn(__dadd_rn(a.x1, a.x2), 2.0), acy = __ddiv_rn(__dadd_rn(a.y1, a.y2), 2.0);
    double bcx = __ddiv_rn(__dadd_rn(b.x1, b.x2), 2.0), bcy = __ddiv_rn(__dadd_rn(b.y1, b.y2), 2.0);
    double dx = __dsub_rn(acx, bcx), dy = __dsub_rn(acy, bcy);
    return __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}

__device__ __forceinline__ double box_iou(const Box &a, const Box &q) {
    // cython_bbox: +1 pixel convention; 0 unless iw > 0 and ih > 0
    double iw = __dadd_rn(__dsub_rn(fmin(a.x2, q.x2), fmax(a.x1, q.x1)), 1.0);
    if (!(iw > 0.0)) return 0.0;
    double ih = __dadd_rn(__dsub_rn(fmin(a.y2, q.y2), fmax(a.y1, q.y1)), 1.0);
    if (!(ih > 0.0)) return 0.0;
    double qa = __dmul_rn(__dadd_rn(__dsub_rn(q.x2, q.x1), 1.0), __dadd_rn(__dsub_rn(q.y2, q.y1), 1.0));
    double aa = __dmul_rn(__dadd_rn(__dsub_rn(a.x2, a.x1), 1.0), __dadd_rn(__dsub_rn(a.y2, a.y1), 1.0));
    double inter = __dmul_rn(iw, ih);
    double ua = __dsub_rn(__dadd_rn(aa, qa), inter);
    return __ddiv_rn(inter, ua);
}

__device__ __forceinline__ bool lex_less(double v1, int i1, double v2, int i2) {
    return (v1 < v2) || (v1 == v2 && i1 < i2);
}

constexpr int GEOM_THREADS = 128;

__global__ void __launch_bounds__(GEOM_THREADS) frame_geometry_kernel(GeomParams p) {
    extern __shared__ double sdist[];            // [D]
    __shared__ Box sbox;
    __shared__ double swv[GEOM_THREADS / 32];
    __shared__ int swi[GEOM_THREADS / 32];
    __shared__ int ssel;

    const int t = blockIdx.x, bz = blockIdx.y;
    const long long row = (long long)bz * p.T + t;
    const int tid = threadIdx.x;

    if (tid == 0) {
        Box b;
        if (p.mean) {
            const double *m = p.mean + row * 8;
            double v7 = (p.tracked && !p.tracked[row]) ? 0.0 : m[7];
            // mean' = mean @ F.T  ==  position += velocity (bit-equal)
            double cx = __dadd_rn(m[0], m[4]), cy = __dadd_rn(m[1], m[5]);
            double a = __dadd_rn(m[2], m[6]), h = __dadd_rn(m[3], v7);
            if (p.mean_out) {
                double *o = p.mean_out + row * 8;
                o[0] = cx; o[1] = cy; o[2] = a; o[3] = h; o[4] = m[4]; o[5] = m[5]; o[6] = m[6]; o[7] = v7;
            }
            // tlwh: ret[2] *= ret[3]; ret[:2] -= ret[2:] / 2
            double w = __dmul_rn(a, h);
            double x = __dsub_rn(cx, __ddiv_rn(w, 2.0)), y = __dsub_rn(cy, __ddiv_rn(h, 2.0));
            b.x1 = x; b.y1 = y; b.x2 = __dadd_rn(w, x); b.y2 = __dadd_rn(h, y);      // tlbr: ret[2:] += ret[:2]
            if (p.tlwh_out) { double *o = p.tlwh_out + row * 4; o[0] = x; o[1] = y; o[2] = w; o[3] = h; }
        } else {
            const double *q = p.trk_tlbr + row * 4;
            b.x1 = q[0]; b.y1 = q[1]; b.x2 = q[2]; b.y2 = q[3];
        }
        if (p.tlbr_out) { double *o = p.tlbr_out + row * 4; o[0] = b.x1; o[1] = b.y1; o[2] = b.x2; o[3] = b.y2; }
        sbox = b;
    }
    __syncthreads();
    const Box tb = sbox;
    const double *dets = p.det_tlbr + (long long)bz * p.D * 4;
    for (int d = tid; d < p.D; d += GEOM_THREADS) {
        const double2 lo = *reinterpret_cast<const double2 *>(dets + 4 * (long long)d);
        const double2 hi = *reinterpret_cast<const double2 *>(dets + 4 * (long long)d + 2);
        Box q{lo.x, lo.y, hi.x, hi.y};
        double dist = p.dists_in ? p.dists_in[row * p.D + d] : center_dist(tb, q);
        sdist[d] = dist;
        if (p.dist_out) p.dist_out[row * p.D + d] = dist;
        if (p.iou_out) p.iou_out[row * p.D + d] = box_iou(tb, q);
    }
    if (!p.cand_out) return;
    __syncthreads();

    // np.argsort(dists[t])[:C] (ties -> lower index), padded with -1; the motion proposal takes slot min(D, C-1).
    int *cand = p.cand_out + row * p.C;
    const int n_sel = min(p.D, p.C);
    for (int k = 0; k < n_sel; ++k) {
        double bv = CUDART_INF;
        int bi = 0x7fffffff;
        for (int d = tid; d < p.D; d += GEOM_THREADS) {
            double v = sdist[d];
            if (v == v && lex_less(v, d, bv, bi)) { bv = v; bi = d; }      // NaN = already taken
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (lex_less(ov, oi, bv, bi)) { bv = ov; bi = oi; }
        }
        if ((tid & 31) == 0) { swv[tid >> 5] = bv; swi[tid >> 5] = bi; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < GEOM_THREADS / 32; ++w)
                if (lex_less(swv[w], swi[w], bv, bi)) { bv = swv[w]; bi = swi[w]; }
            if (bi == 0x7fffffff) bi = -1;                                  // only +inf/NaN left
            cand[k] = bi;
            ssel = bi;
        }
        __syncthreads();
        if (tid == 0 && ssel >= 0) sdist[ssel] = __longlong_as_double(0x7ff8000000000001LL);
        __syncthreads();
    }
    if (tid == 0) {
        for (int k = n_sel; k < p.C; ++k) cand[k] = -1;
        if (p.use_kalman) cand[min(p.D, p.C - 1)] = p.D + t;
    }
}

__global__ void pair_matrix_kernel(const double *__restrict__ a, int na, const double *__restrict__ b, int nb,
                                   double *__restrict__ out, int want_iou) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)na * nb) return;
    int r = (int)(i / nb), c = (int)(i % nb);
    Box A{a[4 * r], a[4 * r + 1], a[4 * r + 2], a[4 * r + 3]};
    Box B{b[4 * c], b[4 * c + 1], b[4 * c + 2], b[4 * c + 3]};
    out[i] = want_iou ? box_iou(A, B) : center_dist(A, B);
}

}  // namespace

cudaError_t launch_frame_geometry(const GeomParams &p, cudaStream_t s) {
    if (p.T <= 0) return cudaSuccess;
    size_t smem = (size_t)(p.D > 0 ? p.D : 1) * sizeof(double);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(frame_geometry_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    dim3 grid(p.T, p.nbatch > 0 ? p.nbatch : 1);
    frame_geometry_kernel<<<grid, GEOM_THREADS, smem, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_pair_matrix(const double *a, int na, const double *b, int nb, double *out, int want_iou, cudaStream_t s) {
    long long n = (long long)na * nb;
    if (n == 0) return cudaSuccess;
    pair_matrix_kernel<<<ceil_div(n, 256), 256, 0, s>>>(a, na, b, nb, out, want_iou);
    return cudaGetLastError();
}
