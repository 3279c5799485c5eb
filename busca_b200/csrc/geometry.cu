// Per-frame geometry: motion proposals, centre-distance and IoU matrices, top-C candidate selection.
// One launch per frame (grid = tracks x batch).  All arithmetic is IEEE fp64 with the *_rn intrinsics so
// that nothing is contracted into FMAs: results are bit-identical to numpy / scipy.cdist / cython_bbox.
//
// Reference: byte_tracker.py:50-61,140-161 (multi_predict, tlwh, tlbr); busca/tracking.py:23-60
// (center_distance); matching.py:53-70 -> cython_bbox.bbox_overlaps; busca/network.py:324-380 (selection).
#include "common.cuh"
#include "kernels.h"
#include "box_math.cuh"

namespace {

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
            double x = __dsub_rn(cx, __dmul_rn(w, 0.5)), y = __dsub_rn(cy, __dmul_rn(h, 0.5));   // v / 2 == v * 0.5 bit for bit
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

// grid (column blocks, row groups): a thread keeps ONE column box in registers and walks PAIR_ROWS rows (row boxes are warp-uniform
// loads); no 64-bit index division, consecutive threads write consecutive doubles
constexpr int PAIR_ROWS = 8;
__global__ void __launch_bounds__(256) pair_matrix_kernel(const double *__restrict__ a, int na, const double *__restrict__ b, int nb,
                                                          double *__restrict__ out, int want_iou) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nb) return;
    const Box B{b[4 * c], b[4 * c + 1], b[4 * c + 2], b[4 * c + 3]};
    const int r0 = blockIdx.y * PAIR_ROWS, r1 = min(na, r0 + PAIR_ROWS);
    for (int r = r0; r < r1; ++r) {
        const Box A{__ldg(a + 4 * r), __ldg(a + 4 * r + 1), __ldg(a + 4 * r + 2), __ldg(a + 4 * r + 3)};
        out[(size_t)r * nb + c] = want_iou ? box_iou(A, B) : center_dist(A, B);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Detection-coverage gate of Step 3b: BYTETracker.get_detection_coverage / is_reliable
// (adapters/ByteTrack/yolox/tracker/byte_tracker.py:574-623, 459-465).  The reference draws every active
// track's box as a filled cv2.rectangle (int()-truncated corners, both inclusive, either order, clipped) on a frame-sized canvas and
// counts the non-black pixels: here one warp owns a canvas row as a bit mask (32 columns per lane and pass), ORs in the column span
// of every rectangle that covers the row and pop-counts - the union area, exactly, without a canvas in memory.  The per-box relative
// areas (fp64, one rounding per operation, the reference's width/height swap included) are written alongside.
// ------------------------------------------------------------------------------------------------------------------
constexpr int COV_WARPS = 8;
__global__ void __launch_bounds__(COV_WARPS * 32) coverage_kernel(const double *__restrict__ tlbr, int n, int H, int W, unsigned long long *__restrict__ nonzero,
                                                                  double *__restrict__ areas) {
    extern __shared__ int4 rects[];                           // clipped inclusive integer corners (xa, ya, xb, yb); xa > xb = empty
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double x1d = tlbr[4 * i], y1d = tlbr[4 * i + 1], x2d = tlbr[4 * i + 2], y2d = tlbr[4 * i + 3];
        const int x1 = (int)x1d, y1 = (int)y1d, x2 = (int)x2d, y2 = (int)y2d;            // truncation toward zero = Python's int()
        int4 r;
        r.x = max(min(x1, x2), 0); r.y = max(min(y1, y2), 0); r.z = min(max(x1, x2), W - 1); r.w = min(max(y1, y2), H - 1);
        rects[i] = r;
        if (areas && blockIdx.x == 0) {
            double v = __dmul_rn(__ddiv_rn(__dsub_rn(x2d, x1d), (double)H), __ddiv_rn(__dsub_rn(y2d, y1d), (double)W));
            v = (1.0 < v) ? 1.0 : v;                           // Python: max(min(v, 1.0), 0.0)
            v = (v < 0.0) ? 0.0 : v;
            areas[i] = v;
        }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned int cnt = 0;
    for (int y = blockIdx.x * COV_WARPS + warp; y < H; y += gridDim.x * COV_WARPS) {
        for (int w0 = 0; w0 < W; w0 += 1024) {                 // 32 lanes x 32 columns per pass
            const int c0 = w0 + lane * 32;                      // this lane's columns c0 .. c0 + 31
            unsigned int m = 0;
            for (int i = 0; i < n; ++i) {
                const int4 r = rects[i];
                if (y < r.y || y > r.w) continue;
                const int a = max(r.x, c0), b = min(r.z, c0 + 31);
                if (a <= b) m |= (0xffffffffu >> (31 - (b - a))) << (a - c0);
            }
            cnt += __popc(m);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0 && cnt) atomicAdd(nonzero, (unsigned long long)cnt);
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
    pair_matrix_kernel<<<dim3(ceil_div(nb, 256), ceil_div(na, PAIR_ROWS)), 256, 0, s>>>(a, na, b, nb, out, want_iou);
    return cudaGetLastError();
}

cudaError_t launch_coverage(const double *tlbr, int n, int H, int W, unsigned long long *nonzero, double *areas, cudaStream_t s) {
    if (H <= 0 || W <= 0 || n < 0) return cudaErrorInvalidValue;
    const size_t smem = (size_t)(n > 0 ? n : 1) * sizeof(int4);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(coverage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    int grid = (H + COV_WARPS - 1) / COV_WARPS;
    if (grid > 148 * 4) grid = 148 * 4;
    coverage_kernel<<<grid, COV_WARPS * 32, smem, s>>>(tlbr, n, H, W, nonzero, areas);
    return cudaGetLastError();
}
