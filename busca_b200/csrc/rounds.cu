// The host tracker's association rounds on the device: what BYTETracker.update does around BUSCA every frame,
// without a host round trip per matrix.
//
//   kalman_predict_kernel   KalmanFilter.multi_predict, mean AND covariance            (mot_online/kalman_filter.py:154-191)
//   kalman_update_kernel    KalmanFilter.project + update (4x4 Cholesky, gain, P - K S K^T)           (kalman_filter.py:126-152, 193-225)
//   match_cost_kernel       matching.iou_distance (+ matching.fuse_score)                            (matching.py:73-91, 165-180)
//   assignment_kernel       matching.linear_assignment = lap.lapjv(cost, extend_cost=True, cost_limit=thresh)   (matching.py:39-50)
//   duplicate_kernel        remove_duplicate_stracks                                                  (byte_tracker.py:685-698)
//
// Exactness: everything built from additions / multiplications with 0-1 matrices (the predict step, the cost matrices, the duplicate
// test) is bit-identical to numpy (IEEE *_rn operations, no FMA contraction).  The Kalman update goes through LAPACK (dpotrf / dpotrs)
// and BLAS in the reference, whose summation order is not specified: here it agrees to ~1e-13 relative, and the tests hold it to 1e-10.
// The assignment is the exact optimum of the same extended problem lapjv solves; the two can differ only when the optimum is not unique.
#include "common.cuh"
#include "kernels.h"
#include "box_math.cuh"

namespace {

// ------------------------------------------------------------------------------------------------------------------
// Kalman predict: mean' = mean F^T, P' = F P F^T + Q with F = I + shift(4): every entry of F P F^T is a sum of at most four entries of
// P, associated as numpy's two dot products associate them: left = F P first (rows i < 4: P[i][j] + P[i+4][j]), then left F^T.
// One thread per covariance entry, 64 threads per track.
// ------------------------------------------------------------------------------------------------------------------
__global__ void kalman_predict_kernel(const double *__restrict__ mean, const double *__restrict__ cov, const uint8_t *__restrict__ tracked,
                                      int n, double *__restrict__ mean_out, double *__restrict__ cov_out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int t = idx >> 6, e = idx & 63, i = e >> 3, j = e & 7;
    if (t >= n) return;
    const double *m = mean + (size_t)t * 8, *P = cov + (size_t)t * 64;
    auto left = [&](int r, int c) { return r < 4 ? __dadd_rn(P[r * 8 + c], P[(r + 4) * 8 + c]) : P[r * 8 + c]; };
    double v = j < 4 ? __dadd_rn(left(i, j), left(i, j + 4)) : left(i, j);
    if (i == j) {
        // std_pos = 1/20 * h, 1e-2 for the aspect ratio; std_vel = 1/160 * h, 1e-5; Q = diag(std^2); h = mean[3] BEFORE the step
        const double h = m[3];
        const int k = i & 3;
        double sd = (i < 4) ? ((k == 2) ? 1e-2 : __dmul_rn(1.0 / 20, h)) : ((k == 2) ? 1e-5 : __dmul_rn(1.0 / 160, h));
        v = __dadd_rn(v, __dmul_rn(sd, sd));
    }
    cov_out[(size_t)t * 64 + e] = v;
    if (e < 8) {
        // STrack.multi_predict zeroes the height velocity of tracks that are not Tracked first (byte_tracker.py:55-56)
        const double v7 = (tracked && !tracked[t]) ? 0.0 : m[7];
        double r = (e < 4) ? __dadd_rn(m[e], e == 3 ? v7 : m[e + 4]) : (e == 7 ? v7 : m[e]);
        mean_out[(size_t)t * 8 + e] = r;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Kalman update, one thread per track, everything in registers / local memory.
// ------------------------------------------------------------------------------------------------------------------
__global__ void kalman_update_kernel(const double *__restrict__ mean, const double *__restrict__ cov, const double *__restrict__ meas, int n,
                                     double *__restrict__ mean_out, double *__restrict__ cov_out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    double m[8], P[64], S[16], L[16], K[32], z[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = mean[(size_t)t * 8 + i];
#pragma unroll
    for (int i = 0; i < 64; ++i) P[i] = cov[(size_t)t * 64 + i];
#pragma unroll
    for (int i = 0; i < 4; ++i) z[i] = meas[(size_t)t * 4 + i];
    // project: S = H P H^T + R = P[:4,:4] + diag(std^2), std = (h/20, h/20, 1e-1, h/20)
    const double sp = (1.0 / 20) * m[3];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) S[i * 4 + j] = P[i * 8 + j] + ((i == j) ? ((i == 2) ? 1e-1 * 1e-1 : sp * sp) : 0.0);
    // lower Cholesky factor, column by column (the unblocked dpotf2 order; the scaling is a multiplication by 1 / L[j][j])
#pragma unroll
    for (int i = 0; i < 16; ++i) L[i] = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        double d = S[j * 4 + j];
#pragma unroll
        for (int k = 0; k < j; ++k) d -= L[j * 4 + k] * L[j * 4 + k];
        d = sqrt(d);
        L[j * 4 + j] = d;
        const double inv = 1.0 / d;
#pragma unroll
        for (int i = j + 1; i < 4; ++i) {
            double s = S[i * 4 + j];
#pragma unroll
            for (int k = 0; k < j; ++k) s -= L[i * 4 + k] * L[j * 4 + k];
            L[i * 4 + j] = s * inv;
        }
    }
    // gain: solve S X = (P H^T)^T for X [4,8]; K = X^T.  Right-hand side r: B[k][r] = P[r][k].
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        double y[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {                      // L y = b
            double s = P[r * 8 + i];
#pragma unroll
            for (int k = 0; k < i; ++k) s -= L[i * 4 + k] * y[k];
            y[i] = s / L[i * 4 + i];
        }
#pragma unroll
        for (int i = 3; i >= 0; --i) {                     // L^T x = y
            double s = y[i];
#pragma unroll
            for (int k = i + 1; k < 4; ++k) s -= L[k * 4 + i] * y[k];
            y[i] = s / L[i * 4 + i];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) K[r * 4 + i] = y[i];
    }
    // new mean = mean + (z - H mean) K^T
    double inn[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) inn[i] = z[i] - m[i];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) s += inn[k] * K[r * 4 + k];
        mean_out[(size_t)t * 8 + r] = m[r] + s;
    }
    // new covariance = P - K (S K^T)        (numpy.linalg.multi_dot associates to the right when both orders cost the same)
    double SK[32];                                          // S K^T  [4,8]
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) s += S[i * 4 + k] * K[c * 4 + k];
            SK[i * 8 + c] = s;
        }
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) s += K[r * 4 + k] * SK[k * 8 + c];
            cov_out[(size_t)t * 64 + r * 8 + c] = P[r * 8 + c] - s;
        }
}

// ------------------------------------------------------------------------------------------------------------------
// cost[i][j] = 1 - IoU(a_i, b_j); with scores: 1 - (1 - cost) * score_j  (fuse_score recomputes the similarity from the cost)
// ------------------------------------------------------------------------------------------------------------------
constexpr int RND_ROWS = 8;       // rows per thread of the pair kernels: grid (column blocks, row groups), no index division
__global__ void match_cost_kernel(const double *__restrict__ a, int na, const double *__restrict__ b, int nb, const double *__restrict__ score,
                                  double *__restrict__ cost) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nb) return;
    const Box B{b[c * 4], b[c * 4 + 1], b[c * 4 + 2], b[c * 4 + 3]};
    const double sc = score ? score[c] : 0.0;
    const int r0 = blockIdx.y * RND_ROWS, r1 = min(na, r0 + RND_ROWS);
    for (int r = r0; r < r1; ++r) {
        const Box A{__ldg(a + r * 4), __ldg(a + r * 4 + 1), __ldg(a + r * 4 + 2), __ldg(a + r * 4 + 3)};
        double v = __dsub_rn(1.0, box_iou(A, B));
        if (score) v = __dsub_rn(1.0, __dmul_rn(__dsub_rn(1.0, v), sc));
        cost[(size_t)r * nb + c] = v;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Linear assignment with a cost limit.  lap.lapjv(extend_cost=True, cost_limit=L) solves the square problem of size N + M in which
// every row / column may instead take a private 'unassigned' partner at L / 2: the optimum minimises sum over matched pairs of
// (c_ij - L).  The same optimum, stated with N rows only: row i may take column j at c_ij or its PRIVATE column M + i at L, columns
// may stay free.  Solved exactly by shortest augmenting paths with dual potentials (Hungarian / Jonker-Volgenant row insertion):
// ONE CTA, one thread per column (strided), per step one fused relax + arg-min (ties -> lowest column) and one potential update.
// Shared memory: u[N+1], v[C+1], minv[C+1] fp64; p[C+1], way[C+1] int32; used[C+1] bytes, C = M + N.
// ------------------------------------------------------------------------------------------------------------------
constexpr int ASG_THREADS = 1024;
constexpr double ASG_INF = 1e300;

__global__ void __launch_bounds__(ASG_THREADS) assignment_kernel(const double *__restrict__ cost, int N, int M, double limit,
                                                                 int *__restrict__ x, int *__restrict__ y) {
    extern __shared__ double asg_smem[];
    const int C = M + N;                                     // real columns 1..M, private columns M+1..M+N (1-based; 0 = virtual root)
    double *u = asg_smem;                                    // [N+1]
    double *v = u + (N + 1);                                 // [C+1]
    double *minv = v + (C + 1);                              // [C+1]
    int *p = (int *)(minv + (C + 1));                        // [C+1] row assigned to the column (0 = free)
    int *way = p + (C + 1);                                  // [C+1]
    unsigned char *used = (unsigned char *)(way + (C + 1));  // [C+1]
    __shared__ double red_v[ASG_THREADS / 32];
    __shared__ int red_j[ASG_THREADS / 32];
    __shared__ int s_j0, s_j1, s_done;
    __shared__ double s_delta;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;

    for (int j = tid; j <= C; j += nt) { v[j] = 0.0; p[j] = 0; }
    for (int i = tid; i <= N; i += nt) u[i] = 0.0;
    __syncthreads();

    for (int r = 1; r <= N; ++r) {
        for (int j = tid; j <= C; j += nt) { minv[j] = ASG_INF; used[j] = 0; }
        if (tid == 0) { p[0] = r; s_j0 = 0; }
        __syncthreads();
        while (true) {
            const int j0 = s_j0;
            const int i0 = p[j0];
            const double ui = u[i0];
            double best = ASG_INF;
            int bestj = 0x7fffffff;
            for (int j = tid + 1; j <= C; j += nt) {
                if (used[j] || j == j0) continue;            // j0 is marked used below (after this read phase)
                double c;
                if (j <= M) c = cost[(size_t)(i0 - 1) * M + (j - 1)];
                else c = (j - M == i0) ? limit : ASG_INF;
                double mv = minv[j];
                if (c < ASG_INF) {
                    const double cur = c - ui - v[j];
                    if (cur < mv) { mv = cur; minv[j] = cur; way[j] = j0; }
                }
                if (mv < best) { best = mv; bestj = j; }     // ascending j per thread: the first minimum is the lowest column
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int oj = __shfl_xor_sync(0xffffffffu, bestj, o);
                if (ov < best || (ov == best && oj < bestj)) { best = ov; bestj = oj; }
            }
            if (lane == 0) { red_v[warp] = best; red_j[warp] = bestj; }
            __syncthreads();
            if (warp == 0) {
                best = (lane < (nt >> 5)) ? red_v[lane] : ASG_INF;
                bestj = (lane < (nt >> 5)) ? red_j[lane] : 0x7fffffff;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double ov = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oj = __shfl_xor_sync(0xffffffffu, bestj, o);
                    if (ov < best || (ov == best && oj < bestj)) { best = ov; bestj = oj; }
                }
                if (lane == 0) { s_delta = best; s_j1 = bestj; used[j0] = 1; }
            }
            __syncthreads();
            const double delta = s_delta;
            const int j1 = s_j1;
            for (int j = tid; j <= C; j += nt) {
                if (used[j]) { u[p[j]] += delta; v[j] -= delta; }
                else minv[j] -= delta;
            }
            if (tid == 0) { s_j0 = j1; s_done = (p[j1] == 0); }
            __syncthreads();
            if (s_done) break;                               // read through shared memory: thread 0 rewrites p right after the loop
        }
        if (tid == 0) {                                      // augment: flip the path back to the root
            int j0 = s_j0;
            do {
                const int j1 = way[j0];
                p[j0] = p[j1];
                j0 = j1;
            } while (j0);
        }
        __syncthreads();
    }
    for (int i = tid; i < N; i += nt) x[i] = -1;
    for (int j = tid; j < M; j += nt) y[j] = -1;
    __syncthreads();
    for (int j = tid + 1; j <= M; j += nt)
        if (p[j]) { x[p[j] - 1] = j - 1; y[j - 1] = p[j] - 1; }
}

// ------------------------------------------------------------------------------------------------------------------
// remove_duplicate_stracks: pairs (p, q) with 1 - IoU < 0.15; the one that has been alive for less time is dropped (ties: the first list's)
// ------------------------------------------------------------------------------------------------------------------
__global__ void duplicate_kernel(const double *__restrict__ a, const int *__restrict__ age_a, int na, const double *__restrict__ b,
                                 const int *__restrict__ age_b, int nb, double thresh, uint8_t *__restrict__ drop_a, uint8_t *__restrict__ drop_b) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nb) return;
    const Box B{b[c * 4], b[c * 4 + 1], b[c * 4 + 2], b[c * 4 + 3]};
    const int gb = age_b[c];
    const int r0 = blockIdx.y * RND_ROWS, r1 = min(na, r0 + RND_ROWS);
    for (int r = r0; r < r1; ++r) {
        const Box A{__ldg(a + r * 4), __ldg(a + r * 4 + 1), __ldg(a + r * 4 + 2), __ldg(a + r * 4 + 3)};
        if (__dsub_rn(1.0, box_iou(A, B)) < thresh) {
            if (__ldg(age_a + r) > gb) drop_b[c] = 1;
            else drop_a[r] = 1;
        }
    }
}

}  // namespace

cudaError_t launch_kalman_predict(const double *mean, const double *cov, const uint8_t *tracked, int n, double *mean_out, double *cov_out,
                                  cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    kalman_predict_kernel<<<ceil_div((long long)n * 64, 256), 256, 0, s>>>(mean, cov, tracked, n, mean_out, cov_out);
    return cudaGetLastError();
}

cudaError_t launch_kalman_update(const double *mean, const double *cov, const double *meas, int n, double *mean_out, double *cov_out,
                                 cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    kalman_update_kernel<<<ceil_div(n, 64), 64, 0, s>>>(mean, cov, meas, n, mean_out, cov_out);
    return cudaGetLastError();
}

cudaError_t launch_match_cost(const double *a, int na, const double *b, int nb, const double *score, double *cost, cudaStream_t s) {
    const long long n = (long long)na * nb;
    if (n <= 0) return cudaSuccess;
    match_cost_kernel<<<dim3(ceil_div(nb, 128), ceil_div(na, RND_ROWS)), 128, 0, s>>>(a, na, b, nb, score, cost);
    return cudaGetLastError();
}

size_t assignment_smem_bytes(int N, int M) {
    const size_t C = (size_t)M + N;
    return (N + 1) * 8 + (C + 1) * 16 + (C + 1) * 8 + (C + 1) + 16;
}

cudaError_t launch_assignment(const double *cost, int N, int M, double limit, int *x, int *y, cudaStream_t s) {
    if (N <= 0 || M <= 0) return cudaErrorInvalidValue;
    const size_t smem = assignment_smem_bytes(N, M);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(assignment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        configured = 200 * 1024;
    }
    int threads = ((M + N + 31) / 32) * 32;
    if (threads > ASG_THREADS) threads = ASG_THREADS;
    if (threads < 32) threads = 32;
    assignment_kernel<<<1, threads, smem, s>>>(cost, N, M, limit, x, y);
    return cudaGetLastError();
}

cudaError_t launch_duplicates(const double *a, const int *age_a, int na, const double *b, const int *age_b, int nb, double thresh,
                              uint8_t *drop_a, uint8_t *drop_b, cudaStream_t s) {
    const long long n = (long long)na * nb;
    if (n <= 0) return cudaSuccess;
    duplicate_kernel<<<dim3(ceil_div(nb, 128), ceil_div(na, RND_ROWS)), 128, 0, s>>>(a, age_a, na, b, age_b, nb, thresh, drop_a, drop_b);
    return cudaGetLastError();
}
