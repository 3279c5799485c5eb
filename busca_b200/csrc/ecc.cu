// Camera-motion compensation on the device: BYTETracker.camera_motion_compensation
// (adapters/ByteTrack/yolox/tracker/byte_tracker.py:626-657) = cv2.cvtColor(BGR2GRAY) of the previous and the current frame +
// cv2.findTransformECC(template = previous, input = current, eye(2,3), MOTION_EUCLIDEAN, (EPS | COUNT, 100, 1e-5)).
//
// OpenCV's algorithm (modules/video/src/ecc.cpp, restated in oracle/ecc.py and pinned there against cv2 itself) re-designed around
// one pass per iteration: the reference materialises the warped image, two warped gradients, the mask, a [H, 3W] Jacobian and an
// error image every iteration and reduces them with six separate dot products.  Everything the update needs is LINEAR in a handful of
// pixel sums, so here one kernel per iteration gathers the four bilinear taps (frame planes stay L2-resident: 4 x 8 MB at 1080p),
// forms the Jacobian row in registers and accumulates 21 fp64 sums; the last CTA to finish folds the per-CTA partials in a fixed
// order and runs the scalar update (3x3 float32 Hessian inverse, lambda, delta p, new map) - no host visit inside the loop.
//
//   ecc_gray_rows_kernel   BGR u8 -> gray (15-bit integer coefficients, bit-exact) -> fp32 -> horizontal [1 4 6 4 1]/16
//   ecc_cols_grad_kernel   vertical [1 4 6 4 1]/16 (BORDER_REFLECT_101) -> smoothed plane; ecc_grad_kernel: central differences
//   ecc_iter_kernel        warp (OpenCV's 1/32-pixel fixed-point coordinates, BORDER_CONSTANT 0) + sums + update
#include "common.cuh"
#include "kernels.h"

namespace {

__device__ __forceinline__ int reflect101(int i, int n) {
    i = i < 0 ? -i : i;
    return i >= n ? 2 * (n - 1) - i : i;
}

__device__ __forceinline__ float gray_of(const uint8_t *p) {
    return (float)((p[0] * 3735 + p[1] * 19235 + p[2] * 9798 + 16384) >> 15);
}

// one thread per pixel: gray of the five horizontal neighbours (reflected), fp32 row filter in OpenCV's order
__global__ void ecc_gray_rows_kernel(const uint8_t *__restrict__ bgr, long long stride, int H, int W, float *__restrict__ rows) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const uint8_t *r = bgr + (size_t)y * stride;
    const float c = gray_of(r + 3 * x);
    const float a1 = gray_of(r + 3 * reflect101(x - 1, W)), b1 = gray_of(r + 3 * reflect101(x + 1, W));
    const float a2 = gray_of(r + 3 * reflect101(x - 2, W)), b2 = gray_of(r + 3 * reflect101(x + 2, W));
    float v = __fmul_rn(c, 0.375f);
    v = __fadd_rn(v, __fmul_rn(__fadd_rn(a1, b1), 0.25f));
    v = __fadd_rn(v, __fmul_rn(__fadd_rn(a2, b2), 0.0625f));
    rows[(size_t)y * W + x] = v;
}

__global__ void ecc_cols_kernel(const float *__restrict__ rows, int H, int W, float *__restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    auto at = [&](int yy) { return rows[(size_t)reflect101(yy, H) * W + x]; };
    float v = __fmul_rn(at(y), 0.375f);
    v = __fadd_rn(v, __fmul_rn(__fadd_rn(at(y - 1), at(y + 1)), 0.25f));
    v = __fadd_rn(v, __fmul_rn(__fadd_rn(at(y - 2), at(y + 2)), 0.0625f));
    out[(size_t)y * W + x] = v;
}

__global__ void ecc_grad_kernel(const float *__restrict__ img, int H, int W, float *__restrict__ gx, float *__restrict__ gy) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const size_t i = (size_t)y * W + x;
    gx[i] = __fmul_rn(__fsub_rn(img[(size_t)y * W + reflect101(x + 1, W)], img[(size_t)y * W + reflect101(x - 1, W)]), 0.5f);
    gy[i] = __fmul_rn(__fsub_rn(img[(size_t)reflect101(y + 1, H) * W + x], img[(size_t)reflect101(y - 1, H) * W + x]), 0.5f);
}

constexpr int ECC_NSUM = 21;      // n Sa Saa St Stt Sta | H00 H01 H02 H11 H12 H22 | Ja[3] | Jm[3] | Jmt[3]
constexpr int ECC_THREADS = 128;

__device__ __forceinline__ float tap(const float *__restrict__ p, int y, int x, int H, int W) {
    return ((unsigned)x < (unsigned)W && (unsigned)y < (unsigned)H) ? __ldg(p + (size_t)y * W + x) : 0.0f;
}

// The scalar part of one iteration (ecc.cpp / oracle.ecc.ecc_step): float32 containers where OpenCV has CV_32F matrices.
__device__ void ecc_update(const double *s, EccState *st, double eps) {
    const double n = s[0], Sa = s[1], Saa = s[2], St = s[3], Stt = s[4], Sta = s[5];
    const double mu_a = Sa / n, mu_t = St / n;
    const double var_a = Saa / n - mu_a * mu_a, var_t = Stt / n - mu_t * mu_t;
    const double img_norm = sqrt(n * var_a), tmp_norm = sqrt(n * var_t);
    const double maf = (double)(float)mu_a, mtf = (double)(float)mu_t;
    const double corr = Sta - mtf * Sa - maf * St + n * maf * mtf;
    float Hf[9] = {(float)s[6], (float)s[7], (float)s[8], (float)s[7], (float)s[9], (float)s[10], (float)s[8], (float)s[10], (float)s[11]};
    double a[9];
    for (int i = 0; i < 9; ++i) a[i] = Hf[i];
    double c[9] = {a[4] * a[8] - a[5] * a[7], a[2] * a[7] - a[1] * a[8], a[1] * a[5] - a[2] * a[4],
                   a[5] * a[6] - a[3] * a[8], a[0] * a[8] - a[2] * a[6], a[2] * a[3] - a[0] * a[5],
                   a[3] * a[7] - a[4] * a[6], a[1] * a[6] - a[0] * a[7], a[0] * a[4] - a[1] * a[3]};
    const double det = a[0] * c[0] + a[1] * c[3] + a[2] * c[6];
    float Hinv[9];
    for (int i = 0; i < 9; ++i) Hinv[i] = det != 0.0 ? (float)(c[i] / det) : 0.0f;
    float ip[3], tp[3], iph[3], ep[3], dp[3];
    for (int k = 0; k < 3; ++k) {
        ip[k] = (float)(s[12 + k] - maf * s[15 + k]);
        tp[k] = (float)(s[18 + k] - mtf * s[15 + k]);
    }
    const double rho = corr / (img_norm * tmp_norm);
    for (int k = 0; k < 3; ++k) iph[k] = (float)((double)Hinv[3 * k] * ip[0] + (double)Hinv[3 * k + 1] * ip[1] + (double)Hinv[3 * k + 2] * ip[2]);
    const double lam_n = img_norm * img_norm - ((double)ip[0] * iph[0] + (double)ip[1] * iph[1] + (double)ip[2] * iph[2]);
    const double lam_d = corr - ((double)tp[0] * iph[0] + (double)tp[1] * iph[1] + (double)tp[2] * iph[2]);
    st->last_rho = st->rho;
    st->rho = rho;
    st->iterations += 1;
    if (!(rho == rho)) { st->status = 2; st->done = 1; return; }             // NaN: cv2 raises StsNoConv
    if (lam_d <= 0.0) { st->rho = -1.0; st->status = 1; st->done = 1; return; }   // "the correlation is going to be minimized": cv2 raises
    const double lam = lam_n / lam_d;
    for (int k = 0; k < 3; ++k) ep[k] = (float)(lam * (double)tp[k] - (double)ip[k]);
    for (int k = 0; k < 3; ++k) dp[k] = (float)((double)Hinv[3 * k] * ep[0] + (double)Hinv[3 * k + 1] * ep[1] + (double)Hinv[3 * k + 2] * ep[2]);
    float *m = st->map;
    const double theta = asin((double)m[3]) + (double)dp[0];
    m[2] += dp[1];
    m[5] += dp[2];
    m[0] = m[4] = (float)cos(theta);
    m[3] = (float)sin(theta);
    m[1] = -m[3];
    // the loop condition of findTransformECC, evaluated for the NEXT iteration
    if (st->iterations >= st->max_iterations || fabs(st->rho - st->last_rho) < eps) st->done = 1;
}

// grid (ceil(W / 128), row chunks): a thread owns one column x and walks its chunk's rows.
__global__ void __launch_bounds__(ECC_THREADS) ecc_iter_kernel(const float *__restrict__ tmpl, const float *__restrict__ img,
                                                               const float *__restrict__ gx, const float *__restrict__ gy, int H, int W,
                                                               int rows_per_block, EccState *st, double *__restrict__ partials,
                                                               unsigned int *__restrict__ ticket, double eps) {
    if (st->done) return;                                   // uniform: written only by the last CTA of an earlier launch
    __shared__ double red[ECC_THREADS / 32][ECC_NSUM];
    __shared__ bool last;
    const int x = blockIdx.x * ECC_THREADS + threadIdx.x;
    const int y0 = blockIdx.y * rows_per_block, y1 = min(H, y0 + rows_per_block);
    const float mf[6] = {st->map[0], st->map[1], st->map[2], st->map[3], st->map[4], st->map[5]};
    const double m00 = mf[0], m01 = mf[1], m02 = mf[2], m10 = mf[3], m11 = mf[4], m12 = mf[5];
    double acc[ECC_NSUM];
#pragma unroll
    for (int k = 0; k < ECC_NSUM; ++k) acc[k] = 0.0;
    if (x < W) {
        // cv::warpAffine: adelta / bdelta per column, X0 / Y0 per row, 10 fractional bits, each term rounded half to even
        const long long ad = llrint(m00 * (double)x * 1024.0), bd = llrint(m10 * (double)x * 1024.0);
        const float h0 = mf[0], h1 = mf[3], xf = (float)x;
        for (int y = y0; y < y1; ++y) {
            const long long X0 = llrint((m01 * (double)y + m02) * 1024.0), Y0 = llrint((m11 * (double)y + m12) * 1024.0);
            // nearest (mask): round_delta 512, shift 10
            const long long Xn = (X0 + 512 + ad) >> 10, Yn = (Y0 + 512 + bd) >> 10;
            const bool in = Xn >= 0 && Xn < W && Yn >= 0 && Yn < H;
            // bilinear: round_delta 16, 5 fractional bits
            const long long Xl = (X0 + 16 + ad) >> 5, Yl = (Y0 + 16 + bd) >> 5;
            const int sx = (int)(Xl >> 5), sy = (int)(Yl >> 5);
            const float fx = (float)(Xl & 31) * 0.03125f, fy = (float)(Yl & 31) * 0.03125f;
            const float w00 = (1.0f - fy) * (1.0f - fx), w01 = (1.0f - fy) * fx, w10 = fy * (1.0f - fx), w11 = fy * fx;   // exact products
            float a = 0.f, gxw = 0.f, gyw = 0.f;
            if (sx >= -1 && sx < W && sy >= -1 && sy < H) {
                auto blend = [&](const float *__restrict__ p) {
                    float v = __fmul_rn(tap(p, sy, sx, H, W), w00);
                    v = __fadd_rn(v, __fmul_rn(tap(p, sy, sx + 1, H, W), w01));
                    v = __fadd_rn(v, __fmul_rn(tap(p, sy + 1, sx, H, W), w10));
                    return __fadd_rn(v, __fmul_rn(tap(p, sy + 1, sx + 1, H, W), w11));
                };
                a = blend(img); gxw = blend(gx); gyw = blend(gy);
            }
            const float t = tmpl[(size_t)y * W + x];
            const float yf = (float)y;
            const float hatX = __fsub_rn(-__fmul_rn(xf, h1), __fmul_rn(yf, h0));
            const float hatY = __fsub_rn(__fmul_rn(xf, h0), __fmul_rn(yf, h1));
            const float j0f = __fadd_rn(__fmul_rn(gxw, hatX), __fmul_rn(gyw, hatY));
            const double J0 = j0f, J1 = gxw, J2 = gyw, ad_ = a, td = t;
            if (in) {
                acc[0] += 1.0; acc[1] += ad_; acc[2] += ad_ * ad_; acc[3] += td; acc[4] += td * td; acc[5] += td * ad_;
                acc[15] += J0; acc[16] += J1; acc[17] += J2;
                acc[18] += J0 * td; acc[19] += J1 * td; acc[20] += J2 * td;
            }
            acc[6] += J0 * J0; acc[7] += J0 * J1; acc[8] += J0 * J2; acc[9] += J1 * J1; acc[10] += J1 * J2; acc[11] += J2 * J2;
            acc[12] += J0 * ad_; acc[13] += J1 * ad_; acc[14] += J2 * ad_;
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < ECC_NSUM; ++k) {
        double v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][k] = v;
    }
    __syncthreads();
    const int nblocks = gridDim.x * gridDim.y, bid = blockIdx.y * gridDim.x + blockIdx.x;
    if (threadIdx.x < ECC_NSUM) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < ECC_THREADS / 32; ++w) v += red[w][threadIdx.x];
        partials[(size_t)bid * ECC_NSUM + threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == (unsigned)nblocks - 1);
    __syncthreads();
    if (!last) return;
    __threadfence();
    __shared__ double total[ECC_NSUM];
    if (threadIdx.x < ECC_NSUM) {                            // fixed order: CTA 0, 1, 2, ... (deterministic)
        double v = 0.0;
        for (int b = 0; b < nblocks; ++b) v += __ldcg(partials + (size_t)b * ECC_NSUM + threadIdx.x);
        total[threadIdx.x] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        ecc_update(total, st, eps);
        *ticket = 0;
    }
}

}  // namespace

cudaError_t launch_ecc_prepare(const uint8_t *bgr, long long stride, int H, int W, float *scratch_rows, float *smooth, float *gx, float *gy,
                               cudaStream_t s) {
    dim3 grid(ceil_div(W, 256), H);
    ecc_gray_rows_kernel<<<grid, 256, 0, s>>>(bgr, stride, H, W, scratch_rows);
    ecc_cols_kernel<<<grid, 256, 0, s>>>(scratch_rows, H, W, smooth);
    if (gx && gy) ecc_grad_kernel<<<grid, 256, 0, s>>>(smooth, H, W, gx, gy);
    return cudaGetLastError();
}

void ecc_grid(int H, int W, int *gx, int *gy, int *rows_per_block) {
    const int bx = ceil_div(W, ECC_THREADS);
    int by = ceil_div(148 * 4, bx);
    if (by > H) by = H;
    if (by < 1) by = 1;
    const int rpb = ceil_div(H, by);
    *gx = bx; *gy = ceil_div(H, rpb); *rows_per_block = rpb;
}

cudaError_t launch_ecc_iteration(const float *tmpl, const float *img, const float *gx, const float *gy, int H, int W, EccState *st,
                                 double *partials, unsigned int *ticket, double eps, cudaStream_t s) {
    int bx, by, rpb;
    ecc_grid(H, W, &bx, &by, &rpb);
    ecc_iter_kernel<<<dim3(bx, by), ECC_THREADS, 0, s>>>(tmpl, img, gx, gy, H, W, rpb, st, partials, ticket, eps);
    return cudaGetLastError();
}
