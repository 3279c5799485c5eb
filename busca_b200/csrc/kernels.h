// Internal launch interface between api.cu and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <cuda_fp16.h>

// ---------------------------------------------------------------- geometry.cu
struct GeomParams {
    int T, D, C, use_kalman, nbatch;
    const double *mean;        // [B,T,8] or null (then trk_tlbr is used as given)
    const uint8_t *tracked;    // [B,T]   or null (= all tracked)
    const double *trk_tlbr;    // [B,T,4] used when mean == null
    const double *det_tlbr;    // [B,D,4]
    const double *dists_in;    // [B,T,D] or null (then computed from the boxes)
    double *mean_out;          // [B,T,8] | null
    double *tlwh_out;          // [B,T,4] | null
    double *tlbr_out;          // [B,T,4] | null
    double *dist_out;          // [B,T,D] | null
    double *iou_out;           // [B,T,D] | null
    int *cand_out;             // [B,T,C] | null
};
cudaError_t launch_frame_geometry(const GeomParams &p, cudaStream_t s);
// union area (pixels) of the filled, int()-truncated, inclusive rectangles of `tlbr` [n,4] on an H x W canvas, ADDED to *nonzero (zero it
// first), and the per-box relative areas [n] (may be null): BYTETracker.get_detection_coverage
cudaError_t launch_coverage(const double *tlbr, int n, int H, int W, unsigned long long *nonzero, double *areas, cudaStream_t s);
cudaError_t launch_pair_matrix(const double *a, int na, const double *b, int nb, double *out, int want_iou, cudaStream_t s);

// ---------------------------------------------------------------- rounds.cu (host-tracker rounds)
cudaError_t launch_kalman_predict(const double *mean, const double *cov, const uint8_t *tracked, int n, double *mean_out, double *cov_out,
                                  cudaStream_t s);
cudaError_t launch_kalman_update(const double *mean, const double *cov, const double *meas_xyah, int n, double *mean_out, double *cov_out,
                                 cudaStream_t s);
// cost [na,nb] = 1 - IoU (+1 convention); score [nb] (may be null): fused with the detection scores as matching.fuse_score does
cudaError_t launch_match_cost(const double *a, int na, const double *b, int nb, const double *score, double *cost, cudaStream_t s);
// optimum of lap.lapjv(cost, extend_cost=True, cost_limit=limit): x [N] column of each row or -1, y [M] row of each column or -1
cudaError_t launch_assignment(const double *cost, int N, int M, double limit, int *x, int *y, cudaStream_t s);
size_t assignment_smem_bytes(int N, int M);
cudaError_t launch_duplicates(const double *a, const int *age_a, int na, const double *b, const int *age_b, int nb, double thresh,
                              uint8_t *drop_a, uint8_t *drop_b, cudaStream_t s);

// ---------------------------------------------------------------- ecc.cu (camera-motion compensation)
struct EccState {
    float map[6];              // the 2x3 warp, row major (cv2's warpMatrix)
    double rho, last_rho;      // enhanced correlation coefficient of the last two iterations
    int iterations, max_iterations;
    int done;                  // the loop condition of cv2.findTransformECC failed (converged or out of iterations) or status != 0
    int status;                // 0 ok; 1 "the correlation is going to be minimized" (cv2 raises); 2 NaN
};
// gray (BGR2GRAY, integer-exact) + 5x5 Gaussian into `smooth`; central-difference gradients into gx / gy unless null
cudaError_t launch_ecc_prepare(const uint8_t *bgr, long long stride, int H, int W, float *scratch_rows, float *smooth, float *gx, float *gy,
                               cudaStream_t s);
void ecc_grid(int H, int W, int *gx, int *gy, int *rows_per_block);
cudaError_t launch_ecc_iteration(const float *tmpl, const float *img, const float *gx, const float *gy, int H, int W, EccState *st,
                                 double *partials, unsigned int *ticket, double eps, cudaStream_t s);

// ---------------------------------------------------------------- crop.cu
// detector tensor [3,H,W] fp32 (RGB, normalised) -> uint8 BGR HWC (mot_evaluator.py:198-204), both in device memory
cudaError_t launch_frame_ingest(const float *chw, int H, int W, const float mean[3], const float sd[3], uint8_t *bgr, cudaStream_t s);
// boxes [n,4] (x1,y1,x2,y2) fp64 device; slots [n] device; bank = patch bank base.
cudaError_t launch_crop_resize(const uint8_t *frame, int H, int W, int64_t row_stride, const double *boxes, int n,
                               const int32_t *slots, uint8_t *bank, cudaStream_t s);

struct CropSmall { double boxes[16]; int32_t slots[4]; int32_t n; };
cudaError_t launch_crop_resize_small(const uint8_t *frame, int H, int W, int64_t row_stride, const CropSmall &sm, uint8_t *bank, cudaStream_t s);

// ---------------------------------------------------------------- reid.cu
struct ConvLayer {
    int cin, cout, k, stride;
    float *w32;                 // [cout][k][k][cin] fp32 (bf16 mode: rounded to bf16 values, so SIMT and tensor-core kernels agree)
    float *w32m;                // unrounded fp32 master of the same layout (source of the folded weights)
    void *w16;                  // same layout, bf16
    void *w16s;                 // bf16(W * |scale of the input BN|): rewritten by launch_bn_fold on every call
    void *w16f;                 // last 1x1 conv of a bottleneck / downsample conv: bf16(W * |scale of the input BN| * scale of ITS OWN BN), the
                                // weights of the FINAL pass (launch_bn_fold_final)
    uint16_t *xf;               // [2*cin] transform of this conv's input: theta = -shift/|scale| (bf16), then sign masks
    float *gamma, *beta;        // BatchNorm affine
    double *stats;              // [2*cout] sum, sum of squares of the raw conv output (this batch)
    float *scale, *shift;       // finalised: y = x*scale + shift
};
struct ConvArgs {
    const void *in;             // NHWC activations (float or bf16)
    void *out;                  // [M, cout] raw conv output
    int N, H, W;                // input spatial size
    int Ho, Wo;
    const float *in_scale, *in_shift;   // SIMT kernel: deferred BN+ReLU of the producer applied on load (null = identity)
    const uint16_t *in_xf;              // tensor-core kernel: the same as an exact max-transform (ConvLayer::xf), weights = w16s
    const float *img_w;                 // tensor-core kernel: per-image multiplicity weighting the batch statistics (null = 1)
};
cudaError_t launch_stem(const uint8_t *bank, const int32_t *slots, int N, const float *lut, const ConvLayer &L, void *out,
                        int bf16, cudaStream_t s);
cudaError_t launch_conv_simt(const ConvLayer &L, const ConvArgs &a, int bf16, cudaStream_t s);
cudaError_t launch_bn_finalize(const ConvLayer &L, long long count, cudaStream_t s);
// Fold relu(scale*x + shift) of the producing BN ([consumer.cin] values) into the consumer conv for the tensor-core kernel:
// consumer.xf = {bf16(-shift/|scale|), sign masks}, consumer.w16s = bf16(consumer.w32m * |scale|)    (see conv_tc.cu)
cudaError_t launch_bn_fold(const float *scale, const float *shift, const ConvLayer &consumer, cudaStream_t s);
// bn_finalize of `producer` (count elements per channel) and bn_fold into `consumer` as ONE launch
cudaError_t launch_bn_finalize_fold(const ConvLayer &producer, long long count, const ConvLayer &consumer, cudaStream_t s);
// conv_tc.cu: tcgen05 / TMEM / TMA implicit GEMM on bf16 NHWC activations.  ConvArgs::in_scale/in_shift (the BN + ReLU
// of the producing conv) are applied to the A tile in shared memory.
enum { TC_MODE_RAW = 0,      // raw bf16 output + batch statistics
       TC_MODE_STATS = 1,    // batch statistics only (nothing is written)
       TC_MODE_FINAL = 2 };  // out = relu(acc*e_scale + e_shift + identity), identity = idt tensor or BN(downsample conv)
struct ConvTcOpts {
    int mode = TC_MODE_RAW;
    // FINAL: out = relu(acc + e_shift + identity).  The BN scales of this conv (and of the downsample conv) are folded into the weights
    // L.w16f / ds->w16f by launch_bn_fold_final, which also writes e_shift = shift of this BN (+ shift of the downsample BN); with `ds`
    // the downsample 1x1 conv accumulates into the SAME TMEM accumulator and replaces the identity tensor.
    const float *e_shift = nullptr;                        // [cout]
    const void *idt = nullptr;                             // identity tensor [M, cout] bf16 (no downsample)
    const ConvLayer *ds = nullptr;                         // downsample 1x1 conv accumulated by the same kernel
    const void *ds_in = nullptr;                           //   its input [N, ds_H, ds_W, ds->cin] bf16 (activated)
    int ds_H = 0, ds_W = 0;
};
// Weights and shift of a FINAL pass, and the end of the Gram-matrix statistics (one launch, reid.cu: gram_fold_final_kernel).
// Per conv (L, and ds if given):
//   statistics: L.stats holds what the weighted statistics-only pass added (or zeros); if gram_G is given the quadratic forms
//               stats[c] += Wq[c] . m, stats[Cout + c] += Wq[c]^T G Wq[c] (fp64; Wq = the bf16 weights the statistics refer to: w16s
//               for L when in_scale is given, else w16) are added and the totals written back;
//   BatchNorm:  from those totals (count > 0; the arithmetic of bn_finalize) or explicit (scale / shift, ds_scale / ds_shift);
//   weights:    L.w16f = bf16(W |in_scale| s_L), ds->w16f = bf16(W_ds s_ds) from the fp32 masters; shift_out = t_L (+ t_ds); L.scale, ds->scale.
// in_scale: scale of the BN whose ReLU output is L's input (|.| with the clamp of bn_fold; null = none).
struct FoldFinalArgs {
    const ConvLayer *L = nullptr;
    const float *in_scale = nullptr;
    long long count = 0;
    const float *scale = nullptr, *shift = nullptr;
    const float *gram_G = nullptr, *gram_m = nullptr;      // [cin*cin], [cin] fp32 (cin <= 256) or null
    const ConvLayer *ds = nullptr;
    const float *ds_scale = nullptr, *ds_shift = nullptr;
    const float *ds_gram_G = nullptr, *ds_gram_m = nullptr;
    float *shift_out = nullptr;
};
cudaError_t launch_bn_fold_final(const FoldFinalArgs &a, cudaStream_t s);
cudaError_t launch_conv_tc(const ConvLayer &L, const ConvArgs &a, const ConvTcOpts &o, cudaStream_t s);
size_t stem_tc_scratch_bytes(int N);
// Gram-matrix statistics of a 1x1 convolution (conv_tc.cu): G = sum_p a a^T and m = sum_p a by the tensor-core kernel; the quadratic
// forms  stats[c] += W[c].m,  stats[Cout + c] += W[c]^T G W[c]  (W bf16 [Cout][Cin]) are part of launch_bn_fold_final
int gram_max_ctas();
cudaError_t launch_gram_stats(const ConvLayer &L, const ConvArgs &a, float *gpart, float *spart, int *grid_out, cudaStream_t s);
cudaError_t launch_umma_gram_probe(const uint16_t *a_dev /* [128][64*nb] bf16 */, int nb, float *out_dev /* [64nb][64nb] */, cudaStream_t s);
cudaError_t launch_umma_rowshift_probe(int shift, int fill, int use_base_offset, float *out_dev /* [128*64] */, cudaStream_t s);
void conv_tc_set_halo(int on);       // experimental halo-box 3x3 kernel on/off (default: BUSCA_HALO env, off)
void conv_tc_set_mc_min_tiles(int n); // paired (weight-multicast) variant from this many pixel tiles on (-1 = default, a huge value = never)
void conv_tc_set_cg2_min_tiles(int n); // cta_group::2 variant from this many pixel tiles on (-1 = only with BUSCA_CG2=1)
const char *conv_tc_last_kernel();   // "conv_tc_kernel<BN, KB, DUAL, RESB>" of the last tensor-core launch (profiling labels)
cudaError_t launch_stem_tc(const uint8_t *bank, const int32_t *slots, int N, const float *lut, const void *wstem_bf16, void *scratch, void *out,
                           double *stats, const float *img_w, cudaStream_t s);
// x = relu(x*scale + shift) in place (bf16 or float), rows x C
cudaError_t launch_bn_relu_inplace(void *x, const float *scale, const float *shift, long long rows, int C, int bf16, cudaStream_t s);
void reid_set_pool_mono(int on);     // experimental monotone max-pool kernel on/off (default: BUSCA_POOL_MONO env, off)
cudaError_t launch_bn_relu_maxpool(const void *raw, void *out, int N, int H, int W, int C, const float *scale,
                                   const float *shift, int bf16, cudaStream_t s);
// out = relu(raw*scale+shift + (idt_scale ? idt*idt_scale+idt_shift : idt)); out may alias idt
cudaError_t launch_bn_add_relu(const void *raw, const float *scale, const float *shift, const void *idt,
                               const float *idt_scale, const float *idt_shift, void *out, long long rows, int C, int bf16,
                               cudaStream_t s);
cudaError_t launch_global_maxpool(const void *x, float *out, int N, int HW, int C, int bf16, cudaStream_t s);
cudaError_t launch_l2norm_rows(float *x, int rows, int cols, cudaStream_t s);
// distinct bank slots of one BatchNorm batch + multiplicities (see reid.cu); table: [bank_slots + 1] ints, all 0x7f7f7f7f
cudaError_t launch_dedup_slots(const int32_t *slots, int n, int *table, int32_t *uniq, int32_t *map, float *weight, int *n_uniq, cudaStream_t s);
cudaError_t launch_dedup_partition(int32_t *uniq, float *weight, int32_t *map, int n, const int *n_uniq, int32_t *tmp_u, float *tmp_w, int32_t *newpos,
                                   int *n_single, cudaStream_t s);
cudaError_t launch_gather_rows(const float *src, const int32_t *map, float *dst, int rows, int cols, cudaStream_t s);

// ---------------------------------------------------------------- linear (reid.cu): out = act((A W^T + b) * alpha) + res
struct LinearArgs {
    const float *A;             // [M,K]
    const float *W;             // [N,K]
    const float *bias;          // [N] | null
    const float *residual;      // [M,N] | null
    float *out;                 // [M,N]
    int M, N, K;
    float alpha;                // applied after bias
    int act;                    // 0 none, 1 relu, 2 gelu(erf)
};
cudaError_t launch_linear_f32(const LinearArgs &a, cudaStream_t s);
cudaError_t launch_linear_tc(const void *A_bf16, const void *W_bf16, const LinearArgs &la, cudaStream_t s);
cudaError_t launch_cast_bf16(const float *in, void *out_bf16, long long n, cudaStream_t s);
// linear_tc.cu: Decision-Transformer linears on the tensor cores with fused epilogues (bf16 A and W, fp32 accumulate)
enum { LE_F32 = 0,           // out_f32 = act((A W^T + bias) * alpha) + residual
       LE_BF16 = 1,          // out_bf16 = the same, rounded to bf16
       LE_LN = 2 };          // N == 512: y = A W^T + bias + residual; LayerNorm(y) * gamma + beta -> out_f32 and / or out_bf16 (may alias residual)
struct LinearFusedArgs {
    int epilogue = LE_F32;
    int M = 0, N = 0, K = 0;
    const float *bias = nullptr, *residual = nullptr, *gamma = nullptr, *beta = nullptr;
    float alpha = 1.f;
    int act = 0;
    float *out_f32 = nullptr;
    void *out_bf16 = nullptr;
};
cudaError_t launch_linear_fused(const void *A_bf16, const void *W_bf16, const LinearFusedArgs &a, cudaStream_t s);

// ---------------------------------------------------------------- transformer.cu
struct PeTables { const __half *xy, *size, *t; };
cudaError_t launch_assemble_candidates(const int *cand, int T, int D, int C, const double *det_ltwh, const int32_t *det_slots,
                                       const double *kal_ltwh, const int32_t *kal_slots, double *can_ltwh, int32_t *can_slots,
                                       int sentinel_fp64, cudaStream_t s);
cudaError_t launch_pe_index(const double *mem_ltwh, const double *can_ltwh, int T, int L, int C, int sentinel_fp64,
                            int32_t *idx /*[T,S,3]*/, cudaStream_t s);
cudaError_t launch_build_tokens(const float *mem_enc, const float *can_enc, const float *sep, const float *non, const float *bad,
                                const int32_t *idx, PeTables pe, int T, int L, int C, float *x, cudaStream_t s);
// out: fp32 [T*S, nhead*dh], or bf16 of the same shape when out_is_bf16 (the A operand of the tensor-core out_proj)
cudaError_t launch_attention(const float *qkv, void *out, int out_is_bf16, int T, int S, int nhead, int dh, cudaStream_t s);
cudaError_t launch_layernorm(const float *x, const float *gamma, const float *beta, float *out, int rows, int cols,
                             cudaStream_t s);
cudaError_t launch_decoder(const float *x, int T, int S, int L, int C, const float *ln_g, const float *ln_b, const float *w,
                           const float *b, float *logits, float *probs, float *cand_rows, float *mem_logits,
                           cudaStream_t s);
cudaError_t launch_decide(const float *probs, const int *cand, const uint8_t *reliable, int T, int D, int C, float thresh,
                          int select_highest, float min_thresh, int keep_value, uint8_t *keep, cudaStream_t s);
