/*
 * busca_b200.h - C ABI of libbusca_b200.so: BUSCA's per-frame association hot path on B200 (sm_100a).
 *
 * The reference (lorenzovaquero/BUSCA @ 88a9ed7e) has no FFI of its own: its boundary is the Python
 * object protocol of busca.network.BUSCA.  Each entry point below names the
 * reference interface it replaces (file:line under /root/reference).  The Python mirror in busca_b200/
 * binds these with ctypes; INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions: plain pointers and sizes, caller-owned buffers, int return code (0 = OK, <0 = error,
 * text via busca_last_error).  One context per tracker; a context is NOT thread-safe; different
 * contexts (and devices) are independent.  Unless a name ends in _dev every pointer is HOST memory and
 * the call returns after its results are on the host.
 */
#ifndef BUSCA_B200_H
#define BUSCA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BUSCA_PATCH_H 384
#define BUSCA_PATCH_W 128
#define BUSCA_PATCH_BYTES (384 * 128 * 3)
#define BUSCA_EMB_DIM 512

typedef struct busca_ctx busca_ctx;

enum { BUSCA_OK = 0, BUSCA_ERR_ARG = -1, BUSCA_ERR_CUDA = -2, BUSCA_ERR_STATE = -3, BUSCA_ERR_NOMEM = -4 };
enum { BUSCA_F32 = 0, BUSCA_F16 = 1, BUSCA_I64 = 2 };                 /* busca_load_tensor dtype */
enum { BUSCA_PREC_FP32 = 0, BUSCA_PREC_BF16 = 1 };                    /* arithmetic of the conv / GEMM path */
enum { BUSCA_ACT_RELU = 0, BUSCA_ACT_GELU = 1 };

/* Mirrors the YAML `transformer:` block read by BUSCA.__init__/build (busca/network.py:12-96). */
typedef struct busca_config {
    int32_t device;          /* CUDA ordinal */
    int32_t d_model;         /* trans_dim == dim_embedding == 512 */
    int32_t nhead;           /* 4 */
    int32_t ff_size;         /* 1024 */
    int32_t num_layers;      /* 4 */
    int32_t activation;      /* what the reference EXECUTES: BUSCA_ACT_RELU (parity traps) */
    int32_t precision;       /* BUSCA_PREC_* */
    int32_t sentinel_fp64;   /* 1: numpy-1.23.5 behaviour of missing_candidate_bbox */
    int64_t bank_slots;      /* initial patch-bank capacity (147,456 B each); grows on demand */
} busca_config;

const char *busca_version(void);
const char *busca_last_error(void);

/* BUSCA(args) + .to(device)            busca/network.py:11-101; byte_tracker.py:217-221 */
int busca_create(const busca_config *cfg, busca_ctx **out);
void busca_destroy(busca_ctx *ctx);

/* load_pretrained                      busca/network.py:432-467; load_trained_net.py:44-62
 * One call per state-dict entry, key names as in model_busca.pth ("encoder.weight",
 * "reid_encoder.model.layer1.0.conv1.weight", ...).  Unknown / ignored keys return BUSCA_OK.
 * Extra names: "pe.tab_xy" [211,172] f16, "pe.tab_size" [211,172] f16, "pe.tab_t" [61,168] f16
 * (the separable factors of encodings.py:23-32) and "norm.lut" [256,3] f32 BGR (network.py:470-478). */
int busca_load_tensor(busca_ctx *ctx, const char *name, const void *data, int32_t dtype, int32_t ndim,
                      const int64_t *shape);
/* Verifies that every tensor of the path is present and builds derived layouts. */
int busca_finalize(busca_ctx *ctx);

/* ---- frame + patch bank --------------------------------------------------------------------- */
/* the `image` argument of get_image_crops: uint8 BGR HWC, row stride in bytes      network.py:492 */
int busca_upload_frame(busca_ctx *ctx, const uint8_t *bgr, int32_t H, int32_t W, int64_t row_stride);
/* Upload only if needed: compares the pixels the given boxes read (all pixels when boxes == NULL or n_boxes > 8) with a page-locked
 * host mirror of the frame already in HBM; *uploaded = 1 if a new frame went up.  For callers that pass the same image to several
 * crop calls per frame without saying so (get_image_crops, network.py:492-507; byte_tracker.py:278-282, 468-479). */
int busca_sync_frame(busca_ctx *ctx, const uint8_t *bgr, int32_t H, int32_t W, int64_t row_stride, const double *boxes, int32_t n_boxes,
                     int32_t *uploaded);
/* Frame ingest on the device: the detector's input tensor [3,H,W] float32 (RGB, normalised) becomes the context's current uint8 BGR
 * frame: v = x*std + mean (fp32), RGB->BGR, clip [0,1], *255, truncate - bit-identical to the evaluator's torch/numpy code.
 * chw_on_device != 0: `chw` is a device pointer (the detector's own tensor: the frame never crosses PCIe).  frame_out: host [H,W,3]
 * uint8 copy for the caller (the adapters pass it on as current_frame), or NULL.       yolox/evaluators/mot_evaluator.py:198-204 */
int busca_ingest_frame(busca_ctx *ctx, const float *chw, int32_t chw_on_device, int32_t H, int32_t W, const float *rgb_mean,
                       const float *rgb_std, uint8_t *frame_out);
int busca_bank_reserve(busca_ctx *ctx, int64_t n_slots);
int64_t busca_bank_capacity(busca_ctx *ctx);
/* get_image_crops -> get_bbox_crop -> _cutout_with_pad + cv2.resize     network.py:492-507; tracking.py:62-113
 * boxes: n x (x1,y1,x2,y2) float64 in frame pixels.  Crop i is written to bank slot slots[i] and, when
 * host_out != NULL, also copied to host_out[i*BUSCA_PATCH_BYTES] (uint8 [384,128,3] BGR). */
int busca_crop(busca_ctx *ctx, const double *boxes, int32_t n, const int32_t *slots, uint8_t *host_out);
/* crops that did not come from busca_crop (an adapter's own arrays): host -> bank */
int busca_bank_upload(busca_ctx *ctx, const uint8_t *patches, int32_t n, const int32_t *slots);
int busca_bank_download(busca_ctx *ctx, const int32_t *slots, int32_t n, uint8_t *host_out);

/* ---- geometry -------------------------------------------------------------------------------- */
/* busca.tracking.center_distance (weight_size=False)                    tracking.py:23-60
 * a [na,4], b [nb,4] tlbr float64 -> out [na,nb] float64 */
int busca_center_distance(busca_ctx *ctx, const double *a, int32_t na, const double *b, int32_t nb, double *out);
/* matching.ious -> cython_bbox.bbox_overlaps (+1 convention)            matching.py:53-70 */
int busca_iou(busca_ctx *ctx, const double *a, int32_t na, const double *b, int32_t nb, double *out);
/* STrack.multi_predict (mean only) + STrack.tlwh/tlbr                   byte_tracker.py:50-61,140-161
 * mean [n,8] float64 (x,y,a,h,vx,vy,va,vh); tracked [n] (0: mean[7] is zeroed first).
 * Outputs (any may be NULL): mean_out [n,8], tlwh [n,4], tlbr [n,4]. */
int busca_motion_proposals(busca_ctx *ctx, const double *mean, const uint8_t *tracked, int32_t n, double *mean_out,
                           double *tlwh, double *tlbr);
/* BYTETracker.get_detection_coverage (the is_reliable gate of Step 3b)    byte_tracker.py:574-623, 459-465
 * tlbr [n,4] float64 = track.tlbr * track.scale of the active tracks; frame size H x W.
 * nonzero_out: pixels of the union of the filled rectangles (int()-truncated corners, both inclusive, clipped) = np.count_nonzero of the
 * reference's canvas; bbox_areas_out [n] (may be NULL): max(min(((x2-x1)/H) * ((y2-y1)/W), 1), 0) per box (the reference's H/W swap kept). */
int busca_detection_coverage(busca_ctx *ctx, const double *tlbr, int32_t n, int32_t H, int32_t W, int64_t *nonzero_out,
                             double *bbox_areas_out);
/* BYTETracker.camera_motion_compensation up to the warp matrix: BGR2GRAY of both frames + cv2.findTransformECC(template = previous,
 * input = current, eye(2,3), MOTION_EUCLIDEAN, (EPS | COUNT, iterations, eps), gaussFiltSize 5).      byte_tracker.py:626-651
 * prev_bgr / cur_bgr: host uint8 [H,W,3] with row_stride bytes per row.  prev_bgr == NULL: the current frame of the previous call (its
 * smoothed plane is kept in HBM); cur_bgr == NULL: the context's current frame (busca_upload_frame / busca_sync_frame / busca_ingest_frame).
 * warp_out [6] float32 row-major 2x3; rho_out: the correlation coefficient cv2 returns; BUSCA_ERR_STATE where cv2 raises StsNoConv.
 * Agrees with cv2 to ~1e-5 px (not bit-wise: cv2's summation order is unspecified). */
int busca_camera_motion(busca_ctx *ctx, const uint8_t *prev_bgr, const uint8_t *cur_bgr, int32_t H, int32_t W, int64_t row_stride,
                        int32_t iterations, double eps, float *warp_out, double *rho_out, int32_t *iterations_out);
/* ---- host-tracker rounds on the device ---------------------------------------------------------------
 * KalmanFilter.multi_predict with the covariance                      mot_online/kalman_filter.py:154-191; byte_tracker.py:50-61
 * mean [n,8], cov [n,8,8] float64; tracked [n] uint8 or NULL (0 = the height velocity is zeroed first). Bit-identical to numpy. */
int busca_kalman_predict(busca_ctx *ctx, const double *mean, const double *cov, const uint8_t *tracked, int32_t n, double *mean_out,
                         double *cov_out);
/* KalmanFilter.update (project, 4x4 Cholesky solve, gain) for n independent tracks   kalman_filter.py:126-152, 193-225
 * xyah [n,4] = the matched detections as (cx, cy, w/h, h). Agrees with scipy/LAPACK to ~1e-13 relative (not bit-wise). */
int busca_kalman_update(busca_ctx *ctx, const double *mean, const double *cov, const double *xyah, int32_t n, double *mean_out,
                        double *cov_out);
/* One association round: matching.iou_distance (+ matching.fuse_score when b_score != NULL) -> matching.linear_assignment
 * (lap.lapjv(cost, extend_cost=True, cost_limit))                      matching.py:39-50, 73-91, 165-180; byte_tracker.py:312-362
 * x [na]: matched column or -1; y [nb]: matched row or -1; cost_out [na,nb] may be NULL. na == 0 or nb == 0: all -1. */
int busca_match_round(busca_ctx *ctx, const double *a_tlbr, int32_t na, const double *b_tlbr, int32_t nb, const double *b_score,
                      double cost_limit, int32_t *x, int32_t *y, double *cost_out);
/* matching.linear_assignment on a caller-supplied cost matrix [n,m] float64       matching.py:39-50 */
int busca_linear_assignment(busca_ctx *ctx, const double *cost, int32_t n, int32_t m, double cost_limit, int32_t *x, int32_t *y);
/* remove_duplicate_stracks: pairs with IoU distance < thresh (0.15); age = frame_id - start_frame; the younger one is flagged, the
 * first list's on equal age.  drop_a [na], drop_b [nb] uint8.                       byte_tracker.py:685-698 */
int busca_duplicate_tracks(busca_ctx *ctx, const double *a_tlbr, const int32_t *a_age, int32_t na, const double *b_tlbr,
                           const int32_t *b_age, int32_t nb, double thresh, uint8_t *drop_a, uint8_t *drop_b);
/* All of the above plus candidate selection in ONE launch (north_star: "one vectorised kernel per frame"):
 * proposals from Kalman means, T x D centre-distance and IoU matrices, per-track top-C detections.
 * Outputs may be NULL. cand [T,C] int32: detection index, D+t for the motion proposal, -1 = missing. */
int busca_frame_geometry(busca_ctx *ctx, const double *mean, const uint8_t *tracked, int32_t T, const double *det_tlbr,
                         int32_t D, int32_t C, int32_t use_kalman, double *tlwh, double *tlbr, double *dist, double *iou,
                         int32_t *cand);
/* The same for B frames of B independent sequences in one launch (one tracker process usually owns several sequences):
 * mean [B,T,8], tracked [B,T], det_tlbr [B,D,4]; outputs [B,T,...]; cand holds per-frame indices (detection id, D + t, -1). */
int busca_frame_geometry_batch(busca_ctx *ctx, int32_t B, const double *mean, const uint8_t *tracked, int32_t T, const double *det_tlbr,
                               int32_t D, int32_t C, int32_t use_kalman, double *tlwh, double *tlbr, double *dist, double *iou,
                               int32_t *cand);

/* ---- appearance embedding -------------------------------------------------------------------- */
/* ReID_Encoder.get_features on ONE BatchNorm batch                      network.py:542-570; resnet.py:266-322
 * slots [n]: bank slots (-1 = the all-zero image, network.py:306,354).  out [n,512] float32 unit norm.
 * Normalisation, BGR->RGB and layout change (network.py:470-478, 397-398) are fused into the stem. */
int busca_reid_embed(busca_ctx *ctx, const int32_t *slots, int32_t n, float *out);

/* ---- association ----------------------------------------------------------------------------- */
typedef struct busca_assoc_args {
    int32_t T, D, L, C;            /* tracks, detections, seq_len, num_candidates */
    int32_t use_kalman;            /* len(extra_kalman_candidates) > 0 */
    const int32_t *mem_slots;      /* [T,L] bank slot per memory token; -1 = zero image */
    const double *mem_ltwh;        /* [T,L,4] tlwh_mem * scale (network.py:277); the [250,250,500,500] filler included */
    const int32_t *det_slots;      /* [D]  det.images_mem[-1] */
    const double *det_ltwh;        /* [D,4] det.tlwh_mem[-1] * det.scale (network.py:349) */
    const double *dists;           /* [T,D] the caller's dists_matrix (network.py:332) */
    const int32_t *kal_slots;      /* [T] or NULL */
    const double *kal_ltwh;        /* [T,4] new_det.tlwh * scale (network.py:370) or NULL */
    /* outputs, any may be NULL */
    float *probs;                  /* [T,C+2] softmax over [C candidates, NON, BAD]  (network.py:403) */
    float *logits;                 /* [T,C+2]                                       (network.py:244) */
    int32_t *cand;                 /* [T,C]   dets_embeddings_inds                  (network.py:327-378) */
    int32_t *pe_index;             /* [T,S,3] (xy, size, t) bins, S = L + 2(C+2)    (encodings.py:150-235) */
    float *mem_emb;                /* [T,L,512]                                     (network.py:196) */
    float *can_emb;                /* [T,C,512]                                     (network.py:197) */
    float *cand_rows;              /* [T,C+2,512] pre-decoder rows = self.logits    (network.py:225) */
    float *mem_logits;             /* [T,512] mean of MEM rows = self.mem_logits    (network.py:226-229) */
    float *input_seq;              /* [T,S,512] pos_encoder output                  (network.py:213) */
} busca_assoc_args;

/* BUSCA.associate_embeddings -> forward                                 network.py:282-429, 176-244 */
int busca_associate(busca_ctx *ctx, const busca_assoc_args *args);

/* Decision Transformer alone, from given embeddings (stage-wise parity)   network.py:203-232 */
int busca_transformer(busca_ctx *ctx, int32_t T, int32_t L, int32_t C, const float *mem_emb, const float *can_emb,
                      const double *mem_ltwh, const double *can_ltwh, float *logits, float *probs, int32_t *pe_index,
                      float *cand_rows, float *input_seq);

/* ---- device-resident step (bench `value`: inputs already in HBM) ------------------------------
 * All boxes are FRAME coordinates, i.e. track.scale == 1 (CenterTrack / StrongSORT / GHOST pass the original frame; an
 * adapter with a letter-boxed frame uses busca_crop + busca_associate, which take the scaled boxes the reference builds). */
typedef struct busca_step_args {
    int32_t T, D, L, C;
    const double *track_mean_dev;   /* [T,8]  Kalman means BEFORE prediction */
    const uint8_t *tracked_dev;     /* [T] */
    const double *det_tlbr_dev;     /* [D,4] frame coordinates */
    const int32_t *mem_slots_dev;   /* [T,L] */
    const double *mem_ltwh_dev;     /* [T,L,4] */
    const int32_t *det_slots_dev;   /* [D]  slots the D detection crops are written to */
    const int32_t *kal_slots_dev;   /* [T]  slots the T motion-proposal crops are written to */
    float busca_thresh;             /* byte_tracker.py:504-526 */
    const uint8_t *reliable_dev;    /* [T] */
    /* outputs (device) */
    float *probs_dev;               /* [T,C+2] */
    uint8_t *keep_dev;              /* [T] decision: reliable && p'[kalman slot] > busca_thresh, p' per the three fields below */
    int32_t *cand_dev;              /* [T,C] proposal table (detection id, D+t for the Kalman slot, -1 = filler) or NULL */
    int32_t select_highest;         /* select_highest_candidate (network.py:415-424): p' = one-hot of the argmax over the C+2 outputs */
    float highest_min_thresh;       /* highest_candidate_minimum_thresh; 0 = none */
    int32_t keep_highest_value;     /* keep_highest_value: the one-hot carries the maximum instead of 1.0 */
    const uint8_t *frame_dev;       /* optional: this sequence's current frame, uint8 BGR [frame_H, frame_W, 3] packed, resident in HBM
                                       (several sequences share one context); NULL = the frame of busca_upload_frame */
    int32_t frame_H, frame_W;
} busca_step_args;
/* One frame of the whole hot path with everything resident: motion proposals + geometry + crops of the
 * D detections and T proposals from the uploaded frame + ReID (2 batches) + Transformer + decision. */
int busca_frame_step_dev(busca_ctx *ctx, const busca_step_args *args);

/* ---- test hooks (tests/test_gpu_conv_tc.py): one ReID convolution on caller-provided bf16 NHWC input ---------- */
int busca_debug_conv(busca_ctx *ctx, int32_t conv_index, const uint16_t *in_bf16, int32_t N, int32_t H, int32_t W, int32_t use_tc,
                     uint16_t *out_bf16, double *stats_out /* [2*cout] or NULL */);
int busca_conv_info(busca_ctx *ctx, int32_t conv_index, int32_t *cin_cout_k_stride);
/* every mode of the tensor-core convolution: mode 0 raw output + statistics (optionally with the producer's BN+ReLU applied
 * to the input on load), 1 statistics only, 2 final: out = relu(conv*e_scale + e_shift + identity | BN(downsample conv)). */
typedef struct busca_debug_conv_args {
    int32_t conv_index, N, H, W, use_tc, mode;
    const uint16_t *in_bf16;                 /* [N,H,W,cin] */
    const float *in_scale, *in_shift;        /* [cin] or NULL */
    const float *e_scale, *e_shift;          /* mode 2: [cout] */
    const uint16_t *idt_bf16;                /* mode 2 without downsample: [N,Ho,Wo,cout] */
    int32_t ds_index, ds_H, ds_W;            /* mode 2 with downsample: conv index (or -1) and its input size */
    const uint16_t *ds_in_bf16;              /* [N,ds_H,ds_W,cin of the downsample conv] */
    const float *ds_scale, *ds_shift;        /* [cout] */
    uint16_t *out_bf16;                      /* [N,Ho,Wo,cout] or NULL */
    double *stats_out;                       /* [2*cout] or NULL */
    const float *img_w;                      /* tensor-core kernel: [N] multiplicity of every image in the batch statistics, or NULL (= 1) */
} busca_debug_conv_args;
int busca_debug_conv_ex(busca_ctx *ctx, const busca_debug_conv_args *args);
/* test hook: relu(x*scale + shift) followed by the 3x3 stride-2 max-pool (padding 1) of the ReID stem, bf16 NHWC in and out */
int busca_debug_maxpool(busca_ctx *ctx, const uint16_t *in_bf16 /* [N,H,W,C] */, int32_t N, int32_t H, int32_t W, int32_t C,
                        const float *scale, const float *shift, uint16_t *out_bf16 /* [N,H/2,W/2,C] */);
/* hardware probe (tests/probe_umma.py): one tcgen05.mma whose A descriptor starts `shift_rows` rows of 128 B inside a
 * SWIZZLE_128B tile, B = identity: out[128][64] must be A[m + shift_rows][n] (fill 0: A = row index, fill 1: A = column index) */
int busca_debug_umma_rowshift(busca_ctx *ctx, int32_t shift_rows, int32_t fill, int32_t use_base_offset, float *out /* [128*64] */);
/* hardware probe (tests/probe_gram.py): out[64nb][64nb] = A^T A of A[128][64nb] (bf16), both tcgen05 operands read MN-major from the
 * K-major-stored SWIZZLE_128B pixel tiles - the operand form of the Gram-matrix batch statistics */
int busca_debug_gram(busca_ctx *ctx, const uint16_t *a_bf16, int32_t nb, float *out);
int busca_debug_stem(busca_ctx *ctx, const int32_t *slots, int32_t N, int32_t use_tc, uint16_t *out_bf16 /* [N,192,64,64] */,
                     double *stats_out /* [128] or NULL */);

/* ---- plumbing -------------------------------------------------------------------------------- */
void *busca_dev_alloc(busca_ctx *ctx, int64_t bytes);
void busca_dev_free(busca_ctx *ctx, void *p);
/* page-locked host memory for the arrays get_image_crops hands to the caller (network.py:492-507): crops reach the
 * host by ONE direct DMA per contiguous slot run instead of a staged pageable copy per crop */
void *busca_host_alloc(busca_ctx *ctx, int64_t bytes);
void busca_host_free(busca_ctx *ctx, void *p);
int busca_memcpy_h2d(busca_ctx *ctx, void *dst_dev, const void *src, int64_t bytes);
int busca_memcpy_d2h(busca_ctx *ctx, void *dst, const void *src_dev, int64_t bytes);
int busca_sync(busca_ctx *ctx);
void *busca_stream(busca_ctx *ctx);               /* cudaStream_t the context launches on */
int64_t busca_kernel_launches(busca_ctx *ctx);    /* kernels launched by this context so far */
/* per-kernel device time of the most recent associate / step, name -> ms, as JSON (CUDA events) */
int busca_set_profiling(busca_ctx *ctx, int32_t on);
const char *busca_last_profile(busca_ctx *ctx);
/* options: "dedup" (default 1, bf16 mode): run the ReID encoder once per DISTINCT patch of a BatchNorm batch and weight
 * the batch statistics by the multiplicities - the batches the reference stacks (network.py:313-316, 383-386) repeat
 * every detection crop for each track that lists it as a candidate, and every incomplete history is the same zero image;
 * "halo" (default 1, process-wide): halo-box kernel for the stride-1 3x3 convolutions (csrc/conv_tc.cu), 0 = tap-by-tap kernel;
 * "pool_mono" (default 1, process-wide): max-pool kernel that pools before BN + ReLU (csrc/reid.cu);
 * "defer_crop_copies" (default 0): busca_crop returns once the gather and its device->host copy are enqueued; the host bytes of the
 * crops are valid after the next call on the context that waits for the stream (busca_associate, busca_sync, ...) - for callers
 * that, like the adapters between get_extra_kalman_candidates and associate_embeddings, only store the crops in between */
int busca_set_option(busca_ctx *ctx, const char *name, int64_t value);
/* counters: "reid_images_run" / "reid_images_total" (encoder images executed / images of the stacked batches), "kernel_launches" */
int64_t busca_counter(busca_ctx *ctx, const char *name);

#ifdef __cplusplus
}
#endif
#endif /* BUSCA_B200_H */
