/*
 * wsovod_b200.h -- C ABI of libwsovod_b200.so: the B200 (sm_100a) region-scoring kernels of WSOVOD.
 *
 * Every entry point takes plain device pointers, sizes and a CUDA stream handle (cudaStream_t passed
 * as void*).  No torch types, no allocation, no host synchronisation, no global state: calls are
 * re-entrant and stream-asynchronous.  Return value: 0 = success, negative = argument error
 * (WSOVOD_B200_E*), positive = cudaError_t reported right after the launch.
 * Scratch memory is caller-provided: ask `*_workspace()` for the size, pass a device buffer.
 *
 * Each function names the reference interface it replaces (paths relative to the WSOVOD tree).
 * All tensors are dense row-major ("contiguous"), fp32 unless stated, index types as stated.
 */
#ifndef WSOVOD_B200_H
#define WSOVOD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WSOVOD_B200_ABI_VERSION 1

/* argument errors */
#define WSOVOD_B200_EINVAL (-1)      /* null pointer / negative size / bad enum                 */
#define WSOVOD_B200_ETOOBIG (-2)     /* a dimension exceeds what the kernels index (see docs)    */
#define WSOVOD_B200_EWORKSPACE (-3)  /* workspace missing or smaller than *_workspace() says      */
#define WSOVOD_B200_EUNSUPPORTED (-4)/* valid request the sm_100a build does not implement        */
#define WSOVOD_B200_EALIGN (-5)      /* pointer not aligned as required (16 B for TMA/vector IO)  */

int wsovod_b200_abi_version(void);
/* static string for any return code of this library (cudaGetErrorString for positive codes) */
const char* wsovod_b200_strerror(int code);
/* number of kernel launches issued by this library since load (bench.py's `gpu_launches`) */
uint64_t wsovod_b200_launch_count(void);
/* Process-wide tuning / test switches: which of several bit-identical kernels runs.  Not part of the
 * reference's interface; results never depend on them.  Returns the previous value (or EINVAL). */
#define WSOVOD_B200_TUNE_POOL_PATH 0  /* 0 library's choice (default), 1 scan kernels, 2 block-max planes */
#define WSOVOD_B200_TUNE_POOL_GROUP 1 /* 1 bank-conflict-aware lane order of the block-max path (default), 0 row-major */
#define WSOVOD_B200_TUNE_ALIGN_PAIR 2 /* 1 CTA-pair (cta_group::2) contraction for K + 1 > 256 (default), 0 one CTA per tile */
int wsovod_b200_tune(int key, int value);

/* ------------------------------------------------------------------------------------------------
 * (1) ROI pooling.  rois: [R,5] = (batch_index, x1, y1, x2, y2) in image pixels, fp32.
 *     input: [N,C,H,W] NCHW.  output: [R,C,PH,PW].  `row_scale` (may be NULL) folds the reference's
 *     `box_features * (objectness_logits + 1)` (wsovod/modeling/roi_heads/roi_heads.py:733-739) into
 *     the store: out = pooled * (row_scale[r] + row_scale_bias), one fp32 add and one fp32 multiply.
 * ---------------------------------------------------------------------------------------------- */

/* replaces torch.ops.torchvision.roi_pool (selected at wsovod/modeling/poolers.py:183-186; semantics
 * restated in wsovod/layers/ROILoopPool/ROILoopPool_cpu.cpp:14-80).  argmax [R,C,PH,PW] int32 =
 * h*W+w of the first maximum in h-major scan order, -1 for an empty bin; pass NULL to skip it
 * (frozen-backbone inference never reads it). */
size_t wsovod_b200_roi_pool_workspace(int64_t N, int64_t R, int pooled_h, int pooled_w);
int wsovod_b200_roi_pool_fwd(const float* input, int64_t N, int64_t C, int64_t H, int64_t W,
                             const float* rois, int64_t R, float spatial_scale,
                             int pooled_h, int pooled_w,
                             const float* row_scale, float row_scale_bias,
                             float* output, int32_t* argmax,
                             void* workspace, size_t workspace_bytes, void* stream);
/* replaces the roi_pool backward (ROILoopPool_cpu.cpp:82-123): grad_input[b,c,argmax] += grad_output.
 * grad_input [N,C,H,W] must be zero-filled by the caller; accumulation uses fp32 atomics. */
int wsovod_b200_roi_pool_bwd(const float* grad_output, const float* rois, const int32_t* argmax,
                             int64_t R, int64_t N, int64_t C, int64_t H, int64_t W,
                             int pooled_h, int pooled_w, float* grad_input, void* stream);

/* replaces wsovod._C.roi_loop_pool_forward (wsovod/layers/vision.cpp:10,
 * wsovod/layers/ROILoopPool/ROILoopPool_cuda.cu:9-204,252-313).  output/argmax: [3R,C,PH,PW] laid out
 * roi | frame | context, maxima initialised at 0 (argmax -1), context ratio 1.8. */
size_t wsovod_b200_roi_loop_pool_workspace(int64_t N, int64_t R, int pooled_h, int pooled_w);
int wsovod_b200_roi_loop_pool_fwd(const float* input, int64_t N, int64_t C, int64_t H, int64_t W,
                                  const float* rois, int64_t R, float spatial_scale,
                                  int pooled_h, int pooled_w,
                                  const float* row_scale, float row_scale_bias,
                                  float* output, int32_t* argmax,
                                  void* workspace, size_t workspace_bytes, void* stream);
/* replaces wsovod._C.roi_loop_pool_backward (vision.cpp:11, ROILoopPool_cuda.cu:206-248,315-388):
 * grad_output/argmax are [3R,C,PH,PW]; row n maps to roi n % R. */
int wsovod_b200_roi_loop_pool_bwd(const float* grad_output, const float* rois, const int32_t* argmax,
                                  int64_t R, int64_t N, int64_t C, int64_t H, int64_t W,
                                  int pooled_h, int pooled_w, float* grad_input, void* stream);

/* the same two entry points for the reference's other dtypes (ROILoopPool_cuda.cu:294,364:
 * AT_DISPATCH_FLOATING_TYPES_AND_HALF): input, rois, output, grad tensors all of `dtype`, box arithmetic rounded the
 * way the reference's template rounds it for that T (c10::Half products and quotients rounded to half, the clamp bound
 * T(1.0 * width / spatial_scale), double throughout for double).  Values and argmax, no fused row scale.  F32 is accepted
 * too and gives the results of wsovod_b200_roi_loop_pool_fwd.  The backward accumulates in T with atomics. */
#define WSOVOD_B200_F32 0
#define WSOVOD_B200_F16 1
#define WSOVOD_B200_F64 2
size_t wsovod_b200_roi_loop_pool_dtype_workspace(int64_t N, int64_t R, int pooled_h, int pooled_w);
int wsovod_b200_roi_loop_pool_dtype_fwd(int dtype, const void* input, int64_t N, int64_t C, int64_t H, int64_t W,
                                        const void* rois, int64_t R, float spatial_scale,
                                        int pooled_h, int pooled_w, void* output, int32_t* argmax,
                                        void* workspace, size_t workspace_bytes, void* stream);
int wsovod_b200_roi_loop_pool_dtype_bwd(int dtype, const void* grad_output, const void* rois, const int32_t* argmax,
                                        int64_t R, int64_t N, int64_t C, int64_t H, int64_t W,
                                        int pooled_h, int pooled_w, void* grad_input, void* stream);

/* replaces torch.ops.torchvision.roi_align reached through detectron2's ROIAlign
 * (wsovod/modeling/poolers.py:169-182): bilinear, sampling_ratio<=0 -> ceil(roi/P) samples per bin,
 * aligned!=0 -> half-pixel offset ("ROIAlignV2"). */
size_t wsovod_b200_roi_align_workspace(int64_t N, int64_t R, int pooled_h, int pooled_w);
/* workspace that also holds the separable tap tables of the 7x7 adaptive-grid kernel for an H x W map (the forward
 * takes that kernel when workspace_bytes allows it, the per-sample kernel otherwise; same results to 1e-5 rel). */
size_t wsovod_b200_roi_align_workspace_hw(int64_t N, int64_t R, int pooled_h, int pooled_w, int64_t H, int64_t W);
int wsovod_b200_roi_align_fwd(const float* input, int64_t N, int64_t C, int64_t H, int64_t W,
                              const float* rois, int64_t R, float spatial_scale,
                              int pooled_h, int pooled_w, int sampling_ratio, int aligned,
                              const float* row_scale, float row_scale_bias,
                              float* output, void* workspace, size_t workspace_bytes, void* stream);

/* backward of roi_align_fwd (torchvision _roi_align_backward): every sample scatters grad * w / count to its four taps.
 * grad_output [R,C,PH,PW]; grad_input [N,C,H,W] must be zero-filled by the caller; fp32 atomics. */
int wsovod_b200_roi_align_bwd(const float* grad_output, const float* rois, int64_t R, int64_t N, int64_t C,
                              int64_t H, int64_t W, float spatial_scale, int pooled_h, int pooled_w,
                              int sampling_ratio, int aligned, float* grad_input, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (2) region x concept alignment and the MIL two-stream score.
 * ---------------------------------------------------------------------------------------------- */

#define WSOVOD_B200_ALIGN_FP32 0 /* fp32 FMA contraction, matches torch.mm(fp32) to ~1e-6 rel       */
#define WSOVOD_B200_ALIGN_TF32 1 /* tcgen05 kind::tf32, fp32 accumulate in TMEM: |dlogit| <= 5e-2   */

/* replaces the contraction part of OpenVocabularyClassifier.forward
 * (wsovod/modeling/class_heads/open_vocabulary_classifier.py:85-104) plus the row softmax of
 * InstanceRefinementOutputLayers.predict_probs (roi_heads/fast_rcnn_open_vocabulary.py:1019-1036):
 *   w_k   = norm_weight==1 ? classifier[k] / max(||classifier[k]||, 1e-12) : classifier[k]
 *   x_r   = norm_weight!=0 ? temperature * x[r] / max(||x[r]||, 1e-12)     : x[r]
 *   (norm_weight: 0 = NORM_WEIGHT False; 1 = True with a classifier passed in, :87-90; 2 = True with the
 *    module's stored weights, which the reference does NOT re-normalise, :91-92)
 *   logits[r,k] = <x_r, w_k> (+ bias[0] if bias != NULL);  background column (all-zero weight) appended
 *   when append_background != 0;  probs = softmax(logits, dim=1) when probs != NULL.
 * x [M,D], classifier [K,D] (the (K,D) text-embedding matrix, as loaded at
 * meta_arch/rcnn_wsovod.py:296-306), logits/probs [M, K + (append_background?1:0)].
 * logits may be NULL when only probs are wanted. */
size_t wsovod_b200_align_workspace(int64_t M, int64_t D, int64_t K, int precision);
int wsovod_b200_align_fwd(const float* x, const float* classifier, int64_t M, int64_t D, int64_t K,
                          float temperature, int norm_weight, int append_background,
                          const float* bias, int precision,
                          float* logits, float* probs,
                          void* workspace, size_t workspace_bytes, void* stream);
size_t wsovod_b200_align_bwd_workspace(int64_t M, int64_t D, int64_t K);
/* backward of align_fwd given grad_logits [M,K+bg], fp32: w.r.t. x (grad_x [M,D], may be NULL) and w.r.t. the
 * classifier (grad_classifier [K,D], may be NULL: the text embeddings are a buffer in every shipped config and a
 * Parameter only with weight_path "rand", open_vocabulary_classifier.py:62-65); includes the Jacobian of the
 * weight normalisation when norm_weight == 1.  Deterministic (fixed-order split-M sums). */
int wsovod_b200_align_bwd(const float* grad_logits, const float* x, const float* classifier,
                          int64_t M, int64_t D, int64_t K, float temperature, int norm_weight,
                          int append_background, float* grad_x, float* grad_classifier,
                          void* workspace, size_t workspace_bytes, void* stream);

/* replaces the per-image loop of ObjectMiningOutputLayers.forward and predict_probs_img
 * (roi_heads/fast_rcnn_open_vocabulary.py:338-357,604-618):
 *   scores[r,k]  = softmax(cls[r,:])[k] * softmax over the image's rows of det[:,k]
 *   img_scores[n,k] = clamp(sum_r scores[r,k], 1e-6, 1-1e-6)
 * cls, det, scores: [M,K]; offsets: device int64 [N+1], image n owns rows offsets[n]..offsets[n+1].
 * img_scores [N,K] may be NULL. */
size_t wsovod_b200_mil_workspace(int64_t M, int64_t N, int64_t K);
int wsovod_b200_mil_fwd(const float* cls, const float* det, const int64_t* offsets,
                        int64_t M, int64_t N, int64_t K, float* scores, float* img_scores,
                        void* workspace, size_t workspace_bytes, void* stream);
/* backward of mil_fwd: grad_cls/grad_det [M,K] from grad_scores [M,K] (may be NULL) and
 * grad_img [N,K] (may be NULL; gradient w.r.t. the *unclamped* image sum, i.e. already masked by the
 * caller where the clamp saturates). */
int wsovod_b200_mil_bwd(const float* grad_scores, const float* grad_img, const float* cls,
                        const float* det, const int64_t* offsets, int64_t M, int64_t N, int64_t K,
                        float* grad_cls, float* grad_det,
                        void* workspace, size_t workspace_bytes, void* stream);

/* North-star kernel 2 in one call: alignment on tensor cores FUSED with the MIL two-stream score -- replaces
 * ObjectMiningOutputLayers.forward + predict_probs_img when `cls` is the open-vocabulary class head
 * (roi_heads/fast_rcnn_open_vocabulary.py:280-285,318-367,604-618; the head variant of roi_heads.py:588-590):
 *   C = align_fwd(x, classifier) without background column (tcgen05 kind::tf32, |dlogit| <= 5e-2),
 *   scores[r,k] = softmax(C[r,:])[k] * softmax over the image's rows of det[:,k],  img_scores as mil_fwd.
 * The TMEM epilogue emits the row softmax and per-(image, class) column (max, sum exp) partials of `det` in one
 * pass; a second light launch applies the column softmax in place and sums the image scores (deterministic).
 * x [M,D] (16-byte aligned, D % 4 == 0), classifier [K,D] with K <= 256, det/scores [M,K], offsets device int64 [N+1],
 * img_scores [N,K] (may be NULL), logits [M,K] (may be NULL; the backward pass is mil_bwd on them, then align_bwd). */
size_t wsovod_b200_align_mil_fused_workspace(int64_t M, int64_t N, int64_t D, int64_t K);
int wsovod_b200_align_mil_fused_fwd(const float* x, const float* classifier, const float* det,
                                    const int64_t* offsets, int64_t M, int64_t N, int64_t D, int64_t K,
                                    float temperature, int norm_weight, const float* bias, float* scores,
                                    float* img_scores, float* logits, void* workspace, size_t workspace_bytes,
                                    void* stream);

/* ------------------------------------------------------------------------------------------------
 * (3) refinement pseudo-label assignment.
 * ---------------------------------------------------------------------------------------------- */

/* replaces the top_k=1 seed selection of WSOVODROIHeads.get_pgt_top_k
 * (wsovod/modeling/roi_heads/roi_heads.py:1079-1207): for image n and each of its image-level classes
 * c = gt_classes[g] (g in gt_offsets[n]..gt_offsets[n+1]), over the image's proposals whose box area
 * is > 20: r* = first argmax_r scores[r,c].  Writes seed_boxes[g] = boxes[r*], seed_scores[g],
 * seed_rows[g] = r* (global row, -1 if the image has no eligible proposal) and
 * seed_weights[g] = img_scores[n,c].  An image with no eligible proposal gets the reference's
 * fallback seed (box (-1e4,-1e4,1e4,1e4), score 1, weight 1; class forced to 0) in its FIRST slot and
 * seed_count[n] = 1; otherwise seed_count[n] = number of classes.  scores: [M,score_stride] (only the
 * first K columns are read). */
int wsovod_b200_pgt_top1(const float* scores, int64_t score_stride, const float* boxes,
                         const int64_t* offsets, const int64_t* gt_classes, const int64_t* gt_offsets,
                         const float* img_scores, int64_t M, int64_t N, int64_t K, int64_t G,
                         float* seed_boxes, int64_t* seed_classes, float* seed_scores,
                         float* seed_weights, int64_t* seed_rows, int64_t* seed_count, void* stream);

/* replaces pairwise_iou + Matcher([thr],[0,1]) + the label/gather part of
 * label_and_sample_proposals_wsl / _sample_proposals_wsl (roi_heads.py:1589-1593,1770-1797):
 *   iou[g,r] = inter>0 ? inter/(area_g + area_r - inter) : 0 (every op fp32, round-to-nearest)
 *   matched_idx[r] = first argmax_g iou[g,r] (index local to the image's seed list), matched_label =
 *   iou >= iou_thresh, gt_class[r] = label ? seed_classes[idx] : num_classes, and the gathers
 *   gt_boxes/gt_scores/gt_weights[r] = seed_*[idx].  An image with zero seeds gets idx 0, label 0,
 *   class num_classes and zero-filled gathers.  seed_offsets: device int64 [N+1] into the seed arrays
 *   (only the first seed_count[n] seeds of image n are used when seed_count != NULL).
 *   Random subsampling to BATCH_SIZE_PER_IMAGE stays with the caller (torch RNG). */
int wsovod_b200_refine_assign(const float* boxes, const int64_t* offsets,
                              const float* seed_boxes, const int64_t* seed_classes,
                              const float* seed_scores, const float* seed_weights,
                              const int64_t* seed_offsets, const int64_t* seed_count,
                              int64_t M, int64_t N, int64_t num_classes, float iou_thresh,
                              int64_t* matched_idx, int8_t* matched_label, float* matched_iou,
                              int64_t* gt_classes, float* gt_boxes, float* gt_scores,
                              float* gt_weights, void* stream);

/* SURVEY 8f-2 -- replaces InstanceRefinementOutputLayers.losses with cross_entropy_weighted and
 * BBOX_REG_LOSS_TYPE "smooth_l1_weighted" (fast_rcnn_open_vocabulary.py:754-892), the consumer of
 * refine_assign's gt_classes / gt_boxes / gt_weights:
 *   w_i = gt_classes_i == -1 ? 0 : gt_weights_i;  valid = #(w_i > 1e-12)
 *   out[0] = sum_i w_i * CE(logits[i, :K1], gt_classes_i, ignore_index = -1) / valid          (:813-820)
 *   out[1] = sum_{0 <= gt_i < num_classes} w_i * sum_j smooth_l1(deltas_ij - target_ij, beta) / max(M, 1),
 *            target = Box2BoxTransform(wx, wy, ww, wh).get_deltas(proposal_boxes_i, gt_boxes_i);
 *            0 if any target is NaN (:869-872)                                                 (:864-892)
 *   out[2] = valid, out[3] = 1 if a target was NaN;  lse[i] = logsumexp(logits[i]) (kept for backward).
 * deltas: [M, dcols], dcols = 4 (class-agnostic), 4 * num_classes (class-specific) or 0 (no box loss:
 * refine_reg false, boxes and deltas may be NULL).  Deterministic: per-CTA partial sums added in a fixed
 * order in double precision.  Workspace: wsovod_b200_refine_loss_workspace(M) bytes. */
size_t wsovod_b200_refine_loss_workspace(int64_t M);
int wsovod_b200_refine_loss_fwd(const float* logits, int64_t K1, const int64_t* gt_classes,
                                const float* gt_weights, const float* proposal_boxes, const float* gt_boxes,
                                const float* deltas, int64_t dcols, int64_t M, int64_t num_classes,
                                float wx, float wy, float ww, float wh, float beta, float* out,
                                float* lse, void* workspace, size_t workspace_bytes, void* stream);
/* backward: grad_out [2] (device: d/d out[0], d/d out[1]), fwd_out = the forward's out[4];
 * grad_logits [M, K1] and grad_deltas [M, dcols] (either may be NULL). */
int wsovod_b200_refine_loss_bwd(const float* grad_out, const float* fwd_out, const float* logits, int64_t K1,
                                const float* lse, const int64_t* gt_classes, const float* gt_weights,
                                const float* proposal_boxes, const float* gt_boxes, const float* deltas,
                                int64_t dcols, int64_t M, int64_t num_classes, float wx, float wy, float ww,
                                float wh, float beta, float* grad_logits, float* grad_deltas, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (4) per-class NMS.
 * ---------------------------------------------------------------------------------------------- */

/* IoU arithmetic of the suppressing test `iou(i,j) > thr` (i = kept, higher score; j = candidate) */
#define WSOVOD_B200_IOU_TV_CPU 0  /* torchvision CPU kernel: (area_i + area_j) - inter, fp32, compared in double */
#define WSOVOD_B200_IOU_TV_CUDA 1 /* torchvision sm_100 SASS: fmaf(w_j,h_j,area_i) - inter, compared in fp32     */

/* replaces detectron2.layers.batched_nms -> torchvision.ops.batched_nms in its "vanilla" strategy on
 * un-offset coordinates (call sites roi_heads/fast_rcnn_open_vocabulary.py:206, roi_heads.py:933,
 * proposal_generator/proposal_utils.py:129,335).  groups: dense class ids in [0,num_groups), int64.
 * iou_thresh is a double because torchvision's CPU kernel compares the fp32 IoU against the python
 * float as a C++ double (mode TV_CPU); mode TV_CUDA narrows it to fp32 once, like the CUDA kernel.
 * keep [M] int64 receives the kept candidate indices ordered by score descending (ties: lower index
 * first); *num_keep (device int64) their count.  Entries past num_keep are set to -1. */
size_t wsovod_b200_batched_nms_workspace(int64_t M, int64_t num_groups);
int wsovod_b200_batched_nms(const float* boxes, const float* scores, const int64_t* groups,
                            int64_t M, int64_t num_groups, double iou_thresh, int iou_mode,
                            int64_t* keep, int64_t* num_keep,
                            void* workspace, size_t workspace_bytes, void* stream);

/* replaces fast_rcnn_inference / fast_rcnn_inference_single_image
 * (roi_heads/fast_rcnn_open_vocabulary.py:52-96,149-217) for class-agnostic boxes:
 *   per image: drop rows with a non-finite box or score, drop the background column, clip boxes to
 *   image_sizes[n] = (height, width), keep (row, class) with score > score_thresh, per-class NMS at
 *   nms_thresh, then the topk highest-scoring survivors (ties: lower row*K+class first).
 * probs [M,K+1], boxes [M,4], offsets device int64 [N+1], image_sizes device fp32 [N,2].
 * Outputs are padded to topk per image: det_boxes [N,topk,4] (clipped), det_scores [N,topk],
 * det_classes [N,topk] int64, det_rows [N,topk] int64 (row local to the image = the reference's
 * pred_inds; equal to its kept_indices whenever no row was dropped as non-finite), det_count [N]
 * int64.  Unused slots: score 0, class/row -1, box 0.  max_rows_per_image: host-known upper bound of
 * offsets[n+1]-offsets[n] (sizes the per-class shared-memory candidate list; <= 16384).  topk must
 * be in [1,4096]; for "keep everything" filter on the host side and call batched_nms. */
size_t wsovod_b200_detections_workspace(int64_t M, int64_t N, int64_t K, int64_t topk);
int wsovod_b200_detections(const float* probs, const float* boxes, const int64_t* offsets,
                           const float* image_sizes, int64_t M, int64_t N, int64_t K,
                           int64_t max_rows_per_image,
                           float score_thresh, double nms_thresh, int64_t topk, int iou_mode,
                           float* det_boxes, float* det_scores, int64_t* det_classes,
                           int64_t* det_rows, int64_t* det_count,
                           void* workspace, size_t workspace_bytes, void* stream);

/* replaces wsovod._C.csc_forward (wsovod/layers/vision.cpp:12, wsovod/layers/csc/csc.h:22-33,
 * csc_cuda.cu:183-531; call site proposal_generator/proposal_utils.py:272-291): W [R,K], the proposals' frame-vs-context
 * contrast of the class peak response maps cpgs [B,K,H,W], for the classes with labels[b,k] >= 0.5, normalised to
 * [-1, 1] and blended with preds [B,K]; columns of classes without a positive label are 1.  rois [R,5] in MAP pixels
 * (the batch index is not read, as upstream: every positive (image, class) pair rewrites the whole column, the last
 * image wins).  The reference's tau / mass_threshold / density_threshold arguments do not reach its arithmetic
 * (csc_cuda.cu:424-426,443-444,322) and are not part of this entry.  Stream-ordered (the reference blocks the host). */
size_t wsovod_b200_csc_workspace(int64_t K, int64_t H, int64_t W);
int wsovod_b200_csc_fwd(const float* cpgs, const float* labels, const float* preds, const float* rois,
                        int64_t B, int64_t K, int64_t H, int64_t W, int64_t R, float fg_threshold, int area_sqrt,
                        float context_scale, float* W_out, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Host-buffer convenience for the reference-facing end-to-end call (bench.py's `e2e`): the inference
 * slice pool -> align+softmax -> detections on pinned HOST buffers; copies in, launches, copies the
 * detections out on `stream`.  `dev_arena` is a caller-owned device buffer of at least
 * wsovod_b200_infer_host_arena() bytes.  The pooled tensor stays on the device (it feeds the
 * box-head FCs there); a pointer to it inside the arena is returned through pooled_dev when non-NULL.
 * copy_stream (may be NULL): a second caller-owned stream; when given, the per-image host->device copies
 * are issued there and overlap the kernels of the previous image (events order the two streams; the
 * caller only synchronises `stream`).  The embedding area of the arena is reused for packed boxes, so a
 * new call must not start on another stream before this one has finished.
 * ---------------------------------------------------------------------------------------------- */
size_t wsovod_b200_infer_host_arena(int64_t N, int64_t C, int64_t H, int64_t W, int64_t R_total,
                                    int64_t D, int64_t K, int pooled, int64_t topk, int with_argmax);
int wsovod_b200_infer_host(const float* h_features, int64_t N, int64_t C, int64_t H, int64_t W,
                           const float* h_rois, const float* h_objectness, int64_t R_total,
                           const int64_t* h_offsets, const float* h_image_sizes,
                           const float* h_region_emb, const float* h_text_emb, int64_t D, int64_t K,
                           float spatial_scale, int pooled, float temperature,
                           float score_thresh, double nms_thresh, int64_t topk,
                           int precision, int iou_mode, int with_argmax,
                           float* h_det_boxes, float* h_det_scores, int64_t* h_det_classes,
                           int64_t* h_det_rows, int64_t* h_det_count,
                           void* dev_arena, size_t arena_bytes, float** pooled_dev, void* stream,
                           void* copy_stream);

#ifdef __cplusplus
}
#endif
#endif /* WSOVOD_B200_H */
