// infer_host.cu -- the reference-facing end-to-end inference slice on HOST buffers:
//   H2D(features, rois, objectness, region embeddings, text embeddings, offsets, image sizes)
//   -> roi_pool(+objectness scale) -> align+softmax -> detections -> D2H(detections)
// all on one stream, no host synchronisation (the caller syncs the stream).  This is what bench.py
// times as `e2e`.  The pooled tensor stays on the device (the box-head FCs consume it there).
#include "common.cuh"

#include <algorithm>

namespace wsovod {
struct Arena {
  size_t feat, rois, obj, offs, sizes, emb, text, pooled, argmax, probs, ws_pool, ws_align, ws_det;
  size_t det_boxes, det_scores, det_classes, det_rows, det_count, bytes;
};
static Arena arena_plan(int64_t N, int64_t C, int64_t H, int64_t W, int64_t R, int64_t D, int64_t K, int P,
                        int64_t topk, int with_argmax, int precision) {
  Arena a;
  size_t o = 0;
  auto take = [&](size_t b) { size_t r = o; o += align_up(b, 1024); return r; };
  a.feat = take(sizeof(float) * (size_t)(N * C * H * W));
  a.rois = take(sizeof(float) * 5 * (size_t)R);
  a.obj = take(sizeof(float) * (size_t)R);
  a.offs = take(sizeof(int64_t) * (size_t)(N + 1));
  a.sizes = take(sizeof(float) * 2 * (size_t)N);
  a.emb = take(sizeof(float) * (size_t)(R * D));
  a.text = take(sizeof(float) * (size_t)(K * D));
  a.pooled = take(sizeof(float) * (size_t)(R * C * P * P));
  a.argmax = take(with_argmax ? sizeof(int32_t) * (size_t)(R * C * P * P) : 0);
  a.probs = take(sizeof(float) * (size_t)(R * (K + 1)));
  a.ws_pool = take(wsovod_b200_roi_pool_workspace(N, R, P, P));
  a.ws_align = take(wsovod_b200_align_workspace(R, D, K, precision));
  a.ws_det = take(wsovod_b200_detections_workspace(R, N, K, topk));
  a.det_boxes = take(sizeof(float) * 4 * (size_t)(N * topk));
  a.det_scores = take(sizeof(float) * (size_t)(N * topk));
  a.det_classes = take(sizeof(int64_t) * (size_t)(N * topk));
  a.det_rows = take(sizeof(int64_t) * (size_t)(N * topk));
  a.det_count = take(sizeof(int64_t) * (size_t)N);
  a.bytes = o;
  return a;
}
}  // namespace wsovod

using namespace wsovod;

WSOVOD_API size_t wsovod_b200_infer_host_arena(int64_t N, int64_t C, int64_t H, int64_t W, int64_t R_total,
                                               int64_t D, int64_t K, int pooled, int64_t topk,
                                               int with_argmax) {
  if (N < 0 || C < 0 || H < 0 || W < 0 || R_total < 0 || D < 0 || K < 0 || pooled <= 0 || topk <= 0) return 0;
  // sized for the larger (TF32) workspace so one arena serves both precisions
  return arena_plan(N, C, H, W, R_total, D, K, pooled, topk, with_argmax, WSOVOD_B200_ALIGN_TF32).bytes;
}

WSOVOD_API int wsovod_b200_infer_host(const float* h_features, int64_t N, int64_t C, int64_t H, int64_t W,
                                      const float* h_rois, const float* h_objectness, int64_t R,
                                      const int64_t* h_offsets, const float* h_image_sizes,
                                      const float* h_region_emb, const float* h_text_emb, int64_t D,
                                      int64_t K, float spatial_scale, int pooled, float temperature,
                                      float score_thresh, double nms_thresh, int64_t topk, int precision,
                                      int iou_mode, int with_argmax, float* h_det_boxes,
                                      float* h_det_scores, int64_t* h_det_classes, int64_t* h_det_rows,
                                      int64_t* h_det_count, void* dev_arena, size_t arena_bytes,
                                      float** pooled_dev, void* stream) {
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0 || R < 0 || D <= 0 || K <= 0 || pooled <= 0 || topk <= 0)
    return WSOVOD_B200_EINVAL;
  if (!h_features || !h_rois || !h_offsets || !h_image_sizes || !h_region_emb || !h_text_emb ||
      !h_det_boxes || !h_det_scores || !h_det_classes || !h_det_rows || !h_det_count || !dev_arena)
    return WSOVOD_B200_EINVAL;
  const Arena a = arena_plan(N, C, H, W, R, D, K, pooled, topk, with_argmax, WSOVOD_B200_ALIGN_TF32);
  if (arena_bytes < a.bytes) return WSOVOD_B200_EWORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  char* d = (char*)dev_arena;
  int64_t max_rows = 0;
  for (int64_t n = 0; n < N; ++n) max_rows = std::max(max_rows, h_offsets[n + 1] - h_offsets[n]);
  auto h2d = [&](size_t off, const void* src, size_t bytes) {
    return bytes ? cudaMemcpyAsync(d + off, src, bytes, cudaMemcpyHostToDevice, st) : cudaSuccess;
  };
  cudaError_t e;
  if ((e = h2d(a.feat, h_features, sizeof(float) * (size_t)(N * C * H * W))) != cudaSuccess) return (int)e;
  if ((e = h2d(a.rois, h_rois, sizeof(float) * 5 * (size_t)R)) != cudaSuccess) return (int)e;
  if (h_objectness && (e = h2d(a.obj, h_objectness, sizeof(float) * (size_t)R)) != cudaSuccess) return (int)e;
  if ((e = h2d(a.offs, h_offsets, sizeof(int64_t) * (size_t)(N + 1))) != cudaSuccess) return (int)e;
  if ((e = h2d(a.sizes, h_image_sizes, sizeof(float) * 2 * (size_t)N)) != cudaSuccess) return (int)e;
  if ((e = h2d(a.emb, h_region_emb, sizeof(float) * (size_t)(R * D))) != cudaSuccess) return (int)e;
  if ((e = h2d(a.text, h_text_emb, sizeof(float) * (size_t)(K * D))) != cudaSuccess) return (int)e;
  int rc = wsovod_b200_roi_pool_fwd((const float*)(d + a.feat), N, C, H, W, (const float*)(d + a.rois), R,
                                    spatial_scale, pooled, pooled,
                                    h_objectness ? (const float*)(d + a.obj) : nullptr, 1.0f,
                                    (float*)(d + a.pooled), with_argmax ? (int32_t*)(d + a.argmax) : nullptr,
                                    d + a.ws_pool, a.ws_align - a.ws_pool, st);
  if (rc) return rc;
  rc = wsovod_b200_align_fwd((const float*)(d + a.emb), (const float*)(d + a.text), R, D, K, temperature, 1, 1,
                             nullptr, precision, nullptr, (float*)(d + a.probs), d + a.ws_align,
                             a.ws_det - a.ws_align, st);
  if (rc) return rc;
  // class-agnostic boxes = the proposal boxes (columns 1..4 of rois) -> need a packed [R,4] copy
  // (rois rows are 20 B apart): reuse the first 16R bytes of the detections' output area? No: pack
  // with a strided 2D copy into the emb buffer, which align_fwd has finished reading on this stream.
  float* dboxes = (float*)(d + a.emb);
  e = cudaMemcpy2DAsync(dboxes, 16, (const char*)(d + a.rois) + 4, 20, 16, (size_t)R, cudaMemcpyDeviceToDevice, st);
  if (e != cudaSuccess) return (int)e;
  rc = wsovod_b200_detections((const float*)(d + a.probs), dboxes, (const int64_t*)(d + a.offs),
                              (const float*)(d + a.sizes), R, N, K, max_rows, score_thresh, nms_thresh, topk,
                              iou_mode, (float*)(d + a.det_boxes), (float*)(d + a.det_scores),
                              (int64_t*)(d + a.det_classes), (int64_t*)(d + a.det_rows),
                              (int64_t*)(d + a.det_count), d + a.ws_det, a.det_boxes - a.ws_det, st);
  if (rc) return rc;
  auto d2h = [&](void* dst, size_t off, size_t bytes) {
    return cudaMemcpyAsync(dst, d + off, bytes, cudaMemcpyDeviceToHost, st);
  };
  if ((e = d2h(h_det_boxes, a.det_boxes, sizeof(float) * 4 * (size_t)(N * topk))) != cudaSuccess) return (int)e;
  if ((e = d2h(h_det_scores, a.det_scores, sizeof(float) * (size_t)(N * topk))) != cudaSuccess) return (int)e;
  if ((e = d2h(h_det_classes, a.det_classes, sizeof(int64_t) * (size_t)(N * topk))) != cudaSuccess) return (int)e;
  if ((e = d2h(h_det_rows, a.det_rows, sizeof(int64_t) * (size_t)(N * topk))) != cudaSuccess) return (int)e;
  if ((e = d2h(h_det_count, a.det_count, sizeof(int64_t) * (size_t)N)) != cudaSuccess) return (int)e;
  if (pooled_dev) *pooled_dev = (float*)(d + a.pooled);
  return 0;
}
