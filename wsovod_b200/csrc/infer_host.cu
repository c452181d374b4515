// infer_host.cu -- the reference-facing end-to-end inference slice on HOST buffers:
//   H2D(features, rois, objectness, region embeddings, text embeddings, offsets, image sizes)
//   -> roi_pool(+objectness scale) -> align+softmax -> detections -> D2H(detections)
// all on one stream, no host synchronisation (the caller syncs the stream).  This is what bench.py
// times as `e2e`.  The pooled tensor stays on the device (the box-head FCs consume it there).
#include "common.cuh"

#include <algorithm>
#include <map>
#include <vector>

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

// Events of this entry point: a per-thread, per-device pool created on first use and reused by every later call (the
// entry point needs up to 2 N + 2 of them per call; creating and destroying them each time cost ~40 us of host time).
// Re-recording an event whose earlier cudaStreamWaitEvent is still pending is fine: a wait binds to the record that
// preceded it.  The pool is never destroyed (process lifetime).
namespace {
struct EventPool {
  std::vector<cudaEvent_t> ev;
  size_t used = 0;
  cudaError_t next(cudaEvent_t* out) {
    if (used == ev.size()) {
      cudaEvent_t e = nullptr;
      const cudaError_t rc = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
      if (rc != cudaSuccess) return rc;
      ev.push_back(e);
    }
    *out = ev[used++];
    return cudaSuccess;
  }
};
EventPool& event_pool() {
  thread_local std::map<int, EventPool> pools;
  int dev = 0;
  cudaGetDevice(&dev);
  EventPool& p = pools[dev];
  p.used = 0;
  return p;
}
}  // namespace

WSOVOD_API int wsovod_b200_infer_host(const float* h_features, int64_t N, int64_t C, int64_t H, int64_t W,
                                      const float* h_rois, const float* h_objectness, int64_t R,
                                      const int64_t* h_offsets, const float* h_image_sizes,
                                      const float* h_region_emb, const float* h_text_emb, int64_t D,
                                      int64_t K, float spatial_scale, int pooled, float temperature,
                                      float score_thresh, double nms_thresh, int64_t topk, int precision,
                                      int iou_mode, int with_argmax, float* h_det_boxes,
                                      float* h_det_scores, int64_t* h_det_classes, int64_t* h_det_rows,
                                      int64_t* h_det_count, void* dev_arena, size_t arena_bytes,
                                      float** pooled_dev, void* stream, void* copy_stream) {
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0 || R < 0 || D <= 0 || K <= 0 || pooled <= 0 || topk <= 0)
    return WSOVOD_B200_EINVAL;
  if (!h_features || !h_rois || !h_offsets || !h_image_sizes || !h_region_emb || !h_text_emb ||
      !h_det_boxes || !h_det_scores || !h_det_classes || !h_det_rows || !h_det_count || !dev_arena)
    return WSOVOD_B200_EINVAL;
  const Arena a = arena_plan(N, C, H, W, R, D, K, pooled, topk, with_argmax, WSOVOD_B200_ALIGN_TF32);
  if (arena_bytes < a.bytes) return WSOVOD_B200_EWORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  // With a second stream the per-image host->device copies run ahead of the kernels of the previous
  // image (PCIe and SMs busy at the same time); without one everything is serial on `stream`.
  cudaStream_t cs = copy_stream ? (cudaStream_t)copy_stream : st;
  const bool piped = copy_stream != nullptr && copy_stream != stream;
  char* d = (char*)dev_arena;
  int64_t max_rows = 0;
  for (int64_t n = 0; n < N; ++n) {
    if (h_offsets[n + 1] < h_offsets[n]) return WSOVOD_B200_EINVAL;
    max_rows = std::max(max_rows, h_offsets[n + 1] - h_offsets[n]);
  }
  if (h_offsets[0] != 0 || h_offsets[N] != R) return WSOVOD_B200_EINVAL;
  cudaError_t e = cudaSuccess;
  int rc = 0;
  cudaEvent_t ev_start = nullptr, ev_small = nullptr;
  std::vector<cudaEvent_t> ev((size_t)N, nullptr);
  EventPool& pool = event_pool();
  auto cleanup = [&]() {};
#define CK(call) do { e = (call); if (e != cudaSuccess) { cleanup(); return (int)e; } } while (0)
#define RC(call) do { rc = (call); if (rc) { cleanup(); return rc; } } while (0)
  auto h2d = [&](size_t off, const void* src, size_t bytes) {
    return bytes ? cudaMemcpyAsync(d + off, src, bytes, cudaMemcpyHostToDevice, cs) : cudaSuccess;
  };
  if (piped) {   // the copy stream must not overwrite the arena before earlier work on `stream` is done
    CK(pool.next(&ev_start));
    CK(cudaEventRecord(ev_start, st));
    CK(cudaStreamWaitEvent(cs, ev_start, 0));
  }
  // small tensors first
  CK(h2d(a.rois, h_rois, sizeof(float) * 5 * (size_t)R));
  if (h_objectness) CK(h2d(a.obj, h_objectness, sizeof(float) * (size_t)R));
  CK(h2d(a.offs, h_offsets, sizeof(int64_t) * (size_t)(N + 1)));
  CK(h2d(a.sizes, h_image_sizes, sizeof(float) * 2 * (size_t)N));
  CK(h2d(a.text, h_text_emb, sizeof(float) * (size_t)(K * D)));
  if (piped) {
    CK(pool.next(&ev_small));
    CK(cudaEventRecord(ev_small, cs));
    CK(cudaStreamWaitEvent(st, ev_small, 0));
  }
  const size_t plane = sizeof(float) * (size_t)(C * H * W);
  const int64_t out_row = C * (int64_t)pooled * pooled;
  const size_t ws_pool_bytes = a.ws_align - a.ws_pool, ws_align_bytes = a.ws_det - a.ws_align;
  // Copies: all feature planes first, then the region embeddings.  Pooling of image n starts when its plane
  // has landed (the big kernels overlap the rest of the copies); alignment only needs the embeddings, so
  // what is left after the LAST byte arrives is one alignment launch + the detections, not a pooling.
  std::vector<cudaEvent_t> ev_emb((size_t)N, nullptr);
  auto cleanup2 = [&]() {};
#define CK2(call) do { e = (call); if (e != cudaSuccess) { cleanup2(); cleanup(); return (int)e; } } while (0)
#define RC2(call) do { rc = (call); if (rc) { cleanup2(); cleanup(); return rc; } } while (0)
  for (int64_t n = 0; n < N; ++n) {
    CK2(h2d(a.feat + plane * (size_t)n, h_features + (size_t)n * (size_t)(C * H * W), plane));
    if (piped) {
      CK2(pool.next(&ev[n]));
      CK2(cudaEventRecord(ev[n], cs));
    }
  }
  for (int64_t n = 0; n < N; ++n) {
    const int64_t r0 = h_offsets[n], rn = h_offsets[n + 1] - r0;
    CK2(h2d(a.emb + sizeof(float) * (size_t)(r0 * D), h_region_emb + r0 * D, sizeof(float) * (size_t)(rn * D)));
    if (piped) {
      CK2(pool.next(&ev_emb[n]));
      CK2(cudaEventRecord(ev_emb[n], cs));
    }
  }
  // One pooling launch per image (N = 1, that image's plane and rois) so it can start as soon as the
  // image has landed.  The rois keep their global batch index; with N = 1 the kernel clamps it to 0.
  for (int64_t n = 0; n < N; ++n) {
    const int64_t r0 = h_offsets[n], rn = h_offsets[n + 1] - r0;
    if (piped) CK2(cudaStreamWaitEvent(st, ev[n], 0));
    if (rn == 0) continue;
    RC2(wsovod_b200_roi_pool_fwd((const float*)(d + a.feat + plane * (size_t)n), 1, C, H, W,
                                 (const float*)(d + a.rois) + 5 * r0, rn,
                                 spatial_scale, pooled, pooled,
                                 h_objectness ? (const float*)(d + a.obj) + r0 : nullptr, 1.0f,
                                 (float*)(d + a.pooled) + r0 * out_row,
                                 with_argmax ? (int32_t*)(d + a.argmax) + r0 * out_row : nullptr,
                                 d + a.ws_pool, ws_pool_bytes, st));
  }
  for (int64_t n = 0; n < N; ++n) {
    const int64_t r0 = h_offsets[n], rn = h_offsets[n + 1] - r0;
    if (piped) CK2(cudaStreamWaitEvent(st, ev_emb[n], 0));
    if (rn == 0) continue;
    RC2(wsovod_b200_align_fwd((const float*)(d + a.emb) + r0 * D, (const float*)(d + a.text), rn, D, K, temperature,
                              1, 1, nullptr, precision, nullptr, (float*)(d + a.probs) + r0 * (K + 1),
                              d + a.ws_align, ws_align_bytes, st));
  }
  cleanup2();
#undef CK2
#undef RC2
  // class-agnostic boxes = the proposal boxes (columns 1..4 of rois) -> packed [R,4] copy (rois rows are
  // 20 B apart) into the embedding buffer, which align_fwd has finished reading on this stream
  float* dboxes = (float*)(d + a.emb);
  CK(cudaMemcpy2DAsync(dboxes, 16, (const char*)(d + a.rois) + 4, 20, 16, (size_t)R, cudaMemcpyDeviceToDevice, st));
  RC(wsovod_b200_detections((const float*)(d + a.probs), dboxes, (const int64_t*)(d + a.offs),
                            (const float*)(d + a.sizes), R, N, K, max_rows, score_thresh, nms_thresh, topk,
                            iou_mode, (float*)(d + a.det_boxes), (float*)(d + a.det_scores),
                            (int64_t*)(d + a.det_classes), (int64_t*)(d + a.det_rows),
                            (int64_t*)(d + a.det_count), d + a.ws_det, a.det_boxes - a.ws_det, st));
  auto d2h = [&](void* dst, size_t off, size_t bytes) {
    return cudaMemcpyAsync(dst, d + off, bytes, cudaMemcpyDeviceToHost, st);
  };
  CK(d2h(h_det_boxes, a.det_boxes, sizeof(float) * 4 * (size_t)(N * topk)));
  CK(d2h(h_det_scores, a.det_scores, sizeof(float) * (size_t)(N * topk)));
  CK(d2h(h_det_classes, a.det_classes, sizeof(int64_t) * (size_t)(N * topk)));
  CK(d2h(h_det_rows, a.det_rows, sizeof(int64_t) * (size_t)(N * topk)));
  CK(d2h(h_det_count, a.det_count, sizeof(int64_t) * (size_t)N));
#undef CK
#undef RC
  cleanup();
  if (pooled_dev) *pooled_dev = (float*)(d + a.pooled);
  return 0;
}
