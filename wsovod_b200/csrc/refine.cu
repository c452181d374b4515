// refine.cu -- kernel (3): OICR-style refinement pseudo-labels, batched over all images.
//
//   pgt_top1      : per (image, image-level class) seed = first argmax of the class score over the
//                   image's proposals with box area > 20 (roi_heads.py:1079-1146), fallbacks :1182-1207
//   refine_assign : per proposal IoU against the image's seeds -> first argmax -> label (>= thr) ->
//                   class / gathered seed box, score, loss weight (roi_heads.py:1587-1593,1770-1797)
// The reference runs a Python loop per image with dozens of micro-kernels; here one launch each for
// the whole batch.  IoU uses explicit round-to-nearest intrinsics so no FMA contraction can change a
// label: iou = inter / ((area_g + area_r) - inter), exactly detectron2.structures.pairwise_iou.
#include "common.cuh"

namespace wsovod {

__device__ __forceinline__ float box_area_rn(const float4 b) {
  return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
}

__device__ __forceinline__ float d2_iou(const float4 a, const float4 b) {
  float w = __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x));
  float h = __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y));
  w = fmaxf(w, 0.f);
  h = fmaxf(h, 0.f);
  const float inter = __fmul_rn(w, h);
  const float den = __fsub_rn(__fadd_rn(box_area_rn(a), box_area_rn(b)), inter);
  return inter > 0.f ? __fdiv_rn(inter, den) : 0.f;
}

// image of global row/seed index i given offsets[N+1] (N is small: linear scan from a guess)
__device__ __forceinline__ int find_segment(const int64_t* __restrict__ offsets, int N, int64_t i) {
  int lo = 0, hi = N - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (offsets[mid] <= i) lo = mid; else hi = mid - 1;
  }
  return lo;
}

constexpr int kSeedThreads = 512;

// one CTA per seed slot g
__global__ void __launch_bounds__(kSeedThreads) pgt_top1_kernel(
    const float* __restrict__ scores, int64_t stride, const float* __restrict__ boxes,
    const int64_t* __restrict__ offsets, const int64_t* __restrict__ gt_classes,
    const int64_t* __restrict__ gt_offsets, const float* __restrict__ img_scores, int N, int K,
    float* __restrict__ seed_boxes, int64_t* __restrict__ seed_classes, float* __restrict__ seed_scores,
    float* __restrict__ seed_weights, int64_t* __restrict__ seed_rows, int64_t* __restrict__ seed_count) {
  __shared__ float s_val[kSeedThreads / 32];
  __shared__ long long s_row[kSeedThreads / 32];
  const int64_t g = blockIdx.x;
  const int n = find_segment(gt_offsets, N, g);
  const int64_t c = gt_classes[g];
  const int64_t r0 = offsets[n], r1 = offsets[n + 1];
  float best = 0.f;
  long long brow = -1;
  const float4* b4 = reinterpret_cast<const float4*>(boxes);
  // four rows per thread and trip, box and score loads of a trip in flight together (the loop is a chain
  // of DRAM round trips otherwise: 20 trips x ~1 us at 5000 proposals)
  for (int64_t rb = r0 + threadIdx.x; rb < r1; rb += 4 * kSeedThreads) {
    float4 b[4];
    float s[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t r = rb + u * kSeedThreads;
      b[u] = r < r1 ? __ldg(b4 + r) : make_float4(0.f, 0.f, 0.f, 0.f);
      s[u] = r < r1 ? __ldg(scores + r * stride + c) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t r = rb + u * kSeedThreads;
      if (r < r1 && box_area_rn(b[u]) > 20.f && (brow < 0 || s[u] > best)) { best = s[u]; brow = r; }   // rows in increasing order
    }
  }
  // (score desc, row asc) reduction
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const long long orow = __shfl_xor_sync(0xffffffffu, brow, o);
    const bool take = orow >= 0 && (brow < 0 || ov > best || (ov == best && orow < brow));
    if (take) { best = ov; brow = orow; }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { s_val[wid] = best; s_row[wid] = brow; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kSeedThreads / 32; ++w) {
      const float ov = s_val[w];
      const long long orow = s_row[w];
      const bool take = orow >= 0 && (brow < 0 || ov > best || (ov == best && orow < brow));
      if (take) { best = ov; brow = orow; }
    }
    const int64_t g0 = gt_offsets[n], g1 = gt_offsets[n + 1];
    float4* sb = reinterpret_cast<float4*>(seed_boxes);
    if (brow >= 0) {
      sb[g] = __ldg(b4 + brow);
      seed_classes[g] = c;
      seed_scores[g] = best;
      seed_weights[g] = img_scores[(int64_t)n * K + c];
      seed_rows[g] = brow;
      if (g == g0) seed_count[n] = g1 - g0;
    } else {
      // no proposal of the image survives the area filter (the filter is class independent, so every
      // slot of the image lands here): the reference falls back to one dummy seed (:1182-1207)
      if (g == g0) {
        sb[g] = make_float4(-10000.f, -10000.f, 10000.f, 10000.f);
        seed_classes[g] = 0; seed_scores[g] = 1.f; seed_weights[g] = 1.f;
        seed_count[n] = 1;
      } else {
        sb[g] = make_float4(0.f, 0.f, 0.f, 0.f);
        seed_classes[g] = c; seed_scores[g] = 0.f; seed_weights[g] = 0.f;
      }
      seed_rows[g] = -1;
    }
  }
}

// one thread per proposal
__global__ void __launch_bounds__(256) refine_assign_kernel(
    const float* __restrict__ boxes, const int64_t* __restrict__ offsets,
    const float* __restrict__ seed_boxes, const int64_t* __restrict__ seed_classes,
    const float* __restrict__ seed_scores, const float* __restrict__ seed_weights,
    const int64_t* __restrict__ seed_offsets, const int64_t* __restrict__ seed_count, int64_t M, int N,
    int64_t num_classes, float thr, int64_t* __restrict__ midx, int8_t* __restrict__ mlabel,
    float* __restrict__ miou, int64_t* __restrict__ gt_classes, float* __restrict__ gt_boxes,
    float* __restrict__ gt_scores, float* __restrict__ gt_weights) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= M) return;
  const int n = find_segment(offsets, N, r);
  const int64_t s0 = seed_offsets[n];
  const int64_t G = seed_count ? seed_count[n] : seed_offsets[n + 1] - s0;
  const float4 b = __ldg(reinterpret_cast<const float4*>(boxes) + r);
  const float4* sb = reinterpret_cast<const float4*>(seed_boxes) + s0;
  int64_t bi = 0;
  float bv = 0.f;
  for (int64_t g = 0; g < G; ++g) {
    const float v = d2_iou(__ldg(sb + g), b);
    if (g == 0 || v > bv) { bv = v; bi = g; }        // max(dim=0): first maximum wins
  }
  const bool lab = G > 0 && bv >= thr;
  midx[r] = bi;
  mlabel[r] = lab ? 1 : 0;
  if (miou) miou[r] = bv;
  float4* gb = reinterpret_cast<float4*>(gt_boxes);
  if (G > 0) {
    gt_classes[r] = lab ? seed_classes[s0 + bi] : num_classes;
    gb[r] = __ldg(sb + bi);
    gt_scores[r] = seed_scores[s0 + bi];
    gt_weights[r] = seed_weights[s0 + bi];
  } else {
    gt_classes[r] = num_classes;
    gb[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    gt_scores[r] = 0.f;
    gt_weights[r] = 0.f;
  }
}

}  // namespace wsovod

using namespace wsovod;

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

WSOVOD_API int wsovod_b200_pgt_top1(const float* scores, int64_t score_stride, const float* boxes,
                                    const int64_t* offsets, const int64_t* gt_classes,
                                    const int64_t* gt_offsets, const float* img_scores, int64_t M,
                                    int64_t N, int64_t K, int64_t G, float* seed_boxes,
                                    int64_t* seed_classes, float* seed_scores, float* seed_weights,
                                    int64_t* seed_rows, int64_t* seed_count, void* stream) {
  if (M < 0 || N < 0 || K < 0 || G < 0 || score_stride < K) return WSOVOD_B200_EINVAL;
  if (G == 0 || N == 0) return 0;
  if (!scores || !boxes || !offsets || !gt_classes || !gt_offsets || !img_scores || !seed_boxes ||
      !seed_classes || !seed_scores || !seed_weights || !seed_rows || !seed_count)
    return WSOVOD_B200_EINVAL;
  if (!aligned16(boxes) || !aligned16(seed_boxes)) return WSOVOD_B200_EALIGN;
  if (N > (1 << 30) || G > (1LL << 31) - 1) return WSOVOD_B200_ETOOBIG;
  pgt_top1_kernel<<<(unsigned)G, kSeedThreads, 0, (cudaStream_t)stream>>>(
      scores, score_stride, boxes, offsets, gt_classes, gt_offsets, img_scores, (int)N, (int)K,
      seed_boxes, seed_classes, seed_scores, seed_weights, seed_rows, seed_count);
  return after_launch();
}

WSOVOD_API int wsovod_b200_refine_assign(const float* boxes, const int64_t* offsets,
                                         const float* seed_boxes, const int64_t* seed_classes,
                                         const float* seed_scores, const float* seed_weights,
                                         const int64_t* seed_offsets, const int64_t* seed_count,
                                         int64_t M, int64_t N, int64_t num_classes, float iou_thresh,
                                         int64_t* matched_idx, int8_t* matched_label,
                                         float* matched_iou, int64_t* gt_classes, float* gt_boxes,
                                         float* gt_scores, float* gt_weights, void* stream) {
  if (M < 0 || N < 0) return WSOVOD_B200_EINVAL;
  if (M == 0 || N == 0) return 0;
  if (!boxes || !offsets || !seed_offsets || !matched_idx || !matched_label || !gt_classes ||
      !gt_boxes || !gt_scores || !gt_weights)
    return WSOVOD_B200_EINVAL;
  if (!aligned16(boxes) || !aligned16(gt_boxes) || (seed_boxes && !aligned16(seed_boxes)))
    return WSOVOD_B200_EALIGN;
  if (N > (1 << 30)) return WSOVOD_B200_ETOOBIG;
  refine_assign_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, (cudaStream_t)stream>>>(
      boxes, offsets, seed_boxes, seed_classes, seed_scores, seed_weights, seed_offsets, seed_count, M,
      (int)N, num_classes, iou_thresh, matched_idx, matched_label, matched_iou, gt_classes, gt_boxes,
      gt_scores, gt_weights);
  return after_launch();
}
