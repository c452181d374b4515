// csc.cu -- csc_forward (wsovod/layers/csc/csc_cuda.cu:183-531; call site
// proposal_generator/proposal_utils.py:272-291): per class with a positive image label, the proposals' contrast
// between the class peak response inside the box's frame and in its context,
//   score = sum_frame / sqrt(area_frame) - sum_context / sqrt(area_context)        (:300-306)
// from an integral image of the binarised response map (:117-147), normalised to [-1, 1] by the largest positive /
// negative score (:466-505) and blended with the image-level prediction (:506-509).
//
// The reference walks (image, class) on the HOST: per pair one device->host copy of the map, a CPU integral image, a
// host->device copy, one kernel, cudaDeviceSynchronize, a device->host copy of W, a CPU normalisation loop and a
// host->device copy back (:397-531).  Every pair rewrites the WHOLE column c of W for ALL rois (the roi's batch
// index is never read, :200-204), so only the LAST image whose label for c is positive survives.  Here: three
// stream-ordered launches, no host round trip --
//   csc_integral_kernel   one CTA per class: picks that image, binarises its map and builds the integral image
//                         (row prefix sums, then a column walk: the same float additions as :117-147);
//   csc_score_kernel      one thread per (roi, class): the score, with the reference's mix of int / float / double
//                         arithmetic statement for statement (:200-306), and the column's max / min (ordered-int atomics);
//   csc_blend_kernel      normalisation cases and the blend with `preds`, every product / sum rounded separately
//                         like the reference's host loop.
#include "common.cuh"

namespace wsovod {

__device__ __forceinline__ int f2ord(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void __launch_bounds__(256) csc_integral_kernel(const float* __restrict__ cpgs, const float* __restrict__ labels, int B,
                                                           int K, int H, int W, float thr, int* __restrict__ sel,
                                                           float* __restrict__ integ, int* __restrict__ minmax) {
  const int c = blockIdx.x;
  __shared__ int s_b;
  if (threadIdx.x == 0) {
    int b = -1;
    for (int i = 0; i < B; ++i)
      if (!(labels[(int64_t)i * K + c] < 0.5f)) b = i;        // `if (label_value < 0.5) continue;` (:422)
    s_b = b;
    sel[c] = b;
    minmax[2 * c] = f2ord(0.f);                               // max_value = 0, min_value = 0 (:469-470)
    minmax[2 * c + 1] = f2ord(0.f);
  }
  __syncthreads();
  const int b = s_b;
  if (b < 0) return;
  const float* m = cpgs + ((int64_t)b * K + c) * H * W;
  float* out = integ + (int64_t)c * H * W;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  // rows: s[y][x] = #(m[y][0..x] >= thr), exact small integers in fp32
  for (int y = warp; y < H; y += nw) {
    float carry = 0.f;
    for (int x0 = 0; x0 < W; x0 += 32) {
      const int x = x0 + lane;
      float v = (x < W && m[(int64_t)y * W + x] >= thr) ? 1.f : 0.f;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
      }
      v += carry;
      if (x < W) out[(int64_t)y * W + x] = v;
      carry = __shfl_sync(0xffffffffu, v, 31);
    }
  }
  __syncthreads();
  // columns: sum[y][x] = sum[y-1][x] + s[y][x] (:139-146)
  for (int x = threadIdx.x; x < W; x += blockDim.x) {
    float acc = out[x];
    for (int y = 1; y < H; ++y) {
      acc = __fadd_rn(acc, out[(int64_t)y * W + x]);
      out[(int64_t)y * W + x] = acc;
    }
  }
}

__device__ __forceinline__ float csc_box_sum(const float* __restrict__ d, int W, int ws, int hs, int we, int he) {
  const float a1 = d[he * W + we];
  const float a2 = (ws - 1 >= 0) ? d[he * W + (ws - 1)] : 0.f;
  const float a3 = (hs - 1 >= 0) ? d[(hs - 1) * W + we] : 0.f;
  const float a4 = (hs - 1 >= 0 && ws - 1 >= 0) ? d[(hs - 1) * W + (ws - 1)] : 0.f;
  return __fadd_rn(__fsub_rn(__fsub_rn(a1, a2), a3), a4);     // a1 - a2 - a3 + a4
}

__global__ void __launch_bounds__(256) csc_score_kernel(const float* __restrict__ integ, const int* __restrict__ sel,
                                                        const float* __restrict__ rois, int R, int K, int H, int W,
                                                        int area_sqrt, float context_scale, float* __restrict__ Wout,
                                                        int* __restrict__ minmax) {
  const int c = blockIdx.y;
  if (sel[c] < 0) return;
  const float* d = integ + (int64_t)c * H * W;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  float score = 0.f;
  if (r < R) {
    const float* roi = rois + (int64_t)r * 5;
    int wstart = (int)roundf(roi[1]), hstart = (int)roundf(roi[2]), wend = (int)roundf(roi[3]), hend = (int)roundf(roi[4]);
    wstart = max(min(wstart, W - 1), 0);
    hstart = max(min(hstart, H - 1), 0);
    wend = max(min(wend, W - 1), 0);
    hend = max(min(hend, H - 1), 0);
    float width_roi = (float)(wend - wstart), height_roi = (float)(hend - hstart);
    // `1.0 * x / s`, `1.0 * x * s`, `1.0 * (a + b) / 2.0`: double arithmetic narrowed to float (:228-235)
    float width_inner = (float)(1.0 * width_roi / context_scale), height_inner = (float)(1.0 * height_roi / context_scale);
    float width_outer = (float)(1.0 * width_roi * context_scale), height_outer = (float)(1.0 * height_roi * context_scale);
    const float wcenter = (float)(1.0 * (wend + wstart) / 2.0), hcenter = (float)(1.0 * (hend + hstart) / 2.0);
    const int ws_in = (int)round(wcenter - width_inner / 2.0), hs_in = (int)round(hcenter - height_inner / 2.0);
    int we_in = (int)round(wcenter + width_inner / 2.0), he_in = (int)round(hcenter + height_inner / 2.0);
    we_in = min(we_in, W - 1); he_in = min(he_in, H - 1);      // memory safety for context_scale < 1 (a no-op otherwise)
    const int ws_out = (int)round(fmax(wcenter - width_outer / 2.0, 0.0)), hs_out = (int)round(fmax(hcenter - height_outer / 2.0, 0.0));
    const int we_out = (int)round(fmin(wcenter + width_outer / 2.0, W - 1.0)), he_out = (int)round(fmin(hcenter + height_outer / 2.0, H - 1.0));
    width_roi = (float)(wend - wstart + 1); height_roi = (float)(hend - hstart + 1);
    width_inner = (float)(we_in - ws_in + 1); height_inner = (float)(he_in - hs_in + 1);
    width_outer = (float)(we_out - ws_out + 1); height_outer = (float)(he_out - hs_out + 1);
    const float sum_roi = csc_box_sum(d, W, wstart, hstart, wend, hend);
    const float sum_inner = csc_box_sum(d, W, ws_in, hs_in, we_in, he_in);
    const float sum_outer = csc_box_sum(d, W, ws_out, hs_out, we_out, he_out);
    const float area_roi = __fmul_rn(height_roi, width_roi), area_inner = __fmul_rn(height_inner, width_inner);
    const float area_outer = __fmul_rn(height_outer, width_outer);
    const float area_frame = fmaxf(__fsub_rn(area_roi, area_inner), 1.f), area_context = fmaxf(__fsub_rn(area_outer, area_roi), 1.f);
    const float sum_frame = __fsub_rn(sum_roi, sum_inner), sum_context = __fsub_rn(sum_outer, sum_roi);
    if (area_sqrt)
      score = __fsub_rn(__fdiv_rn(sum_frame, __fsqrt_rn(area_frame)), __fdiv_rn(sum_context, __fsqrt_rn(area_context)));
    else
      score = __fsub_rn(__fdiv_rn(sum_frame, area_frame), __fdiv_rn(sum_context, area_context));
    Wout[(int64_t)r * K + c] = score;
  }
  // column max / min over the rois (both start at 0, :469-479); a NaN score never wins either compare
  float mx = (r < R && score > 0.f) ? score : 0.f, mn = (r < R && score < 0.f) ? score : 0.f;
  mx = warp_max(mx);
  mn = -warp_max(-mn);
  if ((threadIdx.x & 31) == 0) {
    if (mx > 0.f) atomicMax(&minmax[2 * c], f2ord(mx));
    if (mn < 0.f) atomicMin(&minmax[2 * c + 1], f2ord(mn));
  }
}

__global__ void csc_blend_kernel(const int* __restrict__ sel, const int* __restrict__ minmax, const float* __restrict__ preds,
                                 int R, int K, float* __restrict__ Wout) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)R * K) return;
  const int c = (int)(i % K);
  const int b = sel[c];
  if (b < 0) { Wout[i] = 1.f; return; }                        // W = at::ones (:382)
  const float mx = ord2f(minmax[2 * c]), mn = ord2f(minmax[2 * c + 1]);
  float v = Wout[i];
  if (mx > 0.f && mn < 0.f) v = v > 0.f ? __fdiv_rn(v, mx) : __fdiv_rn(v, -mn);     // :480-489
  else if (mx > 0.f && mn == 0.f) v = __fdiv_rn(v, mx);                             // :490-499
  else v = 1.f;                                                                     // :500-504
  const float p = preds[(int64_t)b * K + c];
  Wout[i] = __fadd_rn(__fmul_rn(p, v), __fmul_rn(__fsub_rn(1.f, p), 1.f));          // pred * W + (1 - pred) * 1 (:506-509)
}

}  // namespace wsovod

using namespace wsovod;

WSOVOD_API size_t wsovod_b200_csc_workspace(int64_t K, int64_t H, int64_t W) {
  if (K < 0 || H < 0 || W < 0) return 0;
  return align_up(sizeof(float) * (size_t)(K * H * W), 256) + align_up(sizeof(int) * 3 * (size_t)K, 256);
}

WSOVOD_API int wsovod_b200_csc_fwd(const float* cpgs, const float* labels, const float* preds, const float* rois,
                                   int64_t B, int64_t K, int64_t H, int64_t W, int64_t R, float fg_threshold, int area_sqrt,
                                   float context_scale, float* Wout, void* workspace, size_t workspace_bytes, void* stream) {
  if (B < 0 || K < 0 || H < 0 || W < 0 || R < 0) return WSOVOD_B200_EINVAL;
  if (R == 0 || K == 0) return 0;
  if (!Wout || !labels || !preds || (B > 0 && (!cpgs || !rois))) return WSOVOD_B200_EINVAL;
  if (H * W >= (1LL << 24) || R * K >= (1LL << 40) || K > 65535) return WSOVOD_B200_ETOOBIG;   // counts stay exact in fp32
  if (!workspace || workspace_bytes < wsovod_b200_csc_workspace(K, H, W)) return WSOVOD_B200_EWORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  float* integ = (float*)workspace;
  int* sel = (int*)((char*)workspace + align_up(sizeof(float) * (size_t)(K * H * W), 256));
  int* minmax = sel + K;
  int rc;
  if (B == 0 || H == 0 || W == 0) {   // no map: nothing selected, W stays at::ones
    cudaError_t e = cudaMemsetAsync(sel, 0xff, sizeof(int) * (size_t)K, st);
    if (e != cudaSuccess) return (int)e;
  } else {
    csc_integral_kernel<<<(unsigned)K, 256, 0, st>>>(cpgs, labels, (int)B, (int)K, (int)H, (int)W, 1.f * fg_threshold, sel, integ, minmax);
    if ((rc = after_launch())) return rc;
    csc_score_kernel<<<dim3((unsigned)ceil_div(R, 256), (unsigned)K), 256, 0, st>>>(integ, sel, rois, (int)R, (int)K, (int)H, (int)W,
                                                                                     area_sqrt, context_scale, Wout, minmax);
    if ((rc = after_launch())) return rc;
  }
  csc_blend_kernel<<<(unsigned)ceil_div(R * K, 256), 256, 0, st>>>(sel, minmax, preds, (int)R, (int)K, Wout);
  return after_launch();
}
