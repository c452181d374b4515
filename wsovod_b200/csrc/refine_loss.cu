// refine_loss.cu -- SURVEY 8f-2: the weighted refinement losses that consume kernel (3)'s labels,
// InstanceRefinementOutputLayers.losses with cross_entropy_weighted and "smooth_l1_weighted"
// (fast_rcnn_open_vocabulary.py:754-892), forward and backward, one pass over the logits each.
//
//   w_i      = gt_classes_i == -1 ? 0 : gt_weights_i                                   (:790-792)
//   valid    = #(w_i > 1e-12)                                                          (:794-795)
//   loss_cls = sum_i w_i * CE(logits_i, gt_i; ignore_index = -1) / valid               (:813-820)
//   loss_box = sum_{i fg} w_i * sum_j smooth_l1(delta_ij - target_ij; beta) / max(M, 1) (:864-892)
//              fg: 0 <= gt_i < num_classes; target = Box2BoxTransform.get_deltas(proposal, gt box)
//              (detectron2 box_regression.py: dx = wx (gcx - pcx) / pw, dw = ww log(gw / pw));
//              a NaN target anywhere makes the reference return zeros(1) (:869-872): loss 0, no gradient.
//
// The reference runs ~25 element-wise / indexing launches for this; here one warp owns a row (max,
// sum-exp, the CE term, the four box terms), CTAs write partial sums, and a second tiny launch adds them
// in a fixed order in double precision (deterministic; no float atomics).
#include "common.cuh"

#include <math.h>

#include <algorithm>

namespace wsovod {

constexpr int kRlThreads = 256;
constexpr int kRlRows = kRlThreads / 32;     // rows per CTA

__device__ __forceinline__ void target_deltas(const float4 p, const float4 g, float wx, float wy, float ww, float wh,
                                              float* t) {
  const float pw = p.z - p.x, ph = p.w - p.y;
  const float pcx = p.x + 0.5f * pw, pcy = p.y + 0.5f * ph;
  const float gw = g.z - g.x, gh = g.w - g.y;
  const float gcx = g.x + 0.5f * gw, gcy = g.y + 0.5f * gh;
  t[0] = wx * (gcx - pcx) / pw;
  t[1] = wy * (gcy - pcy) / ph;
  t[2] = ww * logf(gw / pw);
  t[3] = wh * logf(gh / ph);
}

__device__ __forceinline__ float smooth_l1(float d, float beta) {
  const float n = fabsf(d);
  if (beta < 1e-5f) return n;                                   // fvcore smooth_l1_loss: plain L1
  return n < beta ? 0.5f * n * n / beta : n - 0.5f * beta;
}
__device__ __forceinline__ float smooth_l1_grad(float d, float beta) {
  const float s = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);       // torch.abs: zero gradient at zero
  if (beta < 1e-5f) return s;
  return fabsf(d) < beta ? d / beta : s;
}

// partial[blockIdx][4] = { sum w CE, #(w > 1e-12), sum w smooth_l1, #NaN targets } over the CTA's rows
__global__ void __launch_bounds__(kRlThreads) refine_loss_rows_kernel(
    const float* __restrict__ logits, int K1, const int64_t* __restrict__ gt_classes, const float* __restrict__ gt_weights,
    const float* __restrict__ proposal_boxes, const float* __restrict__ gt_boxes, const float* __restrict__ deltas,
    int dcols, int64_t M, int num_classes, float wx, float wy, float ww, float wh, float beta,
    float* __restrict__ lse, double* __restrict__ partial) {
  __shared__ float s_part[kRlRows][4];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t r = (int64_t)blockIdx.x * kRlRows + wid;
  float ce_w = 0.f, valid = 0.f, box_w = 0.f, nan_t = 0.f;
  if (r < M) {
    const float* row = logits + r * K1;
    float mx = -INFINITY;
    for (int k = lane; k < K1; k += 32) mx = fmaxf(mx, __ldg(row + k));
    mx = warp_max(mx);
    float se = 0.f;
    for (int k = lane; k < K1; k += 32) se += expf(__ldg(row + k) - mx);
    se = warp_sum(se);
    const float l = mx + logf(se);
    if (lane == 0) {
      lse[r] = l;
      const int64_t gt = gt_classes[r];
      const float w = gt == -1 ? 0.f : gt_weights[r];
      valid = w > 1e-12f ? 1.f : 0.f;
      if (gt >= 0 && gt < K1) ce_w = (l - __ldg(row + gt)) * w;
      if (dcols > 0 && gt >= 0 && gt < num_classes) {
        float t[4];
        target_deltas(__ldg(reinterpret_cast<const float4*>(proposal_boxes) + r),
                      __ldg(reinterpret_cast<const float4*>(gt_boxes) + r), wx, wy, ww, wh, t);
        const float* d = deltas + r * dcols + (dcols == 4 ? 0 : 4 * gt);
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (isnan(t[j])) nan_t += 1.f;
          s += smooth_l1(__ldg(d + j) - t[j], beta) * w;
        }
        box_w = s;
      }
    }
  }
  if (lane == 0) { s_part[wid][0] = ce_w; s_part[wid][1] = valid; s_part[wid][2] = box_w; s_part[wid][3] = nan_t; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double a = 0.0;
    for (int w = 0; w < kRlRows; ++w) a += (double)s_part[w][threadIdx.x];
    partial[(int64_t)blockIdx.x * 4 + threadIdx.x] = a;
  }
}

// out[0] = loss_cls, out[1] = loss_box, out[2] = valid count, out[3] = 1 if a target was NaN
__global__ void __launch_bounds__(256) refine_loss_finish_kernel(const double* __restrict__ partial, int64_t nblocks,
                                                                 int64_t M, float* __restrict__ out) {
  __shared__ double s_acc[256][4];
  double a[4] = {0.0, 0.0, 0.0, 0.0};
  for (int64_t b = threadIdx.x; b < nblocks; b += 256)
#pragma unroll
    for (int j = 0; j < 4; ++j) a[j] += partial[b * 4 + j];
#pragma unroll
  for (int j = 0; j < 4; ++j) s_acc[threadIdx.x][j] = a[j];
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o)
#pragma unroll
      for (int j = 0; j < 4; ++j) s_acc[threadIdx.x][j] += s_acc[threadIdx.x + o][j];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double ce = s_acc[0][0], valid = s_acc[0][1], box = s_acc[0][2], nans = s_acc[0][3];
    out[0] = (float)(ce / valid);                                          // 0 / 0 = NaN like the reference
    out[1] = nans > 0.0 ? 0.f : (float)(box / (double)(M > 1 ? M : 1));
    out[2] = (float)valid;
    out[3] = nans > 0.0 ? 1.f : 0.f;
  }
}

// one warp per row: d loss / d logits and d loss / d deltas
__global__ void __launch_bounds__(kRlThreads) refine_loss_bwd_kernel(
    const float* __restrict__ grad_out, const float* __restrict__ fwd_out, const float* __restrict__ logits, int K1,
    const float* __restrict__ lse, const int64_t* __restrict__ gt_classes, const float* __restrict__ gt_weights,
    const float* __restrict__ proposal_boxes, const float* __restrict__ gt_boxes, const float* __restrict__ deltas,
    int dcols, int64_t M, int num_classes, float wx, float wy, float ww, float wh, float beta,
    float* __restrict__ grad_logits, float* __restrict__ grad_deltas) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t r = (int64_t)blockIdx.x * kRlRows + wid;
  if (r >= M) return;
  const int64_t gt = gt_classes[r];
  const float w = gt == -1 ? 0.f : gt_weights[r];
  const bool counted = gt >= 0 && gt < K1;
  if (grad_logits) {
    const float g = counted ? grad_out[0] * w / fwd_out[2] : 0.f;
    const float l = lse[r];
    const float* row = logits + r * K1;
    float* go = grad_logits + r * K1;
    for (int k = lane; k < K1; k += 32) {
      // an ignored row contributes exactly zero, whatever its logits hold
      go[k] = counted ? g * (expf(__ldg(row + k) - l) - (k == gt ? 1.f : 0.f)) : 0.f;
    }
  }
  if (grad_deltas && dcols > 0) {
    float* gd = grad_deltas + r * dcols;
    const bool fg = gt >= 0 && gt < num_classes && fwd_out[3] == 0.f;
    float t[4] = {0.f, 0.f, 0.f, 0.f};
    if (fg) target_deltas(__ldg(reinterpret_cast<const float4*>(proposal_boxes) + r),
                          __ldg(reinterpret_cast<const float4*>(gt_boxes) + r), wx, wy, ww, wh, t);
    const float g = fg ? grad_out[1] * w / (float)(M > 1 ? M : 1) : 0.f;
    const int c0 = dcols == 4 ? 0 : 4 * (int)(fg ? gt : 0);
    for (int c = lane; c < dcols; c += 32) {
      float v = 0.f;
      if (fg && c >= c0 && c < c0 + 4) v = g * smooth_l1_grad(__ldg(deltas + r * dcols + c) - t[c - c0], beta);
      gd[c] = v;
    }
  }
}

}  // namespace wsovod

using namespace wsovod;

WSOVOD_API size_t wsovod_b200_refine_loss_workspace(int64_t M) {
  if (M < 0) return 0;
  return sizeof(double) * 4 * (size_t)std::max<int64_t>(1, ceil_div(M, kRlRows));
}

static int rl_check(const float* logits, int64_t K1, const int64_t* gt_classes, const float* gt_weights,
                    const float* proposal_boxes, const float* gt_boxes, const float* deltas, int64_t dcols, int64_t M,
                    int64_t num_classes) {
  if (M < 0 || K1 < 1 || num_classes < 0 || dcols < 0) return WSOVOD_B200_EINVAL;
  if (dcols != 0 && dcols != 4 && dcols != 4 * num_classes) return WSOVOD_B200_EINVAL;
  if (M > 0 && (!logits || !gt_classes || !gt_weights)) return WSOVOD_B200_EINVAL;
  if (M > 0 && dcols > 0 && (!proposal_boxes || !gt_boxes || !deltas)) return WSOVOD_B200_EINVAL;
  if (dcols > 0 && ((((uintptr_t)proposal_boxes) | ((uintptr_t)gt_boxes)) & 15)) return WSOVOD_B200_EALIGN;
  if (K1 >= (1LL << 30) || M >= (1LL << 40)) return WSOVOD_B200_ETOOBIG;
  return 0;
}

WSOVOD_API int wsovod_b200_refine_loss_fwd(const float* logits, int64_t K1, const int64_t* gt_classes,
                                           const float* gt_weights, const float* proposal_boxes, const float* gt_boxes,
                                           const float* deltas, int64_t dcols, int64_t M, int64_t num_classes,
                                           float wx, float wy, float ww, float wh, float beta, float* out,
                                           float* lse, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = rl_check(logits, K1, gt_classes, gt_weights, proposal_boxes, gt_boxes, deltas, dcols, M, num_classes);
  if (rc) return rc;
  if (!out || (M > 0 && !lse)) return WSOVOD_B200_EINVAL;
  if (!workspace || workspace_bytes < wsovod_b200_refine_loss_workspace(M)) return WSOVOD_B200_EWORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t nblocks = ceil_div(M, kRlRows);
  double* partial = (double*)workspace;
  if (nblocks > 0) {
    refine_loss_rows_kernel<<<(unsigned)nblocks, kRlThreads, 0, st>>>(
        logits, (int)K1, gt_classes, gt_weights, proposal_boxes, gt_boxes, deltas, (int)dcols, M, (int)num_classes,
        wx, wy, ww, wh, beta, lse, partial);
    if ((rc = after_launch())) return rc;
  }
  refine_loss_finish_kernel<<<1, 256, 0, st>>>(partial, nblocks, M, out);
  return after_launch();
}

WSOVOD_API int wsovod_b200_refine_loss_bwd(const float* grad_out, const float* fwd_out, const float* logits, int64_t K1,
                                           const float* lse, const int64_t* gt_classes, const float* gt_weights,
                                           const float* proposal_boxes, const float* gt_boxes, const float* deltas,
                                           int64_t dcols, int64_t M, int64_t num_classes, float wx, float wy, float ww,
                                           float wh, float beta, float* grad_logits, float* grad_deltas, void* stream) {
  int rc = rl_check(logits, K1, gt_classes, gt_weights, proposal_boxes, gt_boxes, deltas, dcols, M, num_classes);
  if (rc) return rc;
  if (!grad_out || !fwd_out || (M > 0 && !lse)) return WSOVOD_B200_EINVAL;
  if (M == 0 || (!grad_logits && !grad_deltas)) return 0;
  refine_loss_bwd_kernel<<<(unsigned)ceil_div(M, kRlRows), kRlThreads, 0, (cudaStream_t)stream>>>(
      grad_out, fwd_out, logits, (int)K1, lse, gt_classes, gt_weights, proposal_boxes, gt_boxes, deltas, (int)dcols, M,
      (int)num_classes, wx, wy, ww, wh, beta, grad_logits, grad_deltas);
  return after_launch();
}
