// mil.cu -- kernel (2b): WSDDN-style two-stream MIL score, batched over all images of the step.
//
//   scores[r,k] = softmax_k(cls[r,:])[k] * softmax_{r in image}(det[:,k])[r]
//   img[n,k]    = clamp(sum_r scores[r,k], 1e-6, 1-1e-6)
// (roi_heads/fast_rcnn_open_vocabulary.py:338-357,604-618: there a Python loop over images issuing ~6
// tiny kernels each.)  Here: four launches for the whole batch, all HBM-streaming:
//   1. mil_colstats : per (image, row-chunk) online (max, sum exp) of every det column   -> partials
//   1b. mil_colmerge: merges the partials once per image                                  -> (max, 1/sum)
//   2. mil_scores   : one warp per proposal row computes
//                     the row softmax with shuffle reductions, multiplies, stores, and accumulates the
//                     per-column score sums in a fixed order                              -> partials
//   3. mil_finalize : img = clamp(sum of partials)
// Deterministic: no floating-point atomics anywhere.
#include "common.cuh"

#include <algorithm>

namespace wsovod {

constexpr int kMilThreads = 256;
constexpr int kMilWarps = kMilThreads / 32;

struct MilPlan {
  int chunks;        // row chunks per image
  size_t off_stats;  // float2 [N, chunks, K]
  size_t off_sums;   // float  [N, chunks, K]
  size_t off_asum;   // float  [N, chunks, K] (backward)
  size_t off_cstat;  // float2 [N, K] merged column statistics (max, 1 / sum exp)
  size_t bytes;
};

static MilPlan mil_plan(int64_t M, int64_t N, int64_t K) {
  MilPlan p;
  int64_t avg = N > 0 ? ceil_div(M, N) : 1;
  // enough chunks to spread one image over the whole GPU (c3: one image of 5024 rows -> 157 CTAs), each
  // chunk at least 32 rows; the per-column partials are merged ONCE per image (mil_colmerge_kernel), not
  // by every CTA of the scores pass
  p.chunks = (int)std::max<int64_t>(1, std::min<int64_t>(160, ceil_div(avg, 32)));
  size_t o = 0;
  p.off_stats = o; o += align_up(sizeof(float2) * (size_t)(N * p.chunks * K), 256);
  p.off_sums = o;  o += align_up(sizeof(float) * (size_t)(N * p.chunks * K), 256);
  p.off_asum = o;  o += align_up(sizeof(float) * (size_t)(N * p.chunks * K), 256);
  p.off_cstat = o; o += align_up(sizeof(float2) * (size_t)(N * K), 256);
  p.bytes = o;
  return p;
}

__device__ __forceinline__ void chunk_rows(const int64_t* offsets, int n, int chunk, int chunks,
                                           int64_t& r0, int64_t& r1) {
  const int64_t a = offsets[n], b = offsets[n + 1];
  const int64_t per = (b - a + chunks - 1) / chunks;
  r0 = a + (int64_t)chunk * per;
  r1 = r0 + per < b ? r0 + per : b;
  if (r0 > b) r0 = b;
}

// online softmax statistics merge
__device__ __forceinline__ void merge_ms(float& m, float& s, float m2, float s2) {
  if (s2 == 0.f) return;
  if (s == 0.f) { m = m2; s = s2; return; }
  const float mx = fmaxf(m, m2);
  s = s * expf(m - mx) + s2 * expf(m2 - mx);
  m = mx;
}

// 1. column statistics of `det`.  Threads are laid out (row-group, column) so that a warp reads
//    consecutive floats of a row.
__global__ void __launch_bounds__(kMilThreads) mil_colstats_kernel(
    const float* __restrict__ det, const int64_t* __restrict__ offsets, int K, int chunks,
    float2* __restrict__ stats) {
  extern __shared__ float2 sh[];   // [RG][KT]
  const int n = blockIdx.y, chunk = blockIdx.x;
  int64_t r0, r1;
  chunk_rows(offsets, n, chunk, chunks, r0, r1);
  const int KT = min(K, kMilThreads);
  const int RG = kMilThreads / KT;
  const int tk = threadIdx.x % KT, rg = threadIdx.x / KT;
  for (int k0 = 0; k0 < K; k0 += KT) {
    const int k = k0 + tk;
    float m = 0.f, s = 0.f;
    if (rg < RG && k < K)
      for (int64_t r = r0 + rg; r < r1; r += RG) {
        const float v = __ldg(det + r * K + k);
        if (s == 0.f) { m = v; s = 1.f; }
        else if (v <= m) s += expf(v - m);
        else { s = s * expf(m - v) + 1.f; m = v; }
      }
    if (rg < RG) sh[rg * KT + tk] = make_float2(m, s);
    __syncthreads();
    if (rg == 0 && k < K) {
      for (int g = 1; g < RG; ++g) merge_ms(m, s, sh[g * KT + tk].x, sh[g * KT + tk].y);
      stats[((int64_t)n * chunks + chunk) * K + k] = make_float2(m, s);
    }
    __syncthreads();
  }
}

// 1b. merge the row-chunk partials of every det column: cstat[n, k] = (max, 1 / sum exp).  One warp per
//     column: lanes take the chunks 32 apart, then a shuffle tree (fixed order: deterministic)
__global__ void mil_colmerge_kernel(const float2* __restrict__ stats, int K, int chunks, float2* __restrict__ cstat) {
  const int n = blockIdx.y;
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (k >= K) return;
  float m = 0.f, s = 0.f;
  for (int c = lane; c < chunks; c += 32) {
    const float2 t = stats[((int64_t)n * chunks + c) * K + k];
    merge_ms(m, s, t.x, t.y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    merge_ms(m, s, m2, s2);
  }
  if (lane == 0) cstat[(int64_t)n * K + k] = make_float2(m, s > 0.f ? 1.f / s : 0.f);
}

// 2. scores.  smem: colm[K], cinv[K] (merged det column stats), wsum[kMilWarps][K] (per-warp sums)
template <bool BWD>
__global__ void __launch_bounds__(kMilThreads) mil_scores_kernel(
    const float* __restrict__ cls, const float* __restrict__ det, const int64_t* __restrict__ offsets,
    int K, int chunks, const float2* __restrict__ stats, float* __restrict__ scores,
    float* __restrict__ sums,
    // backward only
    const float* __restrict__ grad_scores, const float* __restrict__ grad_img,
    const float* __restrict__ asum_total, float* __restrict__ grad_cls, float* __restrict__ grad_det,
    int bwd_phase) {
  extern __shared__ float shf[];
  float* colm = shf;
  float* cinv = shf + K;
  float* wsum = shf + 2 * K;
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int k = threadIdx.x; k < K; k += kMilThreads) {
    const float2 t = stats[(int64_t)n * K + k];     // merged by mil_colmerge_kernel
    colm[k] = t.x;
    cinv[k] = t.y;
  }
  for (int i = threadIdx.x; i < kMilWarps * K; i += kMilThreads) wsum[i] = 0.f;
  __syncthreads();
  int64_t r0, r1;
  chunk_rows(offsets, n, chunk, chunks, r0, r1);
  float* mysum = wsum + wid * K;
  for (int64_t r = r0 + wid; r < r1; r += kMilWarps) {
    const float* c = cls + r * K;
    float mx = K == 1 ? 0.f : -FLT_MAX;    // K==1: implicit extra zero logit (:338-340)
    for (int k = lane; k < K; k += 32) mx = fmaxf(mx, __ldg(c + k));
    mx = warp_max(mx);
    float se = 0.f;
    for (int k = lane; k < K; k += 32) se += expf(__ldg(c + k) - mx);
    se = warp_sum(se);
    if (K == 1) se += expf(0.f - mx);
    const float rinv = 1.f / se;
    if (!BWD) {
      for (int k = lane; k < K; k += 32) {
        const float pc = expf(__ldg(c + k) - mx) * rinv;
        const float pd = expf(__ldg(det + r * K + k) - colm[k]) * cinv[k];
        const float s = pc * pd;
        scores[r * K + k] = s;
        mysum[k] += s;
      }
    } else {
      // a = G * S with G = grad_scores + grad_img[n];  phase 0: column sums of a;
      // phase 1: dC = a - pc * sum_k a ;  dD = a - pd * sum_{rows of image} a
      float rowa = 0.f;
      for (int k = lane; k < K; k += 32) {
        const float pc = expf(__ldg(c + k) - mx) * rinv;
        const float pd = expf(__ldg(det + r * K + k) - colm[k]) * cinv[k];
        float g = grad_scores ? __ldg(grad_scores + r * K + k) : 0.f;
        if (grad_img) g += __ldg(grad_img + (int64_t)n * K + k);
        const float a = g * pc * pd;
        rowa += a;
        if (bwd_phase == 0) mysum[k] += a;
      }
      if (bwd_phase == 1) {
        rowa = warp_sum(rowa);
        for (int k = lane; k < K; k += 32) {
          const float pc = expf(__ldg(c + k) - mx) * rinv;
          const float pd = expf(__ldg(det + r * K + k) - colm[k]) * cinv[k];
          float g = grad_scores ? __ldg(grad_scores + r * K + k) : 0.f;
          if (grad_img) g += __ldg(grad_img + (int64_t)n * K + k);
          const float a = g * pc * pd;
          grad_cls[r * K + k] = a - pc * rowa;
          grad_det[r * K + k] = a - pd * asum_total[(int64_t)n * K + k];
        }
      }
    }
  }
  __syncthreads();
  if (!BWD || bwd_phase == 0)
    for (int k = threadIdx.x; k < K; k += kMilThreads) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kMilWarps; ++w) t += wsum[w * K + k];
      sums[((int64_t)n * chunks + chunk) * K + k] = t;
    }
}

// 3. finalize: out[n,k] = (clamp) sum over chunks.  One warp per (n, k): lanes add the chunks 32 apart, then a
//    shuffle tree -- a fixed summation order, so the result is deterministic
__global__ void mil_finalize_kernel(const float* __restrict__ sums, int64_t NK, int K, int chunks,
                                    int do_clamp, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= NK) return;
  const int64_t n = i / K, k = i - n * K;
  float t = 0.f;
  for (int c = lane; c < chunks; c += 32) t += sums[(n * chunks + c) * K + k];
  t = warp_sum(t);
  if (lane == 0) out[i] = do_clamp ? fminf(fmaxf(t, 1e-6f), 1.0f - 1e-6f) : t;
}

}  // namespace wsovod

using namespace wsovod;

WSOVOD_API size_t wsovod_b200_mil_workspace(int64_t M, int64_t N, int64_t K) {
  if (M < 0 || N < 0 || K < 0) return 0;
  return mil_plan(M, N, K).bytes + align_up(sizeof(float) * (size_t)(N * K), 256);
}

static int mil_check(int64_t M, int64_t N, int64_t K) {
  if (M < 0 || N < 0 || K < 0) return WSOVOD_B200_EINVAL;
  if (K > 4096 || N > 65535) return WSOVOD_B200_ETOOBIG;
  return 0;
}

WSOVOD_API int wsovod_b200_mil_fwd(const float* cls, const float* det, const int64_t* offsets, int64_t M,
                                   int64_t N, int64_t K, float* scores, float* img_scores,
                                   void* workspace, size_t workspace_bytes, void* stream) {
  int rc = mil_check(M, N, K);
  if (rc) return rc;
  if (N == 0 || K == 0) return 0;
  if (!offsets || (M > 0 && (!cls || !det || !scores))) return WSOVOD_B200_EINVAL;
  const MilPlan pl = mil_plan(M, N, K);
  if (!workspace || workspace_bytes < wsovod_b200_mil_workspace(M, N, K)) return WSOVOD_B200_EWORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  float2* stats = (float2*)(ws + pl.off_stats);
  float* sums = (float*)(ws + pl.off_sums);
  dim3 grid(pl.chunks, (unsigned)N);
  float2* cstat = (float2*)(ws + pl.off_cstat);
  mil_colstats_kernel<<<grid, kMilThreads, sizeof(float2) * kMilThreads, st>>>(det, offsets, (int)K, pl.chunks, stats);
  if ((rc = after_launch())) return rc;
  mil_colmerge_kernel<<<dim3((unsigned)ceil_div(K, 8), (unsigned)N), 256, 0, st>>>(stats, (int)K, pl.chunks, cstat);
  if ((rc = after_launch())) return rc;
  const size_t smem = sizeof(float) * (size_t)K * (2 + kMilWarps);
  if (smem > 32 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(mil_scores_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  mil_scores_kernel<false><<<grid, kMilThreads, smem, st>>>(cls, det, offsets, (int)K, pl.chunks, cstat, scores, sums,
                                                         nullptr, nullptr, nullptr, nullptr, nullptr, 0);
  if ((rc = after_launch())) return rc;
  if (img_scores) {
    const int64_t NK = N * K;
    mil_finalize_kernel<<<(unsigned)ceil_div(NK, 8), 256, 0, st>>>(sums, NK, (int)K, pl.chunks, 1, img_scores);
    if ((rc = after_launch())) return rc;
  }
  return 0;
}

WSOVOD_API int wsovod_b200_mil_bwd(const float* grad_scores, const float* grad_img, const float* cls,
                                   const float* det, const int64_t* offsets, int64_t M, int64_t N,
                                   int64_t K, float* grad_cls, float* grad_det, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  int rc = mil_check(M, N, K);
  if (rc) return rc;
  if (N == 0 || K == 0 || M == 0) return 0;
  if (!offsets || !cls || !det || !grad_cls || !grad_det) return WSOVOD_B200_EINVAL;
  const MilPlan pl = mil_plan(M, N, K);
  if (!workspace || workspace_bytes < wsovod_b200_mil_workspace(M, N, K)) return WSOVOD_B200_EWORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  float2* stats = (float2*)(ws + pl.off_stats);
  float* asum = (float*)(ws + pl.off_asum);
  float* atot = (float*)(ws + pl.bytes);
  dim3 grid(pl.chunks, (unsigned)N);
  float2* cstat = (float2*)(ws + pl.off_cstat);
  mil_colstats_kernel<<<grid, kMilThreads, sizeof(float2) * kMilThreads, st>>>(det, offsets, (int)K, pl.chunks, stats);
  if ((rc = after_launch())) return rc;
  mil_colmerge_kernel<<<dim3((unsigned)ceil_div(K, 8), (unsigned)N), 256, 0, st>>>(stats, (int)K, pl.chunks, cstat);
  if ((rc = after_launch())) return rc;
  const size_t smem = sizeof(float) * (size_t)K * (2 + kMilWarps);
  if (smem > 32 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(mil_scores_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  mil_scores_kernel<true><<<grid, kMilThreads, smem, st>>>(cls, det, offsets, (int)K, pl.chunks, cstat, nullptr, asum,
                                                        grad_scores, grad_img, nullptr, nullptr, nullptr, 0);
  if ((rc = after_launch())) return rc;
  const int64_t NK = N * K;
  mil_finalize_kernel<<<(unsigned)ceil_div(NK, 8), 256, 0, st>>>(asum, NK, (int)K, pl.chunks, 0, atot);
  if ((rc = after_launch())) return rc;
  mil_scores_kernel<true><<<grid, kMilThreads, smem, st>>>(cls, det, offsets, (int)K, pl.chunks, cstat, nullptr, asum,
                                                        grad_scores, grad_img, atot, grad_cls, grad_det, 1);
  return after_launch();
}
