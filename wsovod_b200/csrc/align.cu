// align.cu -- kernel (2a): region x concept alignment  logits = T * normalize(x) @ normalize(W)^T
// (+ zero background column, + bias) and its row softmax
// (class_heads/open_vocabulary_classifier.py:85-104, roi_heads/fast_rcnn_open_vocabulary.py:1034-1035).
//
// Two contraction paths behind one entry point:
//   WSOVOD_B200_ALIGN_FP32  CUDA-core fp32 FMA tiles; matches torch.mm(fp32) to ~1e-6 (parity grade).
//   WSOVOD_B200_ALIGN_TF32  tcgen05.mma kind::tf32 with TMA-fed shared-memory operands and fp32
//                           accumulators in TMEM (align_tc.cu); fused normalisation/softmax epilogue.
// In both, the row normalisation is folded into the epilogue: logits = (T / max(||x||,1e-12)) * (x . w^),
// with ||x||^2 accumulated from the very tiles the contraction consumes (x is read exactly once).
#include "align.cuh"

#include <algorithm>

namespace wsovod {

// w^[k,:] = w[k,:] / max(||w[k,:]||, 1e-12)  (F.normalize(..., dim=0) on the D x K transpose, :89-90); rows K..Kp-1 and
// columns D..Dp-1 of the padded output are written as zeros (grid covers Kp rows: no separate memset)
__global__ void align_wnorm_kernel(const float* __restrict__ w, int K, int Kp, int D, int Dp, int norm,
                                   float* __restrict__ out, int* __restrict__ zero, int nzero) {
  // `zero`: ints the caller wants cleared on the same launch (the pair kernel's tickets: one memset node less)
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nzero; i += gridDim.x * blockDim.x) zero[i] = 0;
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (k >= Kp) return;
  if (k >= K) {
    for (int d = lane; d < Dp; d += 32) out[(int64_t)k * Dp + d] = 0.f;
    return;
  }
  float ss = 0.f;
  for (int d = lane; d < D; d += 32) { const float v = w[(int64_t)k * D + d]; ss += v * v; }
  ss = warp_sum(ss);
  const float den = norm ? fmaxf(sqrtf(ss), 1e-12f) : 1.f;
  for (int d = lane; d < Dp; d += 32) out[(int64_t)k * Dp + d] = d < D ? w[(int64_t)k * D + d] / den : 0.f;
}

// fp32 tile GEMM: C[m, n] = scale_m * sum_d A[m,d] * B[n,d]   (both operands reduction-major)
//   SCALE_MODE 0: scale = 1     1: scale = T / max(||A[m,:]||, 1e-12) (norm accumulated on the fly)
constexpr int BM = 64, BN = 64, BK = 16;
template <int SCALE_MODE>
__global__ void __launch_bounds__(256) align_gemm_fp32_kernel(
    const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb, int64_t M, int N,
    int Kred, float temperature, const float* __restrict__ bias, int n_zero_cols, float* __restrict__ C,
    int64_t ldc) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  __shared__ float rowss[BM];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  float acc[4][4] = {};
  float ss = 0.f;
  for (int k0 = 0; k0 < Kred; k0 += BK) {
    float a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = k0 + lk + i;
      a[i] = (m0 + lrow < M && k < Kred) ? __ldg(A + (m0 + lrow) * lda + k) : 0.f;
      b[i] = (n0 + lrow < N && k < Kred) ? __ldg(B + (int64_t)(n0 + lrow) * ldb + k) : 0.f;
      if (SCALE_MODE == 1) ss += a[i] * a[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) { As[lk + i][lrow] = a[i]; Bs[lk + i][lrow] = b[i]; }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
  }
  if (SCALE_MODE == 1) {
    ss += __shfl_xor_sync(0xffffffffu, ss, 1);
    ss += __shfl_xor_sync(0xffffffffu, ss, 2);
    if ((tid & 3) == 0) rowss[lrow] = ss;
    __syncthreads();
  }
  const float bv = bias ? __ldg(bias) : 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const float sc = SCALE_MODE == 1 ? temperature / fmaxf(sqrtf(rowss[ty * 4 + i]), 1e-12f) : 1.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N) C[m * ldc + n] = acc[i][j] * sc + bv;
    }
  }
  // appended all-zero weight columns (background, :97-100): logit = 0 (+ bias)
  if (blockIdx.x == 0 && n_zero_cols > 0)
    for (int i = tid; i < BM * n_zero_cols; i += 256) {
      const int64_t m = m0 + i / n_zero_cols;
      if (m < M) C[m * ldc + N + i % n_zero_cols] = bv;
    }
}

// probs[r,:] = softmax(logits[r,:])   one warp per row
__global__ void row_softmax_kernel(const float* __restrict__ logits, int64_t M, int KO, float* __restrict__ probs) {
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= M) return;
  const float* l = logits + r * KO;
  float mx = -FLT_MAX;
  for (int k = lane; k < KO; k += 32) mx = fmaxf(mx, l[k]);
  mx = warp_max(mx);
  float s = 0.f;
  for (int k = lane; k < KO; k += 32) s += expf(l[k] - mx);
  s = warp_sum(s);
  const float inv = 1.f / s;
  for (int k = lane; k < KO; k += 32) probs[r * KO + k] = expf(l[k] - mx) * inv;
}

// backward helper: per row  dx = s * (dy - xh * <xh, dy>),  xh = x / max(||x||, eps), s = T / max(||x||, eps)
// (for ||x|| < eps the denominator is the constant eps: dx = s * dy)
__global__ void align_bwd_rows_kernel(const float* __restrict__ x, const float* __restrict__ dy, int64_t M, int D,
                                      float temperature, int norm, float* __restrict__ dx) {
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= M) return;
  const float* xr = x + r * D;
  const float* gr = dy + r * D;
  if (!norm) {
    for (int d = lane; d < D; d += 32) dx[r * D + d] = gr[d];
    return;
  }
  float ss = 0.f, dot = 0.f;
  for (int d = lane; d < D; d += 32) { const float v = xr[d]; ss += v * v; dot += v * gr[d]; }
  ss = warp_sum(ss);
  dot = warp_sum(dot);
  const float nrm = sqrtf(ss);
  const float den = fmaxf(nrm, 1e-12f);
  const float s = temperature / den;
  const float proj = nrm > 1e-12f ? dot / (den * den) : 0.f;     // <xh,dy>/den
  for (int d = lane; d < D; d += 32) dx[r * D + d] = s * (gr[d] - xr[d] * proj);
}

// backward w.r.t. the classifier (the "rand" weights of open_vocabulary_classifier.py:62-65 are a Parameter):
//   dW^[k, d] = sum_m g[m, k] * s_m * x[m, d],  s_m = T / max(||x_m||, eps)  (1 without normalisation)
// (1) per-row scale, (2) split-M partial products straight from the row-major operands (a tile of 16 rows of g
// and of x is already reduction-major for this product), (3) fixed-order sum of the partials and, for
// norm_weight == 1, the Jacobian of w^ = w / max(||w||, eps):  dW = (dW^ - w^ <w^, dW^>) / max(||w||, eps).
__global__ void align_rowscale_kernel(const float* __restrict__ x, int64_t M, int D, float temperature, int norm,
                                      float* __restrict__ rowscale) {
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= M) return;
  float ss = 0.f;
  if (norm)
    for (int d = lane; d < D; d += 32) { const float v = x[r * D + d]; ss += v * v; }
  ss = warp_sum(ss);
  if (lane == 0) rowscale[r] = norm ? temperature / fmaxf(sqrtf(ss), 1e-12f) : 1.f;
}

constexpr int WM = 16;   // rows of g / x per shared-memory step
__global__ void __launch_bounds__(256) align_bwd_w_kernel(const float* __restrict__ g, int64_t ldg, const float* __restrict__ x,
                                                          const float* __restrict__ rowscale, int64_t M, int K, int D,
                                                          int64_t rows_per_split, float* __restrict__ part) {
  __shared__ float Gs[WM][64 + 4];
  __shared__ float Xs[WM][64 + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int d0 = blockIdx.x * 64, k0 = blockIdx.y * 64;
  const int64_t m_lo = (int64_t)blockIdx.z * rows_per_split, m_hi = min(M, m_lo + rows_per_split);
  const int lrow = tid >> 4, lcol = (tid & 15) * 4;     // 16 rows x 64 columns, four consecutive columns per thread
  float acc[4][4] = {};
  for (int64_t m0 = m_lo; m0 < m_hi; m0 += WM) {
    const int64_t m = m0 + lrow;
    const bool in = m < m_hi;
    const float sc = in ? __ldg(rowscale + m) : 0.f;
    float gv[4], xv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      gv[i] = (in && k0 + lcol + i < K) ? __ldg(g + m * ldg + k0 + lcol + i) : 0.f;
      xv[i] = (in && d0 + lcol + i < D) ? __ldg(x + m * D + d0 + lcol + i) * sc : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) { Gs[lrow][lcol + i] = gv[i]; Xs[lrow][lcol + i] = xv[i]; }
    __syncthreads();
#pragma unroll
    for (int mm = 0; mm < WM; ++mm) {
      const float4 av = *reinterpret_cast<const float4*>(&Gs[mm][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Xs[mm][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
  }
  float* out = part + (int64_t)blockIdx.z * K * D;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = k0 + ty * 4 + i;
    if (k >= K) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int d = d0 + tx * 4 + j;
      if (d < D) out[(int64_t)k * D + d] = acc[i][j];
    }
  }
}

// one CTA per class: the split-M partials are added in a fixed order (deterministic); <w, dW^> and ||w||^2 by a block
// reduction, then the Jacobian of the weight normalisation
__global__ void __launch_bounds__(256) align_bwd_w_finish_kernel(const float* __restrict__ part, int S, const float* __restrict__ w,
                                                                 int K, int D, int norm, float* __restrict__ grad_w) {
  const int k = blockIdx.x;
  __shared__ float s_ss[8], s_dot[8];
  float ss = 0.f, dot = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float a = 0.f;
    for (int z = 0; z < S; ++z) a += __ldg(part + ((int64_t)z * K + k) * D + d);
    grad_w[(int64_t)k * D + d] = a;
    const float v = w[(int64_t)k * D + d];
    ss += v * v;
    dot += v * a;
  }
  if (norm != 1) return;
  ss = warp_sum(ss);
  dot = warp_sum(dot);
  if ((threadIdx.x & 31) == 0) { s_ss[threadIdx.x >> 5] = ss; s_dot[threadIdx.x >> 5] = dot; }
  __syncthreads();
  ss = dot = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { ss += s_ss[i]; dot += s_dot[i]; }
  const float nrm = sqrtf(ss), den = fmaxf(nrm, 1e-12f);
  const float proj = nrm > 1e-12f ? dot / (den * den) : 0.f;     // <w^, dW^> / den
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float a = grad_w[(int64_t)k * D + d];
    grad_w[(int64_t)k * D + d] = (a - w[(int64_t)k * D + d] * proj) / den;
  }
}

__global__ void transpose_kernel(const float* __restrict__ in, int rows, int cols, int ld, float* __restrict__ out) {
  // out[c, r] = in[r, c]   (tiny: the K x D text matrix)
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * cols) return;
  const int r = (int)(i / cols), c = (int)(i % cols);
  out[(int64_t)c * rows + r] = in[(int64_t)r * ld + c];
}

AlignWs align_plan(int64_t M, int64_t D, int64_t K, int precision, bool backward) {
  AlignWs w;
  size_t o = 0;
  auto take = [&](size_t b) { size_t r = o; o += align_up(b, 1024); return r; };
  w.Dp = precision == WSOVOD_B200_ALIGN_TF32 ? (int64_t)align_up((size_t)D, 32) : D;
  w.Kp = precision == WSOVOD_B200_ALIGN_TF32 ? (int64_t)align_up((size_t)K + 1, 32) : K;
  w.what = take(sizeof(float) * (size_t)(w.Kp * w.Dp));            // normalised text matrix
  // completion tickets of the CTA-pair kernel (K + 1 > 256), zeroed together with `what`
  w.tickets_bytes = precision == WSOVOD_B200_ALIGN_TF32 && K + 1 > 256 ? sizeof(int) * (size_t)(ceil_div(M, 256) * 8 + 1) : 0;
  w.tickets = take(w.tickets_bytes);
  w.wt = take(backward ? sizeof(float) * (size_t)(K * D) : 0);     // its transpose (backward)
  w.dy = take(backward ? sizeof(float) * (size_t)(M * D) : 0);     // backward scratch
  // classifier gradient: per-row scales and split-M partial products
  const int64_t tiles = std::max<int64_t>(1, ceil_div(D, 64) * ceil_div(K, 64));
  w.wsplit = backward ? std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(64, ceil_div(M, 64)), ceil_div(2 * kNumSMs, tiles))) : 0;
  w.rowscale = take(backward ? sizeof(float) * (size_t)M : 0);
  w.wpart = take(backward ? sizeof(float) * (size_t)(w.wsplit * K * D) : 0);
  w.rowstat = take(precision == WSOVOD_B200_ALIGN_TF32 && K + 1 > 256 ? sizeof(float) * 2 * (size_t)M * (size_t)ceil_div(K + 1, 256) : 0);
  w.bytes = o;
  return w;
}

}  // namespace wsovod

using namespace wsovod;

WSOVOD_API size_t wsovod_b200_align_workspace(int64_t M, int64_t D, int64_t K, int precision) {
  if (M < 0 || D < 0 || K < 0) return 0;
  return align_plan(M, D, K, precision, false).bytes;
}

WSOVOD_API size_t wsovod_b200_align_bwd_workspace(int64_t M, int64_t D, int64_t K) {
  if (M < 0 || D < 0 || K < 0) return 0;
  return align_plan(M, D, K, WSOVOD_B200_ALIGN_FP32, true).bytes;
}

WSOVOD_API int wsovod_b200_align_fwd(const float* x, const float* classifier, int64_t M, int64_t D,
                                     int64_t K, float temperature, int norm_weight,
                                     int append_background, const float* bias, int precision,
                                     float* logits, float* probs, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  if (M < 0 || D < 0 || K < 0 || (precision != WSOVOD_B200_ALIGN_FP32 && precision != WSOVOD_B200_ALIGN_TF32))
    return WSOVOD_B200_EINVAL;
  const int64_t KO = K + (append_background ? 1 : 0);
  if (M == 0 || KO == 0) return 0;
  if (!x || (K > 0 && !classifier) || (!logits && !probs) || D == 0) return WSOVOD_B200_EINVAL;
  if (K >= (1 << 24) || D >= (1 << 24) || M >= (1LL << 40)) return WSOVOD_B200_ETOOBIG;
  const AlignWs w = align_plan(M, D, K, precision, false);
  if (!workspace || workspace_bytes < w.bytes) return WSOVOD_B200_EWORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  float* what = (float*)(ws + w.what);
  int rc;
  if (precision == WSOVOD_B200_ALIGN_TF32)
    return align_fwd_tf32(x, classifier, M, D, K, temperature, norm_weight, append_background, bias, logits,
                          probs, w, ws, st);
  if (K > 0) {
    align_wnorm_kernel<<<(unsigned)ceil_div(K, 8), 256, 0, st>>>(classifier, (int)K, (int)K, (int)D, (int)D, norm_weight == 1, what, nullptr, 0);
    if ((rc = after_launch())) return rc;
  }
  // logits go to `logits` when given, else straight into `probs` and are normalised in place
  float* lg = logits ? logits : probs;
  dim3 grid((unsigned)std::max<int64_t>(1, ceil_div(K, BN)), (unsigned)ceil_div(M, BM));
  if (norm_weight)
    align_gemm_fp32_kernel<1><<<grid, 256, 0, st>>>(x, D, what, D, M, (int)K, (int)D, temperature, bias, (int)(KO - K), lg, KO);
  else
    align_gemm_fp32_kernel<0><<<grid, 256, 0, st>>>(x, D, what, D, M, (int)K, (int)D, temperature, bias, (int)(KO - K), lg, KO);
  if ((rc = after_launch())) return rc;
  if (probs) {
    row_softmax_kernel<<<(unsigned)ceil_div(M, 8), 256, 0, st>>>(lg, M, (int)KO, probs);
    if ((rc = after_launch())) return rc;
  }
  return 0;
}

WSOVOD_API int wsovod_b200_align_bwd(const float* grad_logits, const float* x, const float* classifier,
                                     int64_t M, int64_t D, int64_t K, float temperature,
                                     int norm_weight, int append_background, float* grad_x,
                                     float* grad_classifier, void* workspace, size_t workspace_bytes,
                                     void* stream) {
  if (M < 0 || D < 0 || K < 0) return WSOVOD_B200_EINVAL;
  if (M == 0 || D == 0) {
    if (grad_classifier && K > 0 && D > 0) return (int)cudaMemsetAsync(grad_classifier, 0, sizeof(float) * (size_t)(K * D), (cudaStream_t)stream);
    return 0;
  }
  if (!grad_logits || !x || (!grad_x && !grad_classifier) || (K > 0 && !classifier)) return WSOVOD_B200_EINVAL;
  const AlignWs w = align_plan(M, D, K, WSOVOD_B200_ALIGN_FP32, true);
  if (!workspace || workspace_bytes < w.bytes) return WSOVOD_B200_EWORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  float* what = (float*)(ws + w.what);
  float* wt = (float*)(ws + w.wt);
  float* dy = (float*)(ws + w.dy);
  const int64_t KO = K + (append_background ? 1 : 0);
  int rc;
  if (K == 0) return grad_x ? (int)cudaMemsetAsync(grad_x, 0, sizeof(float) * (size_t)(M * D), st) : 0;
  if (grad_classifier) {
    float* rowscale = (float*)(ws + w.rowscale);
    float* part = (float*)(ws + w.wpart);
    const int S = (int)w.wsplit;
    const int64_t per = ceil_div(ceil_div(M, S), WM) * WM;
    align_rowscale_kernel<<<(unsigned)ceil_div(M, 8), 256, 0, st>>>(x, M, (int)D, temperature, norm_weight, rowscale);
    if ((rc = after_launch())) return rc;
    align_bwd_w_kernel<<<dim3((unsigned)ceil_div(D, 64), (unsigned)ceil_div(K, 64), (unsigned)S), 256, 0, st>>>(
        grad_logits, KO, x, rowscale, M, (int)K, (int)D, per, part);
    if ((rc = after_launch())) return rc;
    align_bwd_w_finish_kernel<<<(unsigned)K, 256, 0, st>>>(part, S, classifier, (int)K, (int)D, norm_weight, grad_classifier);
    if ((rc = after_launch())) return rc;
    if (!grad_x) return 0;
  }
  align_wnorm_kernel<<<(unsigned)ceil_div(K, 8), 256, 0, st>>>(classifier, (int)K, (int)K, (int)D, (int)D, norm_weight == 1, what, nullptr, 0);
  if ((rc = after_launch())) return rc;
  transpose_kernel<<<(unsigned)ceil_div(K * D, 256), 256, 0, st>>>(what, (int)K, (int)D, (int)D, wt);   // wt [D,K]
  if ((rc = after_launch())) return rc;
  // dy[m, d] = sum_k g[m,k] * w^[k,d]  == A = grad_logits (lda = KO, reduce over K), B = wt [D, K]
  dim3 grid((unsigned)ceil_div(D, BN), (unsigned)ceil_div(M, BM));
  align_gemm_fp32_kernel<0><<<grid, 256, 0, st>>>(grad_logits, KO, wt, K, M, (int)D, (int)K, 1.f, nullptr, 0, dy, D);
  if ((rc = after_launch())) return rc;
  align_bwd_rows_kernel<<<(unsigned)ceil_div(M, 8), 256, 0, st>>>(x, dy, M, (int)D, temperature, norm_weight, grad_x);
  return after_launch();
}
