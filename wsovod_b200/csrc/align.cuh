// align.cuh -- shared declarations of the alignment kernels (align.cu, align_tc.cu)
#pragma once
#include "common.cuh"

namespace wsovod {

struct AlignWs {
  size_t what, tickets, tickets_bytes, wt, dy, rowstat, rowscale, wpart, bytes;
  int64_t wsplit;   // split-M factor of the classifier-gradient partial products
  int64_t Dp, Kp;   // padded reduction length / padded number of weight rows (TF32 path)
};
AlignWs align_plan(int64_t M, int64_t D, int64_t K, int precision, bool backward);

__global__ void align_wnorm_kernel(const float* __restrict__ w, int K, int Kp, int D, int Dp, int norm,
                                   float* __restrict__ out, int* __restrict__ zero = nullptr, int nzero = 0);
__global__ void row_softmax_kernel(const float* __restrict__ logits, int64_t M, int KO,
                                   float* __restrict__ probs);

// fused alignment + MIL (align_mil.cu): image-aligned row tiles and the detection stream whose column statistics the
// epilogue reduces on the way out
struct AlignMilFuse {
  const int4* tiles;        // [ntiles_max] (first row, rows, image, -)
  const int* ntiles_dev;    // valid entries
  int ntiles_max;
  const float* det;         // [M, K]
  float2* colpart;          // [ntiles_max * 4, K]
};

int align_fwd_tf32(const float* x, const float* classifier, int64_t M, int64_t D, int64_t K,
                   float temperature, int norm_weight, int append_background, const float* bias,
                   float* logits, float* probs, const AlignWs& w, char* ws, cudaStream_t st,
                   const AlignMilFuse* mil = nullptr);

}  // namespace wsovod
