// align_mil.cu -- fused region x concept alignment + MIL two-stream score (north star kernel 2):
//   C = T * normalize(x) @ normalize(W)^T                  (open_vocabulary_classifier.py:85-104, no background)
//   scores[r, k] = softmax_k(C[r, :]) * softmax over the image's rows of det[:, k]
//   img[n, k]    = clamp(sum_r scores[r, k], 1e-6, 1 - 1e-6)
// i.e. ObjectMiningOutputLayers.forward / predict_probs_img with `cls` = the open-vocabulary class head
// (fast_rcnn_open_vocabulary.py:280-285,318-367,604-618; the variant of roi_heads.py:588-590).
//
// Four launches after the weight normalisation:
//   (1) mil_tiles_kernel   -- device-side table of row tiles, every tile inside ONE image (offsets stay on the device)
//   (2) align_tc_kernel    -- tcgen05 TF32 contraction; the TMEM epilogue turns the accumulator into the row softmax
//                             p[r, :] (written to `scores`) and, on the same transposed walk that writes it, reduces
//                             the detection stream's column softmax statistics (max, sum exp) of its <= 32 rows
//   (3) mil_fused_merge    -- one warp per (image, class) merges the partial statistics (fixed tree)
//   (4) mil_fused_finish   -- one CTA per tile: applies the column softmax in place, sums the tile's columns; the last
//                             CTA of an image (ticket) adds the tile sums in tile order and clamps: deterministic, no
//                             float atomics.
#include "align.cuh"

#include <algorithm>

namespace wsovod {

constexpr int MIL_TM = 128;   // rows per tile = TC_BM

// one thread per image walks its tiles; tiles of an image are consecutive: tile_first[n] .. tile_first[n + 1]
__global__ void mil_tiles_kernel(const int64_t* __restrict__ offsets, int N, int64_t M, int4* __restrict__ tiles,
                                 int* __restrict__ tile_first, int* __restrict__ ntiles, int* __restrict__ tickets) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  int t = 0;
  for (int n = 0; n < N; ++n) {
    tile_first[n] = t;
    tickets[n] = 0;
    int64_t a = offsets[n], b = offsets[n + 1];
    a = max((int64_t)0, min(a, M));
    b = max(a, min(b, M));
    for (int64_t r = a; r < b; r += MIL_TM) tiles[t++] = make_int4((int)r, (int)min((int64_t)MIL_TM, b - r), n, 0);
  }
  tile_first[N] = t;
  *ntiles = t;
}

// one warp per (image, class): the image's partial statistics (tile-major, four per tile), 32 at a time, combined with
// the online-softmax merge -- a fixed tree, so the result does not depend on scheduling
__global__ void __launch_bounds__(256) mil_fused_merge_kernel(const int* __restrict__ tile_first, const float2* __restrict__ colpart,
                                                              int N, int K, float2* __restrict__ colstat) {
  const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (wid >= N * K) return;
  const int n = wid / K, k = wid - n * K;
  const int q0 = __ldg(tile_first + n) * 4, q1 = __ldg(tile_first + n + 1) * 4;
  float m = -FLT_MAX, s = 0.f;
  for (int q = q0 + lane; q < q1; q += 32) {
    const float2 pq = __ldg(colpart + (size_t)q * K + k);
    const float nm = fmaxf(m, pq.x);
    s = s * expf(m - nm) + pq.y * expf(pq.x - nm);
    m = nm;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o), os = __shfl_xor_sync(0xffffffffu, s, o);
    const float nm = fmaxf(m, om);
    s = s * expf(m - nm) + os * expf(om - nm);
    m = nm;
  }
  if (lane == 0) colstat[(size_t)n * K + k] = make_float2(m, 1.f / s);
}

__global__ void __launch_bounds__(128) mil_fused_finish_kernel(
    const int4* __restrict__ tiles, const int* __restrict__ tile_first, const int* __restrict__ ntiles,
    const float2* __restrict__ colstat, const float* __restrict__ det, int K, float* __restrict__ scores,
    float* __restrict__ tilesum, int* __restrict__ tickets, float* __restrict__ img) {
  extern __shared__ float sm[];
  float* cmax = sm;             // [K]
  float* cinv = sm + K;         // [K]
  float* wsum = sm + 2 * K;     // [4][K] per-warp column sums
  __shared__ int s_last;
  const int tile = blockIdx.x;
  if (tile >= __ldg(ntiles)) return;
  const int4 t = __ldg(tiles + tile);
  const int n = t.z;
  const int t0 = __ldg(tile_first + n), t1 = __ldg(tile_first + n + 1);
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float2 cs = __ldg(colstat + (size_t)n * K + k);
    cmax[k] = cs.x;
    cinv[k] = cs.y;
  }
  __syncthreads();
  // scores = p * softmax_col(det) in place; lane == column, a warp walks its 32 rows in order
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t wrow0 = (int64_t)t.x + warp * 32;
  const int wrows = min(32, t.y - warp * 32);
  for (int k0 = 0; k0 < K; k0 += 32) {
    const int k = k0 + lane;
    float acc = 0.f;
    if (k < K) {
      const float cm = cmax[k], ci = cinv[k];
      for (int r0 = 0; r0 < wrows; r0 += 8) {          // eight rows in flight: `scores` is read and written in place,
        float pv[8], dv[8];                            // so the loads are issued before any of the stores
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const bool in = r0 + u < wrows;
          const size_t e = (size_t)(wrow0 + r0 + (in ? u : 0)) * K + k;
          pv[u] = in ? scores[e] : 0.f;
          dv[u] = in ? __ldg(det + e) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (r0 + u < wrows) {
            const float v = pv[u] * (expf(dv[u] - cm) * ci);
            scores[(size_t)(wrow0 + r0 + u) * K + k] = v;
            acc += v;
          }
      }
      wsum[warp * K + k] = acc;
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x)
    tilesum[(size_t)tile * K + k] = (wsum[k] + wsum[K + k]) + (wsum[2 * K + k] + wsum[3 * K + k]);
  // the last tile of the image adds the tile sums in tile order
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&tickets[n], 1) == (t1 - t0 - 1);
  __syncthreads();
  if (!s_last || !img) return;
  __threadfence();
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float a = 0.f;
    for (int q = t0; q < t1; ++q) a += __ldcg(tilesum + (size_t)q * K + k);
    img[(size_t)n * K + k] = fminf(fmaxf(a, 1e-6f), 1.f - 1e-6f);     // :604-618
  }
}

// images without rows have no tile: their image score is clamp(0)
__global__ void mil_empty_images_kernel(const int* __restrict__ tile_first, int N, int K, float* __restrict__ img) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * K) return;
  const int n = i / K;
  if (tile_first[n + 1] == tile_first[n]) img[i] = 1e-6f;
}

struct MilFuseWs { size_t tiles, tile_first, ntiles, tickets, colpart, colstat, tilesum, align, bytes; int ntiles_max; };

static MilFuseWs mil_fuse_plan(int64_t M, int64_t N, int64_t D, int64_t K) {
  MilFuseWs w;
  size_t o = 0;
  auto take = [&](size_t b) { size_t r = o; o += align_up(b, 256); return r; };
  w.ntiles_max = (int)(ceil_div(M, MIL_TM) + N);
  w.tiles = take(sizeof(int4) * (size_t)w.ntiles_max);
  w.tile_first = take(sizeof(int) * (size_t)(N + 1));
  w.ntiles = take(sizeof(int));
  w.tickets = take(sizeof(int) * (size_t)std::max<int64_t>(N, 1));
  w.colpart = take(sizeof(float2) * (size_t)w.ntiles_max * 4 * (size_t)K);
  w.colstat = take(sizeof(float2) * (size_t)std::max<int64_t>(N, 1) * (size_t)K);
  w.tilesum = take(sizeof(float) * (size_t)w.ntiles_max * (size_t)K);
  w.align = take(align_plan(M, D, K, WSOVOD_B200_ALIGN_TF32, false).bytes);
  w.bytes = o;
  return w;
}

}  // namespace wsovod

using namespace wsovod;

WSOVOD_API size_t wsovod_b200_align_mil_fused_workspace(int64_t M, int64_t N, int64_t D, int64_t K) {
  if (M < 0 || N < 0 || D < 0 || K < 0) return 0;
  return mil_fuse_plan(M, N, D, K).bytes;
}

WSOVOD_API int wsovod_b200_align_mil_fused_fwd(const float* x, const float* classifier, const float* det,
                                               const int64_t* offsets, int64_t M, int64_t N, int64_t D, int64_t K,
                                               float temperature, int norm_weight, const float* bias, float* scores,
                                               float* img_scores, float* logits, void* workspace,
                                               size_t workspace_bytes, void* stream) {
  if (M < 0 || N < 0 || D < 0 || K < 0) return WSOVOD_B200_EINVAL;
  if (K == 0 || N == 0) return 0;
  if (M == 0) return img_scores ? WSOVOD_B200_EINVAL : 0;                // no proposals at all: nothing to score
  if (!x || !classifier || !det || !offsets || !scores || D == 0) return WSOVOD_B200_EINVAL;
  if (K > 256) return WSOVOD_B200_EUNSUPPORTED;                      // one accumulator chunk (K <= 256 concepts)
  if (M > 0x7fffffffLL - MIL_TM || N > (1 << 24)) return WSOVOD_B200_ETOOBIG;
  const MilFuseWs w = mil_fuse_plan(M, N, D, K);
  if (!workspace || workspace_bytes < w.bytes) return WSOVOD_B200_EWORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  int4* tiles = (int4*)(ws + w.tiles);
  int* tile_first = (int*)(ws + w.tile_first);
  int* ntiles = (int*)(ws + w.ntiles);
  int* tickets = (int*)(ws + w.tickets);
  float2* colpart = (float2*)(ws + w.colpart);
  float2* colstat = (float2*)(ws + w.colstat);
  float* tilesum = (float*)(ws + w.tilesum);
  int rc;
  mil_tiles_kernel<<<1, 32, 0, st>>>(offsets, (int)N, M, tiles, tile_first, ntiles, tickets);
  if ((rc = after_launch())) return rc;
  AlignMilFuse mf{tiles, ntiles, w.ntiles_max, det, colpart};
  const AlignWs aw = align_plan(M, D, K, WSOVOD_B200_ALIGN_TF32, false);
  // row softmax of the alignment logits goes to `scores`; logits (for the backward pass) only when asked for
  rc = align_fwd_tf32(x, classifier, M, D, K, temperature, norm_weight, /*append_background=*/0, bias, logits, scores, aw,
                      ws + w.align, st, &mf);
  if (rc) return rc;
  mil_fused_merge_kernel<<<(unsigned)ceil_div(N * K, 8), 256, 0, st>>>(tile_first, colpart, (int)N, (int)K, colstat);
  if ((rc = after_launch())) return rc;
  const size_t smem = sizeof(float) * (size_t)(6 * K);
  mil_fused_finish_kernel<<<(unsigned)w.ntiles_max, 128, smem, st>>>(tiles, tile_first, ntiles, colstat, det, (int)K, scores,
                                                                      tilesum, tickets, img_scores);
  if ((rc = after_launch())) return rc;
  if (img_scores) {
    mil_empty_images_kernel<<<(unsigned)ceil_div(N * K, 256), 256, 0, st>>>(tile_first, (int)N, (int)K, img_scores);
    if ((rc = after_launch())) return rc;
  }
  return 0;
}
