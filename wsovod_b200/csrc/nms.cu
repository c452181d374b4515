// nms.cu -- kernel (4): per-class NMS and the fused fast_rcnn_inference tail.
//
// Greedy NMS keeps a box iff no higher-scoring KEPT box of its class overlaps it by more than thr, so
// only (candidates x kept) IoUs matter, not the (candidates^2)/2 bitmask torchvision's kernel fills.
//
//   detections (the inference tail): det_rows (finite filter, clip, class-major score copy) ->
//     det_class (one CTA per (class, image): gather, exact pre-selection of the best candidates, sort,
//     all-pairs head stage over the first 128, warp-by-warp chunks for the rest; stops at topk kept boxes)
//     -> det_topk (pairwise parallel merge of the class runs, G CTAs per image).  The inference tail needs
//     only the DETECTIONS_PER_IMAGE best survivors of an image and no class can contribute more than that
//     many, so every class stops after `topk` kept boxes (exact, see DESIGN.md "Kernel 4").
//   batched_nms (generic, any number of survivors): segments by group, one CTA per group iterating
//     [suppress what the last kept box covers + block arg-max of what is still alive] in one pass over
//     shared memory per kept box; the arg-max sequence IS the score-descending kept list.
//
// Keys: 64 bit = (order-inverted score bits << 32) | candidate id, so "smaller key" == "higher score,
// then lower id" -- torchvision's stable descending sort.  IoU arithmetic is bit-exact to either
// torchvision kernel (WSOVOD_B200_IOU_TV_CPU / _TV_CUDA), see suppresses() and its division-free
// screen suppresses_fast().
#include "common.cuh"

#include <math.h>

#include <algorithm>

namespace wsovod {

constexpr int kNmsThreads = 512;
constexpr int kNmsWarps = kNmsThreads / 32;
constexpr unsigned long long kDead = ~0ull;

__device__ __forceinline__ unsigned long long make_key(float score, uint32_t id) {
  uint32_t u = __float_as_uint(score);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);     // ascending float order as unsigned
  return ((unsigned long long)(~u) << 32) | id;        // ascending key = descending score
}
__device__ __forceinline__ float key_score(unsigned long long key) {
  uint32_t u = ~(uint32_t)(key >> 32);
  u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  return __uint_as_float(u);
}

// does kept box i (area ai precomputed) suppress candidate j ?
//   MODE 0 (torchvision cpu/nms_kernel.cpp): ovr = inter / ((ai + aj) - inter); (double)ovr > thr
//   MODE 1 (torchvision cuda/nms_kernel.cu as compiled for sm_100): den = fma(wj, hj, ai) - inter;
//           ovr > (float)thr
// `thr` arrives pre-converted so that both are a single fp32 compare (see cmp_threshold()).
template <int MODE>
__device__ __forceinline__ bool suppresses(const float4 bi, const float ai, const float4 bj,
                                           const float thr) {
  float w = __fsub_rn(fminf(bi.z, bj.z), fmaxf(bi.x, bj.x));
  float h = __fsub_rn(fminf(bi.w, bj.w), fmaxf(bi.y, bj.y));
  w = fmaxf(w, 0.f);
  h = fmaxf(h, 0.f);
  const float inter = __fmul_rn(w, h);
  // disjoint boxes (the common case): 0 / den is +-0 or NaN, never > thr for thr >= 0 -- skip the division
  if (inter == 0.f && thr >= 0.f) return false;
  float den;
  if (MODE == 0) {
    const float aj = __fmul_rn(__fsub_rn(bj.z, bj.x), __fsub_rn(bj.w, bj.y));
    den = __fsub_rn(__fadd_rn(ai, aj), inter);
  } else {
    den = __fsub_rn(__fmaf_rn(__fsub_rn(bj.z, bj.x), __fsub_rn(bj.w, bj.y), ai), inter);
  }
  return __fdiv_rn(inter, den) > thr;
}
// Branch-free screening of the same test without the division: 0 = no, 1 = yes, 2 = too close to call
// (run suppresses()).  With t = fl(thr * den) and e = fl(inter - t), |e| > 2^-18 * t implies that the real
// quotient Q = inter / den is more than 3.7e-6 * thr away from thr, on the side sign(e) says (the two
// roundings in e move it by at most 2^-23 * t), and the IEEE quotient is within 2^-24 * Q of Q -- so it
// compares with thr the same way.  Needs a normal, positive den (zero / negative / NaN denominators of
// degenerate boxes go to the exact test; an infinite one makes t infinite and fails |e| > inf) and a
// threshold in [1e-6, 1e6] (anything else is never screened).  Disjoint boxes have e = -t: a sure "no".
template <int MODE>
__device__ __forceinline__ int suppresses_fast(const float4 bi, const float ai, const float4 bj, const float thr) {
  float w = __fsub_rn(fminf(bi.z, bj.z), fmaxf(bi.x, bj.x));
  float h = __fsub_rn(fminf(bi.w, bj.w), fmaxf(bi.y, bj.y));
  w = fmaxf(w, 0.f);
  h = fmaxf(h, 0.f);
  const float inter = __fmul_rn(w, h);
  float den;
  if (MODE == 0) {
    const float aj = __fmul_rn(__fsub_rn(bj.z, bj.x), __fsub_rn(bj.w, bj.y));
    den = __fsub_rn(__fadd_rn(ai, aj), inter);
  } else {
    den = __fsub_rn(__fmaf_rn(__fsub_rn(bj.z, bj.x), __fsub_rn(bj.w, bj.y), ai), inter);
  }
  const float t = __fmul_rn(thr, den);
  const float e = __fsub_rn(inter, t);
  const bool sure = thr >= 1e-6f && thr <= 1e6f && den >= 8.6736173798840355e-19f /* 2^-60 */ &&
                    fabsf(e) > __fmul_rn(t, 3.814697265625e-06f /* 2^-18 */);
  return sure ? (e > 0.f ? 1 : 0) : 2;
}
__device__ __forceinline__ float area_rn(const float4 b) {
  return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
}

// fp32 threshold t such that the reference's comparison equals `ovr > t` for every fp32 ovr
static float cmp_threshold(double thr, int mode) {
  if (mode == WSOVOD_B200_IOU_TV_CUDA) return (float)thr;       // F2F.F32.F64 (round to nearest)
  float f = (float)thr;                                         // (double)ovr > thr
  if ((double)f > thr) f = nextafterf(f, -INFINITY);            // largest float <= thr
  return f;
}

// block arg-min over (key) with the owning position; double-buffered slots -> one barrier per call
struct ArgMinSlots {
  unsigned long long key[2][kNmsWarps];
  int pos[2][kNmsWarps];
};

__device__ __forceinline__ void block_argmin(unsigned long long& key, int& pos, ArgMinSlots& s, int buf) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long ok = __shfl_xor_sync(0xffffffffu, key, o);
    const int op = __shfl_xor_sync(0xffffffffu, pos, o);
    if (ok < key) { key = ok; pos = op; }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { s.key[buf][wid] = key; s.pos[buf][wid] = pos; }
  __syncthreads();
  key = s.key[buf][0];
  pos = s.pos[buf][0];
#pragma unroll
  for (int w = 1; w < kNmsWarps; ++w) {
    const unsigned long long ok = s.key[buf][w];
    if (ok < key) { key = ok; pos = s.pos[buf][w]; }
  }
}

// Greedy selection over n candidates (keys[], boxes[] in shared or global memory).  Writes the kept
// keys in score-descending order to kept_keys[0..ret).  SUPPRESS=false degenerates to "top-limit".
template <int MODE, bool SUPPRESS>
__device__ int select_greedy(unsigned long long* keys, const float4* boxes, int n, float thr, int limit,
                             unsigned long long* kept_keys, ArgMinSlots& slots) {
  int kept = 0, last = -1, buf = 0;
  float4 kb = make_float4(0.f, 0.f, 0.f, 0.f);
  float ka = 0.f;
  while (limit < 0 || kept < limit) {
    unsigned long long best = kDead;
    int bpos = -1;
    for (int i = threadIdx.x; i < n; i += kNmsThreads) {
      const unsigned long long key = keys[i];
      if (key == kDead) continue;
      if (i == last) { keys[i] = kDead; continue; }
      if (SUPPRESS && last >= 0 && suppresses<MODE>(kb, ka, boxes[i], thr)) { keys[i] = kDead; continue; }
      if (key < best) { best = key; bpos = i; }
    }
    block_argmin(best, bpos, slots, buf);
    buf ^= 1;
    if (bpos < 0) break;
    if (threadIdx.x == 0) kept_keys[kept] = best;
    ++kept;
    last = bpos;
    if (SUPPRESS) { kb = boxes[bpos]; ka = area_rn(kb); }
  }
  return kept;
}

// ------------------------------------------------------------------------------------------------
// fused inference tail (fast_rcnn_inference_single_image, fast_rcnn_open_vocabulary.py:149-217)
// ------------------------------------------------------------------------------------------------
// rows: finite filter (:178-182) + Boxes.clip (:187-188), and the class-major copy of the scores:
// scoresT[k][r] = probs[r][k] for finite rows, -inf otherwise.  det_class then reads its column as one
// contiguous run instead of one 4-byte word out of every 32-byte sector of a row-major matrix that all K
// class CTAs of the image walk at the same time (c2: 82 MB of sector traffic for 10 MB of scores).
// One CTA owns 32 rows: a pass over the full rows settles which are finite, then the rows (L1-resident
// by now) go through a 32 x 128 shared-memory tile per block of classes.
constexpr int kRtRows = 32, kRtCols = 128;
__global__ void __launch_bounds__(256) det_rows_kernel(const float* __restrict__ probs, const float* __restrict__ boxes,
                                                       const int64_t* __restrict__ offsets, const float* __restrict__ image_sizes,
                                                       int64_t M, int N, int K1, float4* __restrict__ cboxes,
                                                       float* __restrict__ scoresT) {
  __shared__ float tile[kRtRows][kRtCols + 1];
  __shared__ uint8_t s_ok[kRtRows];
  const int64_t rb = (int64_t)blockIdx.x * kRtRows;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;      // 8 warps, 4 rows each
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int rl = wid * 4 + q;
    const int64_t r = rb + rl;
    bool ok = false;
    if (r < M) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(boxes) + r);
      ok = isfinite(b.x) && isfinite(b.y) && isfinite(b.z) && isfinite(b.w);
      for (int k = lane; k < K1; k += 32) ok &= isfinite(__ldg(probs + r * K1 + k));
      ok = __all_sync(0xffffffffu, ok);
      if (lane == 0) {
        int lo = 0, hi = N - 1;
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (offsets[mid] <= r) lo = mid; else hi = mid - 1; }
        const float ih = image_sizes[2 * lo], iw = image_sizes[2 * lo + 1];
        cboxes[r] = make_float4(fminf(fmaxf(b.x, 0.f), iw), fminf(fmaxf(b.y, 0.f), ih),
                                fminf(fmaxf(b.z, 0.f), iw), fminf(fmaxf(b.w, 0.f), ih));
      }
    }
    if (lane == 0) s_ok[rl] = ok ? 1 : 0;
  }
  __syncthreads();
  const int K = K1 - 1;
  for (int kb = 0; kb < K; kb += kRtCols) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int rl = wid * 4 + q;
      const int64_t r = rb + rl;
      const bool row_ok = r < M && s_ok[rl];
#pragma unroll
      for (int c = lane; c < kRtCols; c += 32)
        tile[rl][c] = (row_ok && kb + c < K) ? __ldg(probs + r * K1 + kb + c) : -INFINITY;
    }
    __syncthreads();
    const int64_t r = rb + lane;
    if (r < M)
      for (int c = wid; c < kRtCols && kb + c < K; c += 8) scoresT[(int64_t)(kb + c) * M + r] = tile[lane][c];
    __syncthreads();
  }
}

// in-place bitonic sort (ascending) of npad = 2^k keys in shared memory by the whole CTA
__device__ __forceinline__ void bitonic_sort_smem(unsigned long long* keys, int npad) {
  for (int k = 2; k <= npad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < npad; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = keys[i], b = keys[ixj];
          const bool asc = (i & k) == 0;
          if ((a > b) == asc) { keys[i] = b; keys[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
}

// the same sort for npad <= blockDim.x keys held one per thread: compare-exchange distances below 32 are
// shuffles, only the larger ones go through shared memory (npad = 256: 12 barriers instead of 36)
__device__ __forceinline__ void bitonic_sort_regs(unsigned long long* keys, int npad) {
  const int t = threadIdx.x;
  const bool busy = (t & ~31) < npad;            // warps past the keys only keep the barriers company
  unsigned long long v = t < npad ? keys[t] : kDead;
  for (int k = 2; k <= npad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      unsigned long long o = kDead;
      if (j >= 32) {
        __syncthreads();                         // the previous exchange has been read
        if (t < npad) keys[t] = v;
        __syncthreads();
        if (t < npad) o = keys[t ^ j];
      } else if (busy) {
        o = __shfl_xor_sync(0xffffffffu, v, j);
      }
      if (busy) {
        const bool take_min = ((t & j) == 0) == ((t & k) == 0);
        v = take_min ? (o < v ? o : v) : (o > v ? o : v);
      }
    }
  }
  __syncthreads();
  if (t < npad) keys[t] = v;
  __syncthreads();
}

constexpr int kMaxRunClasses = 2048;    // the run table / run merge of det_topk handles up to this many classes
constexpr int kSelectBins = 2048;     // histogram over the 11 leading key bits (quarter octaves of the score)
constexpr int kSelectShift = 53;
constexpr int kSelectCap = 1024;      // selected-prefix capacity (keys) == histogram storage (8 KB)
// ---- front end of the pruned path (K >= topk): no class-major copy of the score matrix ------------------------------------
// With the image threshold tau of det_tau_kernel almost nothing survives (c4: a few hundred of 4.8 M scores per image), so
// the scores are read twice in their own row-major layout instead of being transposed for det_class: det_scan_kernel =
// det_rows' finite filter + box clip + the per-(image, class) maxima tau is made from; det_compact_kernel = the rows that
// survive (score > thr, score >= tau, finite row), appended per (image, class) as row indices into the space the
// transposed matrix would have taken.  det_class then gathers its few candidates' scores from the row-major matrix.
constexpr int kScanRows = 64;
__device__ __forceinline__ int det_image_of(const int64_t* __restrict__ offsets, int N, int64_t r) {
  int lo = 0, hi = N - 1;
  while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (offsets[mid] <= r) lo = mid; else hi = mid - 1; }
  return lo;
}
__global__ void __launch_bounds__(256) det_scan_kernel(const float* __restrict__ probs, const float* __restrict__ boxes,
                                                       const int64_t* __restrict__ offsets, const float* __restrict__ image_sizes,
                                                       int64_t M, int N, int K1, float4* __restrict__ cboxes,
                                                       uint8_t* __restrict__ rowok, float score_thr, unsigned* __restrict__ cmax) {
  extern __shared__ unsigned s_cm[];            // [K] maxima of this tile's rows of ONE image
  __shared__ int s_i0, s_i1, s_bad;
  const int K = K1 - 1;
  const int64_t rb = (int64_t)blockIdx.x * kScanRows, re = min(M, rb + kScanRows);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) { s_i0 = det_image_of(offsets, N, rb); s_i1 = det_image_of(offsets, N, re - 1); }
  __syncthreads();
  for (int img = s_i0; img <= s_i1; ++img) {
    const int64_t a = max(rb, offsets[img]), b = min(re, offsets[img + 1]);
    if (a >= b) continue;                                        // an image without proposals
    const float ih = image_sizes[2 * img], iw = image_sizes[2 * img + 1];
    // One pass per row in the common case: the maxima are accumulated while the row is checked for non-finite entries.
    // A row that fails the check has already contributed, so a tile that saw one (s_bad) clears its maxima and repeats
    // the accumulation with the flags known (second attempt: `speculate` false).
    for (int attempt = 0; attempt < 2; ++attempt) {
      const bool speculate = attempt == 0;
      for (int k = threadIdx.x; k < K; k += blockDim.x) s_cm[k] = 0u;
      if (threadIdx.x == 0) s_bad = 0;
      __syncthreads();
      for (int64_t r = a + wid; r < b; r += 8) {
        const float* row = probs + r * K1;
        bool ok = true;
        if (speculate) {
          const float4 bx = __ldg(reinterpret_cast<const float4*>(boxes) + r);
          ok = isfinite(bx.x) && isfinite(bx.y) && isfinite(bx.z) && isfinite(bx.w);
          if (lane == 0)
            cboxes[r] = make_float4(fminf(fmaxf(bx.x, 0.f), iw), fminf(fmaxf(bx.y, 0.f), ih),
                                    fminf(fmaxf(bx.z, 0.f), iw), fminf(fmaxf(bx.w, 0.f), ih));
        } else if (!rowok[r]) {
          continue;
        }
        for (int k0 = 0; k0 < K1; k0 += 32 * 8) {
          float v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int k = k0 + u * 32 + lane;
            v[u] = k < K1 ? __ldg(row + k) : 0.f;
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int k = k0 + u * 32 + lane;
            ok &= isfinite(v[u]);
            if (k < K && v[u] > score_thr && v[u] > 0.f && isfinite(v[u])) atomicMax(&s_cm[k], __float_as_uint(v[u]));
          }
        }
        if (speculate) {
          ok = __all_sync(0xffffffffu, ok);
          if (lane == 0) { rowok[r] = ok ? 1 : 0; if (!ok) s_bad = 1; }
        }
      }
      __syncthreads();
      if (!s_bad) break;
      __syncthreads();
    }
    for (int k = threadIdx.x; k < K; k += blockDim.x)
      if (s_cm[k]) atomicMax(cmax + (size_t)img * K + k, s_cm[k]);
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) det_compact_kernel(const float* __restrict__ probs, const int64_t* __restrict__ offsets,
                                                          int64_t M, int N, int K1, const uint8_t* __restrict__ rowok,
                                                          float score_thr, const float* __restrict__ tau,
                                                          int32_t* __restrict__ ccnt, int32_t* __restrict__ cand) {
  __shared__ int s_i0;
  const int K = K1 - 1;
  const int64_t rb = (int64_t)blockIdx.x * kScanRows, re = min(M, rb + kScanRows);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_i0 = det_image_of(offsets, N, rb);
  __syncthreads();
  int img = s_i0;
  for (int64_t r = rb + wid; r < re; r += 8) {
    if (!rowok[r]) continue;
    while (r >= offsets[img + 1]) ++img;
    const float t = fmaxf(__ldg(tau + img), 0.f);
    const int64_t i0 = offsets[img];
    const float* row = probs + r * K1;
    for (int k0 = 0; k0 < K; k0 += 32 * 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int k = k0 + u * 32 + lane;
        v[u] = k < K ? __ldg(row + k) : -1.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int k = k0 + u * 32 + lane;
        if (v[u] > score_thr && v[u] >= t) {                       // the padding (-1) fails: t >= 0
          const int slot = atomicAdd(ccnt + (size_t)img * K + k, 1);
          cand[(int64_t)k * M + i0 + slot] = (int32_t)(r - i0);
        }
      }
    }
  }
}

// The topk best detections of an image all score at least tau = the topk-th largest of its per-class maxima: the best
// candidate of a class is never suppressed, so an image with topk or more non-empty classes already has topk kept
// boxes at or above tau, and nothing below tau can reach the final list (:207-208 keeps the topk best).  det_class then
// drops candidates below tau before sorting them -- exact, and at LVIS scale (1203 classes, 100 detections) it removes
// almost all of the 4000 x 1203 candidates.  Fewer than topk non-empty classes give tau = 0 (the zeros of cmax).
// One CTA per image: 4-pass radix select on the float bits (probabilities: non-negative, bit order = value order).
__global__ void __launch_bounds__(1024) det_tau_kernel(const unsigned* __restrict__ cmax, int K, int topk, float* __restrict__ tau) {
  extern __shared__ unsigned s_val[];           // the image's K maxima (one global read)
  __shared__ unsigned hist[256];
  __shared__ unsigned s_prefix, s_remaining;
  const unsigned* v = cmax + (size_t)blockIdx.x * K;
  for (int k = threadIdx.x; k < K; k += blockDim.x) s_val[k] = __ldg(v + k);
  if (threadIdx.x == 0) { s_prefix = 0u; s_remaining = (unsigned)topk; }
  for (int shift = 24; shift >= 0; shift -= 8) {
    if (threadIdx.x < 256) hist[threadIdx.x] = 0u;
    __syncthreads();
    const unsigned prefix = s_prefix, himask = shift == 24 ? 0u : 0xffffffffu << (shift + 8);
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      const unsigned x = s_val[k];
      if ((x & himask) == (prefix & himask)) atomicAdd(&hist[(x >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      // digit = the largest b with (number of values whose digit is >= b) >= remaining; lane l owns bins 8l .. 8l + 7
      const int lane = threadIdx.x;
      unsigned h[8], own = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) { h[i] = hist[lane * 8 + i]; own += h[i]; }
      unsigned suf = own;                         // inclusive suffix sum over the lanes: values in bins >= 8 * lane
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_down_sync(0xffffffffu, suf, o);
        if (lane + o < 32) suf += t;
      }
      const unsigned above = suf - own;           // values in the bins of higher lanes
      const unsigned rem = s_remaining;
      int digit = -1;
      unsigned cum_before = 0;
      if (above < rem && suf >= rem) {            // the answer's bin is one of this lane's eight
        unsigned cum = above;
#pragma unroll
        for (int i = 7; i >= 0; --i) {
          if (digit < 0 && cum + h[i] >= rem) { digit = lane * 8 + i; cum_before = cum; }
          if (digit < 0) cum += h[i];
        }
      }
      const unsigned found = __ballot_sync(0xffffffffu, digit >= 0);
      if (found == 0) {
        if (lane == 0) s_remaining = rem;         // fewer than `rem` values in range: digit 0 (prefix unchanged)
      } else if (digit >= 0) {
        s_prefix = prefix | ((unsigned)digit << shift);
        s_remaining = rem - cum_before;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) tau[blockIdx.x] = __uint_as_float(s_prefix);
}

// det_class runs 384 threads per CTA: with <= 45 KB of shared memory five CTAs fit an SM (1920 threads),
// so the 80 x 8 = 640 (class, image) CTAs of c2 are resident at once (740 slots) instead of taking two
// waves of 4 x 148 = 592
constexpr int kDcThreads = 384;
constexpr int kDcHead = 128;          // candidates settled by the all-pairs head stage
constexpr int kDcWindow = 2;          // warps ahead of the resolving one that keep up with the kept list
constexpr int kSelectTarget = 512;    // aim: at least this many best candidates in the prefix
constexpr int kSelectMin = 512;       // columns shorter than this are simply sorted

// one CTA per (class, image):
//   1. gather the column's candidates (score > thr, finite row) as 64-bit keys in shared memory
//   2. bitonic-sort them once (score descending, row ascending)
//   3. greedy NMS over the sorted list in chunks of one candidate per thread: a chunk is first tested
//      against the boxes kept so far, then resolved warp by warp (shuffle broadcast + ballot inside the
//      warp, one barrier per warp); stops at `limit` kept boxes
//   4. append (score, row*K+class) keys to the image's kept list
template <int MODE>
__global__ void __launch_bounds__(kDcThreads, 5) det_class_kernel(
    const float* __restrict__ scoresT, int64_t M, const int64_t* __restrict__ offsets,
    const float4* __restrict__ cboxes, int K, float score_thr, float thr, int limit, int npad_cap, bool fixed_runs,
    int32_t* __restrict__ img_cnt, unsigned long long* __restrict__ img_kept, int64_t kept_stride,
    int2* __restrict__ runs, const float* __restrict__ tau, const int32_t* __restrict__ cand,
    const int32_t* __restrict__ ccnt, const float* __restrict__ probs) {
  extern __shared__ __align__(16) unsigned char sm[];
  __shared__ int s_n, s_base, s_sel, s_bstar, s_m, s_fitsel, s_fitb;
  __shared__ int s_new[2];
  __shared__ int s_head;
  __shared__ int s_wsum[kNmsWarps];
  unsigned long long* skey = reinterpret_cast<unsigned long long*>(sm);                 // [npad_cap]
  float4* kbox = reinterpret_cast<float4*>(sm + (size_t)npad_cap * sizeof(unsigned long long));   // [limit]
  float* karea = reinterpret_cast<float*>(kbox + limit);                                 // [limit]
  unsigned long long* kkey = reinterpret_cast<unsigned long long*>(karea + ((limit + 1) & ~1));   // [limit]
  unsigned long long* ssel = reinterpret_cast<unsigned long long*>(                      // [kSelectCap] / histogram, 16-byte aligned
      (reinterpret_cast<uintptr_t>(kkey + limit) + 15) & ~(uintptr_t)15);
  const int k = blockIdx.x, n = blockIdx.y;
  const int64_t r0 = offsets[n], r1 = offsets[n + 1];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // nothing below it can be among the image's topk detections (det_tau_kernel); null: no pruning
  const float img_tau = tau ? __ldg(tau + n) : -INFINITY;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  const int nrows = (int)(r1 - r0);
  if (cand) {
    // pruned path: the survivors of this (class, image) were listed by det_compact_kernel; their scores come from the
    // row-major matrix (a handful of gathers)
    const int nc0 = __ldg(ccnt + (int64_t)n * K + k);
    const int32_t* lst = cand + (int64_t)k * M + r0;
    for (int i = threadIdx.x; i < nc0; i += kDcThreads) {
      const int rl = __ldg(lst + i);
      skey[i] = make_key(__ldg(probs + (r0 + rl) * (int64_t)(K + 1) + k), (uint32_t)rl);
    }
    if (threadIdx.x == 0) s_n = nc0;
  } else
  // the column is one contiguous run of scoresT (rows that failed the finite filter hold -inf)
  for (int rb = threadIdx.x; rb - lane < nrows; rb += 4 * kDcThreads) {      // warp-uniform trip count
    const float* col = scoresT + (int64_t)k * M + r0;
    float s[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) s[u] = rb + u * kDcThreads < nrows ? __ldg(col + rb + u * kDcThreads) : -INFINITY;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const bool pass = s[u] > score_thr && s[u] >= img_tau;                  // :194, and the image's pruning threshold
      const unsigned m = __ballot_sync(0xffffffffu, pass);
      if (m == 0) continue;
      int at = 0;
      if (lane == 0) at = atomicAdd(&s_n, __popc(m));                         // one append per warp, not per key
      at = __shfl_sync(0xffffffffu, at, 0) + __popc(m & ((1u << lane) - 1u));
      if (pass) skey[at] = make_key(s[u], (uint32_t)(rb + u * kDcThreads));
    }
  }
  __syncthreads();
  const int nc = s_n;
  if (nc == 0) {
    if (threadIdx.x == 0) runs[(int64_t)n * K + k] = make_int2(0, 0);
    return;
  }

  // Candidate list for the greedy pass.  A class rarely needs more than its best few hundred candidates
  // to collect `limit` survivors (c2: ~120 for 100), so the top of the column is selected exactly with a
  // 2048-bin histogram over the 11 leading key bits (no sort of the tail) -- first a short prefix
  // (>= 2 * limit keys), then a longer one (>= kSelectTarget), then the whole column: a pass whose prefix
  // runs dry before `limit` boxes are kept is repeated on the next level (exact).
  unsigned long long* list = skey;
  int ln = nc;
  bool complete = true;
  // CTA-uniform: selects the smallest histogram prefix holding >= target keys into ssel; false if it does
  // not fit (or is the whole column anyway)
  // `fit` > 0: a prefix of at most `fit` keys (one key per thread of the register sort) is preferred to the
  // smallest one reaching `target` as long as it still holds 3/4 of the target
  auto select_prefix = [&](int target, int fit) -> bool {
    int* hist = reinterpret_cast<int*>(ssel);
    __syncthreads();                                     // ssel / s_* free to be rewritten
    for (int i = threadIdx.x; i < kSelectBins; i += kDcThreads) hist[i] = 0;
    if (threadIdx.x == 0) { s_sel = 0; s_m = 0; s_fitsel = 0; }
    __syncthreads();
    for (int i = threadIdx.x; i < nc; i += kDcThreads) atomicAdd(&hist[(int)(skey[i] >> kSelectShift)], 1);
    __syncthreads();
    constexpr int PER = (kSelectBins + kDcThreads - 1) / kDcThreads;      // bins per thread
    int mine = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j)
      if (threadIdx.x * PER + j < kSelectBins) mine += hist[threadIdx.x * PER + j];
    int inc = mine;                                      // inclusive scan over threads
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) s_wsum[wid] = inc;
    __syncthreads();
    int before = inc - mine;
    for (int w2 = 0; w2 < wid; ++w2) before += s_wsum[w2];
    if (before < target && before + mine >= target) {    // exactly one thread
      int cum = before;
      for (int j = 0; j < PER && threadIdx.x * PER + j < kSelectBins; ++j) {
        cum += hist[threadIdx.x * PER + j];
        if (cum >= target) { s_sel = cum; s_bstar = threadIdx.x * PER + j; break; }
      }
    }
    if (fit > 0 && before <= fit && before + mine > fit) {   // exactly one thread: the last bin with cum <= fit
      int cum = before, b = threadIdx.x * PER - 1;
      for (int j = 0; j < PER && threadIdx.x * PER + j < kSelectBins; ++j) {
        if (cum + hist[threadIdx.x * PER + j] > fit) break;
        cum += hist[threadIdx.x * PER + j];
        b = threadIdx.x * PER + j;
      }
      s_fitsel = cum; s_fitb = b;
    }
    __syncthreads();
    const bool use_fit = fit > 0 && s_fitsel * 4 >= target * 3;
    const int selcount = use_fit ? s_fitsel : s_sel, bstar = use_fit ? s_fitb : s_bstar;
    if (!(selcount > 0 && selcount <= kSelectCap && selcount < nc)) return false;
    __syncthreads();                                     // everyone is done reading the histogram
    for (int i = threadIdx.x; i < nc; i += kDcThreads) {
      const unsigned long long key = skey[i];
      if ((int)(key >> kSelectShift) <= bstar) ssel[atomicAdd(&s_m, 1)] = key;
    }
    __syncthreads();
    list = ssel;
    ln = selcount;
    complete = false;
    return true;
  };
  int kept = 0;
  for (int attempt = 0; attempt < 3; ++attempt) {
    list = skey; ln = nc; complete = true;
    if (attempt < 2 && nc > kSelectMin) {
      const int target = attempt == 0 ? max(2 * limit, 128) : kSelectTarget;
      if ((attempt == 1 && target <= max(2 * limit, 128)) || !select_prefix(target, attempt == 0 ? 256 : 0)) continue;   // next level
    }
    int npad = 2;
    while (npad < ln) npad <<= 1;
    for (int i = ln + threadIdx.x; i < npad; i += kDcThreads) list[i] = kDead;
    __syncthreads();
    if (npad <= 256) bitonic_sort_regs(list, npad);
    else bitonic_sort_smem(list, npad);

    kept = 0;
    int base0 = 0;
    // Head of the list (first 128 candidates -- enough for `limit` = 100 survivors on most columns): all
    // pair tests at once, spread over ten warps (warp <-> 32 candidates x 32 candidates ahead of them, the
    // lower triangle of a 4 x 4 block matrix), then ONE warp settles the four groups in order with ballots.
    // Needs 4 KB of spare shared memory behind the list.
    unsigned char* spare = nullptr;
    if (list != ssel) spare = reinterpret_cast<unsigned char*>(ssel);
    else if (npad <= kSelectCap - 512) spare = reinterpret_cast<unsigned char*>(ssel + npad);
    if (spare != nullptr) {
      const int P = min(ln, kDcHead);
      float4* stage = reinterpret_cast<float4*>(spare);                     // [kDcHead] candidate boxes
      float* sarea = reinterpret_cast<float*>(spare + kDcHead * sizeof(float4));             // [kDcHead] their areas
      unsigned* supby = reinterpret_cast<unsigned*>(sarea + kDcHead);       // [10 tiles][32]: who ahead of me covers me
      if (threadIdx.x < P) {
        const float4 b = __ldg(cboxes + r0 + (uint32_t)list[threadIdx.x]);
        stage[threadIdx.x] = b;
        sarea[threadIdx.x] = area_rn(b);
      }
      __syncthreads();
      if (wid < 10) {
        const int g = wid < 1 ? 0 : wid < 3 ? 1 : wid < 6 ? 2 : 3;          // my candidates: group g
        const int p = wid - (g * (g + 1)) / 2;                              // tested against group p <= g
        const int i = 32 * g + lane;
        if (32 * g < P) {
          const float4 bi = stage[min(i, P - 1)];
          const int jend = min(32, P - 32 * p);
          unsigned sup = 0, uns = 0;
          for (int jj = 0; jj < jend; jj += 4) {                            // the tail repeats the block's last box
            unsigned s4 = 0, u4 = 0;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int j = min(32 * p + jj + u, P - 1);
              const int d = suppresses_fast<MODE>(stage[j], sarea[j], bi, thr);
              if (d == 1) s4 |= 1u << u;
              if (d == 2) u4 |= 1u << u;
            }
            sup |= s4 << jj;
            uns |= u4 << jj;
          }
          const unsigned ahead = (p < g ? 0xffffffffu : ((1u << lane) - 1u)) & (jend < 32 ? (1u << jend) - 1u : 0xffffffffu);
          sup &= ahead;
          for (uns &= ahead; uns; uns &= uns - 1) {                         // rare: settle near-threshold pairs exactly
            const int jj = __ffs(uns) - 1;
            if (suppresses<MODE>(stage[32 * p + jj], sarea[32 * p + jj], bi, thr)) sup |= 1u << jj;
          }
          if (i < P) supby[wid * 32 + lane] = sup;                           // tile (g, p) = g (g + 1) / 2 + p = wid
        }
      }
      __syncthreads();
      if (wid == 0) {
        unsigned km[4] = {0u, 0u, 0u, 0u};
        int total = 0;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (32 * g < P && total < limit) {
            const int i = 32 * g + lane;
            bool alive = i < P;
#pragma unroll
            for (int p = 0; p < g; ++p)
              if (alive && (supby[((g * (g + 1)) / 2 + p) * 32 + lane] & km[p])) alive = false;
            const unsigned diag = alive ? supby[((g * (g + 1)) / 2 + g) * 32 + lane] : 0u;
            // kept = alive and not covered by a kept lane ahead: iterate the ballot to its fixed point (lane l
            // is final after l + 1 rounds at the latest; chains are short, so it takes two or three)
            unsigned kmg = __ballot_sync(0xffffffffu, alive);
            for (;;) {
              const unsigned nk = __ballot_sync(0xffffffffu, alive && !(diag & kmg));
              if (nk == kmg) break;
              kmg = nk;
            }
            const int rank = __popc(kmg & ((1u << lane) - 1u));
            const bool keep = ((kmg >> lane) & 1u) && total + rank < limit;
            kmg = __ballot_sync(0xffffffffu, keep);
            if (keep) {
              const float4 kb = stage[i];
              kbox[total + rank] = kb; karea[total + rank] = area_rn(kb); kkey[total + rank] = list[i];
            }
            total += __popc(kmg);
            km[g] = kmg;
          }
        }
        if (lane == 0) s_head = total;
      }
      __syncthreads();
      kept = s_head;
      base0 = P;
    }
    for (int base = base0; base < ln && kept < limit; base += kDcThreads) {
      const int i = base + threadIdx.x;
      bool alive = i < ln;
      unsigned long long key = 0;
      float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
      int tested = 0;       // kept boxes this candidate has been tested against
      // candidate against kept boxes [from, to): four independent tests per trip
      auto catch_up = [&](int from, int to) {
        int j = from;
        for (; j + 4 <= to && alive; j += 4) {
          int d[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) d[u] = suppresses_fast<MODE>(kbox[j + u], karea[j + u], box, thr);
          if ((d[0] | d[1] | d[2] | d[3]) & 2) {             // rare: a quotient within 1e-5 of the threshold
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (d[u] == 2) d[u] = suppresses<MODE>(kbox[j + u], karea[j + u], box, thr) ? 1 : 0;
          }
          alive = (d[0] | d[1] | d[2] | d[3]) == 0;
        }
        for (; j < to && alive; ++j)
          if (suppresses<MODE>(kbox[j], karea[j], box, thr)) alive = false;
      };
      if (alive) {
        key = list[i];
        box = __ldg(cboxes + r0 + (uint32_t)key);
        if (wid < kDcWindow) { catch_up(0, kept); tested = kept; }
      }
      // The chunk is resolved warp by warp in list order.  The warp whose turn it is settles its 32
      // candidates among themselves -- every still-alive lane, in order, is kept, its box is broadcast
      // with shuffles and the ballot of the lanes it suppresses clears them -- and appends the kept
      // boxes to the shared list; after one barrier the next kDcWindow warps test their alive lanes against
      // the kept boxes they have not seen yet.  Barriers per chunk: one per warp that still had work (typically ~5 until
      // `limit` boxes are kept), not one per kept box.
      const float my_area = area_rn(box);
      for (int w = 0; w < kDcThreads / 32 && kept < limit; ++w) {
        if (wid == w) {
          unsigned am = __ballot_sync(0xffffffffu, alive);
          // all pairs first (independent tests, pipelined): bit j of supby = alive lane j < me covers me
          unsigned supby = 0, unsure = 0;
          const unsigned pairs = alive ? (am & ((1u << lane) - 1u)) : 0u;     // alive lanes ahead of an alive me
#pragma unroll 4
          for (int j = 0; j < 31; ++j) {
            const float4 kb = make_float4(__shfl_sync(0xffffffffu, box.x, j), __shfl_sync(0xffffffffu, box.y, j),
                                          __shfl_sync(0xffffffffu, box.z, j), __shfl_sync(0xffffffffu, box.w, j));
            const float ka = __shfl_sync(0xffffffffu, my_area, j);
            const int d = suppresses_fast<MODE>(kb, ka, box, thr);
            supby |= (unsigned)(d == 1) << j;
            unsure |= (unsigned)(d == 2) << j;
          }
          supby &= pairs;
          unsure &= pairs;
          // rare: quotients within 1e-5 of the threshold are settled by the exact test
          for (unsigned u = __reduce_or_sync(0xffffffffu, unsure); u; u &= u - 1) {
            const int j = __ffs(u) - 1;
            const float4 kb = make_float4(__shfl_sync(0xffffffffu, box.x, j), __shfl_sync(0xffffffffu, box.y, j),
                                          __shfl_sync(0xffffffffu, box.z, j), __shfl_sync(0xffffffffu, box.w, j));
            const float ka = __shfl_sync(0xffffffffu, my_area, j);
            if (((unsure >> j) & 1u) && suppresses<MODE>(kb, ka, box, thr)) supby |= 1u << j;
          }
          // then the serial part is a ballot per kept box
          int nnew = 0;
          unsigned keepm = 0;
          while (am && kept + nnew < limit) {
            const int sl = __ffs(am) - 1;                      // next kept lane
            keepm |= 1u << sl;
            am &= ~__ballot_sync(0xffffffffu, (supby >> sl) & 1u);
            am &= ~(1u << sl);
            ++nnew;
          }
          if ((keepm >> lane) & 1u) {
            const int o = kept + __popc(keepm & ((1u << lane) - 1u));
            kbox[o] = box; karea[o] = my_area; kkey[o] = key;
          }
          alive = false;                                       // kept or suppressed, or past the limit: done either way
          if (lane == 0) s_new[w & 1] = nnew;
        }
        __syncthreads();
        kept += s_new[w & 1];               // the next writer of this slot (step w+2) is two barriers away
        // Only the warps about to take their turn catch up with the kept list: most chunks end after a few
        // steps (`limit` boxes kept), and the warps behind the window never need to test anything.
        if (kept < limit && wid > w && wid <= w + kDcWindow && alive) { catch_up(tested, kept); tested = kept; }
      }
      __syncthreads();      // kbox/karea of this chunk visible before the next chunk's pre-test
    }
    if (complete || kept >= limit) break;
    // the selected prefix ran dry: redo the pass on the next level
    __syncthreads();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // with the run merge downstream every class owns a fixed slot of the image's list (no atomic round trip)
    s_base = fixed_runs ? k * limit : atomicAdd(&img_cnt[n], kept);
    runs[(int64_t)n * K + k] = make_int2(s_base, kept);      // this class's survivors: a score-descending run
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kept; i += kDcThreads) {
    const unsigned long long key = kkey[i];
    const uint32_t row = (uint32_t)key;
    img_kept[(int64_t)n * kept_stride + s_base + i] =
        (key & 0xffffffff00000000ull) | (uint32_t)(row * (uint32_t)K + (uint32_t)k);
  }
}

// Pairwise merge of nl score-descending lists of <= L unique keys each (list t at src[t * L], length ls[t]) down
// to one list of the L best, all pairs of a round at once: an element's place in the merged list is its
// own index plus its lower bound in the partner list, found by a binary search in shared memory.
// ceil(log2 nl) rounds of independent work.  On return the result is src[0 .. ls[0]).
__device__ __forceinline__ void merge_lists(unsigned long long*& src, unsigned long long*& dst, int*& ls, int*& ld,
                                            int nl, int L) {
  while (nl > 1) {
    const int pairs = nl >> 1;
    for (int e = threadIdx.x; e < pairs * 2 * L; e += blockDim.x) {
      const int p = e / (2 * L), rem = e - p * 2 * L;
      const int side = rem >= L ? 1 : 0, i = rem - side * L;
      const int a = 2 * p + side, b = 2 * p + 1 - side;
      if (i < ls[a]) {
        const unsigned long long x = src[(size_t)a * L + i];
        const unsigned long long* B = src + (size_t)b * L;
        int lo = 0, hi = min(ls[b], L - i);                       // places >= L are dropped anyway
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (B[mid] < x) lo = mid + 1; else hi = mid;
        }
        if (i + lo < L) dst[(size_t)p * L + i + lo] = x;
      }
    }
    for (int p = threadIdx.x; p < pairs; p += blockDim.x) ld[p] = min(L, ls[2 * p] + ls[2 * p + 1]);
    if (nl & 1) {                                                 // the odd list moves on unchanged
      for (int i = threadIdx.x; i < ls[nl - 1]; i += blockDim.x) dst[(size_t)pairs * L + i] = src[(size_t)(nl - 1) * L + i];
      if (threadIdx.x == 0) ld[pairs] = ls[nl - 1];
    }
    __syncthreads();
    unsigned long long* t = src; src = dst; dst = t;
    int* tl = ls; ls = ld; ld = tl;
    nl = pairs + (nl & 1);
  }
}

// The topk best kept candidates of an image, in order (:207-208), and the output gather.
//   G > 0: every class left a score-descending run (det_class).  CTA (n, g) merges the runs of its share of
//          the classes to one list of the topk best; the last of an image's G CTAs to finish (ticket counter)
//          merges the G lists and writes the detections.
//   G = 0 (more classes than the run table holds): one CTA per image sorts / selects from the packed list.
__global__ void __launch_bounds__(kNmsThreads) det_topk_kernel(
    const int32_t* __restrict__ img_cnt, unsigned long long* __restrict__ img_kept, int64_t kept_stride,
    const int64_t* __restrict__ offsets, const float4* __restrict__ cboxes, int K, int topk, int sort_cap, int G,
    const int2* __restrict__ runs, unsigned long long* __restrict__ part, int* __restrict__ part_len, int* __restrict__ done,
    float* __restrict__ det_boxes, float* __restrict__ det_scores, int64_t* __restrict__ det_classes,
    int64_t* __restrict__ det_rows, int64_t* __restrict__ det_count) {
  extern __shared__ __align__(16) unsigned char sm[];
  __shared__ ArgMinSlots slots;
  __shared__ int s_ticket;
  unsigned long long* sel = reinterpret_cast<unsigned long long*>(sm);   // [topk]
  const int n = blockIdx.x;
  unsigned long long* keys = img_kept + (int64_t)n * kept_stride;
  int got;
  if (G > 0) {
    const int L = topk, g = blockIdx.y;
    const int rpc = (K + G - 1) / G, cap = max(rpc, G);
    unsigned long long* src = sel + topk;                         // [cap][L]
    unsigned long long* dst = src + (size_t)cap * L;              // [ceil(cap / 2)][L]
    int* ls = reinterpret_cast<int*>(dst + (size_t)((cap + 1) / 2) * L);   // [cap] list lengths
    int* ld = ls + cap;                                           // [cap]
    const int c0 = min(g * rpc, K), nl = min(K, c0 + rpc) - c0;
    const int2* rn = runs + (int64_t)n * K + c0;
    for (int e = threadIdx.x; e < nl * L; e += kNmsThreads) {
      const int c = e / L, q = e - c * L;
      const int2 r = rn[c];
      if (q < r.y) src[e] = keys[r.x + q];
    }
    for (int c = threadIdx.x; c < nl; c += kNmsThreads) ls[c] = min(rn[c].y, L);
    if (nl == 0 && threadIdx.x == 0) ls[0] = 0;
    __syncthreads();
    merge_lists(src, dst, ls, ld, nl, L);
    if (G > 1) {
      unsigned long long* mine = part + ((int64_t)n * G + g) * L;
      const int len = ls[0];
      for (int i = threadIdx.x; i < len; i += kNmsThreads) mine[i] = src[i];
      if (threadIdx.x == 0) part_len[n * G + g] = len;
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) s_ticket = atomicAdd(&done[n], 1);
      __syncthreads();
      if (s_ticket != G - 1) return;                              // someone else finishes the image
      __threadfence();
      src = sel + topk; dst = src + (size_t)cap * L;
      ls = reinterpret_cast<int*>(dst + (size_t)((cap + 1) / 2) * L); ld = ls + cap;
      __syncthreads();
      for (int t = threadIdx.x; t < G; t += kNmsThreads) ls[t] = __ldcg(part_len + n * G + t);
      __syncthreads();
      for (int e = threadIdx.x; e < G * L; e += kNmsThreads) {
        const int t = e / L, q = e - t * L;
        if (q < ls[t]) src[e] = __ldcg(part + ((int64_t)n * G + t) * L + q);
      }
      __syncthreads();
      merge_lists(src, dst, ls, ld, G, L);
    }
    got = ls[0];
    for (int i = threadIdx.x; i < got; i += kNmsThreads) sel[i] = src[i];
  } else if (sort_cap > 0) {
    // the image's kept list fits shared memory: one bitonic sort, the first topk keys are the answer
    const int cnt = img_cnt[n];
    unsigned long long* sk = sel + topk;
    int npad = 2;
    while (npad < cnt) npad <<= 1;
    for (int i = threadIdx.x; i < npad; i += kNmsThreads) sk[i] = i < cnt ? keys[i] : kDead;
    __syncthreads();
    bitonic_sort_smem(sk, npad);
    got = min(cnt, topk);
    for (int i = threadIdx.x; i < got; i += kNmsThreads) sel[i] = sk[i];
  } else {
    got = select_greedy<0, false>(keys, nullptr, img_cnt[n], 0.f, topk, sel, slots);
  }
  __syncthreads();
  const int64_t r0 = offsets[n];
  for (int i = threadIdx.x; i < topk; i += kNmsThreads) {
    const int64_t o = (int64_t)n * topk + i;
    float4* db = reinterpret_cast<float4*>(det_boxes) + o;
    if (i < got) {
      const unsigned long long key = sel[i];
      const uint32_t id = (uint32_t)key;
      const uint32_t row = id / (uint32_t)K, cls = id - row * (uint32_t)K;
      *db = cboxes[r0 + row];
      det_scores[o] = key_score(key);
      det_classes[o] = cls;
      det_rows[o] = row;
    } else {
      *db = make_float4(0.f, 0.f, 0.f, 0.f);
      det_scores[o] = 0.f;
      det_classes[o] = -1;
      det_rows[o] = -1;
    }
  }
  if (threadIdx.x == 0) det_count[n] = got;
}

// ------------------------------------------------------------------------------------------------
// generic batched_nms (vanilla strategy of torchvision/ops/boxes.py:97-120)
// ------------------------------------------------------------------------------------------------
__global__ void nms_hist_kernel(const int64_t* __restrict__ groups, int64_t M, int G, int32_t* __restrict__ counts) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const int64_t g = groups[i];
  if (g >= 0 && g < G) atomicAdd(&counts[g], 1);
}

__global__ void nms_scan_kernel(const int32_t* __restrict__ counts, int G, int32_t* __restrict__ starts) {
  // single CTA exclusive scan (G is the number of classes / levels: small)
  __shared__ int s_part[1024];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int g0 = 0; g0 < G; g0 += 1024) {
    const int g = g0 + threadIdx.x;
    const int v = g < G ? counts[g] : 0;
    s_part[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int t = threadIdx.x >= o ? s_part[threadIdx.x - o] : 0;
      __syncthreads();
      s_part[threadIdx.x] += t;
      __syncthreads();
    }
    if (g < G) starts[g] = s_carry + s_part[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry += s_part[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) starts[G] = s_carry;
}

__global__ void nms_scatter_kernel(const float* __restrict__ boxes, const float* __restrict__ scores,
                                   const int64_t* __restrict__ groups, int64_t M, int G,
                                   const int32_t* __restrict__ starts, int32_t* __restrict__ cursor,
                                   unsigned long long* __restrict__ keys, float4* __restrict__ segboxes) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const int64_t g = groups[i];
  if (g < 0 || g >= G) return;
  const int p = starts[g] + atomicAdd(&cursor[g], 1);
  keys[p] = make_key(scores[i], (uint32_t)i);
  segboxes[p] = __ldg(reinterpret_cast<const float4*>(boxes) + i);
}

template <int MODE>
__global__ void __launch_bounds__(kNmsThreads) nms_segment_kernel(
    const int32_t* __restrict__ starts, unsigned long long* __restrict__ keys,
    const float4* __restrict__ segboxes, unsigned long long* __restrict__ tmp, float thr, int cap,
    int32_t* __restrict__ total, unsigned long long* __restrict__ kept_all, int min_n) {
  extern __shared__ __align__(16) unsigned char sm[];
  __shared__ ArgMinSlots slots;
  __shared__ int s_base;
  const int g = blockIdx.x;
  const int s0 = starts[g], n = starts[g + 1] - s0;
  if (n <= min_n) return;                        // the blocked kernel took this group
  unsigned long long* k = keys + s0;
  const float4* b = segboxes + s0;
  if (n <= cap) {   // stage the segment in shared memory
    float4* sbox = reinterpret_cast<float4*>(sm);
    unsigned long long* skey = reinterpret_cast<unsigned long long*>(sm + (size_t)cap * sizeof(float4));
    for (int i = threadIdx.x; i < n; i += kNmsThreads) { sbox[i] = b[i]; skey[i] = k[i]; }
    __syncthreads();
    k = skey;
    b = sbox;
  }
  const int kept = select_greedy<MODE, true>(k, b, n, thr, -1, tmp + s0, slots);
  __syncthreads();
  if (threadIdx.x == 0) s_base = atomicAdd(total, kept);
  __syncthreads();
  for (int i = threadIdx.x; i < kept; i += kNmsThreads) kept_all[s_base + i] = tmp[s0 + i];
}

// ---- blocked greedy NMS of one group (round 2) ----------------------------------------------------------------
// nms_segment_kernel pays one pass over ALL candidates of the group and one barrier per KEPT box (a group of 6000
// boxes with 2700 survivors: 8.6 ms on a B200, torchvision's bitmask kernel 2.1 ms).  This flavour sorts the group
// once (bitonic, shared memory) and then walks it in blocks of 128 candidates, four barriers per BLOCK:
//   (a) every candidate of the block is tested against the boxes kept so far (four threads per candidate, each a
//       quarter of the kept list, early exit) -> alive;
//   (b) the pairs inside the block: cov[c] = the earlier candidates of the block that would suppress c (4 words);
//   (c) one warp settles the block 32 candidates at a time: kept = alive & !(cov & kept) has a unique fixed point
//       (lane l is final after l + 1 rounds; chains are two or three long), reached by iterating a ballot;
//   (d) the survivors' positions are appended to the kept list.
// Candidate x kept tests total n * kept / 2 instead of n * kept, with no barrier inside them.  Same greedy order and
// the same exact IoU arithmetic as select_greedy (suppresses_fast screen, then suppresses).
constexpr int kBlkThreads = 512;
constexpr int kBlk = 128;            // candidates per block
constexpr int kBlkCap = 8192;        // group size this flavour handles (sorted keys + boxes + kept list in shared memory)

template <int MODE>
__device__ __forceinline__ bool sup_exact(const float4 bi, const float ai, const float4 bj, const float thr) {
  const int f = suppresses_fast<MODE>(bi, ai, bj, thr);
  return f == 2 ? suppresses<MODE>(bi, ai, bj, thr) : (f == 1);
}

template <int MODE>
__global__ void __launch_bounds__(kBlkThreads) nms_blocked_kernel(
    const int32_t* __restrict__ starts, const unsigned long long* __restrict__ keys, const float4* __restrict__ segboxes,
    const float* __restrict__ boxes_all, float thr, int32_t* __restrict__ total, unsigned long long* __restrict__ kept_all,
    unsigned long long* __restrict__ tmp) {
  extern __shared__ __align__(16) unsigned char sm[];
  const int g = blockIdx.x;
  const int s0 = starts[g], n = starts[g + 1] - s0;
  if (n <= 0 || n > kBlkCap) return;             // larger groups: nms_segment_kernel (launched for them by the host)
  int npad = 2;                                    // even: the float4 array behind the keys stays 16-byte aligned
  while (npad < n) npad <<= 1;
  unsigned long long* skey = reinterpret_cast<unsigned long long*>(sm);                       // [npad]
  float4* sbox = reinterpret_cast<float4*>(sm + sizeof(unsigned long long) * (size_t)npad);   // [n] in sorted order
  uint16_t* kpos = reinterpret_cast<uint16_t*>(sbox + n);                                     // [n] kept positions
  __shared__ uint32_t s_alive[kBlk / 32], s_cov[kBlk][kBlk / 32], s_keptw[kBlk / 32];
  __shared__ int s_nkept, s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < npad; i += kBlkThreads) skey[i] = i < n ? keys[s0 + i] : kDead;
  if (tid == 0) s_nkept = 0;
  __syncthreads();
  // bitonic sort, ascending key = descending score, ties by candidate id
  for (int k = 2; k <= npad; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < npad; i += kBlkThreads) {
        const int p = i ^ j;
        if (p > i) {
          const unsigned long long a = skey[i], b = skey[p];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { skey[i] = b; skey[p] = a; }
        }
      }
      __syncthreads();
    }
  // boxes in sorted order (the candidate id is the low word of the key)
  for (int i = tid; i < n; i += kBlkThreads)
    sbox[i] = __ldg(reinterpret_cast<const float4*>(boxes_all) + (uint32_t)skey[i]);
  __syncthreads();
  (void)segboxes;
  for (int p0 = 0; p0 < n; p0 += kBlk) {
    const int nb = min(kBlk, n - p0);
    const int nk = s_nkept;
    // (a) against the kept list: thread = (candidate c, quarter q of the list)
    {
      const int c = tid >> 2, q = tid & 3;
      bool dead = c >= nb;
      if (!dead) {
        const float4 bc = sbox[p0 + c];
        int k = q;
        for (; k + 12 < nk && !dead; k += 16) {       // four kept boxes in flight per step: the loop is latency-bound
          const float4 b0 = sbox[kpos[k]], b1 = sbox[kpos[k + 4]], b2 = sbox[kpos[k + 8]], b3 = sbox[kpos[k + 12]];
          const int f0 = suppresses_fast<MODE>(b0, area_rn(b0), bc, thr), f1 = suppresses_fast<MODE>(b1, area_rn(b1), bc, thr);
          const int f2 = suppresses_fast<MODE>(b2, area_rn(b2), bc, thr), f3 = suppresses_fast<MODE>(b3, area_rn(b3), bc, thr);
          if ((f0 | f1 | f2 | f3) == 0) continue;      // four sure "no": the common case
          dead = (f0 == 2 ? suppresses<MODE>(b0, area_rn(b0), bc, thr) : f0 == 1) ||
                 (f1 == 2 ? suppresses<MODE>(b1, area_rn(b1), bc, thr) : f1 == 1) ||
                 (f2 == 2 ? suppresses<MODE>(b2, area_rn(b2), bc, thr) : f2 == 1) ||
                 (f3 == 2 ? suppresses<MODE>(b3, area_rn(b3), bc, thr) : f3 == 1);
        }
        for (; k < nk && !dead; k += 4) {
          const float4 bk = sbox[kpos[k]];
          dead = sup_exact<MODE>(bk, area_rn(bk), bc, thr);
        }
      }
      dead |= __shfl_xor_sync(0xffffffffu, dead, 1) != 0;
      dead |= __shfl_xor_sync(0xffffffffu, dead, 2) != 0;
      const uint32_t bal = __ballot_sync(0xffffffffu, !dead && q == 0);     // bits 0, 4, 8, ...: 8 candidates per warp
      if (lane == 0) {
        uint32_t m = 0;
#pragma unroll
        for (int u = 0; u < 8; ++u) m |= ((bal >> (4 * u)) & 1u) << u;
        reinterpret_cast<unsigned char*>(s_alive)[warp] = (unsigned char)m;  // candidates 8 * warp .. 8 * warp + 7
      }
    }
    // (b) pairs inside the block: thread = (candidate c, word w of 32 earlier candidates)
    {
      const int c = tid >> 2, w = tid & 3;
      uint32_t word = 0;
      if (c < nb && w * 32 < c) {
        const float4 bc = sbox[p0 + c];
        const int hi = min(32, c - w * 32);
        for (int u = 0; u < hi; ++u) {
          const float4 be = sbox[p0 + w * 32 + u];
          if (sup_exact<MODE>(be, area_rn(be), bc, thr)) word |= 1u << u;
        }
      }
      if (c < kBlk) s_cov[c][w] = word;
    }
    __syncthreads();
    // (c) one warp settles the block, 32 candidates at a time
    if (warp == 0) {
      uint32_t keptw[kBlk / 32];
#pragma unroll
      for (int gq = 0; gq < kBlk / 32; ++gq) {
        const int c = gq * 32 + lane;
        bool alive = c < nb && ((s_alive[gq] >> lane) & 1u);
#pragma unroll
        for (int e = 0; e < kBlk / 32; ++e)
          if (e < gq) alive = alive && !(s_cov[c][e] & keptw[e]);
        const uint32_t mine = s_cov[c][gq];
        uint32_t kept = __ballot_sync(0xffffffffu, alive);
        for (int it = 0; it < 32; ++it) {
          const uint32_t nk2 = __ballot_sync(0xffffffffu, alive && !(mine & kept));
          if (nk2 == kept) break;
          kept = nk2;
        }
        keptw[gq] = kept;
        if (lane == 0) s_keptw[gq] = kept;
      }
    }
    __syncthreads();
    // (d) append the survivors' positions in order
    if (tid < kBlk) {
      const int gq = tid >> 5;
      const uint32_t kw = s_keptw[gq];
      if ((kw >> lane) & 1u) {
        int before = __popc(kw & ((1u << lane) - 1u));
        for (int e = 0; e < gq; ++e) before += __popc(s_keptw[e]);
        kpos[nk + before] = (uint16_t)(p0 + tid);
      }
    }
    __syncthreads();
    if (tid == 0) {
      int add = 0;
      for (int e = 0; e < kBlk / 32; ++e) add += __popc(s_keptw[e]);
      s_nkept = nk + add;
    }
    __syncthreads();
  }
  const int kept = s_nkept;
  if (tid == 0) s_base = atomicAdd(total, kept);
  __syncthreads();
  for (int i = tid; i < kept; i += kBlkThreads) kept_all[s_base + i] = skey[kpos[i]];
  (void)tmp;
}

// --- global bitonic sort of kept_all[0..npad) (unused slots hold kDead and sort to the end) -------
constexpr int kSortTile = 4096;
constexpr int kSortThreads = 512;

__device__ __forceinline__ void cmpxchg(unsigned long long& a, unsigned long long& b, bool asc) {
  if ((a > b) == asc) { const unsigned long long t = a; a = b; b = t; }
}

// all stages with j < kSortTile for level k (k <= tile: full local sort when first==1)
__global__ void __launch_bounds__(kSortThreads) bitonic_local_kernel(unsigned long long* data, int64_t npad,
                                                                    int64_t k_lo, int64_t k_hi) {
  __shared__ unsigned long long s[kSortTile];
  const int64_t base = (int64_t)blockIdx.x * kSortTile;
  const int tile = (int)std::min<int64_t>(kSortTile, npad);
  for (int i = threadIdx.x; i < tile; i += kSortThreads) s[i] = data[base + i];
  __syncthreads();
  for (int64_t k = k_lo; k <= k_hi; k <<= 1) {
    for (int64_t j = std::min<int64_t>(k >> 1, tile >> 1); j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < tile; i += kSortThreads) {
        const int ixj = i ^ (int)j;
        if (ixj > i) {
          const bool asc = ((base + i) & k) == 0;
          unsigned long long a = s[i], b = s[ixj];
          cmpxchg(a, b, asc);
          s[i] = a; s[ixj] = b;
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < tile; i += kSortThreads) data[base + i] = s[i];
}

__global__ void bitonic_global_kernel(unsigned long long* data, int64_t npad, int64_t k, int64_t j) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npad) return;
  const int64_t ixj = i ^ j;
  if (ixj > i) {
    const bool asc = (i & k) == 0;
    unsigned long long a = data[i], b = data[ixj];
    cmpxchg(a, b, asc);
    data[i] = a; data[ixj] = b;
  }
}

__global__ void nms_emit_kernel(const unsigned long long* __restrict__ kept_all, const int32_t* __restrict__ total,
                                int64_t M, int64_t* __restrict__ keep, int64_t* __restrict__ num_keep) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *num_keep = *total;
  if (i >= M) return;
  keep[i] = i < *total ? (int64_t)(uint32_t)kept_all[i] : -1;
}

static int64_t next_pow2(int64_t v) { int64_t p = 1; while (p < v) p <<= 1; return p; }

struct NmsWs {
  size_t counts, starts, cursor, total, keys, segboxes, tmp, kept_all, bytes;
  int64_t npad;
};
static NmsWs nms_plan(int64_t M, int64_t G) {
  NmsWs w;
  size_t o = 0;
  auto take = [&](size_t b) { size_t r = o; o += align_up(b, 256); return r; };
  w.npad = next_pow2(std::max<int64_t>(M, 1));
  w.counts = take(sizeof(int32_t) * (size_t)(G + 1));
  w.cursor = take(sizeof(int32_t) * (size_t)(G + 1));
  w.total = take(sizeof(int32_t) * 4);
  w.starts = take(sizeof(int32_t) * (size_t)(G + 1));
  w.keys = take(sizeof(unsigned long long) * (size_t)M);
  w.segboxes = take(sizeof(float4) * (size_t)M);
  w.tmp = take(sizeof(unsigned long long) * (size_t)M);
  w.kept_all = take(sizeof(unsigned long long) * (size_t)w.npad);
  w.bytes = o;
  return w;
}

// shared memory of the run-merge flavour of det_topk: sel[topk] + cap lists + ceil(cap / 2) merged lists + two
// length arrays, cap = max(runs per CTA, G)
static size_t det_topk_smem(int64_t K, int G, int64_t topk) {
  const int64_t cap = std::max<int64_t>(ceil_div(K, G), G);
  return sizeof(unsigned long long) * ((size_t)topk + (size_t)(cap + (cap + 1) / 2) * (size_t)topk + (size_t)cap);
}

struct DetWs { size_t cboxes, img_cnt, cmax, tau, ccnt, zero_end, img_kept, runs, scoresT, rowok, part, part_len, bytes; int64_t kept_stride; int G; };
static DetWs det_plan(int64_t M, int64_t N, int64_t K, int64_t topk) {
  DetWs w;
  size_t o = 0;
  auto take = [&](size_t b) { size_t r = o; o += align_up(b, 256); return r; };
  w.kept_stride = K * std::max<int64_t>(topk, 0);
  // top-k stage: G CTAs per image, about eight class runs each (0: no run table, K too large)
  w.G = (M > 0 && K > 0 && K <= kMaxRunClasses) ? (int)std::min<int64_t>(32, ceil_div(K, 8)) : 0;
  if (w.G > 0 && det_topk_smem(K, w.G, topk) > 200 * 1024) w.G = 0;     // very large topk: packed list + sort / selection
  w.cboxes = take(sizeof(float4) * (size_t)M);
  w.img_cnt = take(sizeof(int32_t) * (size_t)(2 * N + 1));      // per-image kept counters, then the top-k tickets
  w.cmax = take(sizeof(unsigned) * (size_t)(N * std::max<int64_t>(K, 1)));   // per (image, class) best candidate score
  w.tau = take(sizeof(float) * (size_t)N);                       // per-image pruning threshold
  w.ccnt = take(sizeof(int32_t) * (size_t)(N * std::max<int64_t>(K, 1)));    // survivors per (image, class) (pruned path)
  w.zero_end = o;                                                // img_cnt .. tau are cleared by one memset
  w.img_kept = take(sizeof(unsigned long long) * (size_t)(N * w.kept_stride));
  w.runs = take(sizeof(int2) * (size_t)(N * std::max<int64_t>(K, 1)));
  w.scoresT = take(sizeof(float) * (size_t)(M * K));             // class-major scores, or (pruned path) the survivors' row lists
  w.rowok = take((size_t)M);                                     // pruned path: finite-row flags
  w.part = take(sizeof(unsigned long long) * (size_t)(N * std::max(w.G, 1) * std::max<int64_t>(topk, 0)));
  w.part_len = take(sizeof(int) * (size_t)(N * std::max(w.G, 1)));
  w.bytes = o;
  return w;
}

}  // namespace wsovod

using namespace wsovod;

static bool al16(const void* p) { return ((uintptr_t)p & 15) == 0; }

WSOVOD_API size_t wsovod_b200_batched_nms_workspace(int64_t M, int64_t num_groups) {
  if (M < 0 || num_groups < 0) return 0;
  return nms_plan(M, num_groups).bytes;
}

WSOVOD_API int wsovod_b200_batched_nms(const float* boxes, const float* scores, const int64_t* groups,
                                       int64_t M, int64_t num_groups, double iou_thresh, int iou_mode,
                                       int64_t* keep, int64_t* num_keep, void* workspace,
                                       size_t workspace_bytes, void* stream) {
  if (M < 0 || num_groups < 0 || !num_keep || (iou_mode != 0 && iou_mode != 1)) return WSOVOD_B200_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  if (M == 0 || num_groups == 0) {
    cudaError_t e = cudaMemsetAsync(num_keep, 0, sizeof(int64_t), st);
    if (e == cudaSuccess && M > 0 && keep) e = cudaMemsetAsync(keep, 0xff, sizeof(int64_t) * (size_t)M, st);
    return (int)e;
  }
  if (!boxes || !scores || !groups || !keep) return WSOVOD_B200_EINVAL;
  if (!al16(boxes)) return WSOVOD_B200_EALIGN;
  if (M >= (1LL << 31) || num_groups >= (1LL << 30)) return WSOVOD_B200_ETOOBIG;
  const NmsWs w = nms_plan(M, num_groups);
  if (!workspace || workspace_bytes < w.bytes) return WSOVOD_B200_EWORKSPACE;
  char* ws = (char*)workspace;
  int32_t* counts = (int32_t*)(ws + w.counts);
  int32_t* cursor = (int32_t*)(ws + w.cursor);
  int32_t* total = (int32_t*)(ws + w.total);
  int32_t* starts = (int32_t*)(ws + w.starts);
  unsigned long long* keys = (unsigned long long*)(ws + w.keys);
  float4* segboxes = (float4*)(ws + w.segboxes);
  unsigned long long* tmp = (unsigned long long*)(ws + w.tmp);
  unsigned long long* kept_all = (unsigned long long*)(ws + w.kept_all);
  cudaError_t e = cudaMemsetAsync(ws + w.counts, 0, w.starts - w.counts, st);   // counts, cursor, total
  if (e != cudaSuccess) return (int)e;
  e = cudaMemsetAsync(kept_all, 0xff, sizeof(unsigned long long) * (size_t)w.npad, st);
  if (e != cudaSuccess) return (int)e;
  int rc;
  const int G = (int)num_groups;
  nms_hist_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, st>>>(groups, M, G, counts);
  if ((rc = after_launch())) return rc;
  nms_scan_kernel<<<1, 1024, 0, st>>>(counts, G, starts);
  if ((rc = after_launch())) return rc;
  nms_scatter_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, st>>>(boxes, scores, groups, M, G, starts, cursor, keys, segboxes);
  if ((rc = after_launch())) return rc;
  const float thr = cmp_threshold(iou_thresh, iou_mode);
  const int cap = (int)std::min<int64_t>(M, 8192);
  const size_t smem = (size_t)cap * (sizeof(float4) + sizeof(unsigned long long));
  auto kern = iou_mode == 0 ? nms_segment_kernel<0> : nms_segment_kernel<1>;
  if (smem > 32 * 1024) {   // static slots + dynamic may cross the 48 KB default limit
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  // groups of at most kBlkCap candidates: the blocked flavour; larger ones: one pass per kept box (nms_segment_kernel
  // skips what the blocked kernel took -- `min_n`)
  {
    const int64_t nmax = std::min<int64_t>(M, kBlkCap);
    int64_t npad2 = 2;
    while (npad2 < nmax) npad2 <<= 1;
    const size_t bsmem = (size_t)npad2 * sizeof(unsigned long long) + (size_t)nmax * (sizeof(float4) + sizeof(uint16_t)) + 16;
    auto bk = iou_mode == 0 ? nms_blocked_kernel<0> : nms_blocked_kernel<1>;
    if (bsmem > 40 * 1024) {
      e = cudaFuncSetAttribute(bk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsmem);
      if (e != cudaSuccess) return (int)e;
    }
    bk<<<(unsigned)G, kBlkThreads, bsmem, st>>>(starts, keys, segboxes, boxes, thr, total, kept_all, tmp);
    if ((rc = after_launch())) return rc;
  }
  if (M > kBlkCap) {
    kern<<<(unsigned)G, kNmsThreads, smem, st>>>(starts, keys, segboxes, tmp, thr, cap, total, kept_all, kBlkCap);
    if ((rc = after_launch())) return rc;
  }
  // order all survivors by (score desc, index asc)
  const int64_t npad = w.npad;
  const unsigned tiles = (unsigned)std::max<int64_t>(1, npad / kSortTile);
  bitonic_local_kernel<<<tiles, kSortThreads, 0, st>>>(kept_all, npad, 2, std::min<int64_t>(npad, kSortTile));
  if ((rc = after_launch())) return rc;
  for (int64_t k = 2 * (int64_t)kSortTile; k <= npad; k <<= 1) {
    for (int64_t j = k >> 1; j >= kSortTile; j >>= 1) {
      bitonic_global_kernel<<<(unsigned)ceil_div(npad, 256), 256, 0, st>>>(kept_all, npad, k, j);
      if ((rc = after_launch())) return rc;
    }
    bitonic_local_kernel<<<tiles, kSortThreads, 0, st>>>(kept_all, npad, k, k);
    if ((rc = after_launch())) return rc;
  }
  nms_emit_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, st>>>(kept_all, total, M, keep, num_keep);
  return after_launch();
}

WSOVOD_API size_t wsovod_b200_detections_workspace(int64_t M, int64_t N, int64_t K, int64_t topk) {
  if (M < 0 || N < 0 || K < 0) return 0;
  return det_plan(M, N, K, topk).bytes;
}

WSOVOD_API int wsovod_b200_detections(const float* probs, const float* boxes, const int64_t* offsets,
                                      const float* image_sizes, int64_t M, int64_t N, int64_t K,
                                      int64_t max_rows_per_image, float score_thresh, double nms_thresh,
                                      int64_t topk, int iou_mode, float* det_boxes, float* det_scores,
                                      int64_t* det_classes, int64_t* det_rows, int64_t* det_count,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  if (M < 0 || N < 0 || K < 0 || max_rows_per_image < 0 || (iou_mode != 0 && iou_mode != 1))
    return WSOVOD_B200_EINVAL;
  if (N == 0) return 0;
  if (topk <= 0) return WSOVOD_B200_EUNSUPPORTED;   // "return all": use batched_nms on the filtered set
  if (!offsets || !image_sizes || !det_boxes || !det_scores || !det_classes || !det_rows || !det_count ||
      (M > 0 && (!probs || !boxes)))
    return WSOVOD_B200_EINVAL;
  if (!al16(boxes) || !al16(det_boxes)) return WSOVOD_B200_EALIGN;
  if (M >= (1LL << 31) || K >= 65535 || N >= 65535 || max_rows_per_image * std::max<int64_t>(K, 1) >= (1LL << 32) ||
      topk > 4096)
    return WSOVOD_B200_ETOOBIG;
  const DetWs w = det_plan(M, N, K, topk);
  if (!workspace || workspace_bytes < w.bytes) return WSOVOD_B200_EWORKSPACE;
  int npad_cap = 2;
  while (npad_cap < max_rows_per_image) npad_cap <<= 1;
  const int limit = (int)std::min<int64_t>(topk, std::max<int64_t>(max_rows_per_image, 1));
  const size_t smem = (size_t)npad_cap * sizeof(unsigned long long) +
                      (size_t)limit * (sizeof(float4) + sizeof(unsigned long long)) + sizeof(float) * (size_t)(limit + 2) +
                      sizeof(unsigned long long) * (size_t)kSelectCap + 16;
  if (smem > (size_t)kMaxSmemOptin - 2048) return WSOVOD_B200_EUNSUPPORTED;   // > 16384 proposals per image
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  float4* cboxes = (float4*)(ws + w.cboxes);
  int32_t* img_cnt = (int32_t*)(ws + w.img_cnt);
  unsigned long long* img_kept = (unsigned long long*)(ws + w.img_kept);
  int2* runs = (int2*)(ws + w.runs);
  cudaError_t e = cudaMemsetAsync(img_cnt, 0, w.zero_end - w.img_cnt, st);
  if (e != cudaSuccess) return (int)e;
  int rc;
  const bool use_runs = w.G > 0;
  if (M > 0 && K > 0) {
    float* scoresT = (float*)(ws + w.scoresT);
    // pruning needs at least topk classes (else the threshold is 0) and non-negative candidates (score bits ordered like
    // values: any threshold >= 0, the reference's 1e-5 / 0.05, guarantees that)
    const bool prune = K >= topk && K <= 8192 && score_thresh >= 0.f;
    if (prune) {
      const unsigned tiles = (unsigned)ceil_div(M, kScanRows);
      det_scan_kernel<<<tiles, 256, sizeof(unsigned) * (size_t)K, st>>>(probs, boxes, offsets, image_sizes, M, (int)N, (int)K + 1, cboxes,
                                                                       (uint8_t*)(ws + w.rowok), score_thresh, (unsigned*)(ws + w.cmax));
      if ((rc = after_launch())) return rc;
      det_tau_kernel<<<(unsigned)N, 1024, sizeof(unsigned) * (size_t)K, st>>>((const unsigned*)(ws + w.cmax), (int)K, (int)topk, (float*)(ws + w.tau));
      if ((rc = after_launch())) return rc;
      det_compact_kernel<<<tiles, 256, 0, st>>>(probs, offsets, M, (int)N, (int)K + 1, (const uint8_t*)(ws + w.rowok), score_thresh,
                                                (const float*)(ws + w.tau), (int32_t*)(ws + w.ccnt), (int32_t*)scoresT);
      if ((rc = after_launch())) return rc;
    } else {
      det_rows_kernel<<<(unsigned)ceil_div(M, kRtRows), 256, 0, st>>>(probs, boxes, offsets, image_sizes, M, (int)N, (int)K + 1, cboxes, scoresT);
      if ((rc = after_launch())) return rc;
    }
    auto kern = iou_mode == 0 ? det_class_kernel<0> : det_class_kernel<1>;
    if (smem > 32 * 1024) {   // static slots + dynamic may cross the 48 KB default limit
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
    }
    dim3 grid((unsigned)K, (unsigned)N);
    kern<<<grid, kDcThreads, smem, st>>>(scoresT, M, offsets, cboxes, (int)K, score_thresh,
                                         cmp_threshold(nms_thresh, iou_mode), limit, npad_cap, use_runs, img_cnt, img_kept, w.kept_stride, runs,
                                         prune ? (const float*)(ws + w.tau) : nullptr, prune ? (const int32_t*)scoresT : nullptr,
                                         (const int32_t*)(ws + w.ccnt), probs);
    if ((rc = after_launch())) return rc;
  }
  // final ordering: sort the image's kept list in shared memory when it fits (K * topk <= 16384 keys)
  int sort_cap = 2;
  while (sort_cap < w.kept_stride) sort_cap <<= 1;
  if (w.kept_stride > 16384) sort_cap = 0;
  const int G = w.G;
  const size_t tsmem = G > 0 ? det_topk_smem(K, G, topk) : sizeof(unsigned long long) * ((size_t)topk + (size_t)sort_cap);
  if (tsmem > 32 * 1024) {
    e = cudaFuncSetAttribute(det_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem);
    if (e != cudaSuccess) return (int)e;
  }
  det_topk_kernel<<<dim3((unsigned)N, (unsigned)std::max(G, 1)), kNmsThreads, tsmem, st>>>(
      img_cnt, img_kept, w.kept_stride, offsets, cboxes, (int)std::max<int64_t>(K, 1), (int)topk, sort_cap, G,
      runs, (unsigned long long*)(ws + w.part), (int*)(ws + w.part_len), img_cnt + N + 1, det_boxes, det_scores, det_classes,
      det_rows, det_count);
  return after_launch();
}
