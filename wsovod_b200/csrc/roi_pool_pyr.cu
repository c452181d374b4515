// roi_pool_pyr.cu -- block-max ("pyramid") fast path of kernel family (1): ROI max-pool 7x7, values only.
//
// Same ownership as roi_pool.cu (a CTA owns (image, CB channels) and walks every proposal of the image;
// lanes = consecutive (proposal, bin) outputs; the feature planes cross L2->SM once), but the
// shared-memory plane is not the raw map: it is D_k[h][w] = max of the kh x kw block at (h, w), so a
// bin costs <= 2 x 2 (at most 4 x 4) LDS instead of nh x nw -- see pool_pyr.cuh.  Proposals are grouped
// by (kh, kw) in a prologue; the CTA walks the groups ("phases") and rebuilds D in place between them
// with doubling steps D[i] = max(D[i], D[i + stride]).  Values are exact copies of input cells, so the
// result is bit-identical to the scan of ROILoopPool_cpu.cpp:52-79 (max is order independent; NaN and
// -inf never win against the -FLT_MAX start exactly as `v > maxval` never lets them).
//
// Launch sequence on the caller's stream: memset (histogram + cursors) -> pyr_classify (per proposal: class
// key, 7 + 7 descriptor entries) -> pyr_order (per image: bucket offsets, scatter) -> pyr_bins (per lane
// slot: 32-bit descriptor) -> roi_pool7_pyr_kernel (1024 threads, one CTA per SM, grid = images x channel
// groups [x proposal splits]).  With four channels per CTA the lane-slot stream is padded to 64 slots
// per proposal so that no warp store straddles two proposals (Kernel 1b).
#include "common.cuh"
#include "pool_pyr.cuh"

#include <algorithm>
#include <cstdlib>

namespace wsovod {

using namespace pyr;

struct PyrWs {
  int32_t* hist;        // [N, kBuckets] followed by cursor [N, kBuckets] (zeroed together)
  int32_t* cursor;      // [N, kBuckets] scatter cursors of pyr_order_kernel
  uint32_t* axtab;      // [R, 14] per-proposal row / column descriptor entries (pool_pyr.cuh: axis_entry)
  int32_t* img_start;   // [N + 1]   first sorted position of each image
  int32_t* bucket_off;  // [N, kBuckets + 1] bucket boundaries (positions relative to the image start)
  int32_t* bidx;        // [R]
  uint32_t* pkey;       // [R]
  int32_t* order;       // [R]   proposal id at each sorted position
  uint2* pinfo;         // [R]   (proposal id | (ch-1) << 26 | (cw-1) << 28, bits of row_scale + bias) per sorted position
  uint32_t* desc;       // [R, 64] bin descriptors in sorted position / lane slot order (slots 49..63 idle)
  size_t bytes;
};

static PyrWs pyr_carve(void* ws, int64_t N, int64_t R) {
  PyrWs w;
  size_t off = 0;
  char* base = (char*)ws;
  auto take = [&](size_t b) { size_t o = off; off += align_up(b, 256); return base + o; };
  w.hist = (int32_t*)take(sizeof(int32_t) * (size_t)N * kBuckets * 2);
  w.cursor = w.hist + (size_t)N * kBuckets;
  w.axtab = (uint32_t*)take(sizeof(uint32_t) * (size_t)R * 14);
  w.img_start = (int32_t*)take(sizeof(int32_t) * (size_t)(N + 1));
  w.bucket_off = (int32_t*)take(sizeof(int32_t) * (size_t)N * (kBuckets + 1));
  w.bidx = (int32_t*)take(sizeof(int32_t) * (size_t)R);
  w.pkey = (uint32_t*)take(sizeof(uint32_t) * (size_t)R);
  w.order = (int32_t*)take(sizeof(int32_t) * (size_t)R);
  w.pinfo = (uint2*)take(sizeof(uint2) * (size_t)R);
  w.desc = (uint32_t*)take(sizeof(uint32_t) * (size_t)R * kSlots);
  w.bytes = off;
  return w;
}

size_t pool7_pyr_workspace(int64_t N, int64_t R) { return pyr_carve(nullptr, N, R).bytes; }

// ------------------------------------------------------------------------------------------------
// prologue: classify -> per-image bucket sort -> descriptors in sorted order
// ------------------------------------------------------------------------------------------------
__global__ void pyr_classify_kernel(const float* __restrict__ rois, int64_t R, int N, int H, int W, float scale,
                                    int32_t* __restrict__ bidx, uint32_t* __restrict__ pkey,
                                    int32_t* __restrict__ hist, uint32_t* __restrict__ axtab) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float* roi = rois + r * 5;
  int b = (int)roi[0];
  b = min(max(b, 0), N - 1);
  const float x1 = roi[1], y1 = roi[2], x2 = roi[3], y2 = roi[4];
  const uint32_t key = proposal_key(x1, y1, x2, y2, scale, H, W);
  bidx[r] = b;
  pkey[r] = key;
  atomicAdd(&hist[(int64_t)b * kBuckets + key_bucket(key)], 1);
  const int phase = key_phase(key);
  if (phase != PH_FALLBACK) {     // the 7 row and 7 column entries every bin descriptor of this proposal is made of
    const Axis ah = axis_of(y1, y2, scale), aw = axis_of(x1, x2, scale);
    const int kh = phase_kh(phase), kw = phase_kw(phase);
#pragma unroll
    for (int p = 0; p < 7; ++p) {
      axtab[r * 14 + p] = axis_entry(ah, p, H, kh, W + kPad);
      axtab[r * 14 + 7 + p] = axis_entry(aw, p, W, kw, 1);
    }
  }
}

// grid (image, part): every CTA derives the first sorted position of its image (sum of the histograms before
// it) and the bucket offsets (block scan); part p scatters the proposal ids of its slice of the roi list
// through global per-bucket cursors.  The order inside a bucket is whatever the atomics give: every
// output element is written exactly once from position-independent data, so the result does not
// depend on it.
__global__ void __launch_bounds__(512) pyr_order_kernel(const int32_t* __restrict__ bidx, const uint32_t* __restrict__ pkey,
                                                        const int32_t* __restrict__ hist, int32_t* __restrict__ cursor,
                                                        int N, int64_t R, int32_t* __restrict__ order,
                                                        int32_t* __restrict__ img_start, int32_t* __restrict__ bucket_off) {
  __shared__ int s_cur[kBuckets];
  __shared__ int s_red[16];
  __shared__ int s_base, s_total;
  static_assert(kBuckets <= 512, "one bucket per thread");
  const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // (1) proposals of the images before this one
  int sum = 0;
  for (int64_t i = tid; i < (int64_t)n * kBuckets; i += blockDim.x) sum += hist[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) s_red[wid] = sum;
  // (2) exclusive scan of this image's buckets (kBuckets <= 5 * 32: warps 0..4, one bucket per lane)
  int mine = 0;
  if (tid < kBuckets) mine = hist[(int64_t)n * kBuckets + tid];
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  __shared__ int s_wtot[16];
  if (lane == 31) s_wtot[wid] = incl;
  __syncthreads();
  if (tid == 0) {
    int t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_red[w];
    s_base = t;
  }
  int wbase = 0;
  for (int w = 0; w < wid; ++w) wbase += s_wtot[w];
  if (tid < kBuckets) {
    const int excl = wbase + incl - mine;
    s_cur[tid] = excl;
    if (blockIdx.y == 0) {
      bucket_off[n * (kBuckets + 1) + tid] = excl;
      if (tid == kBuckets - 1) bucket_off[n * (kBuckets + 1) + kBuckets] = excl + mine;
    }
    if (tid == kBuckets - 1) s_total = excl + mine;
  }
  __syncthreads();
  const int base = s_base;
  if (tid == 0 && blockIdx.y == 0) {
    img_start[n] = base;
    if (n == N - 1) img_start[N] = base + s_total;
  }
  const int64_t per = (R + gridDim.y - 1) / gridDim.y;
  const int64_t r_lo = (int64_t)blockIdx.y * per, r_hi = min(R, r_lo + per);
  for (int64_t r = r_lo + tid; r < r_hi; r += blockDim.x) {
    if (bidx[r] != n) continue;
    const int bucket = key_bucket(pkey[r]);
    const int pos = s_cur[bucket] + atomicAdd(&cursor[(int64_t)n * kBuckets + bucket], 1);
    order[base + pos] = (int32_t)r;
  }
}

// Descriptors of the lane slots of every sorted position.  With four channels per CTA a proposal takes
// kSlots = 64 slots = two 32-lane passes of the main kernel: bins 0..31 in the first, bins 32..48 in the
// second (15 idle slots), so that every warp store stays inside ONE proposal -- warp stores that straddle
// two proposals (two pieces 100 KB apart, four partial 32-byte sectors) cost the whole c2 kernel 1.41 ms
// in stores alone, proposal-aligned passes 0.88 ms (tools/ubench/store_pattern.cu, "pattern" vs "V6").
// The two-channel flavour (larger maps, fewer bytes per pass) is not store-bound and keeps the dense
// 49-slot stream.
//
// WHICH lane of a pass serves which of the pass's bins is free: the set of addresses a warp store covers
// does not depend on it.  The hot loop's LDS.128 are issued per quarter-warp (8 lanes, 128 bytes), and a
// quarter is conflict-free when its 8 cells fall into 8 different 16-byte bank groups (cell index mod 8).
// Consecutive bins of a proposal step by a near-constant number of cells, so row-major lanes collide
// whenever that step is even (simulated on c2: 1.73 wavefronts per ideal wavefront, ncu: 1.66).  GROUP
// therefore deals the bins of a pass into quarters by first fit: a bin goes to the quarter where the
// fewest of its (up to four) first block loads meet a bank group some lane of the quarter already uses
// (ties: the fuller quarter).  tools/bank_sim.py replays it: 1.70 -> 1.44 wavefronts per ideal wavefront;
// ncu on c2: 72.8 M -> 40.3 M conflict wavefronts of the main kernel.
template <int SLOTS, bool GROUP>
__global__ void __launch_bounds__(256) pyr_bins_kernel(int64_t R, int H, int W, const int32_t* __restrict__ order,
                                                       const uint32_t* __restrict__ pkey, const uint32_t* __restrict__ axtab,
                                                       const float* __restrict__ row_scale, float row_scale_bias,
                                                       uint2* __restrict__ pinfo, uint32_t* __restrict__ desc) {
  if constexpr (!GROUP) {
    // eight threads per proposal: thread t < 7 writes lane slots 7t .. 7t+6, thread 7 the idle tail and the
    // proposal record (one wave of threads and one chain of dependent loads instead of a thread per slot)
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t gpos = i >> 3;
    const int t = (int)(i & 7);
    if (gpos >= R) return;
    const int r = order[gpos];
    const uint32_t key = pkey[r];
    uint32_t* d = desc + gpos * SLOTS;
    if (t == 7) {
#pragma unroll
      for (int q = 49; q < SLOTS; ++q) d[q] = kDescIdle;
      const float sc = row_scale ? __fadd_rn(row_scale[r], row_scale_bias) : 1.f;   // roi_heads.py:733-739
      pinfo[gpos] = make_uint2((uint32_t)r | (((key >> 4) & 3u) << 26) | (((key >> 6) & 3u) << 28), __float_as_uint(sc));
      return;
    }
    const bool fallback = key_phase(key) == PH_FALLBACK;
    const uint32_t* ax = axtab + (int64_t)r * 14;
#pragma unroll
    for (int u = 0; u < 7; ++u) {
      const int q = t * 7 + u;
      const int bin = slot_bin(key, q);
      const int ph = bin / 7, pw = bin - ph * 7;
      d[q] = fallback ? 0u : combine_desc(__ldg(ax + ph), __ldg(ax + 7 + pw), bin, H, W);
    }
  } else {
    static_assert(!GROUP || SLOTS == kSlots, "grouping needs the proposal-aligned 64-slot stream");
    // two threads per proposal: thread 0 the first pass (bins 0..31, four quarters), thread 1 the second
    // (bins 32..48, three quarters) and the proposal record.  occ[quarter] = bank groups taken in that quarter,
    // one byte per block load (0,0), (0,1), (1,0), (1,1); a bin's conflicts with a quarter = popc(occ & mask).
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t gpos = i >> 1;
    const int half = (int)(i & 1);
    if (gpos >= R) return;
    const int r = order[gpos];
    const uint32_t key = pkey[r];
    uint32_t* d = desc + gpos * SLOTS + half * 32;
    if (half) {
      const float sc = row_scale ? __fadd_rn(row_scale[r], row_scale_bias) : 1.f;   // roi_heads.py:733-739
      pinfo[gpos] = make_uint2((uint32_t)r | (((key >> 4) & 3u) << 26) | (((key >> 6) & 3u) << 28), __float_as_uint(sc));
    }
    const int b_lo = half ? 32 : 0, b_hi = half ? 49 : 32;
    const int phase = key_phase(key);
    if (phase == PH_FALLBACK) {          // lane slot = output bin (the fallback scan recomputes its edges)
      for (int q = 0; q < 32; ++q) d[q] = b_lo + q < 49 ? 0u : kDescIdle;
      return;
    }
    const uint32_t* ax = axtab + (int64_t)r * 14;
    const uint32_t kh = (uint32_t)phase_kh(phase), kw = (uint32_t)phase_kw(phase);
    const uint32_t WP = (uint32_t)(W + kPad);
    uint32_t occ0 = 0, occ1 = 0, occ2 = 0, occ3 = 0;
    uint32_t fill = half ? 0x8000u : 0u;   // four nibbles: lanes taken in each quarter (the second pass has three)
    uint32_t written = 0;                  // lane slots of this pass that received a bin
    int ph = b_lo / 7, pw = b_lo - ph * 7;
    uint32_t re = __ldg(ax + ph);
    for (int bin = b_lo; bin < b_hi; ++bin) {
      const uint32_t dd = combine_desc(re, __ldg(ax + 7 + pw), bin, H, W);
      const uint32_t c0 = dd & 0xffffu, lh = (dd >> 16) & 15u, lw = (dd >> 20) & 15u;
      const uint32_t ro = min(kh, lh) * WP, co = min(kw, lw);
      uint32_t m = 1u << (c0 & 7u);
      if (lw) m |= 0x100u << ((c0 + co) & 7u);
      if (lh) m |= 0x10000u << ((c0 + ro) & 7u);
      if (lw && lh) m |= 0x1000000u << ((c0 + ro + co) & 7u);
      // first fit: fewest occupied bank groups, ties to the fuller quarter (full quarters are out)
      int best_q = 0, best_key = 1 << 30;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t o = q == 0 ? occ0 : q == 1 ? occ1 : q == 2 ? occ2 : occ3;
        const int f = (int)((fill >> (4 * q)) & 15u);
        const int k = f >= 8 ? (1 << 29) : (__popc(o & m) << 4) + (8 - f);
        if (k < best_key) { best_key = k; best_q = q; }
      }
      const int slot = best_q * 8 + (int)((fill >> (4 * best_q)) & 15u);
      d[slot] = dd;
      written |= 1u << slot;
      fill += 1u << (4 * best_q);
      occ0 |= best_q == 0 ? m : 0u; occ1 |= best_q == 1 ? m : 0u; occ2 |= best_q == 2 ? m : 0u; occ3 |= best_q == 3 ? m : 0u;
      if (++pw == 7) { pw = 0; ++ph; re = __ldg(ax + (ph < 7 ? ph : 6)); }
    }
    for (int q = 0; q < 32; ++q)
      if (!((written >> q) & 1u)) d[q] = kDescIdle;
  }
}

// ------------------------------------------------------------------------------------------------
// main kernel
// ------------------------------------------------------------------------------------------------
struct PyrParams {
  const float* input;
  const float* rois;
  float scale;
  float floor_v;        // values-only flavours: cells are staged as max(cell, floor_v); -FLT_MAX = torchvision roi_pool,
                        // 0 = the maxima "starting at 0" of ROILoopPool (ROILoopPool_cuda.cu:107-113)
  float* output;
  int32_t* argmax;      // ARG flavour only
  const int32_t* img_start;
  const int32_t* bucket_off;
  const uint2* pinfo;
  const uint32_t* desc;
  int32_t N, C, H, W;
  int32_t CG, S;
};

// A plane cell holds CB channel values; the ARG flavour (argmax requested: training with a trainable backbone,
// ROILoopPool_cuda.cu:206-248 reads it back) holds CB (value, index) pairs instead, index = h * W + w of the cell the
// value came from.  Blocks are merged by  "greater value, or equal value and smaller index": the maximum over a set of
// cells then carries the FIRST cell of the row-major scan that attains it, which is what `v > maxval` leaves behind in
// the reference's scan (ROILoopPool_cpu.cpp:52-79, torchvision roi_pool) whatever the order the blocks are visited in.
// Cells that can never win the reference's comparison (NaN, -inf, -FLT_MAX itself) are staged as (-FLT_MAX, -1), like
// the pad cells: a bin of only such cells returns (-FLT_MAX, -1) as the reference does.
// Cell size: 4 * CB bytes, ARG: 8 * CB (CB = 2 only: 16-byte cells, the same descriptors and lane order as CB = 4).
template <int CB, bool ARG> struct CellT { static constexpr uint32_t CS = (ARG ? 8u : 4u) * CB; };

template <int CB, bool ARG> __device__ __forceinline__ void p_lds(uint32_t addr, float* f, int* a);
template <> __device__ __forceinline__ void p_lds<4, false>(uint32_t addr, float* f, int*) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f[0]), "=f"(f[1]), "=f"(f[2]), "=f"(f[3]) : "r"(addr));
}
template <> __device__ __forceinline__ void p_lds<2, false>(uint32_t addr, float* f, int*) {
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(f[0]), "=f"(f[1]) : "r"(addr));
}
template <> __device__ __forceinline__ void p_lds<2, true>(uint32_t addr, float* f, int* a) {
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=f"(f[0]), "=f"(f[1]), "=r"(a[0]), "=r"(a[1]) : "r"(addr));
}
template <int CB, bool ARG> __device__ __forceinline__ void p_sts(uint32_t addr, const float* f, const int* a);
template <> __device__ __forceinline__ void p_sts<4, false>(uint32_t addr, const float* f, const int*) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(f[0]), "f"(f[1]), "f"(f[2]), "f"(f[3]) : "memory");
}
template <> __device__ __forceinline__ void p_sts<2, false>(uint32_t addr, const float* f, const int*) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(f[0]), "f"(f[1]) : "memory");
}
template <> __device__ __forceinline__ void p_sts<2, true>(uint32_t addr, const float* f, const int* a) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(f[0]), "f"(f[1]), "r"(a[0]), "r"(a[1]) : "memory");
}
// m <- merge(m, f)
template <int CB, bool ARG> __device__ __forceinline__ void p_max(float* m, int* ma, const float* f, const int* a) {
#pragma unroll
  for (int k = 0; k < CB; ++k) {
    if (ARG) {
      const bool take = f[k] > m[k] || (f[k] == m[k] && a[k] < ma[k]);
      m[k] = take ? f[k] : m[k];
      ma[k] = take ? a[k] : ma[k];
    } else {
      m[k] = fmaxf(m[k], f[k]);
    }
  }
}
template <int CB, bool ARG> __device__ __forceinline__ void p_copy(float* m, int* ma, const float* f, const int* a) {
#pragma unroll
  for (int k = 0; k < CB; ++k) { m[k] = f[k]; if (ARG) ma[k] = a[k]; }
}

// D <- plane of the map (MODE 0), max with the right neighbour (1: block 1x2) or the lower neighbour
// (2: block 2x1), in the padded layout; pad cells and absent channels hold the identity -FLT_MAX, and
// NaN / -inf cells are clamped to it (they can never win a bin: `v > maxval` with maxval = -FLT_MAX).
// A warp takes whole rows of the padded plane (row hh = warp, warp + #warps, ...), lanes walk the columns
// 32 at a time: addresses are (row pointer of the channel) + column, the only predicates are the row /
// column borders, and up to 8 x CB loads per lane are in flight.
template <int CB, int MODE, bool ARG>
__device__ __noinline__ void pyr_stage(uint32_t sbase, const float* __restrict__ src, int nc, int H, int W, float floor_v) {
  constexpr uint32_t CS = CellT<CB, ARG>::CS;
  const int WP = W + kPad, HP = H + kPad, HW = H * W;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int hh = wid; hh < HP; hh += nw) {
    const int h = hh - kPad;
    const uint32_t srow = sbase + (uint32_t)(hh * WP) * CS;
    const bool hin = h >= 0;                                   // a real map row
    const bool h2 = MODE == 2 && h + 1 >= 0 && h + 1 < H;      // its lower neighbour exists
    const float* rowp = src + (int64_t)(hin ? h : 0) * W;
    const float* rowq = src + (int64_t)(h2 ? h + 1 : 0) * W;
    for (int ww0 = 0; ww0 < WP; ww0 += 64) {
      float f[2][CB], g[2][CB];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int ww = ww0 + u * 32 + lane, w = ww - kPad;
        const bool in = hin && ww < WP && w >= 0;
        const bool in2 = ww < WP && (MODE == 1 ? (hin && w + 1 >= 0 && w + 1 < W) : (MODE == 2 && h2 && w >= 0));
#pragma unroll
        for (int k = 0; k < CB; ++k) {
          f[u][k] = (in && k < nc) ? __ldg(rowp + (int64_t)k * HW + w) : -FLT_MAX;
          g[u][k] = (in2 && k < nc) ? __ldg((MODE == 1 ? rowp + 1 : rowq) + (int64_t)k * HW + w) : -FLT_MAX;
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int ww = ww0 + u * 32 + lane;
        if (ww < WP) {
          int fa[CB], ga[CB];
          if (ARG) {
            const int idx = h * W + (ww - kPad);                 // only used where the cell is a real one
#pragma unroll
            for (int k = 0; k < CB; ++k) {
              fa[k] = f[u][k] > -FLT_MAX ? idx : -1;             // false for NaN too
              ga[k] = g[u][k] > -FLT_MAX ? idx + (MODE == 1 ? 1 : W) : -1;
              f[u][k] = fmaxf(f[u][k], -FLT_MAX);
              g[u][k] = fmaxf(g[u][k], -FLT_MAX);
            }
            p_max<CB, ARG>(f[u], fa, g[u], ga);
          } else {
#pragma unroll
            for (int k = 0; k < CB; ++k) f[u][k] = fmaxf(fmaxf(f[u][k], g[u][k]), floor_v);
          }
          p_sts<CB, ARG>(srow + (uint32_t)ww * CS, f[u], fa);
        }
      }
    }
  }
}

// generic in-place doubling (any stride; used where a plane row is wider than the CTA): reads only go
// forward, so chunks are processed front to back with one barrier between a chunk's loads and its
// stores.  Horizontal steps may wrap into the next row's pad columns, which hold the identity for the
// strides used (1, 2 with 3 pad columns); vertical steps run into the identity tail rows.
template <int CB, bool ARG>
__device__ __noinline__ void pyr_double(uint32_t sbase, int ncell, int stride) {
  constexpr uint32_t CS = CellT<CB, ARG>::CS;
  constexpr int U = 4;
  const uint32_t sb = (uint32_t)stride * CS;
  for (int base = 0; base < ncell; base += (int)blockDim.x * U) {
    float v[U][CB];
    int va[U][CB];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int idx = base + u * (int)blockDim.x + (int)threadIdx.x;
      if (idx < ncell) {
        float b[CB];
        int ba[CB];
        p_lds<CB, ARG>(sbase + (uint32_t)idx * CS, v[u], va[u]);
        p_lds<CB, ARG>(sbase + (uint32_t)idx * CS + sb, b, ba);
        p_max<CB, ARG>(v[u], va[u], b, ba);
      }
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int idx = base + u * (int)blockDim.x + (int)threadIdx.x;
      if (idx < ncell) p_sts<CB, ARG>(sbase + (uint32_t)idx * CS, v[u], va[u]);
    }
  }
  __syncthreads();
}

// in-place doubling D[i] = max(D[i], D[i + stride]) over the first `ncell` cells (every output from the ORIGINAL
// values).  Round 1 read both operands and wrote the result per cell, chunk by chunk with a barrier per chunk
// (3 shared-memory accesses per cell).  Now every thread owns a run of cells it walks front to back with the
// operands in a register window: it first saves the `s` cells just behind its run (the head of the next run, which
// that run's owner will overwrite), one barrier, then one load and one store per cell.
//   horizontal (stride 1 or 2 cells): runs of kRun consecutive cells of the linear plane; kRun is odd so the 8 lanes
//     of a quarter-warp (kRun cells apart) hit 8 different 16-byte bank groups.  A step may wrap into the next row's
//     pad columns, which hold the identity for the strides used (3 pad columns).
//   vertical (stride 1 or 2 rows): a thread owns a column segment, lanes are consecutive columns (conflict-free);
//     the last segments run into the identity tail rows.
template <int CB, int S, bool ARG>
__device__ __noinline__ void pyr_double_h(uint32_t sbase, int ncell) {
  constexpr int s = S;
  constexpr uint32_t CS = CellT<CB, ARG>::CS;
  const int run = max(13, (((ncell + (int)blockDim.x - 1) / (int)blockDim.x) | 1));   // odd, >= cells per thread
  const int a = (int)threadIdx.x * run, b = min(a + run, ncell);
  float ov0[CB], ov1[CB];
  int oa0[CB], oa1[CB];
  if (a < ncell) {
    p_lds<CB, ARG>(sbase + (uint32_t)b * CS, ov0, oa0);
    if (S == 2) p_lds<CB, ARG>(sbase + (uint32_t)(b + 1) * CS, ov1, oa1);     // b + 1 <= ncell + 1 < ntot: identity tail rows
  }
  __syncthreads();
  if (a < ncell) {
    float p0[CB], p1[CB], nx[CB];
    int a0[CB], a1[CB], na[CB];
    p_lds<CB, ARG>(sbase + (uint32_t)a * CS, p0, a0);
    if (s == 2) {
      if (a + 1 < b) p_lds<CB, ARG>(sbase + (uint32_t)(a + 1) * CS, p1, a1);
      else p_copy<CB, ARG>(p1, a1, ov0, oa0);
    }
    for (int j = a; j < b; ++j) {
      const int q = j + s;
      if (q < b) p_lds<CB, ARG>(sbase + (uint32_t)q * CS, nx, na);
      else if (S == 2 && (q - b)) p_copy<CB, ARG>(nx, na, ov1, oa1);
      else p_copy<CB, ARG>(nx, na, ov0, oa0);
      float o[CB];
      int oa[CB];
      p_copy<CB, ARG>(o, oa, p0, a0);
      p_max<CB, ARG>(o, oa, nx, na);
      p_sts<CB, ARG>(sbase + (uint32_t)j * CS, o, oa);
      if (s == 2) p_copy<CB, ARG>(p0, a0, p1, a1); else p_copy<CB, ARG>(p0, a0, nx, na);
      p_copy<CB, ARG>(p1, a1, nx, na);
    }
  }
  __syncthreads();
}

template <int CB, int S, bool ARG>
__device__ __noinline__ void pyr_double_v(uint32_t sbase, int rows, int WP) {
  constexpr int s = S;
  constexpr uint32_t CS = CellT<CB, ARG>::CS;
  if (WP > (int)blockDim.x) { pyr_double<CB, ARG>(sbase, rows * WP, S * WP); return; }
  const int nseg = max(1, (int)blockDim.x / WP);                 // column segments
  const int seg = (int)threadIdx.x / WP, w = (int)threadIdx.x - seg * WP;
  const int per = (rows + nseg - 1) / nseg;
  const int r0 = seg * per, r1 = min(r0 + per, rows);
  const bool on = seg < nseg && r0 < rows;
  const uint32_t col = sbase + (uint32_t)w * CS, pitch = (uint32_t)WP * CS;
  float ov0[CB], ov1[CB];
  int oa0[CB], oa1[CB];
  if (on) {
    p_lds<CB, ARG>(col + (uint32_t)r1 * pitch, ov0, oa0);
    if (S == 2) p_lds<CB, ARG>(col + (uint32_t)(r1 + 1) * pitch, ov1, oa1);   // rows + 1 < rows + kTailRows
  }
  __syncthreads();
  if (on) {
    float p0[CB], p1[CB], nx[CB];
    int a0[CB], a1[CB], na[CB];
    p_lds<CB, ARG>(col + (uint32_t)r0 * pitch, p0, a0);
    if (s == 2) {
      if (r0 + 1 < r1) p_lds<CB, ARG>(col + (uint32_t)(r0 + 1) * pitch, p1, a1);
      else p_copy<CB, ARG>(p1, a1, ov0, oa0);
    }
    for (int j = r0; j < r1; ++j) {
      const int q = j + s;
      if (q < r1) p_lds<CB, ARG>(col + (uint32_t)q * pitch, nx, na);
      else if (S == 2 && (q - r1)) p_copy<CB, ARG>(nx, na, ov1, oa1);
      else p_copy<CB, ARG>(nx, na, ov0, oa0);
      float o[CB];
      int oa[CB];
      p_copy<CB, ARG>(o, oa, p0, a0);
      p_max<CB, ARG>(o, oa, nx, na);
      p_sts<CB, ARG>(col + (uint32_t)j * pitch, o, oa);
      if (s == 2) p_copy<CB, ARG>(p0, a0, p1, a1); else p_copy<CB, ARG>(p0, a0, nx, na);
      p_copy<CB, ARG>(p1, a1, nx, na);
    }
  }
  __syncthreads();
}

// One bucket slice = `total` lane slots (kSlots = 64 per proposal, 15 of them idle) whose proposals all need
// at most CH x CW blocks per bin.  A lane handles TWO slots per pass (f and f + stride: independent LDS
// chains hide each other's latency); per slot it reads one descriptor word (coalesced) and its
// proposal's (id, scale) pair, both fetched one pass ahead, issues up to CH*CW LDS (a block is skipped
// where it would repeat the previous one: the bin is not larger than the blocks before it) and stores
// CB scalars (ARG: and CB indices).  Passes are aligned to proposals (f is a multiple of 32, a proposal of 64 slots).
// shared-space loads from a 32-bit address in the hot loop: NOT volatile (ptxas schedules and predicates them freely; the
// wrappers above are volatile), and no generic -> shared conversion of the plane pointer per bin (it cost four uniform
// instructions per lane slot).  The plane does not change between the barriers that bracket a bin phase.
template <class V> __device__ __forceinline__ V lds_plane(uint32_t addr);
template <> __device__ __forceinline__ float4 lds_plane<float4>(uint32_t addr) {
  float4 v;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
template <> __device__ __forceinline__ float2 lds_plane<float2>(uint32_t addr) {
  float2 v;
  asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
  return v;
}
template <int CB, bool ARG> struct PV;
template <> struct PV<4, false> { using T = float4; };
template <> struct PV<2, false> { using T = float2; };
template <> struct PV<2, true> { using T = float4; };
__device__ __forceinline__ void pv_get(float* m, int*, const float4& v, const PV<4, false>&) { m[0] = v.x; m[1] = v.y; m[2] = v.z; m[3] = v.w; }
__device__ __forceinline__ void pv_get(float* m, int*, const float2& v, const PV<2, false>&) { m[0] = v.x; m[1] = v.y; }
__device__ __forceinline__ void pv_get(float* m, int* a, const float4& v, const PV<2, true>&) {
  m[0] = v.x; m[1] = v.y; a[0] = __float_as_int(v.z); a[1] = __float_as_int(v.w);
}

template <int CB, bool ARG, int SLOTS, int CH, int CW, bool FULL>
__device__ __forceinline__ void pyr_run(uint32_t plane, uint32_t pitch, uint32_t khp, uint32_t kwb,
                                        const uint32_t* __restrict__ dsc, const uint2* __restrict__ pin,
                                        int total, float* __restrict__ outc, int32_t* __restrict__ argc, uint32_t c49,
                                        int nc, int flat0, int stride) {
  using V = typename PV<CB, ARG>::T;
  constexpr uint32_t CS = CellT<CB, ARG>::CS;
  auto one = [&](const uint32_t d, const uint2 pi) {
    if (d & kDescIdle) return;                 // slots 49..63 of a proposal
    const uint32_t a0 = (d & 0xffffu) * CS;
    const uint32_t lhp = ((d >> 16) & 15u) * pitch, lwb = ((d >> 20) & 15u) * CS;
    // the plane holds no NaN / -inf (pyr_stage clamps at -FLT_MAX), so the first block seeds the maximum
    float m[CB];
    int ma[CB];
    pv_get(m, ma, lds_plane<V>(plane + a0), PV<CB, ARG>());
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const uint32_t ro = i == 0 ? 0u : min((uint32_t)i * khp, lhp);
      const bool ni = i == 0 || (uint32_t)(i - 1) * khp < lhp;
#pragma unroll
      for (int j = 0; j < CW; ++j) {
        if (i == 0 && j == 0) continue;
        const uint32_t co = j == 0 ? 0u : min((uint32_t)j * kwb, lwb);
        const bool nj = j == 0 || (uint32_t)(j - 1) * kwb < lwb;
        if (ni && nj) {
          float f[CB];
          int fa[CB];
          pv_get(f, fa, lds_plane<V>(plane + a0 + ro + co), PV<CB, ARG>());
          p_max<CB, ARG>(m, ma, f, fa);
        }
      }
    }
    const float sc = __uint_as_float(pi.y);   // 1.0f without a row scale: exact
    const size_t oo = (size_t)(pi.x & 0x3ffffffu) * c49 + ((d >> 24) & 63u);
#pragma unroll
    for (int k = 0; k < CB; ++k)
      if (FULL || k < nc) {
        __stcs(outc + oo + k * 49, __fmul_rn(m[k], sc));
        if (ARG) __stcs(argc + oo + k * 49, ma[k]);
      }
  };
  // 64-slot stream: a warp takes BOTH halves of a proposal (lane l: slots l and 32 + l, i.e. bins l and 32 + l), so every
  // warp -- and every SM sub-partition's scheduler -- carries the same 32 + 17 bins per pass.  (With the halves dealt to
  // alternating warps, the even warps = two of the four schedulers had twice the issue load of the others.)
  const int second = SLOTS == 64 ? 32 : stride;
  const int step = 2 * stride;
  int f = SLOTS == 64 ? 2 * (flat0 & ~31) + (flat0 & 31) : flat0;
  uint32_t d0 = kDescIdle, d1 = kDescIdle;
  uint2 p0 = make_uint2(0u, 0u), p1 = p0;
  if (f < total) { d0 = __ldg(dsc + f); p0 = __ldg(pin + (uint32_t)f / (uint32_t)SLOTS); }
  if (f + second < total) { d1 = __ldg(dsc + f + second); p1 = __ldg(pin + (uint32_t)(f + second) / (uint32_t)SLOTS); }
  while (f < total) {
    const uint32_t c0 = d0, c1 = d1;
    const uint2 q0 = p0, q1 = p1;
    const int g = f + step;
    d0 = kDescIdle;
    d1 = kDescIdle;
    if (g < total) { d0 = __ldg(dsc + g); p0 = __ldg(pin + (uint32_t)g / (uint32_t)SLOTS); }
    if (g + second < total) { d1 = __ldg(dsc + g + second); p1 = __ldg(pin + (uint32_t)(g + second) / (uint32_t)SLOTS); }
    one(c0, q0);
    one(c1, q1);
    f = g;
  }
}

template <int CB, bool ARG>
__global__ void __launch_bounds__(1024, 1) roi_pool7_pyr_kernel(const PyrParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int BINS = 49;
  constexpr uint32_t CS = CellT<CB, ARG>::CS;
  constexpr int SLOTS = CS == 16 ? kSlots : 49;   // lane slots per proposal in the descriptor stream
  const int H = p.H, W = p.W, HW = H * W;
  const int WP = W + kPad, ncell = (H + kPad) * WP, ntot = (H + kPad + kTailRows) * WP;
  const int bid = blockIdx.x;
  const int cg = bid % p.CG;
  const int sidx = (bid / p.CG) % p.S;
  const int n = bid / (p.CG * p.S);
  const int c0 = cg * CB;
  const int nc = min(CB, p.C - c0);
  const int gstart = __ldg(p.img_start + n);
  const int cnt = __ldg(p.img_start + n + 1) - gstart;
  if (cnt <= 0) return;
  uint32_t sbase;
  {
    unsigned long long s64;
    asm volatile("cvta.to.shared.u64 %0, %1;" : "=l"(s64) : "l"((unsigned long long)(uintptr_t)smem_raw));
    sbase = (uint32_t)s64;
  }
  {  // identity tail rows and the all-zero cell empty bins point at (never written again)
    float id[CB];
    int ida[CB];
#pragma unroll
    for (int k = 0; k < CB; ++k) { id[k] = -FLT_MAX; ida[k] = -1; }
    for (int i = ncell + (int)threadIdx.x; i < ntot; i += blockDim.x) p_sts<CB, ARG>(sbase + (uint32_t)i * CS, id, ida);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int k = 0; k < CB; ++k) id[k] = 0.f;
      p_sts<CB, ARG>(sbase + (uint32_t)ntot * CS, id, ida);       // empty bin: value 0, argmax -1
    }
  }
  const float* src = p.input + ((int64_t)n * p.C + c0) * HW;
  float* outc = p.output + (size_t)c0 * BINS;
  int32_t* argc = ARG ? p.argmax + (size_t)c0 * BINS : nullptr;
  const uint32_t c49 = (uint32_t)p.C * BINS;
  const uint32_t pitch = (uint32_t)WP * CS;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int stride = nw * 32;
  const int flat0 = wid * 32 + lane;
  __shared__ int boff[kBuckets + 1];   // this image's bucket boundaries
  for (int i = threadIdx.x; i <= kBuckets; i += blockDim.x) boff[i] = __ldg(p.bucket_off + n * (kBuckets + 1) + i);
  __syncthreads();

  for (int phase = 0; phase < kPhases; ++phase) {
    const int plo = boff[phase * 16];
    const int rem = boff[(chain_end(phase) + 1) * 16] - plo;   // proposals left in this chain
    if (rem > 0 && phase != PH_FALLBACK) {
      __syncthreads();                                                 // everyone is done reading the old plane
      switch (phase) {
        case PH_11: pyr_stage<CB, 0, ARG>(sbase, src, nc, H, W, p.floor_v); __syncthreads(); break;
        case PH_21: pyr_double_v<CB, 1, ARG>(sbase, H + kPad, WP); break;
        case PH_22: pyr_double_h<CB, 1, ARG>(sbase, ncell); break;
        case PH_42: pyr_double_v<CB, 2, ARG>(sbase, H + kPad, WP); break;
        case PH_44: pyr_double_h<CB, 2, ARG>(sbase, ncell); break;
        case PH_12: pyr_stage<CB, 1, ARG>(sbase, src, nc, H, W, p.floor_v); __syncthreads(); break;
        case PH_14: pyr_double_h<CB, 2, ARG>(sbase, ncell); break;
        case PH_24: pyr_double_v<CB, 1, ARG>(sbase, H + kPad, WP); break;
        default:    pyr_stage<CB, 2, ARG>(sbase, src, nc, H, W, p.floor_v); __syncthreads();
                    pyr_double_v<CB, 2, ARG>(sbase, H + kPad, WP); break;   // PH_41
      }
    }
    if (phase != PH_FALLBACK) {
      const uint32_t khp = (uint32_t)phase_kh(phase) * pitch, kwb = (uint32_t)phase_kw(phase) * CS;
      for (int sub = 5; sub < 16; sub += (sub == 7 ? 6 : 2)) {   // block counts (ch, cw) in {2,4}^2: (ch-1) + 4 (cw-1)
        const int lo = boff[phase * 16 + sub], hi = boff[phase * 16 + sub + 1];
        if (hi <= lo) continue;
        const int per = (hi - lo + p.S - 1) / p.S;
        const int slo = lo + sidx * per;
        const int shi = min(hi, slo + per);
        if (shi <= slo) continue;
        const int total = (shi - slo) * SLOTS;
        const uint32_t* dsc = p.desc + (size_t)(gstart + slo) * SLOTS;
        const uint2* pin = p.pinfo + gstart + slo;
#define PYR_CASE(CH, CW) \
  case ((CH - 1) + (CW - 1) * 4): \
    if (nc == CB) pyr_run<CB, ARG, SLOTS, CH, CW, true>(sbase, pitch, khp, kwb, dsc, pin, total, outc, argc, c49, nc, flat0, stride); \
    else pyr_run<CB, ARG, SLOTS, CH, CW, false>(sbase, pitch, khp, kwb, dsc, pin, total, outc, argc, c49, nc, flat0, stride); \
    break;
        switch (sub) {
          PYR_CASE(2, 2) PYR_CASE(4, 2) PYR_CASE(2, 4) PYR_CASE(4, 4)
        }
#undef PYR_CASE
      }
    } else {
      // bins needing more than kMaxLoads blocks per axis: direct scan of the (1,1) plane with edges
      // recomputed from the roi (lane slot = output bin)
      const int lo = plo, hi = boff[(phase + 1) * 16];
      if (hi <= lo) continue;
      const int per = (hi - lo + p.S - 1) / p.S;
      const int slo = lo + sidx * per;
      const int shi = min(hi, slo + per);
      if (shi <= slo) continue;
      const int total = (shi - slo) * BINS;
      const uint2* pin = p.pinfo + gstart + slo;
      for (int flat = flat0; flat < total; flat += stride) {
        const int rp = flat / BINS;
        const int bin = flat - rp * BINS;
        const int ph = bin / 7, pw = bin - ph * 7;
        const uint2 pi = __ldg(pin + rp);
        const float* roi = p.rois + (int64_t)(pi.x & 0x3ffffffu) * 5;
        const Axis ah = axis_of(roi[2], roi[4], p.scale), aw = axis_of(roi[1], roi[3], p.scale);
        int hs, he, ws, we;
        bin_edges(ah, ph, H, hs, he);
        bin_edges(aw, pw, W, ws, we);
        const bool empty = he <= hs || we <= ws;
        float m[CB];
        int ma[CB];
#pragma unroll
        for (int k = 0; k < CB; ++k) { m[k] = empty ? 0.f : -FLT_MAX; ma[k] = -1; }
        if (!empty) {
          for (int h = hs; h < he; ++h) {
            uint32_t a = sbase + (uint32_t)((h + kPad) * WP + ws + kPad) * CS;
            for (int w = ws; w < we; ++w, a += CS) {
              float f[CB];
              int fa[CB];
              p_lds<CB, ARG>(a, f, fa);
              p_max<CB, ARG>(m, ma, f, fa);
            }
          }
        }
        const size_t oo = (size_t)(pi.x & 0x3ffffffu) * c49 + bin;
#pragma unroll
        for (int k = 0; k < CB; ++k)
          if (k < nc) {
            __stcs(outc + oo + k * BINS, __fmul_rn(m[k], __uint_as_float(pi.y)));
            if (ARG) __stcs(argc + oo + k * BINS, ma[k]);
          }
      }
    }
  }
}

// shared memory of the padded plane (+ identity tail rows) at `cell` bytes per cell
static size_t pyr_smem(int64_t H, int64_t W, int cell) {
  return ((size_t)(H + kPad + kTailRows) * (size_t)(W + kPad) + 1) * (size_t)cell;   // + the zero cell
}

// channels per CTA the pyramid path would use for this map (0: does not apply); with argmax a cell holds two
// (value, index) pairs
int pool7_pyr_cb(int64_t C, int64_t H, int64_t W, int64_t R, bool with_argmax) {
  if ((H + kPad + kTailRows) * (W + kPad) >= 65535 || R >= (1 << 26) || C * 49 >= (1LL << 32)) return 0;
  if (with_argmax) return (H * W < (1LL << 31) && pyr_smem(H, W, 16) + 1024 <= (size_t)kMaxSmemOptin) ? 2 : 0;
  if (C >= 3 && pyr_smem(H, W, 16) + 1024 <= (size_t)kMaxSmemOptin) return 4;   // + static shared memory
  if (pyr_smem(H, W, 8) + 1024 <= (size_t)kMaxSmemOptin) return 2;
  return 0;
}

static int64_t pyr_split(int64_t units, int64_t proposals_per_image) {
  if (units >= 100) return 1;
  const int64_t fill = ceil_div(kNumSMs, std::max<int64_t>(units, 1));
  return std::max<int64_t>(1, std::min<int64_t>(fill, proposals_per_image / 1500));
}

template <int CB, bool ARG>
static int pyr_launch_main(PyrParams& p, int64_t R, cudaStream_t st) {
  const size_t smem = pyr_smem(p.H, p.W, (int)CellT<CB, ARG>::CS);
  p.CG = (int)ceil_div(p.C, CB);
  // every CTA of an image rebuilds all planes, so an image's proposals are only split over several CTAs
  // when the (image, channel group) units cannot even fill most of one wave (one image of C = 512 gives
  // 128 units: 0.21 ms unsplit against 0.54 ms split three ways on a B200)
  const int64_t units = (int64_t)p.N * p.CG;
  p.S = (int)pyr_split(units, R / std::max<int64_t>(p.N, 1));
  if (units * p.S > 0x7fffffffLL) return WSOVOD_B200_ETOOBIG;
  auto kern = roi_pool7_pyr_kernel<CB, ARG>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  kern<<<(unsigned)(units * p.S), 1024, smem, st>>>(p);
  return after_launch();
}

// ROI max-pool 7x7 through the block-max planes (argmax: optional); `workspace` holds pool7_pyr_workspace() bytes
int pool7_pyr(const float* input, int64_t N, int64_t C, int64_t H, int64_t W, const float* rois, int64_t R,
              float scale, const float* row_scale, float row_scale_bias, float* output, int32_t* argmax, void* workspace,
              cudaStream_t st, float floor_v) {
  const int cb = pool7_pyr_cb(C, H, W, R, argmax != nullptr);
  if (!cb) return WSOVOD_B200_EINVAL;
  const bool cell16 = cb == 4 || argmax;            // 16-byte cells: 64 lane slots per proposal
  PyrWs w = pyr_carve(workspace, N, R);
  cudaError_t e = cudaMemsetAsync(w.hist, 0, sizeof(int32_t) * (size_t)N * kBuckets * 2, st);   // histogram + cursors
  if (e != cudaSuccess) return (int)e;
  int rc;
  pyr_classify_kernel<<<(unsigned)ceil_div(R, 128), 128, 0, st>>>(rois, R, (int)N, (int)H, (int)W, scale, w.bidx, w.pkey, w.hist, w.axtab);
  if ((rc = after_launch())) return rc;
  const unsigned parts = (unsigned)std::max<int64_t>(1, std::min<int64_t>(32, ceil_div(R, 2048)));
  pyr_order_kernel<<<dim3((unsigned)N, parts), 512, 0, st>>>(w.bidx, w.pkey, w.hist, w.cursor, (int)N, R, w.order, w.img_start, w.bucket_off);
  if ((rc = after_launch())) return rc;
  if (cell16 && tune(TUNE_POOL_GROUP))
    pyr_bins_kernel<kSlots, true><<<(unsigned)ceil_div(R * 2, 128), 128, 0, st>>>(R, (int)H, (int)W, w.order, w.pkey, w.axtab, row_scale,
                                                                                   row_scale_bias, w.pinfo, w.desc);
  else if (cell16)
    pyr_bins_kernel<kSlots, false><<<(unsigned)ceil_div(R * 8, 256), 256, 0, st>>>(R, (int)H, (int)W, w.order, w.pkey, w.axtab, row_scale,
                                                                                      row_scale_bias, w.pinfo, w.desc);
  else
    pyr_bins_kernel<49, false><<<(unsigned)ceil_div(R * 8, 256), 256, 0, st>>>(R, (int)H, (int)W, w.order, w.pkey, w.axtab, row_scale,
                                                                                  row_scale_bias, w.pinfo, w.desc);
  if ((rc = after_launch())) return rc;
  PyrParams p;
  p.input = input; p.rois = rois; p.scale = scale; p.floor_v = floor_v;
  p.output = output; p.argmax = argmax; p.img_start = w.img_start; p.bucket_off = w.bucket_off; p.pinfo = w.pinfo; p.desc = w.desc;
  p.N = (int)N; p.C = (int)C; p.H = (int)H; p.W = (int)W; p.CG = 0; p.S = 1;
  if (argmax) return pyr_launch_main<2, true>(p, R, st);
  return cb == 4 ? pyr_launch_main<4, false>(p, R, st) : pyr_launch_main<2, false>(p, R, st);
}

}  // namespace wsovod
