// pool_pyr.cuh -- geometry of the block-max ("pyramid") ROI max-pool fast path, shared by the device
// kernels (roi_pool_pyr.cu) and the host-side emulation that checks it (tests/host/pyr_emul.cu).
//
// Idea.  max() is idempotent, so a bin [s, e) can be covered by OVERLAPPING blocks of a fixed size k:
// positions s, s+k, ..., and a last one at e-k.  If shared memory holds D_k[h][w] = max of the
// kh x kw block whose top-left cell is (h, w), a bin costs ceil(nh/kh) * ceil(nw/kw) loads instead of
// nh * nw.  All 49 bins of a proposal have (unclipped) sizes {a, a+1} per axis, so one (kh, kw) in
// {1,2,4}^2 per PROPOSAL gives <= 2 x 2 loads for almost every bin (mean 2.9 loads/bin instead of
// 10.5 on SAM-like proposals).  Only one D plane is resident per CTA: proposals are grouped by
// (kh, kw) and the CTA rebuilds D in place between groups ("phases") by doubling steps.
//
// Clipped bins (n < k, possible only where the bin was clamped at a map border) are handled by the
// plane itself: blocks are truncated at the map's far border, and the plane carries k_max-1 = 3 pad
// rows/columns of identity on the near border, so [e-k, e) with e-k < 0 is a legal position.
//
// Bin edges are the exact fp32 sequence of the reference (ROILoopPool_cpu.cpp:29-51).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define WS_HD __host__ __device__ __forceinline__
#else
#define WS_HD inline
#endif

namespace wsovod {
namespace pyr {

constexpr int kPad = 3;        // identity rows above / columns left of the map (k_max - 1)
constexpr int kTailRows = 2;   // identity rows below the plane (largest vertical doubling stride)
constexpr int kPhases = 10;
constexpr int kMaxLoads = 4;   // per axis; more -> fallback phase (direct scan of the k=1 plane)
constexpr int kBuckets = kPhases * 16;   // sort key = phase * 16 + (ch-1) * 4 + (cw-1)

// Phase order = the order a CTA walks the (kh, kw) groups; consecutive phases of a chain differ by one
// in-place doubling step:   chain A: global -> (1,1) -> (2,1) -> (2,2) -> (4,2) -> (4,4)
//                           chain B: global -> (1,2) -> (1,4) -> (2,4)
//                           chain C: global -> (2,1) -> (4,1)
// phase 1 ("fallback") reuses the (1,1) plane for proposals whose bins need more than kMaxLoads blocks.
enum { PH_11 = 0, PH_FALLBACK = 1, PH_21 = 2, PH_22 = 3, PH_42 = 4, PH_44 = 5, PH_12 = 6, PH_14 = 7, PH_24 = 8, PH_41 = 9 };

WS_HD int phase_of(int kh, int kw) {
  // kh, kw in {1, 2, 4}
  const int ih = kh >> 1, iw = kw >> 1;   // 0, 1, 2
  const int t = ih * 3 + iw;
  // (1,1) (1,2) (1,4) (2,1) (2,2) (2,4) (4,1) (4,2) (4,4)
  return t == 0 ? PH_11 : t == 1 ? PH_12 : t == 2 ? PH_14 : t == 3 ? PH_21 : t == 4 ? PH_22 : t == 5 ? PH_24
       : t == 6 ? PH_41 : t == 7 ? PH_42 : PH_44;
}
WS_HD int phase_kh(int phase) {
  return (phase == PH_11 || phase == PH_FALLBACK || phase == PH_12 || phase == PH_14) ? 1
       : (phase == PH_21 || phase == PH_22 || phase == PH_24) ? 2 : 4;
}
WS_HD int phase_kw(int phase) {
  return (phase == PH_11 || phase == PH_FALLBACK || phase == PH_21 || phase == PH_41) ? 1
       : (phase == PH_12 || phase == PH_22 || phase == PH_42) ? 2 : 4;
}
// last phase of the chain `phase` belongs to (builds are skipped when nothing is left in the chain)
WS_HD int chain_end(int phase) { return phase <= PH_44 ? PH_44 : phase <= PH_24 ? PH_24 : PH_41; }

// saturating float->int like the reference's `int x = round(float)` on sane inputs; huge values clamp
WS_HD int round_i(float v) {
  v = roundf(v);
  v = fminf(fmaxf(v, -1.0e6f), 1.0e6f);
  return (int)v;
}

#if defined(__CUDA_ARCH__)
#define WS_FMUL(a, b) __fmul_rn((a), (b))
#define WS_FDIV(a, b) __fdiv_rn((a), (b))
#else
#define WS_FMUL(a, b) ((a) * (b))
#define WS_FDIV(a, b) ((a) / (b))
#endif

// one axis of one proposal: start cell and bin size as ROILoopPool_cpu.cpp:29-38
struct Axis {
  int rs;      // rounded start
  float bin;   // bin size (cells)
};
WS_HD Axis axis_of(float lo, float hi, float scale) {
  Axis a;
  a.rs = round_i(WS_FMUL(lo, scale));
  const int re = round_i(WS_FMUL(hi, scale));
  int n = re - a.rs + 1;
  n = n > 1 ? n : 1;
  a.bin = WS_FDIV((float)n, 7.f);
  return a;
}
// clamped edges of bin p (ROILoopPool_cpu.cpp:42-51)
WS_HD void bin_edges(const Axis& a, int p, int L, int& s, int& e) {
  s = (int)floorf(WS_FMUL((float)p, a.bin)) + a.rs;
  e = (int)ceilf(WS_FMUL((float)(p + 1), a.bin)) + a.rs;
  s = s < 0 ? 0 : (s > L ? L : s);
  e = e < 0 ? 0 : (e > L ? L : e);
}

// block size k in {1,2,4} and loads-per-bin c of one axis: the largest k such that every non-empty bin
// either holds k cells or touches a map border (where the padded / truncated plane covers the rest).
WS_HD void axis_class(const Axis& a, int L, int& k, int& c) {
  bool ok4 = true, ok2 = true;
  int nmax = 0;
#pragma unroll
  for (int p = 0; p < 7; ++p) {
    int s, e;
    bin_edges(a, p, L, s, e);
    const int n = e - s;
    if (n <= 0) continue;
    const bool border = s == 0 || e == L;
    ok4 = ok4 && (n >= 4 || border);
    ok2 = ok2 && (n >= 2 || border);
    nmax = n > nmax ? n : nmax;
  }
  k = ok4 ? 4 : ok2 ? 2 : 1;
  c = (nmax + k - 1) / k;
  c = c < 1 ? 1 : c;
}

// first block position (may be negative: pad rows) and distance to the last block of a non-empty bin
WS_HD void bin_blocks(int s, int e, int k, int L, int& pos0, int& last) {
  const int n = e - s;
  if (n >= k) { pos0 = s; last = n - k; }
  else if (s == 0) { pos0 = e - k; last = 0; }   // near-border clip: block [e-k, e) n [0, L)
  else { pos0 = s; last = 0; }                   // far-border clip: block [s, s+k) truncated at L
}

// per-proposal key (classification result)
//   bits 0-3 phase | 4-5 ch-1 | 6-7 cw-1 | 8 lanes walk the bins column-major
// Bit 8 (column-major lane order: the 8 lanes of a shared-memory phase then differ in their row instead of in
// a column stride that repeats bank groups; simulated 23.9 -> 21.9 LDS wavefronts per 32 bins) is
// supported by the descriptor stream but no longer set: lanes that walk a run with stride 7 scatter every
// warp store over the whole 196-byte run, and on the B200 the extra partial-sector writes cost more than
// the bank conflicts saved (c2: 1.314 ms with the heuristic, 1.303 ms without).
constexpr uint32_t kKeyTransposed = 1u << 8;
WS_HD uint32_t proposal_key(float x1, float y1, float x2, float y2, float scale, int H, int W) {
  const Axis ah = axis_of(y1, y2, scale), aw = axis_of(x1, x2, scale);
  int kh, ch, kw, cw;
  axis_class(ah, H, kh, ch);
  axis_class(aw, W, kw, cw);
  if (ch > kMaxLoads || cw > kMaxLoads) return (uint32_t)PH_FALLBACK;
  // block counts are rounded up to {2, 4}: four buckets per phase keep the per-bucket lane runs long;
  // a bin that needs fewer blocks skips the surplus ones (duplicate-block predicates in the kernels)
  ch = ch <= 2 ? 2 : 4;
  cw = cw <= 2 ? 2 : 4;
  return (uint32_t)phase_of(kh, kw) | ((uint32_t)(ch - 1) << 4) | ((uint32_t)(cw - 1) << 6);
}
// output bin served by lane slot q (0..48) of a proposal
WS_HD int slot_bin(uint32_t key, int q) { return (key & kKeyTransposed) ? (q % 7) * 7 + q / 7 : q; }
WS_HD int key_phase(uint32_t key) { return (int)(key & 15u); }
WS_HD int key_bucket(uint32_t key) { return (int)((key & 15u) * 16u + ((key >> 4) & 15u)); }

// 32-bit bin descriptor: bits 0-15 cell index of the first block in the padded plane,
//   16-19 rows to the last block, 20-23 columns to the last block, 24-29 output bin, 30 idle slot, 31 empty bin.
// An empty bin points at the all-zero cell behind the plane, so the kernel needs no special case.
constexpr uint32_t kDescEmpty = 0x80000000u;
constexpr uint32_t kDescIdle = 0x40000000u;   // lane slot without a bin (slots 49..63 of a proposal)
constexpr int kSlots = 64;                    // lane slots per proposal in the descriptor stream
WS_HD int zero_cell(int H, int W) { return (H + kPad + kTailRows) * (W + kPad); }
// One axis of a descriptor, computed once per (proposal, bin row / bin column): bits 0-15 the first block's
// offset in the padded plane (rows: (p0 + kPad) * (W + kPad) cells, columns: p0 + kPad), 16-19 distance
// to the last block, 31 empty.  A bin descriptor is the sum of its row and column entries.
WS_HD uint32_t axis_entry(const Axis& a, int p, int L, int k, int cell_stride) {
  int s, e;
  bin_edges(a, p, L, s, e);
  if (e <= s) return kDescEmpty;
  int p0, last;
  bin_blocks(s, e, k, L, p0, last);
  return (uint32_t)((p0 + kPad) * cell_stride) | ((uint32_t)last << 16);
}
WS_HD uint32_t combine_desc(uint32_t row_entry, uint32_t col_entry, int bin, int H, int W) {
  const uint32_t binbits = (uint32_t)bin << 24;
  if ((row_entry | col_entry) & kDescEmpty) return kDescEmpty | binbits | (uint32_t)zero_cell(H, W);
  return ((row_entry & 0xffffu) + (col_entry & 0xffffu)) | (row_entry & 0xf0000u) | ((col_entry & 0xf0000u) << 4) | binbits;
}
WS_HD uint32_t bin_desc(float x1, float y1, float x2, float y2, float scale, int H, int W, int phase, int ph, int pw) {
  const Axis ah = axis_of(y1, y2, scale), aw = axis_of(x1, x2, scale);
  return combine_desc(axis_entry(ah, ph, H, phase_kh(phase), W + kPad), axis_entry(aw, pw, W, phase_kw(phase), 1),
                      ph * 7 + pw, H, W);
}

}  // namespace pyr
}  // namespace wsovod
