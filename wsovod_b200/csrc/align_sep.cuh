// align_sep.cuh -- separable tap tables of bilinear ROIAlign (torchvision roi_align_common.h, reached through
// detectron2's ROIAlign / ROIAlignV2, poolers.py:169-182).  __host__ __device__: tests/host/align_sep_emul.cpp replays
// the tables on the CPU against the per-sample loop.
//
// A bin's output is (1 / count) * sum over its gh x gw samples of the bilinear interpolation at (y_i, x_j).  A sample
// is dropped when y OR x lies outside [-1, extent], so the valid samples are a product set and the bilinear weights are
// products too: the sum equals  sum_a WY[a] * sum_b WX[b] * f[a][b]  with WY[a] = sum over valid sample rows of that
// row's weight on map row a (hy on floor(y), ly on floor(y) + 1), WX likewise.  One axis of one proposal is therefore
// seven short weight lists -- (first cell, length, weights) per bin -- shared by every channel; a bin then costs one
// shared-memory load and one FMA per FOOTPRINT cell and channel instead of four loads and ~12 flops per SAMPLE.
// With the adaptive sample grid (sampling_ratio = 0: g = ceil(bin size)) the sample spacing is <= 1 cell, consecutive
// samples touch adjacent cells and the list is exactly the set of cells the reference touches; with a fixed grid on
// large bins the cells between two samples carry weight 0 (the library keeps those launches on the per-sample kernel).
// Sample coordinates follow the reference's fp32 sequence operation by operation; the tap weights are summed in sample
// order.  What differs from the reference is only the association of the final sum (stated tolerance 1e-5).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ASEP_HD __host__ __device__ __forceinline__
#else
#define ASEP_HD inline
#endif

namespace wsovod {
namespace asep {

#if defined(__CUDA_ARCH__)
ASEP_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
ASEP_HD float fsub(float a, float b) { return __fsub_rn(a, b); }
ASEP_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
ASEP_HD float fdiv(float a, float b) { return __fdiv_rn(a, b); }
#else   // host build: compile with -ffp-contract=off
ASEP_HD float fadd(float a, float b) { return a + b; }
ASEP_HD float fsub(float a, float b) { return a - b; }
ASEP_HD float fmul(float a, float b) { return a * b; }
ASEP_HD float fdiv(float a, float b) { return a / b; }
#endif

struct Hdr { uint32_t cell_n; uint32_t off; };   // first cell | length << 16, offset of the list in the axis' weights

ASEP_HD int round4(int v) { return (v + 3) & ~3; }
// capacity (floats) of one axis' weight lists: every bin's list is padded to a multiple of four
ASEP_HD int axis_cap(int extent, int P) { return round4(extent + 6 * P + 4); }

// One axis of one proposal.  start: roi start on this axis (already scaled and offset), bs: bin size, g: samples per
// bin, P: bins, L: map extent.  Writes P headers and at most `cap` weights (lists 16-byte aligned, zero padded).
ASEP_HD void axis_tables(float start, float bs, int g, int P, int L, Hdr* hdr, float* wts, int cap) {
  int off = 0;
  for (int b = 0; b < P; ++b) {
    int c0 = 0, n = 0, cur = -1;
    float wa = 0.f, wb = 0.f;     // weights collected for cells cur and cur + 1
    bool has_b = false;           // some sample touched cell cur + 1
    const float base = fadd(start, fmul((float)b, bs));
    const int room = cap - off;
    const bool rev = bs < 0.f;    // malformed box with aligned != 0: walk the samples backwards so that cells only go up
    for (int j = 0; j < g; ++j) {
      const int i = rev ? g - 1 - j : j;
      float y = fadd(base, fdiv(fmul((float)i + .5f, bs), (float)g));
      if (y < -1.0f || y > (float)L) continue;
      if (y <= 0) y = 0;
      int yl = (int)y;
      bool edge = false;
      if (yl >= L - 1) { yl = L - 1; y = (float)yl; edge = true; }
      const float ly = fsub(y, (float)yl), hy = fsub(1.f, ly);
      if (cur < 0) {
        cur = yl; c0 = yl;
      } else if (yl != cur) {
        // flush cells [cur, yl)
        if (n < room) wts[off + n] = wa;
        ++n;
        if (yl == cur + 1) {
          wa = wb;
        } else {
          if (n < room) wts[off + n] = wb;
          ++n;
          for (int c = cur + 2; c < yl; ++c) { if (n < room) wts[off + n] = 0.f; ++n; }
          wa = 0.f;
        }
        wb = 0.f; has_b = false; cur = yl;
      }
      wa = fadd(wa, hy);
      if (edge) wa = fadd(wa, ly);          // yh == yl on the far border (ly is 0 there)
      else { wb = fadd(wb, ly); has_b = true; }
    }
    if (cur >= 0) {
      if (n < room) wts[off + n] = wa;
      ++n;
      if (has_b) { if (n < room) wts[off + n] = wb; ++n; }
    }
    if (n > room) n = room;       // cannot happen for cap = axis_cap(L, P); keeps the writes inside the table regardless
    if (n > 65532) n = 65532;
    const int padded = round4(n);   // <= room: off, cap and so room are multiples of four
    for (int c = n; c < padded; ++c) wts[off + c] = 0.f;
    hdr[b].cell_n = (uint32_t)c0 | ((uint32_t)n << 16);
    hdr[b].off = (uint32_t)off;
    off += padded;
  }
}

}  // namespace asep
}  // namespace wsovod
