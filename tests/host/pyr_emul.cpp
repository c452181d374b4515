// Host emulation of the block-max pooling path (wsovod_b200/csrc/pool_pyr.cuh + the plane recipes of
// roi_pool_pyr.cu) against a brute-force scan with the reference's bin edges
// (ROILoopPool_cpu.cpp:29-79).  Test infrastructure: g++ -O1 -ffp-contract=off pyr_emul.cpp && ./a.out
#include <float.h>
#include <stdio.h>
#include <stdlib.h>

#include <random>
#include <vector>

#include "../../wsovod_b200/csrc/pool_pyr.cuh"

using namespace wsovod::pyr;

struct Plane {
  int H, W, WP, ncell;
  std::vector<float> d;
  void stage(const std::vector<float>& src, int mode) {
    for (int idx = 0; idx < ncell; ++idx) {
      const int hh = idx / WP, ww = idx % WP, h = hh - kPad, w = ww - kPad;
      float f = -FLT_MAX;
      if (h >= 0 && w >= 0) f = fmaxf(f, src[h * W + w]);
      if (mode == 1 && h >= 0 && w + 1 >= 0 && w + 1 < W) f = fmaxf(f, src[h * W + w + 1]);
      if (mode == 2 && w >= 0 && h + 1 >= 0 && h + 1 < H) f = fmaxf(f, src[(h + 1) * W + w]);
      d[idx] = f;
    }
  }
  void dbl(int stride) {
    for (int idx = 0; idx < ncell; ++idx) d[idx] = fmaxf(d[idx], d[idx + stride]);   // forward reads only
  }
};

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 300;
  std::mt19937 rng(1234);
  std::uniform_real_distribution<float> U(0.f, 1.f);
  long bins = 0, fallback = 0, loads = 0, cells = 0;
  for (int it = 0; it < iters; ++it) {
    const int H = 1 + (int)(U(rng) * (it % 3 == 0 ? 12 : 90)), W = 1 + (int)(U(rng) * (it % 3 == 1 ? 12 : 130));
    const float scale = it % 5 == 0 ? 0.0625f : 0.125f;
    std::vector<float> src((size_t)H * W);
    for (auto& v : src) v = U(rng) < 0.3f ? 0.f : (U(rng) - 0.3f) * 8.f;
    if (it % 7 == 0) src[(size_t)(U(rng) * H * W) % src.size()] = NAN;
    if (it % 11 == 0) src[(size_t)(U(rng) * H * W) % src.size()] = -INFINITY;
    Plane P;
    P.H = H; P.W = W; P.WP = W + kPad; P.ncell = (H + kPad) * P.WP;
    P.d.assign((size_t)(H + kPad + kTailRows) * P.WP + 1, -FLT_MAX);
    P.d[zero_cell(H, W)] = 0.f;   // empty bins point here
    const int R = 120;
    const float iw = W / scale, ih = H / scale;
    std::vector<float> rois((size_t)R * 4);
    for (int r = 0; r < R; ++r) {
      float x1 = (U(rng) * 1.3f - 0.15f) * iw, y1 = (U(rng) * 1.3f - 0.15f) * ih;
      float w = expf(U(rng) * logf(iw * 1.2f + 2.f)), h = expf(U(rng) * logf(ih * 1.2f + 2.f));
      if (r % 2 == 0) { w = U(rng) * iw * 0.9f; h = U(rng) * ih * 0.9f; }
      if (r % 9 == 0) w = 0.f;
      if (r % 13 == 0) { x1 = floorf(x1); y1 = floorf(y1); w = floorf(w); h = floorf(h); }
      rois[r * 4 + 0] = x1; rois[r * 4 + 1] = y1; rois[r * 4 + 2] = x1 + w; rois[r * 4 + 3] = y1 + h;
    }
    for (int phase = 0; phase < kPhases; ++phase) {
      switch (phase) {
        case PH_11: P.stage(src, 0); break;
        case PH_FALLBACK: break;
        case PH_21: P.dbl(P.WP); break;
        case PH_22: P.dbl(1); break;
        case PH_42: P.dbl(2 * P.WP); break;
        case PH_44: P.dbl(2); break;
        case PH_12: P.stage(src, 1); break;
        case PH_14: P.dbl(2); break;
        case PH_24: P.dbl(P.WP); break;
        default: P.stage(src, 2); P.dbl(2 * P.WP); break;
      }
      const int kh = phase_kh(phase), kw = phase_kw(phase);
      for (int r = 0; r < R; ++r) {
        const float x1 = rois[r * 4], y1 = rois[r * 4 + 1], x2 = rois[r * 4 + 2], y2 = rois[r * 4 + 3];
        const uint32_t key = proposal_key(x1, y1, x2, y2, scale, H, W);
        if (key_phase(key) != phase) continue;
        if (phase != PH_FALLBACK && phase_of(kh, kw) != phase) { printf("phase table broken\n"); return 1; }
        const int ch = (int)((key >> 4) & 3) + 1, cw = (int)((key >> 6) & 3) + 1;
        const Axis ah = axis_of(y1, y2, scale), aw = axis_of(x1, x2, scale);
        for (int ph = 0; ph < 7; ++ph)
          for (int pw = 0; pw < 7; ++pw) {
            int hs, he, ws, we;
            bin_edges(ah, ph, H, hs, he);
            bin_edges(aw, pw, W, ws, we);
            const bool empty = he <= hs || we <= ws;
            float ref = empty ? 0.f : -FLT_MAX;
            for (int h = hs; h < he && !empty; ++h)
              for (int w = ws; w < we; ++w)
                if (src[h * W + w] > ref) ref = src[h * W + w];
            ++bins;
            if (!empty) cells += (he - hs) * (we - ws);
            if (phase == PH_FALLBACK) { ++fallback; continue; }
            const uint32_t d = bin_desc(x1, y1, x2, y2, scale, H, W, phase, ph, pw);
            float got;
            if (((d >> 24) & 63) != (uint32_t)(ph * 7 + pw)) { printf("bin bits broken\n"); return 1; }
            {
              got = -FLT_MAX;
              const int cell = d & 0xffff, lh = (d >> 16) & 15, lw = (d >> 20) & 15;
              for (int i = 0; i < ch; ++i)
                for (int j = 0; j < cw; ++j) {
                  const int ro = i * kh < lh ? i * kh : lh, co = j * kw < lw ? j * kw : lw;
                  const bool need = (i == 0 || (i - 1) * kh < lh) && (j == 0 || (j - 1) * kw < lw);
                  if (!need) continue;   // the kernel skips blocks that repeat the previous one
                  got = fmaxf(got, P.d[cell + ro * P.WP + co]);
                  ++loads;
                }
            }
            if (!(got == ref) || (d >> 31) != (uint32_t)empty) {
              printf("MISMATCH it=%d H=%d W=%d r=%d phase=%d ch=%d cw=%d bin=(%d,%d) edges h[%d,%d) w[%d,%d) got=%g ref=%g\n",
                     it, H, W, r, phase, ch, cw, ph, pw, hs, he, ws, we, got, ref);
              return 1;
            }
          }
      }
    }
  }
  printf("ok bins=%ld fallback=%ld mean_loads=%.3f mean_cells=%.3f\n", bins, fallback,
         (double)loads / (double)(bins - fallback), (double)cells / (double)bins);
  return 0;
}
