// Host replay of the separable ROIAlign tap tables (wsovod_b200/csrc/align_sep.cuh, shared with the CUDA kernel)
// against the per-sample loop of torchvision's roi_align (roi_align_common.h pre_calc_for_bilinear_interpolate +
// roi_align_kernel.cpp: the algorithm detectron2's ROIAlign reaches, poolers.py:169-182).  Checks, per bin:
//   * the table's cells are exactly the cells the samples touch (adaptive grid: contiguous, nothing extra);
//   * sum_a WY[a] sum_b WX[b] f[a][b] / count == the sample loop's value within 1e-5 (rel + abs);
//   * the lists stay inside axis_cap().
// Build: g++ -O1 -ffp-contract=off.  Prints "ok ..." on success.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <set>
#include <vector>

#include "../../wsovod_b200/csrc/align_sep.cuh"

using namespace wsovod::asep;

struct Geo { float sw, sh, bw, bh; int gh, gw; float count; };

static Geo geometry(const float* roi, float scale, int P, int sampling_ratio, bool aligned) {
  float off = aligned ? 0.5f : 0.f;
  float sw = roi[0] * scale - off, sh = roi[1] * scale - off, ew = roi[2] * scale - off, eh = roi[3] * scale - off;
  float rw = ew - sw, rh = eh - sh;
  if (!aligned) { rw = std::fmax(rw, 1.f); rh = std::fmax(rh, 1.f); }
  Geo g;
  g.sw = sw; g.sh = sh; g.bh = rh / (float)P; g.bw = rw / (float)P;
  g.gh = sampling_ratio > 0 ? sampling_ratio : (int)std::ceil(rh / (float)P);
  g.gw = sampling_ratio > 0 ? sampling_ratio : (int)std::ceil(rw / (float)P);
  g.count = (float)std::max(g.gh * g.gw, 1);
  return g;
}

int main(int argc, char** argv) {
  const int rounds = argc > 1 ? atoi(argv[1]) : 50;
  std::mt19937 rng(7);
  std::uniform_real_distribution<float> U(0.f, 1.f);
  long bins = 0, taps = 0, cells = 0;
  double worst = 0;
  const int P = 7;
  for (int round = 0; round < rounds; ++round) {
    const int H = 1 + (int)(U(rng) * 60), W = 1 + (int)(U(rng) * 90);
    std::vector<float> f((size_t)H * W);
    for (auto& v : f) v = U(rng) < 0.5f ? 0.f : U(rng) * 4.f - (round % 3 == 0 ? 1.f : 0.f);
    const float scale = 0.125f;
    const int capy = axis_cap(H, P), capx = axis_cap(W, P);
    std::vector<float> wy(capy), wx(capx);
    Hdr hy[P], hx[P];
    for (int it = 0; it < 400; ++it) {
      float roi[4];
      const float iw = W / scale, ih = H / scale;
      const int kind = it % 8;
      float x1 = U(rng) * iw, y1 = U(rng) * ih;
      float w = std::exp(U(rng) * std::log(iw)) , h = std::exp(U(rng) * std::log(ih));
      if (kind == 1) { x1 -= iw * 0.6f; y1 -= ih * 0.6f; }            // partly / fully outside
      if (kind == 2) { w = U(rng) * 3.f; h = U(rng) * 3.f; }          // tiny
      if (kind == 3) { w = iw * 1.5f; h = ih * 1.5f; x1 = -U(rng) * iw * 0.4f; y1 = -U(rng) * ih * 0.4f; }
      if (kind == 4) { w = 0.f; }
      if (kind == 5) { x1 = std::floor(x1 / 8) * 8; y1 = std::floor(y1 / 8) * 8; w = std::floor(w / 8) * 8; h = std::floor(h / 8) * 8; }
      roi[0] = x1; roi[1] = y1; roi[2] = x1 + w; roi[3] = y1 + h;
      if (kind == 6) { roi[2] = x1 - 5.f; }                            // malformed
      for (int aligned = 0; aligned < 2; ++aligned)
        for (int sr : {0, 2}) {
          const Geo g = geometry(roi, scale, P, sr, aligned);
          for (auto& v : wy) v = NAN;
          for (auto& v : wx) v = NAN;
          axis_tables(g.sh, g.bh, g.gh, P, H, hy, wy.data(), capy);
          axis_tables(g.sw, g.bw, g.gw, P, W, hx, wx.data(), capx);
          for (int ph = 0; ph < P; ++ph)
            for (int pw = 0; pw < P; ++pw) {
              // the sample loop
              float acc = 0.f;
              std::set<int> ty, tx;
              for (int iy = 0; iy < g.gh; ++iy) {
                const float yy = g.sh + (float)ph * g.bh + ((float)iy + .5f) * g.bh / (float)g.gh;
                for (int ix = 0; ix < g.gw; ++ix) {
                  const float xx = g.sw + (float)pw * g.bw + ((float)ix + .5f) * g.bw / (float)g.gw;
                  float y = yy, x = xx;
                  if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) continue;
                  if (y <= 0) y = 0;
                  if (x <= 0) x = 0;
                  int yl = (int)y, xl = (int)x, yh, xh;
                  if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else yh = yl + 1;
                  if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
                  const float ly = y - (float)yl, lx = x - (float)xl, hy_ = 1.f - ly, hx_ = 1.f - lx;
                  acc += hy_ * hx_ * f[yl * W + xl] + hy_ * lx * f[yl * W + xh] + ly * hx_ * f[yh * W + xl] + ly * lx * f[yh * W + xh];
                  ty.insert(yl); ty.insert(yh); tx.insert(xl); tx.insert(xh);
                  ++taps;
                }
              }
              const float ref = acc / g.count;
              // the tables
              const int y0 = hy[ph].cell_n & 0xffff, ny = hy[ph].cell_n >> 16;
              const int x0 = hx[pw].cell_n & 0xffff, nx = hx[pw].cell_n >> 16;
              if ((int)hy[ph].off + ny > capy || (int)hx[pw].off + nx > capx || (hy[ph].off & 3) || (hx[pw].off & 3)) {
                printf("FAIL capacity/alignment H=%d W=%d\n", H, W);
                return 1;
              }
              if (y0 + ny > H || x0 + nx > W) { printf("FAIL range H=%d W=%d y0=%d ny=%d x0=%d nx=%d\n", H, W, y0, ny, x0, nx); return 1; }
              float sum = 0.f;
              for (int xc = 0; xc < nx; xc += 4)        // the kernel's association: column chunks of four, rows inside
                for (int a = 0; a < ny; ++a) {
                  float rs = 0.f;
                  for (int b = xc; b < nx && b < xc + 4; ++b) rs = std::fmaf(wx[hx[pw].off + b], f[(y0 + a) * W + x0 + b], rs);
                  sum = std::fmaf(wy[hy[ph].off + a], rs, sum);
                }
              cells += (long)nx * ny;
              const float got = sum * (1.f / g.count);   // the kernel multiplies by the prologue's reciprocal
              const double err = std::fabs((double)got - ref) / (1e-5 + 1e-5 * std::fabs((double)ref));
              if (err > worst) worst = err;
              if (!(err <= 1.0)) {
                printf("FAIL value H=%d W=%d roi=(%g %g %g %g) aligned=%d sr=%d bin=(%d,%d): %g vs %g\n", H, W, roi[0], roi[1], roi[2], roi[3], aligned, sr, ph, pw, got, ref);
                return 1;
              }
              const bool any = !ty.empty() && !tx.empty();
              if (sr == 0 && any) {   // adaptive grid: the lists are exactly the touched cells
                if (ny != (int)ty.size() || nx != (int)tx.size() || *ty.begin() != y0 || *tx.begin() != x0) {
                  printf("FAIL footprint H=%d W=%d aligned=%d bin=(%d,%d): rows %d@%d vs %zu@%d, cols %d@%d vs %zu@%d\n", H, W, aligned, ph, pw,
                         ny, y0, ty.size(), *ty.begin(), nx, x0, tx.size(), *tx.begin());
                  return 1;
                }
              }
              for (int b = nx; b < ((nx + 3) & ~3); ++b)
                if (wx[hx[pw].off + b] != 0.f) { printf("FAIL padding\n"); return 1; }
              ++bins;
            }
        }
    }
  }
  printf("ok bins=%ld samples=%ld footprint_cells=%ld worst_err_over_tol=%.3f\n", bins, taps, cells, worst);
  return 0;
}
