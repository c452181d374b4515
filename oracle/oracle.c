/*
 * oracle.c -- TEST INFRASTRUCTURE ONLY.  A plain-C CPU restatement of the reference algorithms on
 * WSOVOD's region-scoring path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this; the product (wsovod_b200/) never does.
 *
 * Every function cites the reference lines it follows (paths relative to the WSOVOD tree, or to the
 * installed torchvision 0.26 / the detectron2 call sites for arithmetic that lives in those
 * un-vendored dependencies -- see DESIGN.md "Oracle").  Pinned by tests/test_oracle_*.py against
 * (a) torchvision's compiled CPU ops, (b) the reference's own ROILoopPool_cpu.cpp compiled into
 * oracle/_ref, (c) goldens produced by importing the reference Python verbatim (tests/golden/).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (see oracle/Makefile).  Contraction
 * is off on purpose: every fp32 operation is individually rounded, like ATen's eager CPU ops.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline float fminf_(float a, float b) { return a < b ? a : b; }
static inline float fmaxf_(float a, float b) { return a > b ? a : b; }

/* ------------------------------------------------------------------------------------------------
 * ROIPool forward -- wsovod/layers/ROILoopPool/ROILoopPool_cpu.cpp:14-80 (== torchvision roi_pool).
 * ---------------------------------------------------------------------------------------------- */
ORC_API void orc_roi_pool_fwd(const float* input, int64_t N, int64_t C, int64_t H, int64_t W,
                              const float* rois, int64_t R, float scale, int PH, int PW,
                              float* out, int32_t* argmax) {
  (void)N;
  for (int64_t n = 0; n < R; ++n) {
    const float* roi = rois + n * 5;
    int b = (int)roi[0];
    int rsw = (int)roundf(roi[1] * scale);           /* :29-32 */
    int rsh = (int)roundf(roi[2] * scale);
    int rew = (int)roundf(roi[3] * scale);
    int reh = (int)roundf(roi[4] * scale);
    int rw = imax(rew - rsw + 1, 1);                 /* :35-36 */
    int rh = imax(reh - rsh + 1, 1);
    float bh = (float)rh / (float)PH;                /* :37-38 */
    float bw = (float)rw / (float)PW;
    for (int ph = 0; ph < PH; ++ph)
      for (int pw = 0; pw < PW; ++pw) {
        int hs = (int)floorf((float)ph * bh);        /* :42-45 */
        int ws = (int)floorf((float)pw * bw);
        int he = (int)ceilf((float)(ph + 1) * bh);
        int we = (int)ceilf((float)(pw + 1) * bw);
        hs = imin(imax(hs + rsh, 0), (int)H);        /* :48-51 */
        he = imin(imax(he + rsh, 0), (int)H);
        ws = imin(imax(ws + rsw, 0), (int)W);
        we = imin(imax(we + rsw, 0), (int)W);
        int empty = (he <= hs) || (we <= ws);
        for (int64_t c = 0; c < C; ++c) {
          float m = empty ? 0.f : -FLT_MAX;          /* :56 */
          int mi = -1;
          const float* p = input + ((int64_t)b * C + c) * H * W;
          for (int h = hs; h < he; ++h)
            for (int w = ws; w < we; ++w) {
              int idx = h * (int)W + w;
              if (p[idx] > m) { m = p[idx]; mi = idx; }   /* strict >, :63-70 */
            }
          int64_t o = ((n * C + c) * PH + ph) * PW + pw;
          out[o] = m;
          if (argmax) argmax[o] = mi;
        }
      }
  }
}

/* ROIPool backward -- ROILoopPool_cpu.cpp:82-123.  rows_per_roi = 1 for roi_pool; for the 3-way
 * op (ROILoopPool_cuda.cu:206-248) grad rows n map to roi n % R and `rows` = 3R. */
ORC_API void orc_roi_pool_bwd(const float* grad_out, const float* rois, const int32_t* argmax,
                              int64_t rows, int64_t R, int64_t N, int64_t C, int64_t H, int64_t W,
                              int PH, int PW, float* grad_in) {
  memset(grad_in, 0, sizeof(float) * (size_t)(N * C * H * W));
  for (int64_t n = 0; n < rows; ++n) {
    int b = (int)rois[(n % R) * 5];
    for (int64_t c = 0; c < C; ++c) {
      float* g = grad_in + ((int64_t)b * C + c) * H * W;
      for (int k = 0; k < PH * PW; ++k) {
        int64_t o = (n * C + c) * PH * PW + k;
        if (argmax[o] != -1) g[argmax[o]] += grad_out[o];
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * ROILoopPool (roi | frame | context) forward -- ROILoopPool_cuda.cu:24-203 (there is no CPU 3-way
 * version upstream).  Arithmetic follows the kernel's C++ types: box geometry in float, the clamp
 * bound `T(1.0 * width / spatial_scale)` in double then narrowed (:66-73).
 * ---------------------------------------------------------------------------------------------- */
ORC_API void orc_roi_loop_pool_fwd(const float* input, int64_t N, int64_t C, int64_t H, int64_t W,
                                   const float* rois, int64_t R, float scale, int PH, int PW,
                                   float* out, int32_t* argmax) {
  (void)N;
  const float ratio = 1.8f;                                 /* :309 (double 1.8 -> float param) */
  const int64_t block = R * C * PH * PW;                    /* :139,198 */
  const float xmax = (float)(1.0 * (double)W / (double)scale);
  const float ymax = (float)(1.0 * (double)H / (double)scale);
  for (int64_t n = 0; n < R; ++n) {
    const float* roi = rois + n * 5;
    int b = (int)roi[0];
    float x1 = roi[1], y1 = roi[2], x2 = roi[3], y2 = roi[4];
    float rw_ = x2 - x1, rh_ = y2 - y1;                     /* :40-41 */
    float iw = rw_ / ratio, ih = rh_ / ratio;               /* :43-44 */
    float ow = rw_ * ratio, oh = rh_ * ratio;               /* :46-47 */
    float irw = rw_ - iw, irh = rh_ - ih;                   /* :49-50 */
    float orw = ow - rw_, orh = oh - rh_;                   /* :52-53 */
    float x1i = x1 + irw / 2, y1i = y1 + irh / 2, x2i = x2 - irw / 2, y2i = y2 - irh / 2;
    float x1o = x1 - orw / 2, y1o = y1 - orh / 2, x2o = x2 + orw / 2, y2o = y2 + orh / 2;
    x1i = fminf_(fmaxf_(x1i, 0.f), xmax); y1i = fminf_(fmaxf_(y1i, 0.f), ymax);   /* :66-69 */
    x2i = fminf_(fmaxf_(x2i, 0.f), xmax); y2i = fminf_(fmaxf_(y2i, 0.f), ymax);
    x1o = fminf_(fmaxf_(x1o, 0.f), xmax); y1o = fminf_(fmaxf_(y1o, 0.f), ymax);   /* :71-74 */
    x2o = fminf_(fmaxf_(x2o, 0.f), xmax); y2o = fminf_(fmaxf_(y2o, 0.f), ymax);

    for (int pass = 0; pass < 2; ++pass) {
      /* pass 0: grid of the ROI, exclusion = inner box (:76-142);
         pass 1: grid of the outer box, exclusion = the ROI itself (:144-202) */
      int rsw, rsh, rew, reh, isw, ish, iew, ieh;
      if (pass == 0) {
        rsw = (int)roundf(roi[1] * scale); rsh = (int)roundf(roi[2] * scale);
        rew = (int)roundf(roi[3] * scale); reh = (int)roundf(roi[4] * scale);
        isw = (int)roundf(x1i * scale); ish = (int)roundf(y1i * scale);
        iew = (int)roundf(x2i * scale); ieh = (int)roundf(y2i * scale);
      } else {
        rsw = (int)roundf(x1o * scale); rsh = (int)roundf(y1o * scale);
        rew = (int)roundf(x2o * scale); reh = (int)roundf(y2o * scale);
        isw = (int)roundf(roi[1] * scale); ish = (int)roundf(roi[2] * scale);
        iew = (int)roundf(roi[3] * scale); ieh = (int)roundf(roi[4] * scale);
      }
      int rw = imax(rew - rsw + 1, 1), rh = imax(reh - rsh + 1, 1);
      float bh = (float)rh / (float)PH, bw = (float)rw / (float)PW;
      for (int ph = 0; ph < PH; ++ph)
        for (int pw = 0; pw < PW; ++pw) {
          int hs = (int)floorf((float)ph * bh), ws = (int)floorf((float)pw * bw);
          int he = (int)ceilf((float)(ph + 1) * bh), we = (int)ceilf((float)(pw + 1) * bw);
          hs = imin(imax(hs + rsh, 0), (int)H); he = imin(imax(he + rsh, 0), (int)H);
          ws = imin(imax(ws + rsw, 0), (int)W); we = imin(imax(we + rsw, 0), (int)W);
          for (int64_t c = 0; c < C; ++c) {
            float m = 0.f, mf = 0.f;                      /* "assum all input is >=0", :107-113 */
            int mi = -1, mfi = -1;
            const float* p = input + ((int64_t)b * C + c) * H * W;
            for (int h = hs; h < he; ++h)
              for (int w = ws; w < we; ++w) {
                int idx = h * (int)W + w;
                float v = p[idx];
                if (pass == 0 && v > m) { m = v; mi = idx; }
                int in_h = h > ish && h < ieh, in_w = w > isw && w < iew;   /* :124-128,184-190 */
                if (in_h && in_w) continue;
                if (v > mf) { mf = v; mfi = idx; }
              }
            int64_t o = ((n * C + c) * PH + ph) * PW + pw;
            if (pass == 0) {
              out[o] = m; out[o + block] = mf;
              if (argmax) { argmax[o] = mi; argmax[o + block] = mfi; }
            } else {
              out[o + 2 * block] = mf;
              if (argmax) argmax[o + 2 * block] = mfi;
            }
          }
        }
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * ROIAlign forward -- torchvision 0.26 csrc/ops/cpu/roi_align_kernel.cpp + roi_align_common.h
 * (un-vendored dependency; call site wsovod/modeling/poolers.py:169-182; semantics in SURVEY A.8).
 * ---------------------------------------------------------------------------------------------- */
ORC_API void orc_roi_align_fwd(const float* input, int64_t N, int64_t C, int64_t H, int64_t W,
                               const float* rois, int64_t R, float scale, int PH, int PW,
                               int sampling_ratio, int aligned, float* out) {
  (void)N;
  for (int64_t n = 0; n < R; ++n) {
    const float* roi = rois + n * 5;
    int b = (int)roi[0];
    float off = aligned ? 0.5f : 0.f;
    float sw = roi[1] * scale - off, sh = roi[2] * scale - off;
    float ew = roi[3] * scale - off, eh = roi[4] * scale - off;
    float rw = ew - sw, rh = eh - sh;
    if (!aligned) { rw = fmaxf_(rw, 1.f); rh = fmaxf_(rh, 1.f); }
    float bh = rh / (float)PH, bw = rw / (float)PW;
    int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rh / (float)PH);
    int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rw / (float)PW);
    float count = (float)imax(gh * gw, 1);
    for (int64_t c = 0; c < C; ++c) {
      const float* p = input + ((int64_t)b * C + c) * H * W;
      for (int ph = 0; ph < PH; ++ph)
        for (int pw = 0; pw < PW; ++pw) {
          float acc = 0.f;
          for (int iy = 0; iy < gh; ++iy) {
            float yy = sh + ph * bh + ((float)iy + .5f) * bh / (float)gh;
            for (int ix = 0; ix < gw; ++ix) {
              float xx = sw + pw * bw + ((float)ix + .5f) * bw / (float)gw;
              float y = yy, x = xx;
              if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) continue;  /* -> 0 */
              if (y <= 0) y = 0;
              if (x <= 0) x = 0;
              int yl = (int)y, xl = (int)x, yh, xh;
              if (yl >= H - 1) { yh = yl = (int)H - 1; y = (float)yl; } else yh = yl + 1;
              if (xl >= W - 1) { xh = xl = (int)W - 1; x = (float)xl; } else xh = xl + 1;
              float ly = y - yl, lx = x - xl, hy = 1.f - ly, hx = 1.f - lx;
              float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
              acc += w1 * p[yl * W + xl] + w2 * p[yl * W + xh] + w3 * p[yh * W + xl] +
                     w4 * p[yh * W + xh];
            }
          }
          out[((n * C + c) * PH + ph) * PW + pw] = acc / count;
        }
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * Region x concept alignment -- wsovod/modeling/class_heads/open_vocabulary_classifier.py:85-104
 * (contraction part; the projection MLP at :83 is out of scope) and the row softmax of
 * roi_heads/fast_rcnn_open_vocabulary.py:1034-1035.  F.normalize(x, p=2, dim) = x / max(||x||, 1e-12)
 * with the norm accumulated in fp32 by ATen; here the dot products are accumulated in double and
 * rounded once, which is within 1 ulp-ish of any fp32 summation order (tolerance 1e-5 rel in tests).
 * ---------------------------------------------------------------------------------------------- */
ORC_API void orc_align_fwd(const float* x, const float* cls, int64_t M, int64_t D, int64_t K,
                           float temperature, int norm_weight, int append_bg, const float* bias,
                           float* logits, float* probs) {
  int64_t KO = K + (append_bg ? 1 : 0);
  float* w = (float*)malloc(sizeof(float) * (size_t)(K * D));
  for (int64_t k = 0; k < K; ++k) {
    double s = 0;
    for (int64_t d = 0; d < D; ++d) s += (double)cls[k * D + d] * cls[k * D + d];
    float nrm = (float)sqrt(s);
    float den = norm_weight ? fmaxf_(nrm, 1e-12f) : 1.f;       /* :89-90 */
    for (int64_t d = 0; d < D; ++d) w[k * D + d] = norm_weight ? cls[k * D + d] / den : cls[k * D + d];
  }
  float* xn = (float*)malloc(sizeof(float) * (size_t)D);
  float* row = (float*)malloc(sizeof(float) * (size_t)KO);
  for (int64_t r = 0; r < M; ++r) {
    const float* xr = x + r * D;
    if (norm_weight) {                                          /* :94-95 */
      double s = 0;
      for (int64_t d = 0; d < D; ++d) s += (double)xr[d] * xr[d];
      float den = fmaxf_((float)sqrt(s), 1e-12f);
      for (int64_t d = 0; d < D; ++d) xn[d] = temperature * (xr[d] / den);
    } else {
      memcpy(xn, xr, sizeof(float) * (size_t)D);
    }
    for (int64_t k = 0; k < KO; ++k) {
      double s = 0;
      if (k < K) for (int64_t d = 0; d < D; ++d) s += (double)xn[d] * w[k * D + d];   /* :102 */
      float v = (float)s;
      if (bias) v = v + bias[0];                                /* :103-104 */
      row[k] = v;
      if (logits) logits[r * KO + k] = v;
    }
    if (probs) {
      float mx = row[0];
      for (int64_t k = 1; k < KO; ++k) mx = fmaxf_(mx, row[k]);
      double sum = 0;
      for (int64_t k = 0; k < KO; ++k) sum += exp((double)row[k] - mx);
      for (int64_t k = 0; k < KO; ++k) probs[r * KO + k] = (float)(exp((double)row[k] - mx) / sum);
    }
  }
  free(w); free(xn); free(row);
}

/* ------------------------------------------------------------------------------------------------
 * MIL two-stream score -- roi_heads/fast_rcnn_open_vocabulary.py:338-357 (scores) and :604-618
 * (image-level sum + clamp).  Softmaxes evaluated in double and rounded (tolerance 1e-5 rel).
 * ---------------------------------------------------------------------------------------------- */
ORC_API void orc_mil_fwd(const float* cls, const float* det, const int64_t* offsets, int64_t M,
                         int64_t N, int64_t K, float* scores, float* img_scores) {
  (void)M;
  for (int64_t n = 0; n < N; ++n) {
    int64_t r0 = offsets[n], r1 = offsets[n + 1];
    for (int64_t k = 0; k < K; ++k) {
      double mx = -INFINITY, sum = 0;
      /* K==1 special case (:338-340,356-357): the extra zero column only changes the row softmax */
      for (int64_t r = r0; r < r1; ++r) mx = fmax(mx, (double)det[r * K + k]);
      for (int64_t r = r0; r < r1; ++r) sum += exp((double)det[r * K + k] - mx);
      double tot = 0;
      for (int64_t r = r0; r < r1; ++r) {
        double rmx = (K == 1) ? fmax(0.0, (double)cls[r * K]) : -INFINITY, rs = 0;
        for (int64_t j = 0; j < K; ++j) rmx = fmax(rmx, (double)cls[r * K + j]);
        for (int64_t j = 0; j < K; ++j) rs += exp((double)cls[r * K + j] - rmx);
        if (K == 1) rs += exp(0.0 - rmx);
        float pc = (float)(exp((double)cls[r * K + k] - rmx) / rs);
        float pd = (float)(exp((double)det[r * K + k] - mx) / sum);
        float s = pc * pd;
        scores[r * K + k] = s;
        tot += s;
      }
      if (img_scores) {
        float t = (float)tot;
        img_scores[n * K + k] = fminf_(fmaxf_(t, 1e-6f), 1.0f - 1e-6f);   /* :617 */
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * Seed selection -- roi_heads.py:1079-1207 with top_k=1, thres=0 (the call at :801-807).
 * Per image, per image-level class (sorted unique, roi_heads.py:162): proposals with box area > 20
 * (:1090-1096), first argmax of the class score (torch.topk(1) on CPU: first index on ties), weight =
 * image-level score (:1140-1146).  Empty -> fallback seed (:1182-1207).
 * ---------------------------------------------------------------------------------------------- */
ORC_API void orc_pgt_top1(const float* scores, int64_t stride, const float* boxes,
                          const int64_t* offsets, const int64_t* gt_classes,
                          const int64_t* gt_offsets, const float* img_scores, int64_t N, int64_t K,
                          float* seed_boxes, int64_t* seed_classes, float* seed_scores,
                          float* seed_weights, int64_t* seed_rows, int64_t* seed_count) {
  for (int64_t n = 0; n < N; ++n) {
    int64_t g0 = gt_offsets[n], g1 = gt_offsets[n + 1];
    int any = 0;
    for (int64_t g = g0; g < g1; ++g) {
      int64_t c = gt_classes[g];
      int64_t best = -1; float bs = 0.f;
      for (int64_t r = offsets[n]; r < offsets[n + 1]; ++r) {
        const float* bx = boxes + r * 4;
        float area = (bx[2] - bx[0]) * (bx[3] - bx[1]);
        if (!(area > 20.f)) continue;
        float s = scores[r * stride + c];
        if (best < 0 || s > bs) { best = r; bs = s; }
      }
      seed_rows[g] = best; seed_classes[g] = c;
      if (best >= 0) {
        any = 1;
        memcpy(seed_boxes + g * 4, boxes + best * 4, 4 * sizeof(float));
        seed_scores[g] = bs;
        seed_weights[g] = img_scores[n * K + c];
      } else {
        memset(seed_boxes + g * 4, 0, 4 * sizeof(float));
        seed_scores[g] = 0.f; seed_weights[g] = 0.f;
      }
    }
    if (!any && g1 > g0) {   /* no eligible proposal: the area filter is per proposal, so all-or-none */
      float fb[4] = {-10000.f, -10000.f, 10000.f, 10000.f};
      memcpy(seed_boxes + g0 * 4, fb, sizeof(fb));
      seed_classes[g0] = 0; seed_scores[g0] = 1.f; seed_weights[g0] = 1.f; seed_rows[g0] = -1;
      seed_count[n] = 1;
    } else {
      seed_count[n] = g1 - g0;
    }
  }
}

/* pairwise IoU -- detectron2.structures.pairwise_iou (call site roi_heads.py:1770-1772):
 * every op is a separately rounded fp32 ATen op. */
static float d2_iou(const float* a, const float* b) {
  float w = fminf_(a[2], b[2]) - fmaxf_(a[0], b[0]);
  float h = fminf_(a[3], b[3]) - fmaxf_(a[1], b[1]);
  if (w < 0.f) w = 0.f;
  if (h < 0.f) h = 0.f;
  float inter = w * h;
  float aa = (a[2] - a[0]) * (a[3] - a[1]);
  float ab = (b[2] - b[0]) * (b[3] - b[1]);
  return inter > 0.f ? inter / (aa + ab - inter) : 0.f;
}

/* Assignment -- roi_heads.py:1770-1797 + _sample_proposals_wsl :1587-1593 + Matcher([thr],[0,1]). */
ORC_API void orc_refine_assign(const float* boxes, const int64_t* offsets, const float* seed_boxes,
                               const int64_t* seed_classes, const float* seed_scores,
                               const float* seed_weights, const int64_t* seed_offsets,
                               const int64_t* seed_count, int64_t N, int64_t num_classes,
                               float thr, int64_t* midx, int8_t* mlabel, float* miou,
                               int64_t* gt_classes, float* gt_boxes, float* gt_scores,
                               float* gt_weights) {
  for (int64_t n = 0; n < N; ++n) {
    int64_t s0 = seed_offsets[n];
    int64_t G = seed_count ? seed_count[n] : seed_offsets[n + 1] - s0;
    for (int64_t r = offsets[n]; r < offsets[n + 1]; ++r) {
      int64_t bi = 0; float bv = 0.f;
      for (int64_t g = 0; g < G; ++g) {
        float v = d2_iou(seed_boxes + (s0 + g) * 4, boxes + r * 4);
        if (g == 0 || v > bv) { bv = v; bi = g; }     /* max(dim=0): first index on ties */
      }
      int lab = (G > 0) && (bv >= thr);
      midx[r] = bi; mlabel[r] = (int8_t)lab; if (miou) miou[r] = bv;
      if (G > 0) {
        gt_classes[r] = lab ? seed_classes[s0 + bi] : num_classes;
        memcpy(gt_boxes + r * 4, seed_boxes + (s0 + bi) * 4, 4 * sizeof(float));
        gt_scores[r] = seed_scores[s0 + bi];
        gt_weights[r] = seed_weights[s0 + bi];
      } else {
        gt_classes[r] = num_classes;
        memset(gt_boxes + r * 4, 0, 4 * sizeof(float));
        gt_scores[r] = 0.f; gt_weights[r] = 0.f;
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * NMS -- torchvision 0.26 (un-vendored): csrc/ops/cpu/nms_kernel.cpp (mode 0) and the sm_100 SASS of
 * csrc/ops/cuda/nms_kernel.cu (mode 1); batched "vanilla" strategy of
 * torchvision/ops/boxes.py:97-120.  Call site roi_heads/fast_rcnn_open_vocabulary.py:206.
 * ---------------------------------------------------------------------------------------------- */
static int suppresses(const float* bi, const float* bj, float thr_f, double thr_d, int mode) {
  float w = fminf_(bi[2], bj[2]) - fmaxf_(bi[0], bj[0]);
  float h = fminf_(bi[3], bj[3]) - fmaxf_(bi[1], bj[1]);
  if (!(w > 0.f)) w = 0.f;   /* max(0, .) : NaN -> 0 like std::max((T)0, x) */
  if (!(h > 0.f)) h = 0.f;
  float inter = w * h;
  float ai = (bi[2] - bi[0]) * (bi[3] - bi[1]);
  if (mode == 0) {
    float aj = (bj[2] - bj[0]) * (bj[3] - bj[1]);
    float ovr = inter / (ai + aj - inter);
    return (double)ovr > thr_d;
  } else {
    float den = fmaf(bj[2] - bj[0], bj[3] - bj[1], ai) - inter;
    float ovr = inter / den;
    return ovr > thr_f;
  }
}

typedef struct { float s; int64_t i; } sitem;
static int cmp_desc(const void* a, const void* b) {   /* stable: score desc, index asc */
  const sitem* x = (const sitem*)a; const sitem* y = (const sitem*)b;
  if (x->s > y->s) return -1;
  if (x->s < y->s) return 1;
  return (x->i > y->i) - (x->i < y->i);
}

/* greedy NMS over the candidates listed in `order` (already score-desc); writes kept flags */
static void greedy(const float* boxes, const sitem* order, int64_t n, double thr, int mode,
                   uint8_t* kept_flag) {
  uint8_t* sup = (uint8_t*)calloc((size_t)n + 1, 1);
  double thr_d = thr;   /* torchvision passes the python float through as a C++ double */
  float thr_f = (float)thr;  /* the CUDA kernel narrows it once (F2F.F32.F64) */
  for (int64_t a = 0; a < n; ++a) {
    if (sup[a]) continue;
    kept_flag[order[a].i] = 1;
    const float* bi = boxes + order[a].i * 4;
    for (int64_t b = a + 1; b < n; ++b)
      if (!sup[b] && suppresses(bi, boxes + order[b].i * 4, thr_f, thr_d, mode)) sup[b] = 1;
  }
  free(sup);
}

/* returns number kept; keep[] = kept indices ordered by score desc (ties: index asc) */
ORC_API int64_t orc_batched_nms(const float* boxes, const float* scores, const int64_t* groups,
                                int64_t M, double thr, int mode, int64_t* keep) {
  if (M == 0) return 0;
  uint8_t* kept = (uint8_t*)calloc((size_t)M, 1);
  sitem* items = (sitem*)malloc(sizeof(sitem) * (size_t)M);
  uint8_t* done = (uint8_t*)calloc((size_t)M, 1);
  for (int64_t s = 0; s < M; ++s) {
    if (done[s]) continue;
    int64_t g = groups[s], n = 0;
    for (int64_t j = s; j < M; ++j)
      if (!done[j] && groups[j] == g) { done[j] = 1; items[n].s = scores[j]; items[n].i = j; ++n; }
    qsort(items, (size_t)n, sizeof(sitem), cmp_desc);
    greedy(boxes, items, n, thr, mode, kept);
  }
  int64_t nk = 0;
  for (int64_t j = 0; j < M; ++j) if (kept[j]) { items[nk].s = scores[j]; items[nk].i = j; ++nk; }
  qsort(items, (size_t)nk, sizeof(sitem), cmp_desc);
  for (int64_t j = 0; j < nk; ++j) keep[j] = items[j].i;
  free(kept); free(items); free(done);
  return nk;
}

/* fast_rcnn_inference_single_image -- roi_heads/fast_rcnn_open_vocabulary.py:149-217, class-agnostic
 * boxes.  probs [R,K+1], boxes [R,4]; image size (h, w).  Outputs padded to topk. */
ORC_API int64_t orc_detections_image(const float* probs, const float* boxes, int64_t R, int64_t K,
                                     float img_h, float img_w, float score_thr, double nms_thr,
                                     int64_t topk, int mode, float* det_boxes, float* det_scores,
                                     int64_t* det_classes, int64_t* det_rows) {
  int64_t cap = R * K, m = 0;
  float* cb = (float*)malloc(sizeof(float) * 4 * (size_t)(cap + 1));
  float* cs = (float*)malloc(sizeof(float) * (size_t)(cap + 1));
  int64_t* cg = (int64_t*)malloc(sizeof(int64_t) * (size_t)(cap + 1));
  int64_t* cr = (int64_t*)malloc(sizeof(int64_t) * (size_t)(cap + 1));
  for (int64_t r = 0; r < R; ++r) {
    int ok = 1;                                                    /* :178-182 */
    for (int j = 0; j < 4; ++j) ok &= isfinite(boxes[r * 4 + j]) != 0;
    for (int64_t k = 0; k <= K; ++k) ok &= isfinite(probs[r * (K + 1) + k]) != 0;
    if (!ok) continue;
    float b[4];                                                    /* Boxes.clip, :187-188 */
    b[0] = fminf_(fmaxf_(boxes[r * 4 + 0], 0.f), img_w);
    b[1] = fminf_(fmaxf_(boxes[r * 4 + 1], 0.f), img_h);
    b[2] = fminf_(fmaxf_(boxes[r * 4 + 2], 0.f), img_w);
    b[3] = fminf_(fmaxf_(boxes[r * 4 + 3], 0.f), img_h);
    for (int64_t k = 0; k < K; ++k)
      if (probs[r * (K + 1) + k] > score_thr) {                    /* :194-203 */
        memcpy(cb + m * 4, b, sizeof(b)); cs[m] = probs[r * (K + 1) + k]; cg[m] = k; cr[m] = r; ++m;
      }
  }
  int64_t* keep = (int64_t*)malloc(sizeof(int64_t) * (size_t)(m + 1));
  int64_t nk = orc_batched_nms(cb, cs, cg, m, nms_thr, mode, keep);   /* :206 */
  if (topk >= 0 && nk > topk) nk = topk;                              /* :207-208 */
  for (int64_t j = 0; j < nk; ++j) {
    memcpy(det_boxes + j * 4, cb + keep[j] * 4, 4 * sizeof(float));
    det_scores[j] = cs[keep[j]]; det_classes[j] = cg[keep[j]]; det_rows[j] = cr[keep[j]];
  }
  free(cb); free(cs); free(cg); free(cr); free(keep);
  return nk;
}
