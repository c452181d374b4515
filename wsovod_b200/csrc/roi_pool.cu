// roi_pool.cu -- kernel family (1): ROI max-pool (+argmax), the 3-way ROILoopPool and bilinear ROIAlign.
//
// Design (DESIGN.md "Kernel 1"): a proposal's window is read ~50x less often than it is re-read by
// the 100+ other proposals covering the same cells, so the unit of work is NOT a proposal.  A CTA
// owns (image n, a group of CB<=4 channels): it stages those CB feature planes ONCE into shared memory,
// channel-interleaved ([cell][CB] -> one LDS.128 fetches a cell for 4 channels), and then walks every
// proposal of that image.  Lanes map to consecutive flattened (proposal, bin) outputs, so a warp store
// is 32 consecutive floats of the (R,C,7,7) output (coalesced with no staging buffer) and all lanes of a
// warp share at most two proposals (no trip-count divergence to speak of).  Feature maps cross
// L2->SM exactly once per channel group; the only HBM stream left is the output itself.
//
// Bin edges are integer data shared by all channel groups, so a tiny prologue kernel computes them
// once per proposal (exact fp32 sequence of ROILoopPool_cpu.cpp:29-51) into an int16 table and
// builds a stable per-image ordering of the proposals (rois may arrive in any batch order).
#include "common.cuh"

#include <algorithm>
#include <cstdlib>

namespace wsovod {

enum { MODE_POOL = 0, MODE_LOOP = 1, MODE_ALIGN = 2 };

// roi_pool_pyr.cu: block-max fast path (7x7, values only)
size_t pool7_pyr_workspace(int64_t N, int64_t R);
int pool7_pyr_cb(int64_t C, int64_t H, int64_t W, int64_t R, bool with_argmax);
int pool7_pyr(const float* input, int64_t N, int64_t C, int64_t H, int64_t W, const float* rois, int64_t R,
              float scale, const float* row_scale, float row_scale_bias, float* output, int32_t* argmax, void* workspace,
              cudaStream_t st, float floor_v = -FLT_MAX);

// roi_align_sep.cu: ROIAlign 7x7, adaptive sample grid, separable tap tables
size_t align7_sep_workspace(int64_t R, int64_t H, int64_t W);
int align7_sep_cb(int64_t C, int64_t H, int64_t W);
int align7_sep(const float* input, int64_t N, int64_t C, int64_t H, int64_t W, int64_t R, const int32_t* counts,
               const int32_t* order, const float* alignp, const float* row_scale, float row_scale_bias, float* output,
               void* workspace, cudaStream_t st);

struct PoolParams {
  const float* input;
  const float* rois;
  const float* row_scale;
  float row_scale_bias;
  float* output;
  int32_t* argmax;
  // workspace
  const int32_t* counts;   // [N]   proposals per image
  const int32_t* order;    // [R]   proposal ids grouped by image (stable)
  const int16_t* edges;    // [R, EW] bin edge table
  const float* alignp;     // [R, 8] ROIAlign parameters
  int32_t N, C, H, W;
  int64_t R;
  int32_t PH, PW;
  int32_t CG;              // channel groups  = ceil(C / CB)
  int32_t S;               // proposal chunks per image
  int32_t sampling_ratio, aligned;
  int32_t P;               // 7x7 scan kernels: cells per staged row = W | 1 (odd: equal columns of different rows fall into
                           // different bank groups; simulated conflict wavefronts -8 % with four channels, -27 % with two)
};

// ------------------------------------------------------------------------------------------------
// prologue 1: per-proposal geometry (one thread per proposal)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void write_edges(int16_t* e, int rsh, int rsw, int reh, int rew, int PH,
                                            int PW, int H, int W) {
  // ROILoopPool_cpu.cpp:35-51
  int rw = max(rew - rsw + 1, 1);
  int rh = max(reh - rsh + 1, 1);
  float bh = __fdiv_rn((float)rh, (float)PH);
  float bw = __fdiv_rn((float)rw, (float)PW);
  for (int ph = 0; ph < PH; ++ph) {
    int hs = (int)floorf(__fmul_rn((float)ph, bh));
    int he = (int)ceilf(__fmul_rn((float)(ph + 1), bh));
    e[ph] = (int16_t)min(max(hs + rsh, 0), H);
    e[PH + ph] = (int16_t)min(max(he + rsh, 0), H);
  }
  for (int pw = 0; pw < PW; ++pw) {
    int ws = (int)floorf(__fmul_rn((float)pw, bw));
    int we = (int)ceilf(__fmul_rn((float)(pw + 1), bw));
    e[2 * PH + pw] = (int16_t)min(max(ws + rsw, 0), W);
    e[2 * PH + PW + pw] = (int16_t)min(max(we + rsw, 0), W);
  }
}

// saturating float->int like the reference's `int x = round(float)` on sane inputs; huge values clamp
__device__ __forceinline__ int round_i(float v) {
  v = roundf(v);
  v = fminf(fmaxf(v, -1.0e6f), 1.0e6f);
  return (int)v;
}

template <int MODE>
__global__ void roi_prepare_kernel(const float* __restrict__ rois, int64_t R, int N, int H, int W,
                                   float scale, int PH, int PW, int sampling_ratio, int aligned,
                                   int32_t* __restrict__ bidx, int32_t* __restrict__ counts,
                                   int16_t* __restrict__ edges, float* __restrict__ alignp,
                                   uint2* __restrict__ bins) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float* roi = rois + r * 5;
  int b = (int)roi[0];
  b = min(max(b, 0), N - 1);
  bidx[r] = b;
  atomicAdd(&counts[b], 1);
  if (r > 0) {   // counts[N] doubles as the "rois are not grouped by image" flag
    int pb = (int)rois[(r - 1) * 5];
    pb = min(max(pb, 0), N - 1);
    if (pb > b) atomicOr(&counts[N], 1);
  }
  const float x1 = roi[1], y1 = roi[2], x2 = roi[3], y2 = roi[4];
  if (MODE == MODE_POOL) {
    int16_t* e = edges + r * (2 * (PH + PW));
    write_edges(e, round_i(__fmul_rn(y1, scale)), round_i(__fmul_rn(x1, scale)),
                round_i(__fmul_rn(y2, scale)), round_i(__fmul_rn(x2, scale)), PH, PW, H, W);
    if (bins) {   // one 8-byte word per output bin: (hs | he << 16, ws | we << 16) -- coalesced per-lane fetch
      uint2* b = bins + r * (PH * PW);
      for (int ph = 0; ph < PH; ++ph)
        for (int pw = 0; pw < PW; ++pw)
          b[ph * PW + pw] = make_uint2((uint32_t)(uint16_t)e[ph] | ((uint32_t)(uint16_t)e[PH + ph] << 16),
                                       (uint32_t)(uint16_t)e[2 * PH + pw] | ((uint32_t)(uint16_t)e[2 * PH + PW + pw] << 16));
    }
  } else if (MODE == MODE_LOOP) {
    // ROILoopPool_cuda.cu:34-74: inner (/1.8) and outer (x1.8) boxes in image space, clamped.
    // The expressions are kept in the reference's own form and compiled with the same default
    // contraction rules (nvcc -fmad=true) the reference build uses.
    const float ratio = 1.8f;
    float rw_ = x2 - x1, rh_ = y2 - y1;
    float iw = rw_ / ratio, ih = rh_ / ratio;
    float ow = rw_ * ratio, oh = rh_ * ratio;
    float irw = rw_ - iw, irh = rh_ - ih;
    float orw = ow - rw_, orh = oh - rh_;
    float x1i = x1 + irw / 2, y1i = y1 + irh / 2, x2i = x2 - irw / 2, y2i = y2 - irh / 2;
    float x1o = x1 - orw / 2, y1o = y1 - orh / 2, x2o = x2 + orw / 2, y2o = y2 + orh / 2;
    const float xmax = (float)(1.0 * W / scale), ymax = (float)(1.0 * H / scale);
    x1i = fminf(fmaxf(x1i, 0.f), xmax); y1i = fminf(fmaxf(y1i, 0.f), ymax);
    x2i = fminf(fmaxf(x2i, 0.f), xmax); y2i = fminf(fmaxf(y2i, 0.f), ymax);
    x1o = fminf(fmaxf(x1o, 0.f), xmax); y1o = fminf(fmaxf(y1o, 0.f), ymax);
    x2o = fminf(fmaxf(x2o, 0.f), xmax); y2o = fminf(fmaxf(y2o, 0.f), ymax);
    const int EW = 4 * (PH + PW) + 8;
    int16_t* e = edges + r * EW;
    int rsw = round_i(x1 * scale), rsh = round_i(y1 * scale);
    int rew = round_i(x2 * scale), reh = round_i(y2 * scale);
    write_edges(e, rsh, rsw, reh, rew, PH, PW, H, W);                        // grid of the ROI
    int osw = round_i(x1o * scale), osh = round_i(y1o * scale);
    int oew = round_i(x2o * scale), oeh = round_i(y2o * scale);
    write_edges(e + 2 * (PH + PW), osh, osw, oeh, oew, PH, PW, H, W);        // grid of the outer box
    int16_t* q = e + 4 * (PH + PW);
    auto sat = [](int v) { return (int16_t)min(max(v, -32768), 32767); };
    q[0] = sat(round_i(y1i * scale)); q[1] = sat(round_i(y2i * scale));      // inner box (h range)
    q[2] = sat(round_i(x1i * scale)); q[3] = sat(round_i(x2i * scale));      // inner box (w range)
    q[4] = sat(rsh); q[5] = sat(reh); q[6] = sat(rsw); q[7] = sat(rew);      // the ROI itself
  } else {
    // torchvision roi_align
    float off = aligned ? 0.5f : 0.f;
    float sw = __fsub_rn(__fmul_rn(x1, scale), off), sh = __fsub_rn(__fmul_rn(y1, scale), off);
    float ew = __fsub_rn(__fmul_rn(x2, scale), off), eh = __fsub_rn(__fmul_rn(y2, scale), off);
    float rw = __fsub_rn(ew, sw), rh = __fsub_rn(eh, sh);
    if (!aligned) { rw = fmaxf(rw, 1.f); rh = fmaxf(rh, 1.f); }
    float bh = __fdiv_rn(rh, (float)PH), bw = __fdiv_rn(rw, (float)PW);
    int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(rh, (float)PH));
    int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(rw, (float)PW));
    float* a = alignp + r * 8;
    a[0] = sw; a[1] = sh; a[2] = bw; a[3] = bh;
    a[4] = __int_as_float(gh); a[5] = __int_as_float(gw);
    a[6] = (float)max(gh * gw, 1); a[7] = __fdiv_rn(1.f, a[6]);
  }
}

// prologue of the 7x7 fast path: one thread per (proposal, bin) -> coalesced 8-byte bin words.
// Same fp32 sequence as write_edges(); thread 0 of a proposal also does the per-image bookkeeping.
__global__ void roi_bins7_kernel(const float* __restrict__ rois, int64_t R, int N, int H, int W, float scale,
                                 int32_t* __restrict__ bidx, int32_t* __restrict__ counts,
                                 uint2* __restrict__ bins) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * 49) return;
  const int64_t r = i / 49;
  const int bin = (int)(i - r * 49);
  const int ph = bin / 7, pw = bin - ph * 7;
  const float* roi = rois + r * 5;
  const int rsw = round_i(__fmul_rn(roi[1], scale)), rsh = round_i(__fmul_rn(roi[2], scale));
  const int rew = round_i(__fmul_rn(roi[3], scale)), reh = round_i(__fmul_rn(roi[4], scale));
  const float bh = __fdiv_rn((float)max(reh - rsh + 1, 1), 7.f);
  const float bw = __fdiv_rn((float)max(rew - rsw + 1, 1), 7.f);
  const int hs = min(max((int)floorf(__fmul_rn((float)ph, bh)) + rsh, 0), H);
  const int he = min(max((int)ceilf(__fmul_rn((float)(ph + 1), bh)) + rsh, 0), H);
  const int ws = min(max((int)floorf(__fmul_rn((float)pw, bw)) + rsw, 0), W);
  const int we = min(max((int)ceilf(__fmul_rn((float)(pw + 1), bw)) + rsw, 0), W);
  bins[i] = make_uint2((uint32_t)hs | ((uint32_t)he << 16), (uint32_t)ws | ((uint32_t)we << 16));
  if (bin == 0) {
    int b = (int)roi[0];
    b = min(max(b, 0), N - 1);
    bidx[r] = b;
    atomicAdd(&counts[b], 1);
    if (r > 0) {
      int pb = (int)rois[(r - 1) * 5];
      pb = min(max(pb, 0), N - 1);
      if (pb > b) atomicOr(&counts[N], 1);
    }
  }
}

// prologue of the ROILoopPool fast path: one thread per (proposal, bin).  bins[2*i] = bin of the ROI's own
// grid, bins[2*i+1] = bin of the outer (x1.8) box's grid; rects[2*r] = inner (/1.8) box, rects[2*r+1] =
// the ROI, both as (h_lo | h_hi << 16, w_lo | w_hi << 16) with int16 fields (exclusion tests are strict).
// Geometry: ROILoopPool_cuda.cu:34-103,144-166, same expressions as roi_prepare_kernel<MODE_LOOP>.
// iboxes (optional): [2, R, 5] float rois holding the INTEGER boxes (batch, rsw, rsh, rew, reh) of the ROI and of the outer
// box -- fed to the block-max pooling path with spatial_scale 1, whose round(x * 1) gives the same integers back.
__global__ void roi_loopbins7_kernel(const float* __restrict__ rois, int64_t R, int N, int H, int W, float scale,
                                     int32_t* __restrict__ bidx, int32_t* __restrict__ counts,
                                     uint2* __restrict__ bins, uint2* __restrict__ rects, float* __restrict__ iboxes) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * 49) return;
  const int64_t r = i / 49;
  const int bin = (int)(i - r * 49);
  const int ph = bin / 7, pw = bin - ph * 7;
  const float* roi = rois + r * 5;
  const float x1 = roi[1], y1 = roi[2], x2 = roi[3], y2 = roi[4];
  const float ratio = 1.8f;
  float rw_ = x2 - x1, rh_ = y2 - y1;
  float iw = rw_ / ratio, ih = rh_ / ratio;
  float ow = rw_ * ratio, oh = rh_ * ratio;
  float irw = rw_ - iw, irh = rh_ - ih;
  float orw = ow - rw_, orh = oh - rh_;
  float x1i = x1 + irw / 2, y1i = y1 + irh / 2, x2i = x2 - irw / 2, y2i = y2 - irh / 2;
  float x1o = x1 - orw / 2, y1o = y1 - orh / 2, x2o = x2 + orw / 2, y2o = y2 + orh / 2;
  const float xmax = (float)(1.0 * W / scale), ymax = (float)(1.0 * H / scale);
  x1i = fminf(fmaxf(x1i, 0.f), xmax); y1i = fminf(fmaxf(y1i, 0.f), ymax);
  x2i = fminf(fmaxf(x2i, 0.f), xmax); y2i = fminf(fmaxf(y2i, 0.f), ymax);
  x1o = fminf(fmaxf(x1o, 0.f), xmax); y1o = fminf(fmaxf(y1o, 0.f), ymax);
  x2o = fminf(fmaxf(x2o, 0.f), xmax); y2o = fminf(fmaxf(y2o, 0.f), ymax);
  auto one = [&](int rsh, int rsw, int reh, int rew) {
    const float bh = __fdiv_rn((float)max(reh - rsh + 1, 1), 7.f);
    const float bw = __fdiv_rn((float)max(rew - rsw + 1, 1), 7.f);
    const int hs = min(max((int)floorf(__fmul_rn((float)ph, bh)) + rsh, 0), H);
    const int he = min(max((int)ceilf(__fmul_rn((float)(ph + 1), bh)) + rsh, 0), H);
    const int ws = min(max((int)floorf(__fmul_rn((float)pw, bw)) + rsw, 0), W);
    const int we = min(max((int)ceilf(__fmul_rn((float)(pw + 1), bw)) + rsw, 0), W);
    return make_uint2((uint32_t)hs | ((uint32_t)he << 16), (uint32_t)ws | ((uint32_t)we << 16));
  };
  const int rsw = round_i(x1 * scale), rsh = round_i(y1 * scale), rew = round_i(x2 * scale), reh = round_i(y2 * scale);
  const int osh = round_i(y1o * scale), osw = round_i(x1o * scale), oeh = round_i(y2o * scale), oew = round_i(x2o * scale);
  bins[2 * i] = one(rsh, rsw, reh, rew);
  bins[2 * i + 1] = one(osh, osw, oeh, oew);
  if (bin == 0 && iboxes) {
    float* a = iboxes + r * 5;
    float* b = iboxes + (R + r) * 5;
    a[0] = b[0] = roi[0];
    a[1] = (float)rsw; a[2] = (float)rsh; a[3] = (float)rew; a[4] = (float)reh;
    b[1] = (float)osw; b[2] = (float)osh; b[3] = (float)oew; b[4] = (float)oeh;
  }
  if (bin == 0) {
    auto sat = [](int v) { return (uint32_t)(uint16_t)(int16_t)min(max(v, -32768), 32767); };
    rects[2 * r] = make_uint2(sat(round_i(y1i * scale)) | (sat(round_i(y2i * scale)) << 16),
                              sat(round_i(x1i * scale)) | (sat(round_i(x2i * scale)) << 16));
    rects[2 * r + 1] = make_uint2(sat(rsh) | (sat(reh) << 16), sat(rsw) | (sat(rew) << 16));
    int b = (int)roi[0];
    b = min(max(b, 0), N - 1);
    bidx[r] = b;
    atomicAdd(&counts[b], 1);
    if (r > 0) {
      int pb = (int)rois[(r - 1) * 5];
      pb = min(max(pb, 0), N - 1);
      if (pb > b) atomicOr(&counts[N], 1);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// prologue 2: stable grouping of proposal ids by image (one CTA per image)
// ------------------------------------------------------------------------------------------------
__global__ void roi_order_kernel(const int32_t* __restrict__ bidx, const int32_t* __restrict__ counts,
                                 int64_t R, int N, int32_t* __restrict__ order) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nw = blockDim.x >> 5;
  if (counts[N] == 0) {   // already grouped by image (what ROIPooler produces): the order is the identity
    int s = 0;
    for (int m = 0; m < n; ++m) s += counts[m];
    const int c = counts[n];
    for (int i = tid; i < c; i += blockDim.x) order[s + i] = s + i;
    return;
  }
  if (tid == 0) {
    int s = 0;
    for (int m = 0; m < n; ++m) s += counts[m];
    s_base = s;
  }
  __syncthreads();
  for (int64_t c0 = 0; c0 < R; c0 += blockDim.x) {
    int64_t r = c0 + tid;
    bool f = r < R && bidx[r] == n;
    unsigned bal = __ballot_sync(0xffffffffu, f);
    if (lane == 0) s_warp[wid] = __popc(bal);
    __syncthreads();
    int pre = 0, tot = 0;
    for (int w = 0; w < nw; ++w) {
      int c = s_warp[w];
      if (w < wid) pre += c;
      tot += c;
    }
    if (f) order[s_base + pre + __popc(bal & ((1u << lane) - 1))] = (int32_t)r;
    __syncthreads();
    if (tid == 0) s_base += tot;
    __syncthreads();
  }
}

// for roi_loop_dtype.cu
int launch_roi_order(const int32_t* bidx, const int32_t* counts, int64_t R, int N, int32_t* order, cudaStream_t st) {
  roi_order_kernel<<<(unsigned)N, 256, 0, st>>>(bidx, counts, R, N, order);
  return after_launch();
}

// ------------------------------------------------------------------------------------------------
// main kernel
// ------------------------------------------------------------------------------------------------
template <int CB> struct Vec;
template <> struct Vec<4> { using T = float4; };
template <> struct Vec<2> { using T = float2; };
template <> struct Vec<1> { using T = float; };

template <int CB> __device__ __forceinline__ void unpack(const typename Vec<CB>::T& v, float* f);
template <> __device__ __forceinline__ void unpack<4>(const float4& v, float* f) { f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w; }
template <> __device__ __forceinline__ void unpack<2>(const float2& v, float* f) { f[0] = v.x; f[1] = v.y; }
template <> __device__ __forceinline__ void unpack<1>(const float& v, float* f) { f[0] = v; }
template <int CB> __device__ __forceinline__ typename Vec<CB>::T pack(const float* f);
template <> __device__ __forceinline__ float4 pack<4>(const float* f) { return make_float4(f[0], f[1], f[2], f[3]); }
template <> __device__ __forceinline__ float2 pack<2>(const float* f) { return make_float2(f[0], f[1]); }
template <> __device__ __forceinline__ float pack<1>(const float* f) { return f[0]; }

// CB: channels per CTA (interleaved in smem); SMEM=false reads the (single) plane from global memory
// (fallback for maps that do not fit shared memory); ARG: track/emit argmax.
template <int CB, int MODE, bool SMEM, bool ARG>
__global__ void __launch_bounds__(1024, 1) roi_plane_kernel(const PoolParams p) {
  using V = typename Vec<CB>::T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int H = p.H, W = p.W, HW = H * W;
  const int PH = p.PH, PW = p.PW, BINS = PH * PW;
  const int bid = blockIdx.x;
  const int cg = bid % p.CG;
  const int s = (bid / p.CG) % p.S;
  const int n = bid / (p.CG * p.S);
  const int c0 = cg * CB;
  const int nc = min(CB, p.C - c0);

  // proposals of image n handled by this CTA: positions [pos0, pos0 + nroi) of `order`
  int start = 0;
  for (int m = 0; m < n; ++m) start += p.counts[m];
  const int cnt = p.counts[n];
  const int per = (cnt + p.S - 1) / p.S;
  const int pos0 = s * per;
  const int nroi = min(cnt, pos0 + per) - pos0;
  if (nroi <= 0) return;

  const float* src = p.input + ((int64_t)n * p.C + c0) * HW;
  const V* plane;
  if (SMEM) {
    V* sp = reinterpret_cast<V*>(smem_raw);
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
      float f[CB];
#pragma unroll
      for (int k = 0; k < CB; ++k) f[k] = k < nc ? __ldg(src + (int64_t)k * HW + i) : 0.f;
      sp[i] = pack<CB>(f);
    }
    __syncthreads();
    plane = sp;
  } else {
    plane = reinterpret_cast<const V*>(src);   // CB == 1
  }

  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int total = nroi * BINS;
  const int EW = MODE == MODE_LOOP ? 4 * (PH + PW) + 8 : 2 * (PH + PW);
  const int64_t block_stride = p.R * (int64_t)p.C * BINS;   // ROILoopPool: roi | frame | context

  for (int flat0 = wid * 32; flat0 < total; flat0 += nw * 32) {
    const int flat = flat0 + lane;
    if (flat >= total) break;
    const int rpos = flat / BINS;
    const int bin = flat - rpos * BINS;
    const int ph = bin / PW, pw = bin - ph * PW;
    const int r = __ldg(p.order + start + pos0 + rpos);
    float scale = 1.f;
    if (p.row_scale) scale = __fadd_rn(__ldg(p.row_scale + r), p.row_scale_bias);
    const int64_t obase = ((int64_t)r * p.C + c0) * BINS + bin;

    if (MODE == MODE_POOL) {
      const int16_t* e = p.edges + (int64_t)r * EW;
      const int hs = e[ph], he = e[PH + ph], ws = e[2 * PH + pw], we = e[2 * PH + PW + pw];
      const bool empty = (he <= hs) || (we <= ws);
      float m[CB];
      int mi[CB];
#pragma unroll
      for (int k = 0; k < CB; ++k) { m[k] = empty ? 0.f : -FLT_MAX; mi[k] = -1; }
      for (int h = hs; h < he; ++h) {
        const V* row = plane + h * W;
#pragma unroll 2
        for (int w = ws; w < we; ++w) {
          float f[CB];
          unpack<CB>(row[w], f);
#pragma unroll
          for (int k = 0; k < CB; ++k)
            if (f[k] > m[k]) { m[k] = f[k]; if (ARG) mi[k] = h * W + w; }
        }
      }
#pragma unroll
      for (int k = 0; k < CB; ++k)
        if (k < nc) {
          p.output[obase + (int64_t)k * BINS] = p.row_scale ? __fmul_rn(m[k], scale) : m[k];
          if (ARG) p.argmax[obase + (int64_t)k * BINS] = mi[k];
        }
    } else if (MODE == MODE_LOOP) {
      const int16_t* e = p.edges + (int64_t)r * EW;
      const int16_t* q = e + 4 * (PH + PW);
      {  // roi + frame on the ROI's own grid (ROILoopPool_cuda.cu:76-142)
        const int hs = e[ph], he = e[PH + ph], ws = e[2 * PH + pw], we = e[2 * PH + PW + pw];
        const int ish = q[0], ieh = q[1], isw = q[2], iew = q[3];
        float m[CB], mf[CB];
        int mi[CB], mfi[CB];
#pragma unroll
        for (int k = 0; k < CB; ++k) { m[k] = 0.f; mf[k] = 0.f; mi[k] = -1; mfi[k] = -1; }
        for (int h = hs; h < he; ++h) {
          const V* row = plane + h * W;
          const bool in_h = h > ish && h < ieh;
          for (int w = ws; w < we; ++w) {
            float f[CB];
            unpack<CB>(row[w], f);
            const bool inside = in_h && (w > isw && w < iew);
#pragma unroll
            for (int k = 0; k < CB; ++k) {
              if (f[k] > m[k]) { m[k] = f[k]; mi[k] = h * W + w; }
              if (!inside && f[k] > mf[k]) { mf[k] = f[k]; mfi[k] = h * W + w; }
            }
          }
        }
#pragma unroll
        for (int k = 0; k < CB; ++k)
          if (k < nc) {
            const int64_t o = obase + (int64_t)k * BINS;
            p.output[o] = p.row_scale ? __fmul_rn(m[k], scale) : m[k];
            p.output[o + block_stride] = p.row_scale ? __fmul_rn(mf[k], scale) : mf[k];
            if (ARG) { p.argmax[o] = mi[k]; p.argmax[o + block_stride] = mfi[k]; }
          }
      }
      {  // context on the outer box's grid, excluding the ROI (ROILoopPool_cuda.cu:144-202)
        const int16_t* e2 = e + 2 * (PH + PW);
        const int hs = e2[ph], he = e2[PH + ph], ws = e2[2 * PH + pw], we = e2[2 * PH + PW + pw];
        const int ish = q[4], ieh = q[5], isw = q[6], iew = q[7];
        float mc[CB];
        int mci[CB];
#pragma unroll
        for (int k = 0; k < CB; ++k) { mc[k] = 0.f; mci[k] = -1; }
        for (int h = hs; h < he; ++h) {
          const V* row = plane + h * W;
          const bool in_h = h > ish && h < ieh;
          for (int w = ws; w < we; ++w) {
            if (in_h && (w > isw && w < iew)) continue;
            float f[CB];
            unpack<CB>(row[w], f);
#pragma unroll
            for (int k = 0; k < CB; ++k)
              if (f[k] > mc[k]) { mc[k] = f[k]; mci[k] = h * W + w; }
          }
        }
#pragma unroll
        for (int k = 0; k < CB; ++k)
          if (k < nc) {
            const int64_t o = obase + (int64_t)k * BINS + 2 * block_stride;
            p.output[o] = p.row_scale ? __fmul_rn(mc[k], scale) : mc[k];
            if (ARG) p.argmax[o] = mci[k];
          }
      }
    } else {  // MODE_ALIGN
      const float* a = p.alignp + (int64_t)r * 8;
      const float sw = a[0], sh = a[1], bw = a[2], bh = a[3], count = a[6];
      const int gh = __float_as_int(a[4]), gw = __float_as_int(a[5]);
      float acc[CB];
#pragma unroll
      for (int k = 0; k < CB; ++k) acc[k] = 0.f;
      // every operation individually rounded, in torchvision's CPU order (roi_align_common.h), so the
      // result does not depend on FMA contraction
      for (int iy = 0; iy < gh; ++iy) {
        const float yy = __fadd_rn(__fadd_rn(sh, __fmul_rn((float)ph, bh)),
                                   __fdiv_rn(__fmul_rn((float)iy + .5f, bh), (float)gh));
        for (int ix = 0; ix < gw; ++ix) {
          const float xx = __fadd_rn(__fadd_rn(sw, __fmul_rn((float)pw, bw)),
                                     __fdiv_rn(__fmul_rn((float)ix + .5f, bw), (float)gw));
          float y = yy, x = xx;
          if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) continue;
          if (y <= 0) y = 0;
          if (x <= 0) x = 0;
          int yl = (int)y, xl = (int)x, yh, xh;
          if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else yh = yl + 1;
          if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
          const float ly = __fsub_rn(y, (float)yl), lx = __fsub_rn(x, (float)xl);
          const float hy = __fsub_rn(1.f, ly), hx = __fsub_rn(1.f, lx);
          const float w1 = __fmul_rn(hy, hx), w2 = __fmul_rn(hy, lx), w3 = __fmul_rn(ly, hx), w4 = __fmul_rn(ly, lx);
          float f1[CB], f2[CB], f3[CB], f4[CB];
          unpack<CB>(plane[yl * W + xl], f1);
          unpack<CB>(plane[yl * W + xh], f2);
          unpack<CB>(plane[yh * W + xl], f3);
          unpack<CB>(plane[yh * W + xh], f4);
#pragma unroll
          for (int k = 0; k < CB; ++k) {
            const float val = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w1, f1[k]), __fmul_rn(w2, f2[k])),
                                                  __fmul_rn(w3, f3[k])), __fmul_rn(w4, f4[k]));
            acc[k] = __fadd_rn(acc[k], val);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < CB; ++k)
        if (k < nc) {
          float v = __fdiv_rn(acc[k], count);
          p.output[obase + (int64_t)k * BINS] = p.row_scale ? __fmul_rn(v, scale) : v;
        }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// fast path: ROIPool 7x7, four channels per CTA (the shape every shipped WSOVOD config uses)
// ------------------------------------------------------------------------------------------------
template <int CB> __device__ __forceinline__ void lds_cell(uint32_t addr, float* f);
template <> __device__ __forceinline__ void lds_cell<4>(uint32_t addr, float* f) {
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f[0]), "=f"(f[1]), "=f"(f[2]), "=f"(f[3]) : "r"(addr));
}
template <> __device__ __forceinline__ void lds_cell<2>(uint32_t addr, float* f) {
  asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(f[0]), "=f"(f[1]) : "r"(addr));
}

// Lanes own consecutive flattened (proposal, bin) outputs.  Per pass a lane fetches ONE 8-byte bin word
// (coalesced), scans its bin in shared memory and stores CB (+CB) scalars that are contiguous across the
// warp.  CB = 4 channels per CTA (16-byte cells, LDS.128) when four planes fit shared memory, else 2
// (8-byte cells, LDS.64).  ARG=false: max only (FMNMX3) and, because max is order independent, each
// lane starts its column scan at an offset that puts the G = 32/CB lanes sharing a shared-memory phase
// on G different bank groups (bank group of a cell = cell index mod G): conflict-free for bins >= G
// wide; narrower bins spread the lanes that would start on the same group.  ARG=true keeps
// torchvision's h-major scan (strict >, first maximum wins).
template <int CB, bool ARG>
__global__ void __launch_bounds__(1024, 1) roi_pool7_kernel(const PoolParams p, const uint2* __restrict__ bins) {
  using V = typename Vec<CB>::T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int BINS = 49;
  constexpr uint32_t CS = 4u * CB;          // bytes per interleaved cell
  constexpr int G = 32 / CB;                // lanes per shared-memory phase == bank groups
  const int H = p.H, W = p.W, HW = H * W;
  const int bid = blockIdx.x;
  const int cg = bid % p.CG;
  const int sidx = (bid / p.CG) % p.S;
  const int n = bid / (p.CG * p.S);
  const int c0 = cg * CB;
  const int nc = min(CB, p.C - c0);
  int start = 0;
  for (int m = 0; m < n; ++m) start += __ldg(p.counts + m);
  const int cnt = __ldg(p.counts + n);
  const int per = (cnt + p.S - 1) / p.S;
  const int pos0 = sidx * per;
  const int nroi = min(cnt, pos0 + per) - pos0;
  if (nroi <= 0) return;
  {
    const float* src = p.input + ((int64_t)n * p.C + c0) * HW;
    V* sp = reinterpret_cast<V*>(smem_raw);
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
      float f[CB];
#pragma unroll
      for (int k = 0; k < CB; ++k) f[k] = k < nc ? __ldg(src + (int64_t)k * HW + i) : 0.f;
      sp[i + (i / W) * (p.P - W)] = pack<CB>(f);
    }
  }
  __syncthreads();
  uint32_t sbase;
  {  // volatile: computed once, never re-materialised inside the loops
    unsigned long long s64;
    asm volatile("cvta.to.shared.u64 %0, %1;" : "=l"(s64) : "l"((unsigned long long)(uintptr_t)smem_raw));
    sbase = (uint32_t)s64;
  }
  const int P = p.P;
  const uint32_t pitch = (uint32_t)P * CS;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int total = nroi * BINS;
  const int32_t* order = p.order + start + pos0;
  const int stride = nw * 32;

  // software pipeline: the (proposal id, bin word, scale) of the NEXT pass are fetched while this one runs
  int flat = wid * 32 + lane;
  int r_n = 0, bin_n = 0;
  uint2 e_n = make_uint2(0, 0);
  float sc_n = 1.f;
  auto fetch = [&](int f) {
    if (f < total) {
      const int rpos = f / BINS;
      bin_n = f - rpos * BINS;
      r_n = __ldg(order + rpos);
      e_n = __ldg(bins + (int64_t)r_n * BINS + bin_n);
      if (p.row_scale) sc_n = __fadd_rn(__ldg(p.row_scale + r_n), p.row_scale_bias);
    }
  };
  fetch(flat);
  for (; flat < total; flat += stride) {
    const int r = r_n, bin = bin_n;
    const uint2 e = e_n;
    const float scale = sc_n;
    fetch(flat + stride);
    const int hs = e.x & 0xffff, he = e.x >> 16, ws = e.y & 0xffff, we = e.y >> 16;
    const int bw = we - ws;
    const bool empty = (he <= hs) || (bw <= 0);
    float m[CB];
    int mi[CB];
#pragma unroll
    for (int k = 0; k < CB; ++k) { m[k] = empty ? 0.f : -FLT_MAX; mi[k] = -1; }
    if (!empty) {
      const int cell0 = hs * P + ws;
      if (!ARG) {
        int rot = ((lane & (G - 1)) - cell0) & (G - 1);
        if (bw < G) {
          const unsigned grp = __match_any_sync(__activemask(), ((lane / G) * G) | (cell0 & (G - 1)));
          rot = __popc(grp & ((1u << lane) - 1));
          rot = rot < bw ? rot : rot % bw;
        }
        uint32_t lo = sbase + (uint32_t)cell0 * CS;              // first cell of the bin row
        const uint32_t span = (uint32_t)bw * CS;
        for (int h = hs; h < he; ++h, lo += pitch) {
          const uint32_t hi = lo + span;
          uint32_t a = lo + (uint32_t)rot * CS;
#pragma unroll 2
          for (int t = 0; t < bw; ++t) {
            float f[CB];
            lds_cell<CB>(a, f);
#pragma unroll
            for (int k = 0; k < CB; ++k) m[k] = fmaxf(m[k], f[k]);
            a += CS;
            a = a == hi ? lo : a;
          }
        }
      } else {
        int rowi = hs * W + ws;                                   // the argmax is an index into the H x W map
        uint32_t lo = sbase + (uint32_t)cell0 * CS;
        for (int h = hs; h < he; ++h, lo += pitch, rowi += W) {
          uint32_t a = lo;
          int idx = rowi;
#pragma unroll 2
          for (int t = 0; t < bw; ++t, a += CS, ++idx) {
            float f[CB];
            lds_cell<CB>(a, f);
#pragma unroll
            for (int k = 0; k < CB; ++k)
              if (f[k] > m[k]) { m[k] = f[k]; mi[k] = idx; }
          }
        }
      }
    }
    const int64_t o = ((int64_t)r * p.C + c0) * BINS + bin;
#pragma unroll
    for (int k = 0; k < CB; ++k)
      if (k < nc) {
        __stcs(p.output + o + k * BINS, p.row_scale ? __fmul_rn(m[k], scale) : m[k]);
        if (ARG) __stcs(p.argmax + o + k * BINS, mi[k]);
      }
  }
}

// ROILoopPool fast path (values only; argmax requests use the generic kernel): roi | frame | context.  All maxima start
// at 0 ("assum all input is >=0", ROILoopPool_cuda.cu:107-113).
// One bin minus a hole (an open rectangle): `out` collects the cells outside it and, with ALL, `all` every cell.  ONE
// loop nest for every lane with the hole as a per-cell predicate: a warp's lanes hold bins of the same grid, so their
// trip counts differ by a cell, and they stay converged whatever the hole cuts.  (The segment form -- up to three
// scan loops per row, chosen per row -- serialised the lanes over its call sites: ncu counted 12 of 32 threads
// active per instruction and 8.1 G warp instructions for the 3-way pool at c2.)
template <int CB, bool ALL>
__device__ __forceinline__ void scan_holed(float* all, float* out, uint32_t sbase, uint32_t pitch, int P, const uint2 e,
                                           const uint2 x) {
  constexpr uint32_t CS = 4u * CB;
  const int hs = e.x & 0xffff, he = e.x >> 16, ws = e.y & 0xffff, we = e.y >> 16;
  const int ish = (int16_t)(x.x & 0xffff), ieh = (int16_t)(x.x >> 16);
  const int isw = (int16_t)(x.y & 0xffff), iew = (int16_t)(x.y >> 16);
  const int l1 = min(we, max(ws, isw + 1));        // hole columns of this bin: [l1, r0)
  const int r0 = max(l1, min(we, iew));
  const int t1 = min(he, max(hs, ish + 1));        // hole rows of this bin: [t1, b0)
  const int b0 = max(t1, min(he, ieh));
  if (!ALL && t1 == hs && b0 == he && l1 == ws && r0 == we) return;   // the bin lies inside the hole: nothing to collect
  const uint32_t hole_w = (uint32_t)(r0 - l1);
  uint32_t row = sbase + (uint32_t)(hs * P + ws) * CS;
  for (int h = hs; h < he; ++h, row += pitch) {
    const uint32_t hw = (h >= t1 && h < b0) ? hole_w : 0u;            // 0: no cell of this row is inside
    uint32_t a = row;
#pragma unroll 2
    for (int w = ws; w < we; ++w, a += CS) {
      const bool inside = (uint32_t)(w - l1) < hw;
      if (ALL || !inside) {
        float f[CB];
        lds_cell<CB>(a, f);
#pragma unroll
        for (int k = 0; k < CB; ++k) {
          if (ALL) all[k] = fmaxf(all[k], f[k]);
          if (!inside) out[k] = fmaxf(out[k], f[k]);
        }
      }
    }
  }
}

template <int CB>
__global__ void __launch_bounds__(1024, 1) roi_loop7_kernel(const PoolParams p, const uint2* __restrict__ bins,
                                                             const uint2* __restrict__ rects) {
  using V = typename Vec<CB>::T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int BINS = 49;
  constexpr uint32_t CS = 4u * CB;
  const int H = p.H, W = p.W, HW = H * W;
  const int bid = blockIdx.x;
  const int cg = bid % p.CG;
  const int sidx = (bid / p.CG) % p.S;
  const int n = bid / (p.CG * p.S);
  const int c0 = cg * CB;
  const int nc = min(CB, p.C - c0);
  int start = 0;
  for (int m = 0; m < n; ++m) start += __ldg(p.counts + m);
  const int cnt = __ldg(p.counts + n);
  const int per = (cnt + p.S - 1) / p.S;
  const int pos0 = sidx * per;
  const int nroi = min(cnt, pos0 + per) - pos0;
  if (nroi <= 0) return;
  {
    const float* src = p.input + ((int64_t)n * p.C + c0) * HW;
    V* sp = reinterpret_cast<V*>(smem_raw);
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
      float f[CB];
#pragma unroll
      for (int k = 0; k < CB; ++k) f[k] = k < nc ? __ldg(src + (int64_t)k * HW + i) : 0.f;
      sp[i + (i / W) * (p.P - W)] = pack<CB>(f);
    }
  }
  __syncthreads();
  uint32_t sbase;
  {
    unsigned long long s64;
    asm volatile("cvta.to.shared.u64 %0, %1;" : "=l"(s64) : "l"((unsigned long long)(uintptr_t)smem_raw));
    sbase = (uint32_t)s64;
  }
  const uint32_t pitch = (uint32_t)p.P * CS;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int total = nroi * BINS;
  const int32_t* order = p.order + start + pos0;
  const int64_t block_stride = p.R * (int64_t)p.C * BINS;
  // the (proposal id, two bin words, two hole rectangles, scale) of the NEXT pass are fetched while this one runs
  const int stride = nw * 32;
  int r_n = 0, bin_n = 0;
  uint4 e_n = make_uint4(0, 0, 0, 0), x_n = make_uint4(0, 0, 0, 0);
  float sc_n = 1.f;
  auto fetch = [&](int f) {
    if (f < total) {
      const int rpos = f / BINS;
      bin_n = f - rpos * BINS;
      r_n = __ldg(order + rpos);
      e_n = __ldg(reinterpret_cast<const uint4*>(bins + 2 * ((int64_t)r_n * BINS + bin_n)));   // ROI-grid bin | outer-grid bin
      x_n = __ldg(reinterpret_cast<const uint4*>(rects + 2 * (int64_t)r_n));                    // inner box | the ROI
      if (p.row_scale) sc_n = __fadd_rn(__ldg(p.row_scale + r_n), p.row_scale_bias);
    }
  };
  fetch(wid * 32 + lane);
  for (int flat = wid * 32 + lane; flat < total; flat += stride) {
    const int bin = bin_n, r = r_n;
    const uint2 ea = make_uint2(e_n.x, e_n.y), eb = make_uint2(e_n.z, e_n.w);
    const uint2 ri = make_uint2(x_n.x, x_n.y), rr = make_uint2(x_n.z, x_n.w);
    const float scale = sc_n;
    fetch(flat + stride);
    float roi[CB], com[CB], ctx[CB];
#pragma unroll
    for (int k = 0; k < CB; ++k) { roi[k] = 0.f; com[k] = 0.f; ctx[k] = 0.f; }
    scan_holed<CB, true>(roi, com, sbase, pitch, p.P, ea, ri);     // ROI grid: every cell, and the cells outside the inner box
    scan_holed<CB, false>(nullptr, ctx, sbase, pitch, p.P, eb, rr); // outer grid without the ROI's interior (never loaded)
    const int64_t o = ((int64_t)r * p.C + c0) * BINS + bin;
#pragma unroll
    for (int k = 0; k < CB; ++k)
      if (k < nc) {
        const float vr = roi[k];
        __stcs(p.output + o + k * BINS, p.row_scale ? __fmul_rn(vr, scale) : vr);
        __stcs(p.output + o + k * BINS + block_stride, p.row_scale ? __fmul_rn(com[k], scale) : com[k]);
        __stcs(p.output + o + k * BINS + 2 * block_stride, p.row_scale ? __fmul_rn(ctx[k], scale) : ctx[k]);
      }
  }
}

// ROILoopPool on the block-max planes, second half.  The three streams are maxima (starting at 0) over a bin, over a
// bin minus the inner box's open interior, and over a bin of the outer grid minus the ROI's open interior.  The
// block-max pooling kernel (roi_pool_pyr.cu, floor 0) writes plain bin maxima for all three -- right for stream 0, and
// right for every bin of streams 1 / 2 that does not meet the excluded interior (the outer ring of the grid: typically
// 24 of 49).  This kernel rewrites the others: a bin wholly inside the interior becomes 0 without a load, a bin cut
// by it is scanned on the raw plane in row segments left and right of the hole, like roi_loop7_kernel does for every
// bin.  Lanes = consecutive (proposal, bin); a lane serves that bin in both grids.
template <int CB>
__global__ void __launch_bounds__(1024, 1) roi_loop7_fix_kernel(const PoolParams p, const uint2* __restrict__ bins,
                                                                 const uint2* __restrict__ rects) {
  using V = typename Vec<CB>::T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int BINS = 49;
  constexpr uint32_t CS = 4u * CB;
  const int H = p.H, W = p.W, HW = H * W;
  const int bid = blockIdx.x;
  const int cg = bid % p.CG;
  const int sidx = (bid / p.CG) % p.S;
  const int n = bid / (p.CG * p.S);
  const int c0 = cg * CB;
  const int nc = min(CB, p.C - c0);
  int start = 0;
  for (int m = 0; m < n; ++m) start += __ldg(p.counts + m);
  const int cnt = __ldg(p.counts + n);
  const int per = (cnt + p.S - 1) / p.S;
  const int pos0 = sidx * per;
  const int nroi = min(cnt, pos0 + per) - pos0;
  if (nroi <= 0) return;
  {
    const float* src = p.input + ((int64_t)n * p.C + c0) * HW;
    V* sp = reinterpret_cast<V*>(smem_raw);
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
      float f[CB];
#pragma unroll
      for (int k = 0; k < CB; ++k) f[k] = k < nc ? __ldg(src + (int64_t)k * HW + i) : 0.f;
      sp[i + (i / W) * (p.P - W)] = pack<CB>(f);
    }
  }
  __syncthreads();
  uint32_t sbase;
  {
    unsigned long long s64;
    asm volatile("cvta.to.shared.u64 %0, %1;" : "=l"(s64) : "l"((unsigned long long)(uintptr_t)smem_raw));
    sbase = (uint32_t)s64;
  }
  const uint32_t pitch = (uint32_t)p.P * CS;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int total = nroi * BINS;
  const int32_t* order = p.order + start + pos0;
  const int64_t block_stride = p.R * (int64_t)p.C * BINS;
  // Only ~16 of a grid's 49 bins are CUT by the hole, so lanes that walked consecutive bins would mostly idle next to
  // one that scans.  Each warp therefore compacts its cut bins into a small queue (behind the planes in shared memory)
  // and scans them 32 at a time: every lane of a scanning pass has a bin, and a pass holds bins of one grid (similar
  // sizes).  Bins wholly inside the hole are zeroed on the spot.
  uint32_t* queue = reinterpret_cast<uint32_t*>(smem_raw + (size_t)CB * H * p.P * sizeof(float)) + wid * 64;

  auto geometry = [&](int r, int bin, int g, int& hs, int& he, int& ws, int& we, int& l1, int& r0, int& t1, int& b0) {
    const uint2 e = __ldg(bins + 2 * ((int64_t)r * BINS + bin) + g);
    const uint2 x = __ldg(rects + 2 * (int64_t)r + g);
    hs = e.x & 0xffff; he = e.x >> 16; ws = e.y & 0xffff; we = e.y >> 16;
    const int ish = (int16_t)(x.x & 0xffff), ieh = (int16_t)(x.x >> 16);
    const int isw = (int16_t)(x.y & 0xffff), iew = (int16_t)(x.y >> 16);
    l1 = min(we, max(ws, isw + 1));        // [ws, l1) left of / on the hole's left edge
    r0 = max(l1, min(we, iew));            // [r0, we) on / right of its right edge
    t1 = min(he, max(hs, ish + 1));        // rows [hs, t1) above / on its top edge
    b0 = max(t1, min(he, ieh));            // rows [b0, he) on / below its bottom edge
  };
  auto scan_item = [&](uint32_t item) {
    const int g = item & 1, bin = (item >> 1) & 63, rpos = item >> 7;
    const int r = __ldg(order + rpos);
    float acc[CB];
#pragma unroll
    for (int k = 0; k < CB; ++k) acc[k] = 0.f;
    scan_holed<CB, false>(nullptr, acc, sbase, pitch, p.P, __ldg(bins + 2 * ((int64_t)r * BINS + bin) + g),
                          __ldg(rects + 2 * (int64_t)r + g));
    float scale = 1.f;
    if (p.row_scale) scale = __fadd_rn(__ldg(p.row_scale + r), p.row_scale_bias);
    const int64_t o = ((int64_t)r * p.C + c0) * BINS + bin + (g + 1) * block_stride;
#pragma unroll
    for (int k = 0; k < CB; ++k)
      if (k < nc) __stcs(p.output + o + k * BINS, p.row_scale ? __fmul_rn(acc[k], scale) : acc[k]);
  };

  const int per_warp = (total + nw - 1) / nw;                      // a contiguous run of (proposal, bin) slots per warp
  const int f0 = wid * per_warp, f1 = min(total, f0 + per_warp);
#pragma unroll 1
  for (int g = 0; g < 2; ++g) {     // 0: ROI grid without the inner box (frame), 1: outer grid without the ROI (context)
    int qn = 0;
#pragma unroll 1
    for (int base = f0; base < f1; base += 32) {
      const int flat = base + lane;
      bool cut = false;
      uint32_t item = 0;
      if (flat < f1) {
        const int rpos = flat / BINS;
        const int bin = flat - rpos * BINS;
        const int r = __ldg(order + rpos);
        int hs, he, ws, we, l1, r0, t1, b0;
        geometry(r, bin, g, hs, he, ws, we, l1, r0, t1, b0);
        if (t1 < b0 && l1 < r0) {                                   // the bin meets the interior
          if (t1 == hs && b0 == he && l1 == ws && r0 == we) {       // ... and lies inside it: nothing is left, 0
            const int64_t o = ((int64_t)r * p.C + c0) * BINS + bin + (g + 1) * block_stride;
#pragma unroll
            for (int k = 0; k < CB; ++k)
              if (k < nc) __stcs(p.output + o + k * BINS, 0.f);
          } else {
            cut = true;
            item = ((uint32_t)rpos << 7) | ((uint32_t)bin << 1) | (uint32_t)g;
          }
        }
      }
      const unsigned bal = __ballot_sync(0xffffffffu, cut);
      if (cut) queue[qn + __popc(bal & ((1u << lane) - 1))] = item;
      qn += __popc(bal);
      __syncwarp();
      if (qn >= 32) {
        const uint32_t it = queue[lane];
        const uint32_t rest = lane < qn - 32 ? queue[32 + lane] : 0u;
        __syncwarp();
        if (lane < qn - 32) queue[lane] = rest;
        qn -= 32;
        __syncwarp();
        scan_item(it);
      }
    }
    if (qn > 0) {
      const uint32_t it = queue[lane];
      __syncwarp();
      if (lane < qn) scan_item(it);
    }
    __syncwarp();
  }
}

// backward: grad_input[b, c, argmax] += grad_output (ROILoopPool_cuda.cu:206-248)
__global__ void roi_pool_bwd_kernel(const float* __restrict__ grad_out, const float* __restrict__ rois,
                                    const int32_t* __restrict__ argmax, int64_t total, int64_t R, int N,
                                    int C, int HW, int BINS, float* __restrict__ grad_in) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int a = argmax[i];
    if (a < 0) continue;
    const int64_t nc = i / BINS;
    const int c = (int)(nc % C);
    const int64_t row = nc / C;
    int b = (int)rois[(row % R) * 5];
    b = min(max(b, 0), N - 1);
    atomicAdd(grad_in + ((int64_t)b * C + c) * HW + a, grad_out[i]);
  }
}

// backward of ROIAlign (torchvision roi_align backward, reached through detectron2's ROIAlign, poolers.py:169-182):
// every sample of every bin scatters grad * w_i / count to its four taps.  Same sample positions and weights as the
// forward above (each operation individually rounded); accumulation with fp32 atomics like torchvision's kernel.
__global__ void roi_align_bwd_kernel(const float* __restrict__ grad_out, const float* __restrict__ rois, int64_t total,
                                     int N, int C, int H, int W, int PH, int PW, float scale, int sampling_ratio,
                                     int aligned, float* __restrict__ grad_in) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int pw = (int)(i % PW), ph = (int)((i / PW) % PH);
    const int c = (int)((i / ((int64_t)PW * PH)) % C);
    const int64_t r = i / ((int64_t)PW * PH * C);
    const float* roi = rois + r * 5;
    int b = (int)roi[0];
    b = min(max(b, 0), N - 1);
    const float off = aligned ? 0.5f : 0.f;
    const float sw = __fsub_rn(__fmul_rn(roi[1], scale), off), sh = __fsub_rn(__fmul_rn(roi[2], scale), off);
    const float ew = __fsub_rn(__fmul_rn(roi[3], scale), off), eh = __fsub_rn(__fmul_rn(roi[4], scale), off);
    float rw = __fsub_rn(ew, sw), rh = __fsub_rn(eh, sh);
    if (!aligned) { rw = fmaxf(rw, 1.f); rh = fmaxf(rh, 1.f); }
    const float bh = __fdiv_rn(rh, (float)PH), bw = __fdiv_rn(rw, (float)PW);
    const int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(rh, (float)PH));
    const int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(rw, (float)PW));
    const float count = (float)max(gh * gw, 1);
    const float g = grad_out[i];
    float* gi = grad_in + ((int64_t)b * C + c) * H * W;
    for (int iy = 0; iy < gh; ++iy) {
      const float yy = __fadd_rn(__fadd_rn(sh, __fmul_rn((float)ph, bh)), __fdiv_rn(__fmul_rn((float)iy + .5f, bh), (float)gh));
      for (int ix = 0; ix < gw; ++ix) {
        const float xx = __fadd_rn(__fadd_rn(sw, __fmul_rn((float)pw, bw)), __fdiv_rn(__fmul_rn((float)ix + .5f, bw), (float)gw));
        float y = yy, x = xx;
        if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) continue;
        if (y <= 0) y = 0;
        if (x <= 0) x = 0;
        int yl = (int)y, xl = (int)x, yh, xh;
        if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else yh = yl + 1;
        if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
        const float ly = __fsub_rn(y, (float)yl), lx = __fsub_rn(x, (float)xl);
        const float hy = __fsub_rn(1.f, ly), hx = __fsub_rn(1.f, lx);
        atomicAdd(gi + yl * W + xl, __fdiv_rn(__fmul_rn(g, __fmul_rn(hy, hx)), count));
        atomicAdd(gi + yl * W + xh, __fdiv_rn(__fmul_rn(g, __fmul_rn(hy, lx)), count));
        atomicAdd(gi + yh * W + xl, __fdiv_rn(__fmul_rn(g, __fmul_rn(ly, hx)), count));
        atomicAdd(gi + yh * W + xh, __fdiv_rn(__fmul_rn(g, __fmul_rn(ly, lx)), count));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// Which values-only 7x7 kernel?  The block-max path pays ~10 plane rebuilds per (image, channel group),
// so it needs enough proposals per image to win: on a B200 the two cross at 1000-1200 proposals per
// image for every map / batch shape tried (tools/kbench_pool_sweep.py: 8 x 4000 proposals 1.45 vs 2.51 ms,
// 1 x 2000 on a 60x80 map 0.135 vs 0.166 ms, 8 x 500 0.63 vs 0.40 ms).  Test / bench hook:
// wsovod_b200_tune(WSOVOD_B200_TUNE_POOL_PATH, 1) forces the scan kernels, 2 the block-max path wherever it applies.
static bool pool_use_blockmax(int64_t N, int64_t R, int64_t C, bool with_argmax) {
  const int v = tune(TUNE_POOL_PATH);
  if (v == 1) return false;
  if (v == 2) return true;
  // with argmax a CTA owns two channels: one image of C = 512 is 256 CTAs = 1.7 waves of 148, where the scan kernel's
  // 128 four-channel CTAs are faster (c1: 0.199 vs 0.216 ms); from three waves on the planes win (c2: 2.5 vs 3.7 ms)
  if (with_argmax && N * ceil_div(C, 2) < 3 * kNumSMs) return false;
  return R >= 1200 * N;
}

struct PoolWs {
  int32_t* counts; int32_t* bidx; int32_t* order; int16_t* edges; float* alignp; uint2* bins; uint2* rects;
  void* pyr; void* asep; float* iboxes; size_t bytes;
};

// H, W > 0 (ROIAlign only): room for the separable tap tables of roi_align_sep.cu behind the common part
static PoolWs carve(void* ws, int mode, int64_t N, int64_t R, int PH, int PW, int64_t H = 0, int64_t W = 0) {
  PoolWs w;
  size_t off = 0;
  auto take = [&](size_t b) { size_t o = off; off += align_up(b, 256); return o; };
  char* base = (char*)ws;
  size_t o_counts = take(sizeof(int32_t) * (size_t)(N + 1));
  size_t o_bidx = take(sizeof(int32_t) * (size_t)R);
  size_t o_order = take(sizeof(int32_t) * (size_t)R);
  size_t ew = mode == MODE_LOOP ? 4 * (PH + PW) + 8 : 2 * (PH + PW);
  size_t o_edges = take(mode == MODE_ALIGN ? 0 : sizeof(int16_t) * ew * (size_t)R);
  size_t o_align = take(mode == MODE_ALIGN ? sizeof(float) * 8 * (size_t)R : 0);
  const bool seven = PH == 7 && PW == 7;
  size_t o_bins = take(seven && mode == MODE_POOL ? sizeof(uint2) * 49 * (size_t)R
                       : seven && mode == MODE_LOOP ? sizeof(uint2) * 98 * (size_t)R : 0);
  size_t o_rects = take(seven && mode == MODE_LOOP ? sizeof(uint2) * 2 * (size_t)R : 0);
  size_t o_pyr = take(seven && (mode == MODE_POOL || mode == MODE_LOOP) ? pool7_pyr_workspace(N, R) : 0);
  size_t o_iboxes = take(seven && mode == MODE_LOOP ? sizeof(float) * 10 * (size_t)R : 0);
  size_t o_asep = take(seven && mode == MODE_ALIGN && H > 0 && W > 0 ? align7_sep_workspace(R, H, W) : 0);
  w.counts = (int32_t*)(base + o_counts);
  w.bidx = (int32_t*)(base + o_bidx);
  w.order = (int32_t*)(base + o_order);
  w.edges = (int16_t*)(base + o_edges);
  w.alignp = (float*)(base + o_align);
  w.bins = (uint2*)(base + o_bins);
  w.rects = (uint2*)(base + o_rects);
  w.pyr = base + o_pyr;
  w.asep = base + o_asep;
  w.iboxes = (float*)(base + o_iboxes);
  w.bytes = off;
  return w;
}

template <int CB, int MODE, bool SMEM, bool ARG>
static int launch_plane(const PoolParams& p, int threads, size_t smem, cudaStream_t st) {
  auto kern = roi_plane_kernel<CB, MODE, SMEM, ARG>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  const int64_t grid = (int64_t)p.N * p.S * p.CG;
  kern<<<(unsigned)grid, threads, smem, st>>>(p);
  return after_launch();
}

template <int MODE, bool ARG>
static int dispatch_plane(PoolParams& p, int64_t R, cudaStream_t st) {
  const size_t plane = (size_t)p.H * p.W * sizeof(float);
  int cb = 0;
  if (MODE == MODE_LOOP) {
    // three accumulator sets: two channels per lane keep the kernel at 64 registers
    if (2 * plane <= (size_t)kMaxSmemOptin && p.C >= 2) cb = 2;
    else if (plane <= (size_t)kMaxSmemOptin) cb = 1;
  } else {
    if (4 * plane <= (size_t)kMaxSmemOptin && p.C >= 3) cb = 4;
    else if (2 * plane <= (size_t)kMaxSmemOptin && p.C >= 2) cb = 2;
    else if (plane <= (size_t)kMaxSmemOptin) cb = 1;
  }
  const int cbe = cb ? cb : 1;
  p.CG = (int)ceil_div(p.C, cbe);
  const size_t smem = cb ? cbe * plane : 0;
  // CTAs per SM the shared-memory footprint allows (<=4), threads so that one SM holds <=2048
  int per_sm = cb ? (int)std::min<size_t>(4, (size_t)kMaxSmemOptin / std::max<size_t>(smem + 1024, 1)) : 2;
  per_sm = max(per_sm, 1);
  const int threads = per_sm == 1 ? 1024 : 512;
  // proposal chunks per image: aim for >= 4 waves of CTAs, each with a few hundred outputs at least
  const int64_t slots = (int64_t)kNumSMs * per_sm;
  const int64_t base = (int64_t)p.N * p.CG;
  int64_t S = ceil_div(4 * slots, base);
  const int64_t avg = std::max<int64_t>(R / max(p.N, 1), 1);
  S = std::max<int64_t>(1, std::min<int64_t>(S, ceil_div(avg, 96)));
  p.S = (int)S;
  if ((int64_t)p.N * p.S * p.CG > 0x7fffffffLL) return WSOVOD_B200_ETOOBIG;
  switch (cb) {
    case 4:
      // never selected for MODE_LOOP (cb <= 2 there); instantiate the pooling flavour to keep ptxas quiet
      if (MODE == MODE_LOOP) return WSOVOD_B200_EINVAL;
      return launch_plane<4, MODE == MODE_LOOP ? MODE_POOL : MODE, true, ARG>(p, threads, smem, st);
    case 2: return launch_plane<2, MODE, true, ARG>(p, threads, smem, st);
    case 1: return launch_plane<1, MODE, true, ARG>(p, threads, smem, st);
    default: return launch_plane<1, MODE, false, ARG>(p, 512, 0, st);
  }
}

template <int CB>
static int launch_pool7(PoolParams& p, const uint2* bins, int64_t R, bool arg, cudaStream_t st) {
  const size_t smem = CB * (size_t)p.H * p.P * sizeof(float);
  p.CG = (int)ceil_div(p.C, CB);
  int per_sm = (int)std::min<size_t>(4, (size_t)kMaxSmemOptin / (smem + 1024));
  per_sm = std::max(per_sm, 1);
  const int threads = per_sm == 1 ? 1024 : 512;
  const int64_t slots = (int64_t)kNumSMs * per_sm;
  const int64_t base = (int64_t)p.N * p.CG;
  int64_t S = ceil_div(4 * slots, base);
  const int64_t avg = std::max<int64_t>(R / std::max(p.N, 1), 1);
  S = std::max<int64_t>(1, std::min<int64_t>(S, ceil_div(avg, 96)));
  p.S = (int)S;
  if ((int64_t)p.N * p.S * p.CG > 0x7fffffffLL) return WSOVOD_B200_ETOOBIG;
  auto kern = arg ? roi_pool7_kernel<CB, true> : roi_pool7_kernel<CB, false>;
  if (smem > 32 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  kern<<<(unsigned)((int64_t)p.N * p.S * p.CG), threads, smem, st>>>(p, bins);
  return after_launch();
}

constexpr size_t kLoopFixQueueBytes = 32 * 64 * sizeof(uint32_t);

template <int CB, bool FIX>
static int launch_loop7(PoolParams& p, const uint2* bins, const uint2* rects, int64_t R, cudaStream_t st) {
  // FIX: + the per-warp queues of cut bins (32 warps x 64 entries)
  const size_t smem = CB * (size_t)p.H * p.P * sizeof(float) + (FIX ? kLoopFixQueueBytes : 0);
  p.CG = (int)ceil_div(p.C, CB);
  int per_sm = (int)std::min<size_t>(4, (size_t)kMaxSmemOptin / (smem + 1024));
  per_sm = std::max(per_sm, 1);
  const int threads = per_sm == 1 ? 1024 : 512;
  const int64_t slots = (int64_t)kNumSMs * per_sm;
  int64_t S = ceil_div(4 * slots, (int64_t)p.N * p.CG);
  const int64_t avg = std::max<int64_t>(R / std::max(p.N, 1), 1);
  S = std::max<int64_t>(1, std::min<int64_t>(S, ceil_div(avg, 96)));
  p.S = (int)S;
  if ((int64_t)p.N * p.S * p.CG > 0x7fffffffLL) return WSOVOD_B200_ETOOBIG;
  auto kern = FIX ? roi_loop7_fix_kernel<CB> : roi_loop7_kernel<CB>;
  if (smem > 32 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  kern<<<(unsigned)((int64_t)p.N * p.S * p.CG), threads, smem, st>>>(p, bins, rects);
  return after_launch();
}

static int pool_common(int mode, const float* input, int64_t N, int64_t C, int64_t H, int64_t W,
                       const float* rois, int64_t R, float scale, int PH, int PW, int sampling_ratio,
                       int aligned, const float* row_scale, float row_scale_bias, float* output,
                       int32_t* argmax, void* workspace, size_t ws_bytes, void* stream) {
  if (N < 0 || C < 0 || H < 0 || W < 0 || R < 0 || PH <= 0 || PW <= 0) return WSOVOD_B200_EINVAL;
  if (R == 0 || C == 0) return 0;
  if (!input || !rois || !output || N == 0 || H == 0 || W == 0) return WSOVOD_B200_EINVAL;
  if (H > 32767 || W > 32767 || H * W >= (1LL << 31) || PH > 64 || PW > 64 || N > 65535 ||
      R >= (1LL << 31) / (PH * PW) || C >= (1 << 30))
    return WSOVOD_B200_ETOOBIG;
  PoolWs w = carve(nullptr, mode, N, R, PH, PW);
  if (!workspace || ws_bytes < w.bytes) return WSOVOD_B200_EWORKSPACE;
  w = carve(workspace, mode, N, R, PH, PW);
  cudaStream_t st = (cudaStream_t)stream;
  // 7x7 max-pool (+ argmax): block-max planes (roi_pool_pyr.cu) when the padded plane fits shared memory
  if (mode == MODE_POOL && PH == 7 && PW == 7 && pool_use_blockmax(N, R, C, argmax != nullptr) && pool7_pyr_cb(C, H, W, R, argmax != nullptr))
    return pool7_pyr(input, N, C, H, W, rois, R, scale, row_scale, row_scale_bias, output, argmax, w.pyr, st);
  cudaError_t e = cudaMemsetAsync(w.counts, 0, sizeof(int32_t) * (size_t)(N + 1), st);
  if (e != cudaSuccess) return (int)e;
  const int pt = 128;
  const unsigned pg = (unsigned)ceil_div(R, pt);
  // specialised kernel: 7x7 bins, four interleaved planes fit shared memory, at least 3 channels
  const size_t plane_bytes = (size_t)H * (W | 1) * sizeof(float);     // the 7x7 scan kernels stage rows at an odd pitch
  const bool fast7 = mode == MODE_POOL && PH == 7 && PW == 7 && C >= 2 && 2 * plane_bytes <= (size_t)kMaxSmemOptin;
  const bool fast7_cb4 = fast7 && C >= 3 && 4 * plane_bytes <= (size_t)kMaxSmemOptin;
  // 3-way fast path: values only (an argmax request -- trainable backbone -- takes the generic kernel)
  const bool loop7 = mode == MODE_LOOP && PH == 7 && PW == 7 && !argmax && C >= 2 && 2 * plane_bytes <= (size_t)kMaxSmemOptin;
  const bool loop7_cb4 = loop7 && C >= 3 && 4 * plane_bytes <= (size_t)kMaxSmemOptin;
  // ... and on the block-max planes where they pay (same crossover as the plain max-pool): three pooling passes with
  // floor 0 (ROI grid twice, outer grid once) + roi_loop7_fix_kernel for the bins the excluded interior touches
  const int fix_cb = (C >= 3 && 4 * plane_bytes + kLoopFixQueueBytes <= (size_t)kMaxSmemOptin) ? 4
                     : (2 * plane_bytes + kLoopFixQueueBytes <= (size_t)kMaxSmemOptin) ? 2 : 0;
  // measured (tools/kbench_loop_pool.py, after the scan kernel's loops were made convergent): 8 x 4000 proposals 8.8 vs
  // 7.7 ms, 1 x 5000 on a 100x152 map 2.64 vs 2.19 ms, 1 x 2000 on a 60x80 map 0.62 vs 0.36 ms -- the scan kernel wins
  // everywhere (the fix-up's partial-sector stores are DRAM read-modify-writes), so this path runs only when forced
  // (wsovod_b200_tune(WSOVOD_B200_TUNE_POOL_PATH, 2): tests, tools/kbench_loop_pool.py)
  const bool loop7_pyr = loop7 && fix_cb && tune(TUNE_POOL_PATH) == 2 && pool7_pyr_cb(C, H, W, R, false);
  if (fast7)
    roi_bins7_kernel<<<(unsigned)ceil_div(R * 49, 256), 256, 0, st>>>(rois, R, (int)N, (int)H, (int)W, scale, w.bidx, w.counts, w.bins);
  else if (loop7)
    roi_loopbins7_kernel<<<(unsigned)ceil_div(R * 49, 256), 256, 0, st>>>(rois, R, (int)N, (int)H, (int)W, scale, w.bidx, w.counts, w.bins, w.rects,
                                                                          loop7_pyr ? w.iboxes : nullptr);
  else if (mode == MODE_POOL)
    roi_prepare_kernel<MODE_POOL><<<pg, pt, 0, st>>>(rois, R, (int)N, (int)H, (int)W, scale, PH, PW, 0, 0, w.bidx, w.counts, w.edges, w.alignp, fast7 ? w.bins : nullptr);
  else if (mode == MODE_LOOP)
    roi_prepare_kernel<MODE_LOOP><<<pg, pt, 0, st>>>(rois, R, (int)N, (int)H, (int)W, scale, PH, PW, 0, 0, w.bidx, w.counts, w.edges, w.alignp, nullptr);
  else
    roi_prepare_kernel<MODE_ALIGN><<<pg, pt, 0, st>>>(rois, R, (int)N, (int)H, (int)W, scale, PH, PW, sampling_ratio, aligned, w.bidx, w.counts, w.edges, w.alignp, nullptr);
  int rc = after_launch();
  if (rc) return rc;
  roi_order_kernel<<<(unsigned)N, 256, 0, st>>>(w.bidx, w.counts, R, (int)N, w.order);
  rc = after_launch();
  if (rc) return rc;

  PoolParams p;
  p.input = input; p.rois = rois; p.row_scale = row_scale; p.row_scale_bias = row_scale_bias;
  p.output = output; p.argmax = argmax;
  p.counts = w.counts; p.order = w.order; p.edges = w.edges; p.alignp = w.alignp;
  p.N = (int)N; p.C = (int)C; p.H = (int)H; p.W = (int)W; p.R = R; p.PH = PH; p.PW = PW;
  p.CG = 0; p.S = 1; p.sampling_ratio = sampling_ratio; p.aligned = aligned; p.P = (int)(W | 1);
  // ROIAlign 7x7 with the adaptive grid: separable tap tables when the caller's workspace has room for them
  // (wsovod_b200_roi_align_workspace_hw) and the scan kernels are not forced
  if (mode == MODE_ALIGN && PH == 7 && PW == 7 && sampling_ratio <= 0 && tune(TUNE_POOL_PATH) != 1 &&
      align7_sep_cb(C, H, W) && ws_bytes >= carve(nullptr, mode, N, R, PH, PW, H, W).bytes) {
    const PoolWs w2 = carve(workspace, mode, N, R, PH, PW, H, W);
    return align7_sep(input, N, C, H, W, R, w.counts, w.order, w.alignp, row_scale, row_scale_bias, output, w2.asep, st);
  }
  if (loop7_pyr) {
    const int64_t block_stride = R * C * (int64_t)49;
    for (int s = 0; s < 3; ++s) {
      rc = pool7_pyr(input, N, C, H, W, w.iboxes + (s == 2 ? 5 * R : 0), R, 1.0f, row_scale, row_scale_bias, output + s * block_stride,
                     nullptr, w.pyr, st, 0.f);
      if (rc) return rc;
    }
    return fix_cb == 4 ? launch_loop7<4, true>(p, w.bins, w.rects, R, st) : launch_loop7<2, true>(p, w.bins, w.rects, R, st);
  }
  if (loop7) return loop7_cb4 ? launch_loop7<4, false>(p, w.bins, w.rects, R, st) : launch_loop7<2, false>(p, w.bins, w.rects, R, st);
  if (fast7) return fast7_cb4 ? launch_pool7<4>(p, w.bins, R, argmax != nullptr, st)
                             : launch_pool7<2>(p, w.bins, R, argmax != nullptr, st);
  if (mode == MODE_POOL) return argmax ? dispatch_plane<MODE_POOL, true>(p, R, st) : dispatch_plane<MODE_POOL, false>(p, R, st);
  if (mode == MODE_LOOP) return argmax ? dispatch_plane<MODE_LOOP, true>(p, R, st) : dispatch_plane<MODE_LOOP, false>(p, R, st);
  return dispatch_plane<MODE_ALIGN, false>(p, R, st);
}

static int bwd_common(const float* grad_output, const float* rois, const int32_t* argmax, int64_t rows,
                      int64_t R, int64_t N, int64_t C, int64_t H, int64_t W, int PH, int PW,
                      float* grad_input, void* stream) {
  if (rows < 0 || R < 0 || N <= 0 || C < 0 || H <= 0 || W <= 0 || PH <= 0 || PW <= 0) return WSOVOD_B200_EINVAL;
  if (rows == 0 || C == 0) return 0;
  if (!grad_output || !rois || !argmax || !grad_input) return WSOVOD_B200_EINVAL;
  if (H * W >= (1LL << 31)) return WSOVOD_B200_ETOOBIG;
  const int64_t total = rows * C * PH * PW;
  const int threads = 256;
  const int64_t grid = std::min<int64_t>(ceil_div(total, threads), (int64_t)kNumSMs * 32);
  roi_pool_bwd_kernel<<<(unsigned)grid, threads, 0, (cudaStream_t)stream>>>(
      grad_output, rois, argmax, total, R, (int)N, (int)C, (int)(H * W), PH * PW, grad_input);
  return after_launch();
}

}  // namespace wsovod

using namespace wsovod;

WSOVOD_API size_t wsovod_b200_roi_pool_workspace(int64_t N, int64_t R, int PH, int PW) {
  return carve(nullptr, MODE_POOL, N, R, PH, PW).bytes;
}
WSOVOD_API size_t wsovod_b200_roi_loop_pool_workspace(int64_t N, int64_t R, int PH, int PW) {
  return carve(nullptr, MODE_LOOP, N, R, PH, PW).bytes;
}
WSOVOD_API size_t wsovod_b200_roi_align_workspace(int64_t N, int64_t R, int PH, int PW) {
  return carve(nullptr, MODE_ALIGN, N, R, PH, PW).bytes;
}

WSOVOD_API size_t wsovod_b200_roi_align_workspace_hw(int64_t N, int64_t R, int PH, int PW, int64_t H, int64_t W) {
  if (H <= 0 || W <= 0 || H > 32767 || W > 32767) return carve(nullptr, MODE_ALIGN, N, R, PH, PW).bytes;
  return carve(nullptr, MODE_ALIGN, N, R, PH, PW, H, W).bytes;
}

WSOVOD_API int wsovod_b200_roi_pool_fwd(const float* input, int64_t N, int64_t C, int64_t H, int64_t W,
                                        const float* rois, int64_t R, float spatial_scale, int PH,
                                        int PW, const float* row_scale, float row_scale_bias,
                                        float* output, int32_t* argmax, void* workspace,
                                        size_t workspace_bytes, void* stream) {
  return pool_common(MODE_POOL, input, N, C, H, W, rois, R, spatial_scale, PH, PW, 0, 0, row_scale,
                     row_scale_bias, output, argmax, workspace, workspace_bytes, stream);
}

WSOVOD_API int wsovod_b200_roi_loop_pool_fwd(const float* input, int64_t N, int64_t C, int64_t H,
                                             int64_t W, const float* rois, int64_t R,
                                             float spatial_scale, int PH, int PW,
                                             const float* row_scale, float row_scale_bias,
                                             float* output, int32_t* argmax, void* workspace,
                                             size_t workspace_bytes, void* stream) {
  return pool_common(MODE_LOOP, input, N, C, H, W, rois, R, spatial_scale, PH, PW, 0, 0, row_scale,
                     row_scale_bias, output, argmax, workspace, workspace_bytes, stream);
}

WSOVOD_API int wsovod_b200_roi_align_fwd(const float* input, int64_t N, int64_t C, int64_t H, int64_t W,
                                         const float* rois, int64_t R, float spatial_scale, int PH,
                                         int PW, int sampling_ratio, int aligned,
                                         const float* row_scale, float row_scale_bias, float* output,
                                         void* workspace, size_t workspace_bytes, void* stream) {
  return pool_common(MODE_ALIGN, input, N, C, H, W, rois, R, spatial_scale, PH, PW, sampling_ratio,
                     aligned, row_scale, row_scale_bias, output, nullptr, workspace, workspace_bytes,
                     stream);
}

WSOVOD_API int wsovod_b200_roi_align_bwd(const float* grad_output, const float* rois, int64_t R, int64_t N, int64_t C,
                                         int64_t H, int64_t W, float spatial_scale, int PH, int PW, int sampling_ratio,
                                         int aligned, float* grad_input, void* stream) {
  if (R < 0 || N <= 0 || C < 0 || H <= 0 || W <= 0 || PH <= 0 || PW <= 0) return WSOVOD_B200_EINVAL;
  const int64_t total = R * C * PH * PW;
  if (total == 0) return 0;
  if (!grad_output || !rois || !grad_input) return WSOVOD_B200_EINVAL;
  if (H > 32767 || W > 32767 || N >= (1LL << 31) || C >= (1LL << 31)) return WSOVOD_B200_ETOOBIG;
  const int64_t blocks = std::min<int64_t>(ceil_div(total, 256), (int64_t)kNumSMs * 32);
  roi_align_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(grad_output, rois, total, (int)N, (int)C, (int)H, (int)W,
                                                                           PH, PW, spatial_scale, sampling_ratio, aligned, grad_input);
  return after_launch();
}

WSOVOD_API int wsovod_b200_roi_pool_bwd(const float* grad_output, const float* rois,
                                        const int32_t* argmax, int64_t R, int64_t N, int64_t C,
                                        int64_t H, int64_t W, int PH, int PW, float* grad_input,
                                        void* stream) {
  return bwd_common(grad_output, rois, argmax, R, R, N, C, H, W, PH, PW, grad_input, stream);
}

WSOVOD_API int wsovod_b200_roi_loop_pool_bwd(const float* grad_output, const float* rois,
                                             const int32_t* argmax, int64_t R, int64_t N, int64_t C,
                                             int64_t H, int64_t W, int PH, int PW, float* grad_input,
                                             void* stream) {
  return bwd_common(grad_output, rois, argmax, 3 * R, R, N, C, H, W, PH, PW, grad_input, stream);
}
