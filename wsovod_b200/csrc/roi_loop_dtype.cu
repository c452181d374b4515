// roi_loop_dtype.cu -- the 3-way ROILoopPool for the reference's other two dtypes, half and double
// (ROILoopPool_cuda.cu:294,364 dispatch AT_DISPATCH_FLOATING_TYPES_AND_HALF; fp32 lives in roi_pool.cu), forward with
// argmax and backward.  fp32 is instantiated too: tests run it against the fp32 kernels of roi_pool.cu.
//
// The reference's kernel is one template over T, so its BOX arithmetic changes with the dtype: products and quotients
// of two T values are rounded to T (c10::Half: fp32 operation, then round to half), mixed float x T products stay in
// float (double for T = double), the clamp bound `T(1.0 * width / spatial_scale)` is rounded to T, the bin size is a
// T quotient and `floor(T(ph) * bin)` a T product.  `Arith<T>` below states those rules once per dtype; everything
// derived from them is integer data (bin edges, inner / outer boxes), computed once per proposal by a prologue into the
// same int16 table the fp32 kernels use.  The VALUE side has no arithmetic at all -- maxima starting at 0 with a strict
// `>` (ROILoopPool_cuda.cu:107-133) -- so it runs on the library's usual decomposition: a CTA stages CB planes of
// (image, channel group) in shared memory (half is widened to fp32 there, exactly) and walks all proposals of the
// image, lanes = consecutive (proposal, bin) outputs.
#include <cuda_fp16.h>

#include "common.cuh"

#include <algorithm>

namespace wsovod {

// roi_pool.cu
int launch_roi_order(const int32_t* bidx, const int32_t* counts, int64_t R, int N, int32_t* order, cudaStream_t st);

namespace {

__device__ __forceinline__ int sat_i(float v) { return (int)fminf(fmaxf(v, -1.0e6f), 1.0e6f); }
__device__ __forceinline__ int sat_i(double v) { return (int)fmin(fmax(v, -1.0e6), 1.0e6); }
__device__ __forceinline__ float hrn(float v) { return __half2float(__float2half_rn(v)); }   // round to half, keep as float

template <typename T> struct Arith;

template <> struct Arith<float> {
  using S = float;    // storage type of input / rois / output
  using E = float;    // element type of the staged planes
  using B = float;    // type of the bin size
  __device__ static float f(S v) { return v; }
  __device__ static E elem(S v) { return v; }
  __device__ static S store(E v) { return v; }
  __device__ static double bound(int extent, S s) { return (double)(float)(1.0 * extent / s); }
  __device__ static int round_ts(S v, S s) { return sat_i(roundf(v * s)); }
  __device__ static int round_fs(float v, S s) { return sat_i(roundf(v * s)); }
  __device__ static B bin(int len, int P) { return __fdiv_rn((float)len, (float)P); }
  __device__ static int lo(int p, B b) { return (int)floorf(__fmul_rn((float)p, b)); }
  __device__ static int hi(int p, B b) { return (int)ceilf(__fmul_rn((float)p, b)); }
};

template <> struct Arith<__half> {
  using S = __half;
  using E = float;
  using B = float;    // a half value held in a float
  __device__ static float f(S v) { return __half2float(v); }
  __device__ static E elem(S v) { return __half2float(v); }
  __device__ static S store(E v) { return __float2half_rn(v); }   // exact: v is one of the staged half values or 0
  __device__ static double bound(int extent, S s) { return (double)hrn((float)(1.0 * extent / (double)__half2float(s))); }
  __device__ static int round_ts(S v, S s) { return sat_i(roundf(hrn(__fmul_rn(__half2float(v), __half2float(s))))); }
  __device__ static int round_fs(float v, S s) { return sat_i(roundf(__fmul_rn(v, __half2float(s)))); }
  __device__ static B bin(int len, int P) { return hrn(__fdiv_rn(hrn((float)len), hrn((float)P))); }
  __device__ static int lo(int p, B b) { return (int)floorf(hrn(__fmul_rn(hrn((float)p), b))); }
  __device__ static int hi(int p, B b) { return (int)ceilf(hrn(__fmul_rn(hrn((float)p), b))); }
};

template <> struct Arith<double> {
  using S = double;
  using E = double;
  using B = double;
  __device__ static float f(S v) { return (float)v; }
  __device__ static E elem(S v) { return v; }
  __device__ static S store(E v) { return v; }
  __device__ static double bound(int extent, S s) { return 1.0 * extent / s; }
  __device__ static int round_ts(S v, S s) { return sat_i(round(__dmul_rn(v, s))); }
  __device__ static int round_fs(float v, S s) { return sat_i(round(__dmul_rn((double)v, s))); }
  __device__ static B bin(int len, int P) { return __ddiv_rn((double)len, (double)P); }
  __device__ static int lo(int p, B b) { return (int)floor(__dmul_rn((double)p, b)); }
  __device__ static int hi(int p, B b) { return (int)ceil(__dmul_rn((double)p, b)); }
};

__device__ __forceinline__ float clampd(float v, double hi) { return (float)fmin(fmax((double)v, 0.0), hi); }

template <typename T>
__device__ __forceinline__ void edges_of(int16_t* e, int rsh, int rsw, int reh, int rew, int PH, int PW, int H, int W) {
  using A = Arith<T>;
  const int rw = max(rew - rsw + 1, 1), rh = max(reh - rsh + 1, 1);
  const typename A::B bh = A::bin(rh, PH), bw = A::bin(rw, PW);
  for (int ph = 0; ph < PH; ++ph) {
    e[ph] = (int16_t)min(max(A::lo(ph, bh) + rsh, 0), H);
    e[PH + ph] = (int16_t)min(max(A::hi(ph + 1, bh) + rsh, 0), H);
  }
  for (int pw = 0; pw < PW; ++pw) {
    e[2 * PH + pw] = (int16_t)min(max(A::lo(pw, bw) + rsw, 0), W);
    e[2 * PH + PW + pw] = (int16_t)min(max(A::hi(pw + 1, bw) + rsw, 0), W);
  }
}

// prologue: one thread per proposal.  Table layout of roi_prepare_kernel<MODE_LOOP> (roi_pool.cu): grid of the ROI,
// grid of the outer box, inner box, the ROI itself.
template <typename T>
__global__ void loop_prepare_kernel(const T* __restrict__ rois, int64_t R, int N, int H, int W, T scale, int PH, int PW,
                                    int32_t* __restrict__ bidx, int32_t* __restrict__ counts, int16_t* __restrict__ edges) {
  using A = Arith<T>;
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const T* roi = rois + r * 5;
  int b = (int)A::f(roi[0]);
  b = min(max(b, 0), N - 1);
  bidx[r] = b;
  atomicAdd(&counts[b], 1);
  if (r > 0) {
    int pb = (int)A::f(rois[(r - 1) * 5]);
    pb = min(max(pb, 0), N - 1);
    if (pb > b) atomicOr(&counts[N], 1);
  }
  // ROILoopPool_cuda.cu:34-74, `float` locals whatever T is
  const float x1 = A::f(roi[1]), y1 = A::f(roi[2]), x2 = A::f(roi[3]), y2 = A::f(roi[4]);
  const float ratio = 1.8f;
  float rw_ = x2 - x1, rh_ = y2 - y1;
  float iw = rw_ / ratio, ih = rh_ / ratio;
  float ow = rw_ * ratio, oh = rh_ * ratio;
  float irw = rw_ - iw, irh = rh_ - ih;
  float orw = ow - rw_, orh = oh - rh_;
  float x1i = x1 + irw / 2, y1i = y1 + irh / 2, x2i = x2 - irw / 2, y2i = y2 - irh / 2;
  float x1o = x1 - orw / 2, y1o = y1 - orh / 2, x2o = x2 + orw / 2, y2o = y2 + orh / 2;
  const double xmax = A::bound(W, scale), ymax = A::bound(H, scale);
  x1i = clampd(x1i, xmax); y1i = clampd(y1i, ymax); x2i = clampd(x2i, xmax); y2i = clampd(y2i, ymax);
  x1o = clampd(x1o, xmax); y1o = clampd(y1o, ymax); x2o = clampd(x2o, xmax); y2o = clampd(y2o, ymax);
  const int EW = 4 * (PH + PW) + 8;
  int16_t* e = edges + r * EW;
  const int rsw = A::round_ts(roi[1], scale), rsh = A::round_ts(roi[2], scale);
  const int rew = A::round_ts(roi[3], scale), reh = A::round_ts(roi[4], scale);
  edges_of<T>(e, rsh, rsw, reh, rew, PH, PW, H, W);
  const int osw = A::round_fs(x1o, scale), osh = A::round_fs(y1o, scale);
  const int oew = A::round_fs(x2o, scale), oeh = A::round_fs(y2o, scale);
  edges_of<T>(e + 2 * (PH + PW), osh, osw, oeh, oew, PH, PW, H, W);
  int16_t* q = e + 4 * (PH + PW);
  auto sat = [](int v) { return (int16_t)min(max(v, -32768), 32767); };
  q[0] = sat(A::round_fs(y1i, scale)); q[1] = sat(A::round_fs(y2i, scale));
  q[2] = sat(A::round_fs(x1i, scale)); q[3] = sat(A::round_fs(x2i, scale));
  q[4] = sat(rsh); q[5] = sat(reh); q[6] = sat(rsw); q[7] = sat(rew);
}

struct LoopParams {
  const void* input;
  void* output;
  int32_t* argmax;
  const int32_t* counts;
  const int32_t* order;
  const int16_t* edges;
  int32_t N, C, H, W, PH, PW, CG, S;
  int64_t R;
};

template <typename E, int CB> struct __align__(sizeof(E) * CB) Cell { E v[CB]; };

// CB channels per CTA staged in shared memory; SMEM = false: one channel, cells read from global memory (maps whose
// plane does not fit).  The scan is the reference's: row-major, strict `>` against maxima that start at 0, so the
// argmax is the first maximal cell and -1 when nothing exceeds 0.
template <typename T, int CB, bool SMEM>
__global__ void __launch_bounds__(512) loop_plane_kernel(const LoopParams p) {
  using A = Arith<T>;
  using E = typename A::E;
  using C_ = Cell<E, CB>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int H = p.H, W = p.W, HW = H * W;
  const int PH = p.PH, PW = p.PW, BINS = PH * PW;
  const int bid = blockIdx.x;
  const int cg = bid % p.CG;
  const int s = (bid / p.CG) % p.S;
  const int n = bid / (p.CG * p.S);
  const int c0 = cg * CB;
  const int nc = min(CB, p.C - c0);
  int start = 0;
  for (int m = 0; m < n; ++m) start += p.counts[m];
  const int cnt = p.counts[n];
  const int per = (cnt + p.S - 1) / p.S;
  const int pos0 = s * per;
  const int nroi = min(cnt, pos0 + per) - pos0;
  if (nroi <= 0) return;
  const T* src = reinterpret_cast<const T*>(p.input) + ((int64_t)n * p.C + c0) * HW;
  C_* plane = reinterpret_cast<C_*>(smem_raw);
  if (SMEM) {
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
      C_ c;
#pragma unroll
      for (int k = 0; k < CB; ++k) c.v[k] = k < nc ? A::elem(src[(int64_t)k * HW + i]) : (E)0;
      plane[i] = c;
    }
    __syncthreads();
  }
  auto cell = [&](int i) {
    if (SMEM) return plane[i];
    C_ c;
    c.v[0] = A::elem(src[i]);
    return c;
  };
  T* out = reinterpret_cast<T*>(p.output);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int total = nroi * BINS;
  const int EW = 4 * (PH + PW) + 8;
  const int64_t block_stride = p.R * (int64_t)p.C * BINS;   // roi | frame | context
  for (int flat = wid * 32 + lane; flat < total; flat += nw * 32) {
    const int rpos = flat / BINS;
    const int bin = flat - rpos * BINS;
    const int ph = bin / PW, pw = bin - ph * PW;
    const int r = __ldg(p.order + start + pos0 + rpos);
    const int16_t* e = p.edges + (int64_t)r * EW;
    const int16_t* q = e + 4 * (PH + PW);
    const int64_t obase = ((int64_t)r * p.C + c0) * BINS + bin;
    {  // roi and frame on the ROI's own grid (ROILoopPool_cuda.cu:76-142)
      const int hs = e[ph], he = e[PH + ph], ws = e[2 * PH + pw], we = e[2 * PH + PW + pw];
      const int ish = q[0], ieh = q[1], isw = q[2], iew = q[3];
      E m[CB], mf[CB];
      int mi[CB], mfi[CB];
#pragma unroll
      for (int k = 0; k < CB; ++k) { m[k] = 0; mf[k] = 0; mi[k] = -1; mfi[k] = -1; }
      for (int h = hs; h < he; ++h) {
        const bool in_h = h > ish && h < ieh;
        for (int w = ws; w < we; ++w) {
          const C_ c = cell(h * W + w);
          const bool inside = in_h && (w > isw && w < iew);
#pragma unroll
          for (int k = 0; k < CB; ++k) {
            if (c.v[k] > m[k]) { m[k] = c.v[k]; mi[k] = h * W + w; }
            if (!inside && c.v[k] > mf[k]) { mf[k] = c.v[k]; mfi[k] = h * W + w; }
          }
        }
      }
#pragma unroll
      for (int k = 0; k < CB; ++k)
        if (k < nc) {
          const int64_t o = obase + (int64_t)k * BINS;
          out[o] = A::store(m[k]);
          out[o + block_stride] = A::store(mf[k]);
          if (p.argmax) { p.argmax[o] = mi[k]; p.argmax[o + block_stride] = mfi[k]; }
        }
    }
    {  // context on the outer box's grid without the ROI's interior (ROILoopPool_cuda.cu:144-202)
      const int16_t* e2 = e + 2 * (PH + PW);
      const int hs = e2[ph], he = e2[PH + ph], ws = e2[2 * PH + pw], we = e2[2 * PH + PW + pw];
      const int ish = q[4], ieh = q[5], isw = q[6], iew = q[7];
      E mc[CB];
      int mci[CB];
#pragma unroll
      for (int k = 0; k < CB; ++k) { mc[k] = 0; mci[k] = -1; }
      for (int h = hs; h < he; ++h) {
        const bool in_h = h > ish && h < ieh;
        for (int w = ws; w < we; ++w) {
          if (in_h && (w > isw && w < iew)) continue;
          const C_ c = cell(h * W + w);
#pragma unroll
          for (int k = 0; k < CB; ++k)
            if (c.v[k] > mc[k]) { mc[k] = c.v[k]; mci[k] = h * W + w; }
        }
      }
#pragma unroll
      for (int k = 0; k < CB; ++k)
        if (k < nc) {
          const int64_t o = obase + (int64_t)k * BINS + 2 * block_stride;
          out[o] = A::store(mc[k]);
          if (p.argmax) p.argmax[o] = mci[k];
        }
    }
  }
}

// backward: grad_input[b, c, argmax] += grad_output, accumulated in T like the reference's atomicAdd<T>
// (ROILoopPool_cuda.cu:206-248)
template <typename T>
__global__ void loop_bwd_kernel(const T* __restrict__ grad_out, const T* __restrict__ rois, const int32_t* __restrict__ argmax,
                                int64_t total, int64_t R, int N, int C, int HW, int BINS, T* __restrict__ grad_in) {
  using A = Arith<T>;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int a = argmax[i];
    if (a < 0) continue;
    const int64_t nc = i / BINS;
    const int c = (int)(nc % C);
    const int64_t row = nc / C;
    int b = (int)A::f(rois[(row % R) * 5]);
    b = min(max(b, 0), N - 1);
    atomicAdd(grad_in + ((int64_t)b * C + c) * HW + a, grad_out[i]);
  }
}

struct LoopWs { int32_t* counts; int32_t* bidx; int32_t* order; int16_t* edges; size_t bytes; };

LoopWs carve_loop(void* ws, int64_t N, int64_t R, int PH, int PW) {
  LoopWs w;
  size_t off = 0;
  auto take = [&](size_t b) { size_t o = off; off += align_up(b, 256); return o; };
  char* base = (char*)ws;
  const size_t o_counts = take(sizeof(int32_t) * (size_t)(N + 1));
  const size_t o_bidx = take(sizeof(int32_t) * (size_t)R);
  const size_t o_order = take(sizeof(int32_t) * (size_t)R);
  const size_t o_edges = take(sizeof(int16_t) * (size_t)(4 * (PH + PW) + 8) * (size_t)R);
  w.counts = (int32_t*)(base + o_counts);
  w.bidx = (int32_t*)(base + o_bidx);
  w.order = (int32_t*)(base + o_order);
  w.edges = (int16_t*)(base + o_edges);
  w.bytes = off;
  return w;
}

template <typename T, int CB, bool SMEM>
int launch_loop_plane(LoopParams& p, int64_t R, cudaStream_t st) {
  using E = typename Arith<T>::E;
  const size_t smem = SMEM ? CB * (size_t)p.H * p.W * sizeof(E) : 0;
  p.CG = (int)ceil_div(p.C, CB);
  int per_sm = SMEM ? (int)std::min<size_t>(4, (size_t)kMaxSmemOptin / (smem + 1024)) : 4;
  per_sm = std::max(per_sm, 1);
  int64_t S = ceil_div(4 * (int64_t)kNumSMs * per_sm, (int64_t)p.N * p.CG);
  const int64_t avg = std::max<int64_t>(R / std::max(p.N, 1), 1);
  S = std::max<int64_t>(1, std::min<int64_t>(S, ceil_div(avg, 96)));
  p.S = (int)S;
  if ((int64_t)p.N * p.S * p.CG > 0x7fffffffLL) return WSOVOD_B200_ETOOBIG;
  auto kern = loop_plane_kernel<T, CB, SMEM>;
  if (smem > 32 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  kern<<<(unsigned)((int64_t)p.N * p.S * p.CG), 512, smem, st>>>(p);
  return after_launch();
}

template <typename T>
int loop_fwd_t(const void* input, int64_t N, int64_t C, int64_t H, int64_t W, const void* rois, int64_t R, T scale, int PH,
               int PW, void* output, int32_t* argmax, void* workspace, cudaStream_t st) {
  using E = typename Arith<T>::E;
  LoopWs w = carve_loop(workspace, N, R, PH, PW);
  cudaError_t e = cudaMemsetAsync(w.counts, 0, sizeof(int32_t) * (size_t)(N + 1), st);
  if (e != cudaSuccess) return (int)e;
  loop_prepare_kernel<T><<<(unsigned)ceil_div(R, 128), 128, 0, st>>>((const T*)rois, R, (int)N, (int)H, (int)W, scale, PH, PW,
                                                                      w.bidx, w.counts, w.edges);
  int rc = after_launch();
  if (rc) return rc;
  rc = launch_roi_order(w.bidx, w.counts, R, (int)N, w.order, st);
  if (rc) return rc;
  LoopParams p;
  p.input = input; p.output = output; p.argmax = argmax; p.counts = w.counts; p.order = w.order; p.edges = w.edges;
  p.N = (int)N; p.C = (int)C; p.H = (int)H; p.W = (int)W; p.PH = PH; p.PW = PW; p.CG = 0; p.S = 1; p.R = R;
  const size_t plane = (size_t)H * W * sizeof(E);
  if (C >= 2 && 2 * plane <= (size_t)kMaxSmemOptin) return launch_loop_plane<T, 2, true>(p, R, st);
  if (plane <= (size_t)kMaxSmemOptin) return launch_loop_plane<T, 1, true>(p, R, st);
  return launch_loop_plane<T, 1, false>(p, R, st);
}

template <typename T>
int loop_bwd_t(const void* grad_output, const void* rois, const int32_t* argmax, int64_t R, int64_t N, int64_t C, int64_t H,
               int64_t W, int PH, int PW, void* grad_input, cudaStream_t st) {
  const int64_t total = 3 * R * C * PH * PW;
  const int64_t grid = std::min<int64_t>(ceil_div(total, 256), (int64_t)kNumSMs * 32);
  loop_bwd_kernel<T><<<(unsigned)grid, 256, 0, st>>>((const T*)grad_output, (const T*)rois, argmax, total, R, (int)N, (int)C,
                                                    (int)(H * W), PH * PW, (T*)grad_input);
  return after_launch();
}

}  // namespace
}  // namespace wsovod

using namespace wsovod;

WSOVOD_API size_t wsovod_b200_roi_loop_pool_dtype_workspace(int64_t N, int64_t R, int PH, int PW) {
  return carve_loop(nullptr, N, R, PH, PW).bytes;
}

WSOVOD_API int wsovod_b200_roi_loop_pool_dtype_fwd(int dtype, const void* input, int64_t N, int64_t C, int64_t H, int64_t W,
                                                   const void* rois, int64_t R, float spatial_scale, int PH, int PW,
                                                   void* output, int32_t* argmax, void* workspace, size_t workspace_bytes,
                                                   void* stream) {
  if (N < 0 || C < 0 || H < 0 || W < 0 || R < 0 || PH <= 0 || PW <= 0) return WSOVOD_B200_EINVAL;
  if (dtype != WSOVOD_B200_F32 && dtype != WSOVOD_B200_F16 && dtype != WSOVOD_B200_F64) return WSOVOD_B200_EUNSUPPORTED;
  if (R == 0 || C == 0) return 0;
  if (!input || !rois || !output || N == 0 || H == 0 || W == 0) return WSOVOD_B200_EINVAL;
  if (H > 32767 || W > 32767 || H * W >= (1LL << 31) || PH > 64 || PW > 64 || N > 65535 ||
      R >= (1LL << 31) / (3 * PH * PW) || C >= (1 << 30))
    return WSOVOD_B200_ETOOBIG;
  if (!workspace || workspace_bytes < carve_loop(nullptr, N, R, PH, PW).bytes) return WSOVOD_B200_EWORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  // the reference hands `spatial_scale` to the kernel as T (ROILoopPool_cuda.cu:300)
  if (dtype == WSOVOD_B200_F16)
    return loop_fwd_t<__half>(input, N, C, H, W, rois, R, __float2half_rn(spatial_scale), PH, PW, output, argmax, workspace, st);
  if (dtype == WSOVOD_B200_F64)
    return loop_fwd_t<double>(input, N, C, H, W, rois, R, (double)spatial_scale, PH, PW, output, argmax, workspace, st);
  return loop_fwd_t<float>(input, N, C, H, W, rois, R, spatial_scale, PH, PW, output, argmax, workspace, st);
}

WSOVOD_API int wsovod_b200_roi_loop_pool_dtype_bwd(int dtype, const void* grad_output, const void* rois, const int32_t* argmax,
                                                   int64_t R, int64_t N, int64_t C, int64_t H, int64_t W, int PH, int PW,
                                                   void* grad_input, void* stream) {
  if (R < 0 || N <= 0 || C < 0 || H <= 0 || W <= 0 || PH <= 0 || PW <= 0) return WSOVOD_B200_EINVAL;
  if (dtype != WSOVOD_B200_F32 && dtype != WSOVOD_B200_F16 && dtype != WSOVOD_B200_F64) return WSOVOD_B200_EUNSUPPORTED;
  if (R == 0 || C == 0) return 0;
  if (!grad_output || !rois || !argmax || !grad_input) return WSOVOD_B200_EINVAL;
  if (H * W >= (1LL << 31)) return WSOVOD_B200_ETOOBIG;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == WSOVOD_B200_F16) return loop_bwd_t<__half>(grad_output, rois, argmax, R, N, C, H, W, PH, PW, grad_input, st);
  if (dtype == WSOVOD_B200_F64) return loop_bwd_t<double>(grad_output, rois, argmax, R, N, C, H, W, PH, PW, grad_input, st);
  return loop_bwd_t<float>(grad_output, rois, argmax, R, N, C, H, W, PH, PW, grad_input, st);
}
