// roi_align_sep.cu -- bilinear ROIAlign 7x7 with the adaptive sample grid (sampling_ratio = 0, detectron2's default,
// poolers.py:169-182 -> torchvision roi_align) from separable tap tables (align_sep.cuh).
//
// Same ownership as the pooling kernels (DESIGN.md "Kernel 1"): a CTA stages CB channel planes of one image in shared
// memory, channel-interleaved, and walks every proposal of the image; lanes are consecutive (proposal, bin) outputs,
// so a warp store is 32 consecutive floats of [R, C, 7, 7].  What changes is the work per bin.  The per-sample kernel
// (roi_plane_kernel<MODE_ALIGN>) recomputes each sample's coordinates and four weights and issues four LDS per sample
// (~36 loads and ~500 instructions for a 3 x 3 grid).  Here a prologue collapses the samples of a bin into one weight
// per footprint row and column (align_sep.cuh: 14 short lists per proposal, shared by all channel groups), and the
// bin is  sum_a WY[a] * (sum_b WX[b] * cell[a][b])  -- one LDS.128 and CB FMAs per footprint cell, ~16 cells for the
// same bin.  Column weights sit in registers four at a time; the row weight is one L1-resident load per row.
#include "align_sep.cuh"
#include "common.cuh"

#include <algorithm>

namespace wsovod {

struct AlignSepParams {
  const float* input;
  const float* row_scale;
  float row_scale_bias;
  float* output;
  const int32_t* counts;   // [N]   proposals per image
  const int32_t* order;    // [R]   proposal ids grouped by image
  const float* alignp;     // [R, 8] sw, sh, bw, bh, gh, gw, count (roi_prepare_kernel<MODE_ALIGN>)
  const uint2* hdr;        // [R, 14] list headers: rows 0..6, columns 7..13
  const float* wts;        // [R, capy + capx]
  int32_t capy, capx;
  int32_t N, C, H, W;
  int32_t CG, S;
  int32_t pitch;           // cells per staged row: W made odd, so that equal columns of different rows fall into different bank groups
};

// prologue: one thread per (proposal, axis)
__global__ void roi_align_tables_kernel(const float* __restrict__ alignp, int64_t R, int H, int W, int capy, int capx,
                                        uint2* __restrict__ hdr, float* __restrict__ wts) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * R) return;
  const int64_t r = t >> 1;
  const int axis = (int)(t & 1);   // 0 = rows (y), 1 = columns (x)
  const float* a = alignp + r * 8;
  asep::Hdr* h = reinterpret_cast<asep::Hdr*>(hdr + r * 14 + axis * 7);
  float* w = wts + r * (int64_t)(capy + capx) + (axis ? capy : 0);
  if (axis == 0) asep::axis_tables(a[1], a[3], __float_as_int(a[4]), 7, H, h, w, capy);
  else asep::axis_tables(a[0], a[2], __float_as_int(a[5]), 7, W, h, w, capx);
}

template <int CB> struct CellVec;
template <> struct CellVec<4> { using T = float4; };
template <> struct CellVec<2> { using T = float2; };

template <int CB> __device__ __forceinline__ void lds_cells(uint32_t addr, float* f);
template <> __device__ __forceinline__ void lds_cells<4>(uint32_t addr, float* f) {
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f[0]), "=f"(f[1]), "=f"(f[2]), "=f"(f[3]) : "r"(addr));
}
template <> __device__ __forceinline__ void lds_cells<2>(uint32_t addr, float* f) {
  asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(f[0]), "=f"(f[1]) : "r"(addr));
}

template <int CB>
__global__ void __launch_bounds__(1024, 1) roi_align7_sep_kernel(const AlignSepParams p) {
  using V = typename CellVec<CB>::T;
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
      const int y = i / W;
      const int j = i + y * (p.pitch - W);
      if (CB == 4) reinterpret_cast<float4*>(sp)[j] = make_float4(f[0], f[1], f[2 % CB], f[3 % CB]);
      else reinterpret_cast<float2*>(sp)[j] = make_float2(f[0], f[1]);
    }
  }
  __syncthreads();
  uint32_t sbase;
  {
    unsigned long long s64;
    asm volatile("cvta.to.shared.u64 %0, %1;" : "=l"(s64) : "l"((unsigned long long)(uintptr_t)smem_raw));
    sbase = (uint32_t)s64;
  }
  const uint32_t pitch = (uint32_t)p.pitch * CS;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int total = nroi * BINS;
  const int32_t* order = p.order + start + pos0;
  const int stride = nw * 32;
  const int cap = p.capy + p.capx;

  // the (proposal, headers, scale, count) of the NEXT pass are fetched while this one runs
  int flat = wid * 32 + lane;
  int r_n = 0, bin_n = 0;
  uint2 hy_n = make_uint2(0, 0), hx_n = make_uint2(0, 0);
  float sc_n = 1.f, cnt_n = 1.f;
  auto fetch = [&](int f) {
    if (f < total) {
      const int rpos = f / BINS;
      bin_n = f - rpos * BINS;
      r_n = __ldg(order + rpos);
      const int ph = bin_n / 7, pw = bin_n - ph * 7;
      hy_n = __ldg(p.hdr + (int64_t)r_n * 14 + ph);
      hx_n = __ldg(p.hdr + (int64_t)r_n * 14 + 7 + pw);
      cnt_n = __ldg(p.alignp + (int64_t)r_n * 8 + 7);   // 1 / count
      if (p.row_scale) sc_n = __fadd_rn(__ldg(p.row_scale + r_n), p.row_scale_bias);
    }
  };
  fetch(flat);
  for (; flat < total; flat += stride) {
    const int r = r_n, bin = bin_n;
    const uint2 hy = hy_n, hx = hx_n;
    const float scale = sc_n, count = cnt_n;
    fetch(flat + stride);
    const int y0 = hy.x & 0xffff, ny = hy.x >> 16;
    const int x0 = hx.x & 0xffff, nx = hx.x >> 16;
    const float* wyp = p.wts + (int64_t)r * cap + hy.y;
    const float* wxp = p.wts + (int64_t)r * cap + p.capy + hx.y;
    float acc[CB];
#pragma unroll
    for (int k = 0; k < CB; ++k) acc[k] = 0.f;
    const uint32_t cell = sbase + (uint32_t)(y0 * p.pitch + x0) * CS;
    for (int xc = 0; xc < nx; xc += 4) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(wxp + xc));
      const int m = nx - xc;
      uint32_t a = cell + (uint32_t)xc * CS;
#pragma unroll 2
      for (int t = 0; t < ny; ++t, a += pitch) {
        const float wy = __ldg(wyp + t);
        float f[CB], rs[CB];
        lds_cells<CB>(a, f);
#pragma unroll
        for (int k = 0; k < CB; ++k) rs[k] = w.x * f[k];
        if (m > 1) {
          lds_cells<CB>(a + CS, f);
#pragma unroll
          for (int k = 0; k < CB; ++k) rs[k] = fmaf(w.y, f[k], rs[k]);
        }
        if (m > 2) {
          lds_cells<CB>(a + 2 * CS, f);
#pragma unroll
          for (int k = 0; k < CB; ++k) rs[k] = fmaf(w.z, f[k], rs[k]);
        }
        if (m > 3) {
          lds_cells<CB>(a + 3 * CS, f);
#pragma unroll
          for (int k = 0; k < CB; ++k) rs[k] = fmaf(w.w, f[k], rs[k]);
        }
#pragma unroll
        for (int k = 0; k < CB; ++k) acc[k] = fmaf(wy, rs[k], acc[k]);
      }
    }
    const int64_t o = ((int64_t)r * p.C + c0) * BINS + bin;
#pragma unroll
    for (int k = 0; k < CB; ++k)
      if (k < nc) {
        const float v = __fmul_rn(acc[k], count);        // count holds 1 / (gh * gw): one more rounding, inside the stated 1e-5
        __stcs(p.output + o + k * BINS, p.row_scale ? __fmul_rn(v, scale) : v);
      }
  }
}

// ------------------------------------------------------------------------------------------------
// host side (called from roi_pool.cu: pool_common)
// ------------------------------------------------------------------------------------------------
size_t align7_sep_workspace(int64_t R, int64_t H, int64_t W) {
  const size_t cap = (size_t)asep::axis_cap((int)H, 7) + (size_t)asep::axis_cap((int)W, 7);
  return align_up(sizeof(uint2) * 14 * (size_t)R, 256) + align_up(sizeof(float) * cap * (size_t)R + 16, 256);
}

// channels per CTA the separable kernel would use for this map (0: it does not apply)
static int sep_pitch(int64_t W) { return (int)(W | 1); }

int align7_sep_cb(int64_t C, int64_t H, int64_t W) {
  const size_t plane = (size_t)H * sep_pitch(W) * sizeof(float);
  if (C >= 3 && 4 * plane <= (size_t)kMaxSmemOptin) return 4;
  if (C >= 2 && 2 * plane <= (size_t)kMaxSmemOptin) return 2;
  return 0;
}

template <int CB>
static int launch_sep(AlignSepParams& p, int64_t R, cudaStream_t st) {
  p.pitch = sep_pitch(p.W);
  const size_t smem = CB * (size_t)p.H * p.pitch * sizeof(float);
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
  auto kern = roi_align7_sep_kernel<CB>;
  if (smem > 32 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  kern<<<(unsigned)((int64_t)p.N * p.S * p.CG), threads, smem, st>>>(p);
  return after_launch();
}

// counts / order / alignp come from roi_prepare_kernel<MODE_ALIGN> + roi_order_kernel (already enqueued on `st`)
int align7_sep(const float* input, int64_t N, int64_t C, int64_t H, int64_t W, int64_t R, const int32_t* counts,
               const int32_t* order, const float* alignp, const float* row_scale, float row_scale_bias, float* output,
               void* workspace, cudaStream_t st) {
  const int capy = asep::axis_cap((int)H, 7), capx = asep::axis_cap((int)W, 7);
  uint2* hdr = (uint2*)workspace;
  float* wts = (float*)((char*)workspace + align_up(sizeof(uint2) * 14 * (size_t)R, 256));
  roi_align_tables_kernel<<<(unsigned)ceil_div(2 * R, 128), 128, 0, st>>>(alignp, R, (int)H, (int)W, capy, capx, hdr, wts);
  int rc = after_launch();
  if (rc) return rc;
  AlignSepParams p;
  p.input = input; p.row_scale = row_scale; p.row_scale_bias = row_scale_bias; p.output = output;
  p.counts = counts; p.order = order; p.alignp = alignp; p.hdr = hdr; p.wts = wts;
  p.capy = capy; p.capx = capx; p.N = (int)N; p.C = (int)C; p.H = (int)H; p.W = (int)W; p.CG = 0; p.S = 1;
  return align7_sep_cb(C, H, W) == 4 ? launch_sep<4>(p, R, st) : launch_sep<2>(p, R, st);
}

}  // namespace wsovod
