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
#include "common.cuh"
#include "pool_pyr.cuh"

#include <algorithm>

namespace wsovod {

using namespace pyr;

struct PyrWs {
  int32_t* hist;        // [N, kBuckets]
  int32_t* img_start;   // [N + 1]   first sorted position of each image
  int32_t* bucket_off;  // [N, kBuckets + 1] bucket boundaries (positions relative to the image start)
  int32_t* bidx;        // [R]
  uint32_t* pkey;       // [R]
  int32_t* order;       // [R]   proposal id at each sorted position
  uint2* pinfo;         // [R]   (proposal id, bits of row_scale + bias) at each sorted position
  uint32_t* desc;       // [R, 49] bin descriptors in sorted position / lane slot order
  size_t bytes;
};

static PyrWs pyr_carve(void* ws, int64_t N, int64_t R) {
  PyrWs w;
  size_t off = 0;
  char* base = (char*)ws;
  auto take = [&](size_t b) { size_t o = off; off += align_up(b, 256); return base + o; };
  w.hist = (int32_t*)take(sizeof(int32_t) * (size_t)N * kBuckets);
  w.img_start = (int32_t*)take(sizeof(int32_t) * (size_t)(N + 1));
  w.bucket_off = (int32_t*)take(sizeof(int32_t) * (size_t)N * (kBuckets + 1));
  w.bidx = (int32_t*)take(sizeof(int32_t) * (size_t)R);
  w.pkey = (uint32_t*)take(sizeof(uint32_t) * (size_t)R);
  w.order = (int32_t*)take(sizeof(int32_t) * (size_t)R);
  w.pinfo = (uint2*)take(sizeof(uint2) * (size_t)R);
  w.desc = (uint32_t*)take(sizeof(uint32_t) * (size_t)R * 49);
  w.bytes = off;
  return w;
}

size_t pool7_pyr_workspace(int64_t N, int64_t R) { return pyr_carve(nullptr, N, R).bytes; }

// ------------------------------------------------------------------------------------------------
// prologue: classify -> per-image bucket sort -> descriptors in sorted order
// ------------------------------------------------------------------------------------------------
__global__ void pyr_classify_kernel(const float* __restrict__ rois, int64_t R, int N, int H, int W, float scale,
                                    int32_t* __restrict__ bidx, uint32_t* __restrict__ pkey,
                                    int32_t* __restrict__ hist) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float* roi = rois + r * 5;
  int b = (int)roi[0];
  b = min(max(b, 0), N - 1);
  const uint32_t key = proposal_key(roi[1], roi[2], roi[3], roi[4], scale, H, W);
  bidx[r] = b;
  pkey[r] = key;
  atomicAdd(&hist[(int64_t)b * kBuckets + key_bucket(key)], 1);
}

// exclusive scan of the per-image totals (one CTA; thread t owns a contiguous run of images)
__global__ void pyr_scan_kernel(const int32_t* __restrict__ hist, int N, int32_t* __restrict__ img_start) {
  __shared__ int s_part[1024];
  const int tid = threadIdx.x;
  const int q = (N + blockDim.x - 1) / blockDim.x;
  const int n0 = min(tid * q, N), n1 = min(n0 + q, N);
  int sum = 0;
  for (int n = n0; n < n1; ++n)
    for (int k = 0; k < kBuckets; ++k) sum += hist[(int64_t)n * kBuckets + k];
  s_part[tid] = sum;
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int t = 0; t < (int)blockDim.x; ++t) { const int v = s_part[t]; s_part[t] = run; run += v; }
    img_start[N] = run;
  }
  __syncthreads();
  int run = s_part[tid];
  for (int n = n0; n < n1; ++n) {
    img_start[n] = run;
    for (int k = 0; k < kBuckets; ++k) run += hist[(int64_t)n * kBuckets + k];
  }
}

// one CTA per image: bucket offsets and scatter of proposal ids.  The order inside a bucket is whatever
// the shared-memory atomics give: every output element is written exactly once from
// position-independent data, so the result does not depend on it.
__global__ void pyr_order_kernel(const int32_t* __restrict__ bidx, const uint32_t* __restrict__ pkey,
                                 const int32_t* __restrict__ hist, const int32_t* __restrict__ img_start,
                                 int64_t R, int32_t* __restrict__ order, int32_t* __restrict__ bucket_off) {
  __shared__ int s_cur[kBuckets];
  const int n = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) {
    int run = 0;
    for (int k = 0; k < kBuckets; ++k) {
      bucket_off[n * (kBuckets + 1) + k] = run;
      s_cur[k] = run;
      run += hist[(int64_t)n * kBuckets + k];
    }
    bucket_off[n * (kBuckets + 1) + kBuckets] = run;
  }
  __syncthreads();
  const int base = img_start[n];
  for (int64_t r = tid; r < R; r += blockDim.x) {
    if (bidx[r] != n) continue;
    const int pos = atomicAdd(&s_cur[key_bucket(pkey[r])], 1);
    order[base + pos] = (int32_t)r;
  }
}

// one thread per (sorted position, lane slot): the descriptor of the bin that slot serves
__global__ void pyr_bins_kernel(const float* __restrict__ rois, int64_t R, int H, int W, float scale,
                                const int32_t* __restrict__ order, const uint32_t* __restrict__ pkey,
                                const float* __restrict__ row_scale, float row_scale_bias,
                                uint2* __restrict__ pinfo, uint32_t* __restrict__ desc) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * 49) return;
  const int64_t gpos = i / 49;
  const int q = (int)(i - gpos * 49);
  const int r = order[gpos];
  const uint32_t key = pkey[r];
  const int bin = slot_bin(key, q);
  const int ph = bin / 7, pw = bin - ph * 7;
  const float* roi = rois + (int64_t)r * 5;
  const int phase = key_phase(key);
  desc[i] = phase == PH_FALLBACK ? 0u : bin_desc(roi[1], roi[2], roi[3], roi[4], scale, H, W, phase, ph, pw);
  if (q == 0) {
    const float sc = row_scale ? __fadd_rn(row_scale[r], row_scale_bias) : 1.f;   // roi_heads.py:733-739
    pinfo[gpos] = make_uint2((uint32_t)r, __float_as_uint(sc));
  }
}

// ------------------------------------------------------------------------------------------------
// main kernel
// ------------------------------------------------------------------------------------------------
struct PyrParams {
  const float* input;
  const float* rois;
  float scale;
  float* output;
  const int32_t* img_start;
  const int32_t* bucket_off;
  const uint2* pinfo;
  const uint32_t* desc;
  int32_t N, C, H, W;
  int32_t CG, S;
};

template <int CB> __device__ __forceinline__ void p_lds(uint32_t addr, float* f);
template <> __device__ __forceinline__ void p_lds<4>(uint32_t addr, float* f) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f[0]), "=f"(f[1]), "=f"(f[2]), "=f"(f[3]) : "r"(addr));
}
template <> __device__ __forceinline__ void p_lds<2>(uint32_t addr, float* f) {
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(f[0]), "=f"(f[1]) : "r"(addr));
}
template <int CB> __device__ __forceinline__ void p_sts(uint32_t addr, const float* f);
template <> __device__ __forceinline__ void p_sts<4>(uint32_t addr, const float* f) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(f[0]), "f"(f[1]), "f"(f[2]), "f"(f[3]) : "memory");
}
template <> __device__ __forceinline__ void p_sts<2>(uint32_t addr, const float* f) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(f[0]), "f"(f[1]) : "memory");
}

// D <- plane of the map (mode 0), max with the right neighbour (1: block 1x2) or the lower neighbour
// (2: block 2x1), in the padded layout; pad cells and absent channels hold the identity -FLT_MAX.
template <int CB>
__device__ __noinline__ void pyr_stage(uint32_t sbase, const float* __restrict__ src, int nc, int H, int W,
                                       int mode) {
  constexpr uint32_t CS = 4u * CB;
  const int WP = W + kPad, ncell = (H + kPad) * WP, HW = H * W;
  for (int idx = threadIdx.x; idx < ncell; idx += blockDim.x) {
    const int hh = idx / WP, ww = idx - hh * WP;
    const int h = hh - kPad, w = ww - kPad;
    float f[CB];
#pragma unroll
    for (int k = 0; k < CB; ++k) f[k] = -FLT_MAX;
    if (h >= 0 && w >= 0) {
#pragma unroll
      for (int k = 0; k < CB; ++k)
        if (k < nc) f[k] = fmaxf(f[k], __ldg(src + (int64_t)k * HW + h * W + w));
    }
    if (mode == 1 && h >= 0 && w + 1 >= 0 && w + 1 < W) {
#pragma unroll
      for (int k = 0; k < CB; ++k)
        if (k < nc) f[k] = fmaxf(f[k], __ldg(src + (int64_t)k * HW + h * W + w + 1));
    }
    if (mode == 2 && w >= 0 && h + 1 >= 0 && h + 1 < H) {
#pragma unroll
      for (int k = 0; k < CB; ++k)
        if (k < nc) f[k] = fmaxf(f[k], __ldg(src + (int64_t)k * HW + (h + 1) * W + w));
    }
    p_sts<CB>(sbase + (uint32_t)idx * CS, f);
  }
}

// in-place doubling D[i] = max(D[i], D[i + stride]) over the first `ncell` cells.  Reads only go
// forward, so chunks are processed front to back with one barrier between a chunk's loads and its
// stores.  Horizontal steps may wrap into the next row's pad columns, which hold the identity for the
// strides used (1, 2 with 3 pad columns); vertical steps run into the identity tail rows.
template <int CB>
__device__ __noinline__ void pyr_double(uint32_t sbase, int ncell, int stride) {
  constexpr uint32_t CS = 4u * CB;
  constexpr int U = 4;
  const uint32_t sb = (uint32_t)stride * CS;
  for (int base = 0; base < ncell; base += (int)blockDim.x * U) {
    float v[U][CB];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int idx = base + u * (int)blockDim.x + (int)threadIdx.x;
      if (idx < ncell) {
        float b[CB];
        p_lds<CB>(sbase + (uint32_t)idx * CS, v[u]);
        p_lds<CB>(sbase + (uint32_t)idx * CS + sb, b);
#pragma unroll
        for (int k = 0; k < CB; ++k) v[u][k] = fmaxf(v[u][k], b[k]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int idx = base + u * (int)blockDim.x + (int)threadIdx.x;
      if (idx < ncell) p_sts<CB>(sbase + (uint32_t)idx * CS, v[u]);
    }
  }
  __syncthreads();
}

// One bucket slice = `total` lane slots (49 per proposal) whose proposals all need CH x CW blocks per bin.
// Per pass a lane reads one descriptor word (coalesced) and its proposal's (id, scale) pair, both
// fetched one pass ahead, issues up to CH*CW LDS (a block is skipped where it would repeat the
// previous one: the bin is not larger than the blocks before it) and stores CB scalars.
template <int CB, int CH, int CW, bool FULL>
__device__ __forceinline__ void pyr_run(uint32_t sbase, uint32_t pitch, uint32_t khp, uint32_t kwb,
                                        const uint32_t* __restrict__ dsc, const uint2* __restrict__ pin,
                                        int total, float* __restrict__ outc, uint32_t c49, int nc, int flat0,
                                        int stride) {
  constexpr uint32_t CS = 4u * CB;
  int f = flat0;
  uint32_t d_n = 0;
  uint2 pi_n = make_uint2(0u, 0u);
  if (f < total) {
    d_n = __ldg(dsc + f);
    pi_n = __ldg(pin + (uint32_t)f / 49u);
  }
  while (f < total) {
    const uint32_t d = d_n;
    const uint2 pi = pi_n;
    const int fn = f + stride;
    if (fn < total) {
      d_n = __ldg(dsc + fn);
      pi_n = __ldg(pin + (uint32_t)fn / 49u);
    }
    const uint32_t a0 = sbase + (d & 0xffffu) * CS;
    const uint32_t lhp = ((d >> 16) & 15u) * pitch, lwb = ((d >> 20) & 15u) * CS;
    // the plane holds no NaN / -inf (pyr_stage clamps at -FLT_MAX), so the first block seeds the maximum
    float m[CB];
    p_lds<CB>(a0, m);
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
          float v[CB];
          p_lds<CB>(a0 + ro + co, v);
#pragma unroll
          for (int k = 0; k < CB; ++k) m[k] = fmaxf(m[k], v[k]);
        }
      }
    }
    const float sc = __uint_as_float(pi.y);   // 1.0f without a row scale: exact
    float* o = outc + (size_t)pi.x * c49 + ((d >> 24) & 63u);
#pragma unroll
    for (int k = 0; k < CB; ++k)
      if (FULL || k < nc) __stcs(o + k * 49, __fmul_rn(m[k], sc));
    f = fn;
  }
}

template <int CB>
__global__ void __launch_bounds__(1024, 1) roi_pool7_pyr_kernel(const PyrParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int BINS = 49;
  constexpr uint32_t CS = 4u * CB;
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
#pragma unroll
    for (int k = 0; k < CB; ++k) id[k] = -FLT_MAX;
    for (int i = ncell + (int)threadIdx.x; i < ntot; i += blockDim.x) p_sts<CB>(sbase + (uint32_t)i * CS, id);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int k = 0; k < CB; ++k) id[k] = 0.f;
      p_sts<CB>(sbase + (uint32_t)ntot * CS, id);
    }
  }
  const float* src = p.input + ((int64_t)n * p.C + c0) * HW;
  float* outc = p.output + (size_t)c0 * BINS;
  const uint32_t c49 = (uint32_t)p.C * BINS;
  const uint32_t pitch = (uint32_t)WP * CS;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int stride = nw * 32;
  const int flat0 = wid * 32 + lane;
  const int32_t* boff = p.bucket_off + n * (kBuckets + 1);

  for (int phase = 0; phase < kPhases; ++phase) {
    const int plo = __ldg(boff + phase * 16);
    const int rem = __ldg(boff + (chain_end(phase) + 1) * 16) - plo;   // proposals left in this chain
    if (rem > 0 && phase != PH_FALLBACK) {
      __syncthreads();                                                 // everyone is done reading the old plane
      switch (phase) {
        case PH_11: pyr_stage<CB>(sbase, src, nc, H, W, 0); __syncthreads(); break;
        case PH_21: pyr_double<CB>(sbase, ncell, WP); break;
        case PH_22: pyr_double<CB>(sbase, ncell, 1); break;
        case PH_42: pyr_double<CB>(sbase, ncell, 2 * WP); break;
        case PH_44: pyr_double<CB>(sbase, ncell, 2); break;
        case PH_12: pyr_stage<CB>(sbase, src, nc, H, W, 1); __syncthreads(); break;
        case PH_14: pyr_double<CB>(sbase, ncell, 2); break;
        case PH_24: pyr_double<CB>(sbase, ncell, WP); break;
        default:    pyr_stage<CB>(sbase, src, nc, H, W, 2); __syncthreads();
                    pyr_double<CB>(sbase, ncell, 2 * WP); break;       // PH_41
      }
    }
    if (phase != PH_FALLBACK) {
      const uint32_t khp = (uint32_t)phase_kh(phase) * pitch, kwb = (uint32_t)phase_kw(phase) * CS;
      for (int sub = 0; sub < 16; ++sub) {
        const int lo = __ldg(boff + phase * 16 + sub), hi = __ldg(boff + phase * 16 + sub + 1);
        if (hi <= lo) continue;
        const int per = (hi - lo + p.S - 1) / p.S;
        const int slo = lo + sidx * per;
        const int shi = min(hi, slo + per);
        if (shi <= slo) continue;
        const int total = (shi - slo) * BINS;
        const uint32_t* dsc = p.desc + (size_t)(gstart + slo) * BINS;
        const uint2* pin = p.pinfo + gstart + slo;
#define PYR_CASE(CH, CW) \
  case ((CH - 1) + (CW - 1) * 4): \
    if (nc == CB) pyr_run<CB, CH, CW, true>(sbase, pitch, khp, kwb, dsc, pin, total, outc, c49, nc, flat0, stride); \
    else pyr_run<CB, CH, CW, false>(sbase, pitch, khp, kwb, dsc, pin, total, outc, c49, nc, flat0, stride); \
    break;
        switch (sub) {
          PYR_CASE(1, 1) PYR_CASE(1, 2) PYR_CASE(1, 3) PYR_CASE(1, 4)
          PYR_CASE(2, 1) PYR_CASE(2, 2) PYR_CASE(2, 3) PYR_CASE(2, 4)
          PYR_CASE(3, 1) PYR_CASE(3, 2) PYR_CASE(3, 3) PYR_CASE(3, 4)
          PYR_CASE(4, 1) PYR_CASE(4, 2) PYR_CASE(4, 3) PYR_CASE(4, 4)
        }
#undef PYR_CASE
      }
    } else {
      // bins needing more than kMaxLoads blocks per axis: direct scan of the (1,1) plane with edges
      // recomputed from the roi (lane slot = output bin)
      const int lo = plo, hi = __ldg(boff + (phase + 1) * 16);
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
        const float* roi = p.rois + (int64_t)pi.x * 5;
        const Axis ah = axis_of(roi[2], roi[4], p.scale), aw = axis_of(roi[1], roi[3], p.scale);
        int hs, he, ws, we;
        bin_edges(ah, ph, H, hs, he);
        bin_edges(aw, pw, W, ws, we);
        const bool empty = he <= hs || we <= ws;
        float m[CB];
#pragma unroll
        for (int k = 0; k < CB; ++k) m[k] = empty ? 0.f : -FLT_MAX;
        if (!empty) {
          for (int h = hs; h < he; ++h) {
            uint32_t a = sbase + (uint32_t)((h + kPad) * WP + ws + kPad) * CS;
            for (int w = ws; w < we; ++w, a += CS) {
              float f[CB];
              p_lds<CB>(a, f);
#pragma unroll
              for (int k = 0; k < CB; ++k) m[k] = fmaxf(m[k], f[k]);
            }
          }
        }
        float* o = outc + (size_t)pi.x * c49 + bin;
#pragma unroll
        for (int k = 0; k < CB; ++k)
          if (k < nc) __stcs(o + k * BINS, __fmul_rn(m[k], __uint_as_float(pi.y)));
      }
    }
  }
}

// shared memory of the padded plane (+ identity tail rows)
static size_t pyr_smem(int64_t H, int64_t W, int cb) {
  return ((size_t)(H + kPad + kTailRows) * (size_t)(W + kPad) + 1) * 4u * (size_t)cb;   // + the zero cell
}

// channels per CTA the pyramid path would use for this map (0: does not apply)
int pool7_pyr_cb(int64_t C, int64_t H, int64_t W, int64_t R) {
  if ((H + kPad + kTailRows) * (W + kPad) >= 65535 || C * 49 >= (1LL << 32)) return 0;
  if (C >= 3 && pyr_smem(H, W, 4) <= (size_t)kMaxSmemOptin) return 4;
  if (pyr_smem(H, W, 2) <= (size_t)kMaxSmemOptin) return 2;
  return 0;
}

template <int CB>
static int pyr_launch_main(PyrParams& p, int64_t R, cudaStream_t st) {
  const size_t smem = pyr_smem(p.H, p.W, CB);
  p.CG = (int)ceil_div(p.C, CB);
  // every CTA of an image rebuilds all planes, so proposals are only split when (image, channel group)
  // units alone cannot fill the machine twice over
  const int64_t units = (int64_t)p.N * p.CG;
  int64_t S = units >= 2 * kNumSMs ? 1 : ceil_div(2 * kNumSMs, units);
  const int64_t avg = std::max<int64_t>(R / std::max(p.N, 1), 1);
  S = std::max<int64_t>(1, std::min<int64_t>(S, ceil_div(avg, 512)));
  p.S = (int)S;
  if (units * S > 0x7fffffffLL) return WSOVOD_B200_ETOOBIG;
  auto kern = roi_pool7_pyr_kernel<CB>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  kern<<<(unsigned)(units * S), 1024, smem, st>>>(p);
  return after_launch();
}

// values-only ROI max-pool 7x7 through the block-max planes; `workspace` holds pool7_pyr_workspace() bytes
int pool7_pyr(const float* input, int64_t N, int64_t C, int64_t H, int64_t W, const float* rois, int64_t R,
              float scale, const float* row_scale, float row_scale_bias, float* output, void* workspace,
              cudaStream_t st) {
  const int cb = pool7_pyr_cb(C, H, W, R);
  if (!cb) return WSOVOD_B200_EINVAL;
  PyrWs w = pyr_carve(workspace, N, R);
  cudaError_t e = cudaMemsetAsync(w.hist, 0, sizeof(int32_t) * (size_t)N * kBuckets, st);
  if (e != cudaSuccess) return (int)e;
  int rc;
  pyr_classify_kernel<<<(unsigned)ceil_div(R, 128), 128, 0, st>>>(rois, R, (int)N, (int)H, (int)W, scale, w.bidx, w.pkey, w.hist);
  if ((rc = after_launch())) return rc;
  pyr_scan_kernel<<<1, 1024, 0, st>>>(w.hist, (int)N, w.img_start);
  if ((rc = after_launch())) return rc;
  pyr_order_kernel<<<(unsigned)N, 512, 0, st>>>(w.bidx, w.pkey, w.hist, w.img_start, R, w.order, w.bucket_off);
  if ((rc = after_launch())) return rc;
  pyr_bins_kernel<<<(unsigned)ceil_div(R * 49, 256), 256, 0, st>>>(rois, R, (int)H, (int)W, scale, w.order, w.pkey, row_scale,
                                                                    row_scale_bias, w.pinfo, w.desc);
  if ((rc = after_launch())) return rc;
  PyrParams p;
  p.input = input; p.rois = rois; p.scale = scale;
  p.output = output; p.img_start = w.img_start; p.bucket_off = w.bucket_off; p.pinfo = w.pinfo; p.desc = w.desc;
  p.N = (int)N; p.C = (int)C; p.H = (int)H; p.W = (int)W; p.CG = 0; p.S = 1;
  return cb == 4 ? pyr_launch_main<4>(p, R, st) : pyr_launch_main<2>(p, R, st);
}

}  // namespace wsovod
