// align_tc.cu -- the tensor-core path of kernel (2a): tcgen05.mma kind::tf32, operands fed by TMA into
// 128B-swizzled shared memory, fp32 accumulators in TMEM, fused normalise / temperature / bias /
// row-softmax epilogue read back with tcgen05.ld.
//
// Persistent, warp-specialised, one CTA per SM (grid = min(#row tiles, 148)):
//   warp 0   TMA producer : x tile [128 rows x 32 fp32] + W^ tile [BN rows x 32 fp32] per pipeline stage
//   warp 1   MMA issuer   : one elected thread, 4 x (M=128, N=BN, K=8) tcgen05.mma per stage,
//                           tcgen05.commit releases the stage / publishes the accumulator
//   warp 2   TMEM allocator (2 accumulator buffers so the epilogue of tile t overlaps the MMAs of t+1)
//   warps 4-7 epilogue    : thread == accumulator row (TMEM lane); ||x_r|| from the (L2-hot) rows of x,
//                           logits = acc * T/||x_r|| (+bias), online softmax across N chunks.
// x is consumed as fp32 straight from HBM (the tensor core truncates the low 13 mantissa bits: TF32),
// no conversion pass, no extra copy.  K (classes) > 256 is processed in chunks of 256 accumulator
// columns with a running (max, sum) per row.
#include <cuda.h>

#include "align.cuh"
#include "tc_ptx.cuh"

#include <algorithm>
#include <mutex>

namespace wsovod {

struct TcParams {
  const float* x;
  const float* bias;
  float* logits;      // [M, KO] (never null inside the kernel: falls back to `probs` storage)
  float* probs;       // [M, KO] or null
  float* rowstat;     // [M, 2] (row max, sum exp) when the softmax is finished by normalize_rows_kernel
  int64_t M;
  int D, KO, Kp;      // Kp: padded weight rows (multiple of 32)
  int BN;             // accumulator columns per chunk (multiple of 32, <= 256)
  int nchunks, kblocks, stages, ntiles;
  int norm;
  int write_logits;   // 0: only probabilities are wanted and they come straight from TMEM
  float temperature;
  uint32_t tmem_cols;
  // fused alignment + MIL (align_mil.cu): row tiles come from a table so that every tile lies inside ONE image
  // (x = first row, y = rows, z = image); the epilogue then also reduces the detection stream's column softmax
  // statistics of its rows: colpart[(tile * 4 + warp) * KO + k] = (max, sum exp) over the warp's <= 32 rows
  const int4* tiles;          // null: tile t = rows [128 t, 128 t + 128)
  const int* ntiles_dev;      // number of valid table entries (device-side: offsets never visit the host)
  const float* det;           // [M, KO] detection-stream logits (MIL mode)
  float2* colpart;
};

// exp of the TF32 path's softmax: ex2.approx (1e-6 relative on probabilities whose logits carry 1e-3)
__device__ __forceinline__ float tc_exp(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
  return r;
}

struct TcTile { int row0, rows; };
template <bool MIL>
__device__ __forceinline__ TcTile tc_tile(const TcParams& p, int tile) {
  if (MIL) { const int4 t = __ldg(p.tiles + tile); return TcTile{t.x, t.y}; }
  const int64_t r0 = (int64_t)tile * TC_BM;
  return TcTile{(int)r0, (int)min((int64_t)TC_BM, p.M - r0)};
}

// MIL = fused alignment + MIL flavour (tile table, detection-stream column statistics in the epilogue); the plain
// flavour compiles to the same code as before the fusion existed
// REGROW: the single-chunk, <= 96-column flavour whose epilogue keeps the thread's logit row in registers
template <bool MIL, bool REGROW>
__global__ void __launch_bounds__(TC_THREADS, 1)
align_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const TcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // 128B-swizzled operand stages must start on a 1024-byte boundary of the shared window
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t a_bytes = TC_BM * TC_BK * 4, b_bytes = (uint32_t)p.BN * TC_BK * 4;
  unsigned char* sa = smem;                                   // [stages][16 KB]
  unsigned char* sb = smem + (size_t)p.stages * a_bytes;      // [stages][BN*128 B]
  float* stage_out = reinterpret_cast<float*>(sb + (size_t)p.stages * b_bytes);   // [4 warps][32][33] epilogue staging
  float* snorm = stage_out + 4 * 32 * 33;                     // [2][128] sum of squares per row, per tile parity
  uint64_t* bars = reinterpret_cast<uint64_t*>(snorm + 2 * TC_BM);
  uint64_t* full = bars;
  uint64_t* empty = bars + p.stages;
  uint64_t* tfull = bars + 2 * p.stages;
  uint64_t* tempty = tfull + 2;
  uint64_t* nfull = tempty + 2;
  uint64_t* nempty = nfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(nempty + 2);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
  }
  if (warp == 1 && lane == 0) {
    // a stage is released by the MMA commit and by the two norm warps that read it
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 3); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 128);
      mbar_init(&nfull[b], 2); mbar_init(&nempty[b], 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int ntiles = MIL ? min(__ldg(p.ntiles_dev), p.ntiles) : p.ntiles;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int row0 = tc_tile<MIL>(p, tile).row0;
        for (int c = 0; c < p.nchunks; ++c)
          for (int kb = 0; kb < p.kblocks; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_expect_tx(&full[stage], a_bytes + b_bytes);
            tma_load_2d(&map_x, sa + (size_t)stage * a_bytes, &full[stage], kb * TC_BK, row0);
            tma_load_2d(&map_w, sb + (size_t)stage * b_bytes, &full[stage], kb * TC_BK, c * p.BN);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D=F32, A=B=TF32, both K-major, N=BN, M=128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      int stage = 0; uint32_t phase = 0; uint32_t it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
        for (int c = 0; c < p.nchunks; ++c, ++it) {
          const uint32_t buf = it & 1, aphase = (it >> 1) & 1;
          mbar_wait(&tempty[buf], aphase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * (uint32_t)p.BN;
          for (int kb = 0; kb < p.kblocks; ++kb) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint64_t adesc = umma_desc_sw128(smem_u32(sa + (size_t)stage * a_bytes));
            const uint64_t bdesc = umma_desc_sw128(smem_u32(sb + (size_t)stage * b_bytes));
#pragma unroll
            for (int k = 0; k < TC_BK / 8; ++k)   // UMMA_K = 8 tf32 = 32 B: advance the start address inside the atom
              tc_mma_tf32(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
            tc_commit(&empty[stage]);            // frees the smem stage once these MMAs have read it
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
          tc_commit(&tfull[buf]);                // accumulator complete
        }
    }
  } else if (warp < 4) {
    // ===== norm warps: ||x_r||^2 from the very stages the MMA consumes (x crosses HBM once) =====
    const int t = (warp - 2) * 32 + lane;        // 0..63, rows t and t + 64 of the tile
    int stage = 0; uint32_t phase = 0; uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
      float ss0 = 0.f, ss1 = 0.f;
      for (int c = 0; c < p.nchunks; ++c) {
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(&full[stage], phase);
          if (c == 0 && p.norm) {
            // a row is 128 B = 8 chunks of 16 B (swizzled among themselves: irrelevant for a sum);
            // lane l starts at chunk l & 7 so a quarter-warp touches 8 different bank groups
            const unsigned char* base = sa + (size_t)stage * a_bytes;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int ch = (j + lane) & 7;
              const float4 v0 = *reinterpret_cast<const float4*>(base + t * 128 + ch * 16);
              const float4 v1 = *reinterpret_cast<const float4*>(base + (t + 64) * 128 + ch * 16);
              ss0 += v0.x * v0.x + v0.y * v0.y + v0.z * v0.z + v0.w * v0.w;
              ss1 += v1.x * v1.x + v1.y * v1.y + v1.z * v1.z + v1.w * v1.w;
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (c == 0) {
          // publish as soon as the first N chunk has streamed all of x's tile: the epilogue needs the
          // norms to drain accumulator 0 while the MMA warp is already on the next chunks
          const uint32_t nb = tcount & 1, nphase = (tcount >> 1) & 1;
          mbar_wait(&nempty[nb], nphase ^ 1);    // the epilogue has read the previous use of this buffer
          snorm[nb * TC_BM + t] = ss0;
          snorm[nb * TC_BM + t + 64] = ss1;
          __syncwarp();
          if (lane == 0) mbar_arrive(&nfull[nb]);
        }
      }
    }
  } else {
    // ===== epilogue warps: thread == accumulator row; stores staged per warp for coalescing =====
    const int wq = warp - 4;                     // TMEM lane quarter this warp may access (warp % 4)
    float* st = stage_out + wq * 32 * 33;        // [32 rows][33]
    const float bias = p.bias ? __ldg(p.bias) : 0.f;
    uint32_t it = 0, tcount = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
      const TcTile tt = tc_tile<MIL>(p, tile);
      const int64_t wrow0 = (int64_t)tt.row0 + wq * 32;          // first row of this warp
      const int wrows = min(32, tt.rows - wq * 32);              // rows of the tile this warp owns (<= 0: none)
      const uint32_t nb = tcount & 1, nphase = (tcount >> 1) & 1;
      mbar_wait(&nfull[nb], nphase);
      float scale = 1.f;
      if (p.norm) scale = p.temperature / fmaxf(sqrtf(snorm[nb * TC_BM + wq * 32 + lane]), 1e-12f);
      mbar_arrive(&nempty[nb]);
      float run_m = -FLT_MAX, run_s = 0.f;
      for (int c = 0; c < p.nchunks; ++c, ++it) {
        const uint32_t buf = it & 1, aphase = (it >> 1) & 1;
        mbar_wait(&tfull[buf], aphase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + buf * (uint32_t)p.BN;
        const int col0 = c * p.BN;
        const int ncols = min(p.BN, p.KO - col0);        // valid output columns in this chunk
        if (REGROW) {
          // Up to 96 classes (VOC 20, COCO 80 + background): the whole logit row of this thread lives in registers.
          // One pass over TMEM, the accumulator buffer goes back to the MMA warp at once, and max / exp / sum /
          // normalise never leave the register file.  (The three-sweep code below with accurate expf was what bounded
          // the K = 80 kernel: ~10 us of epilogue per 128-row tile against 9 us of HBM time for its x rows.)  The
          // exponentials of this TF32 path are ex2.approx: ~1e-6 relative on probabilities whose logits carry 1e-3.
          float w[3][32];
          const int ns = (ncols + 31) >> 5;
#pragma unroll
          for (int q = 0; q < 3; ++q)
            if (q < ns) tmem_ld32(taddr + 32 * q, w[q]);
          tc_fence_before();
          mbar_arrive(&tempty[buf]);
          constexpr float L2E = 1.4426950408889634f;
          float m4[4] = {-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
#pragma unroll
          for (int q = 0; q < 3; ++q)
            if (q < ns) {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                w[q][i] = fmaf(w[q][i], scale, bias);
                if (q * 32 + i < ncols) m4[i & 3] = fmaxf(m4[i & 3], w[q][i]);
              }
            }
          const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
          if (p.write_logits) {
#pragma unroll
            for (int q = 0; q < 3; ++q)
              if (q < ns) {
#pragma unroll
                for (int i = 0; i < 32; ++i) st[lane * 33 + i] = w[q][i];
                __syncwarp();
                const int cc = q * 32 + lane;
                if (cc < p.KO)
                  for (int rr = 0; rr < 32; ++rr)
                    if (rr < wrows) p.logits[(wrow0 + rr) * p.KO + cc] = st[rr * 33 + lane];
                __syncwarp();
              }
          }
          if (p.probs) {
            const float nb = -mx * L2E;
            float a4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int q = 0; q < 3; ++q)
              if (q < ns) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                  float e;
                  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(w[q][i], L2E, nb)));
                  w[q][i] = e;
                  if (q * 32 + i < ncols) a4[i & 3] += e;
                }
              }
            const float inv = 1.f / ((a4[0] + a4[1]) + (a4[2] + a4[3]));
#pragma unroll
            for (int q = 0; q < 3; ++q)
              if (q < ns) {
#pragma unroll
                for (int i = 0; i < 32; ++i) st[lane * 33 + i] = w[q][i] * inv;
                __syncwarp();
                const int cc = q * 32 + lane;
                if (cc < p.KO) {
                  if (MIL) {
                    // lane == output column: the same walk reduces the detection stream's column statistics over this
                    // warp's rows (32 independent loads, then max and sum of exp as two passes)
                    float dv[32];
#pragma unroll
                    for (int rr = 0; rr < 32; ++rr) dv[rr] = rr < wrows ? __ldg(p.det + (wrow0 + rr) * p.KO + cc) : -FLT_MAX;
                    float dm = -FLT_MAX, ds = 0.f;
#pragma unroll
                    for (int rr = 0; rr < 32; ++rr) dm = fmaxf(dm, dv[rr]);
#pragma unroll
                    for (int rr = 0; rr < 32; ++rr) ds += rr < wrows ? expf(dv[rr] - dm) : 0.f;
                    p.colpart[((size_t)tile * 4 + wq) * p.KO + cc] = make_float2(dm, ds);
                  }
                  for (int rr = 0; rr < 32; ++rr)
                    if (rr < wrows) p.probs[(wrow0 + rr) * p.KO + cc] = st[rr * 33 + lane];
                }
                __syncwarp();
              }
          }
          continue;
        }
        float v[32];
        // sweep A: chunk maximum
        float cm = -FLT_MAX;
        for (int j = 0; j < ncols; j += 32) {
          tmem_ld32(taddr + j, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) if (j + i < ncols) cm = fmaxf(cm, fmaf(v[i], scale, bias));
        }
        const float new_m = fmaxf(run_m, cm);
        run_s *= tc_exp(run_m - new_m);
        run_m = new_m;
        // sweep B: logits (staged, then written as 128-byte row segments), running sum of exp
        for (int j = 0; j < ncols; j += 32) {
          tmem_ld32(taddr + j, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float l = fmaf(v[i], scale, bias);
            if (j + i < ncols) run_s += tc_exp(l - run_m);
            st[lane * 33 + i] = l;
          }
          __syncwarp();
          const int cc = col0 + j + lane;
          if (p.write_logits && cc < p.KO)
            for (int rr = 0; rr < 32; ++rr)
              if (rr < wrows) p.logits[(wrow0 + rr) * p.KO + cc] = st[rr * 33 + lane];
          __syncwarp();
        }
        if (p.probs && p.nchunks == 1) {
          // sweep C (single chunk): probabilities straight from TMEM
          const float inv = 1.f / run_s;
          for (int j = 0; j < ncols; j += 32) {
            tmem_ld32(taddr + j, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) st[lane * 33 + i] = tc_exp(fmaf(v[i], scale, bias) - run_m) * inv;
            __syncwarp();
            const int cc = j + lane;
            if (cc < p.KO) {
              // lane == output column: 128-byte row segments; in MIL mode the same walk reduces the detection
              // stream's column statistics over this warp's rows (online max / sum of exp, row order)
              if (MIL) {
                // 32 independent loads, then max and sum-exp as two passes without a dependent rescale chain
                float dv[32];
#pragma unroll
                for (int rr = 0; rr < 32; ++rr) dv[rr] = rr < wrows ? __ldg(p.det + (wrow0 + rr) * p.KO + cc) : -FLT_MAX;
                float dm = -FLT_MAX, ds = 0.f;
#pragma unroll
                for (int rr = 0; rr < 32; ++rr) dm = fmaxf(dm, dv[rr]);
#pragma unroll
                for (int rr = 0; rr < 32; ++rr) ds += rr < wrows ? expf(dv[rr] - dm) : 0.f;
                p.colpart[((size_t)tile * 4 + wq) * p.KO + cc] = make_float2(dm, ds);
              }
              for (int rr = 0; rr < 32; ++rr)
                if (rr < wrows) p.probs[(wrow0 + rr) * p.KO + cc] = st[rr * 33 + lane];
            }
            __syncwarp();
          }
        }
        tc_fence_before();
        mbar_arrive(&tempty[buf]);               // accumulator buffer may be overwritten
      }
      if (p.rowstat && lane < wrows) {
        // multi-chunk: (max, sum) per row for the streaming normalisation kernel that follows
        p.rowstat[2 * (wrow0 + lane)] = run_m;
        p.rowstat[2 * (wrow0 + lane) + 1] = run_s;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// probs[r, :] = exp(logits[r, :] - max_r) / sum_r   (K > 256: the row spans several accumulator chunks).  `rowstat`
// holds `ns` (max, sum of exp(logit - max)) pairs per row -- one for the whole row from align_tc_kernel, one per chunk
// from the CTA-pair kernel -- merged here.  One warp per row, 16-byte accesses when the row pitch allows.
__global__ void __launch_bounds__(256) normalize_rows_kernel(const float* __restrict__ logits, const float* __restrict__ rowstat,
                                                             int ns, int64_t M, int KO, float* __restrict__ probs) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < M; r += nwarps) {
    const float* rs = rowstat + 2 * r * ns;
    float mx = __ldg(rs), sum = __ldg(rs + 1);
    for (int c = 1; c < ns; ++c) {
      const float m2 = __ldg(rs + 2 * c), s2 = __ldg(rs + 2 * c + 1);
      const float nm = fmaxf(mx, m2);
      sum = sum * expf(mx - nm) + s2 * expf(m2 - nm);
      mx = nm;
    }
    const float inv = 1.f / sum;
    const float* src = logits + r * KO;
    float* dst = probs + r * KO;
    if ((KO & 3) == 0 && (((uintptr_t)logits | (uintptr_t)probs) & 15) == 0) {
      const float4* s4 = reinterpret_cast<const float4*>(src);
      float4* d4 = reinterpret_cast<float4*>(dst);
      for (int i = lane; i < KO / 4; i += 32) {
        float4 v = s4[i];
        v.x = expf(v.x - mx) * inv; v.y = expf(v.y - mx) * inv; v.z = expf(v.z - mx) * inv; v.w = expf(v.w - mx) * inv;
        d4[i] = v;
      }
    } else {
      for (int i = lane; i < KO; i += 32) dst[i] = expf(src[i] - mx) * inv;
    }
  }
}

// ---- host ----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

int tc_make_map(CUtensorMap* m, const float* base, uint64_t inner, uint64_t rows, uint64_t row_stride_elems,
                uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return WSOVOD_B200_EUNSUPPORTED;
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {row_stride_elems * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : WSOVOD_B200_EINVAL;
}

int align_fwd_tf32(const float* x, const float* classifier, int64_t M, int64_t D, int64_t K, float temperature,
                   int norm_weight, int append_background, const float* bias, float* logits, float* probs,
                   const AlignWs& w, char* ws, cudaStream_t st, const AlignMilFuse* mil) {
  if ((D & 3) || ((uintptr_t)x & 15)) return WSOVOD_B200_EALIGN;   // TMA: 16-byte rows
  if (M > 0x7fffffffLL - TC_BM) return WSOVOD_B200_ETOOBIG;
  const int64_t KO = K + (append_background ? 1 : 0);
  float* what = (float*)(ws + w.what);                              // [Kp, Dp], rows >= K and cols >= D are zero
  // the normalised text matrix, zero-padded to [Kp, Dp] (the appended background column is a zero row, :97-100)
  const int64_t nch2 = ceil_div(KO, 256);
  const bool pair = !mil && nch2 > 1 && tune(TUNE_ALIGN_PAIR);
  align_wnorm_kernel<<<(unsigned)ceil_div(w.Kp, 8), 256, 0, st>>>(classifier, (int)K, (int)w.Kp, (int)D, (int)w.Dp, norm_weight == 1, what,
                                                                  pair ? (int*)(ws + w.tickets) : nullptr,
                                                                  pair ? (int)(w.tickets_bytes / sizeof(int)) : 0);
  int rc;
  if ((rc = after_launch())) return rc;
  if (pair) {
    // large vocabularies: CTA pairs (align_tc2.cu), row softmax finished inside the kernel where the row pitch allows
    float* lg = logits ? logits : probs;
    if ((rc = align_tc2_launch(x, what, M, D, KO, w.Kp, w.Dp, temperature, norm_weight, bias, lg, probs,
                               (int*)(ws + w.tickets), (float*)(ws + w.rowstat), st))) return rc;
    return 0;
  }
  TcParams p;
  p.x = x; p.bias = bias; p.logits = logits ? logits : probs; p.probs = probs;
  p.rowstat = nullptr;
  p.tiles = nullptr; p.ntiles_dev = nullptr; p.det = nullptr; p.colpart = nullptr;
  p.M = M; p.D = (int)D; p.KO = (int)KO; p.Kp = (int)w.Kp;
  p.BN = (int)std::min<int64_t>(w.Kp, 256);
  p.nchunks = (int)ceil_div(KO, p.BN);
  p.kblocks = (int)ceil_div(D, TC_BK);
  p.ntiles = (int)ceil_div(M, TC_BM);
  if (mil) {
    if (p.nchunks != 1 || !probs) return WSOVOD_B200_EUNSUPPORTED;   // the column statistics ride on the single-chunk sweep
    p.tiles = mil->tiles; p.ntiles_dev = mil->ntiles_dev; p.ntiles = mil->ntiles_max;
    p.det = mil->det; p.colpart = mil->colpart;
  }
  p.norm = norm_weight; p.temperature = temperature;
  p.write_logits = (logits != nullptr || p.nchunks > 1) ? 1 : 0;
  if (probs && p.nchunks > 1) p.rowstat = (float*)(ws + w.rowstat);
  const size_t stage_bytes = (size_t)TC_BM * TC_BK * 4 + (size_t)p.BN * TC_BK * 4;
  const size_t extra = 4 * 32 * 33 * sizeof(float) + 2 * TC_BM * sizeof(float) + 512;   // staging, norms, barriers
  p.stages = (int)std::max<size_t>(2, std::min<size_t>(8, ((size_t)kMaxSmemOptin - 1024 - extra) / stage_bytes));
  const size_t smem = (size_t)p.stages * stage_bytes + extra + 1024;                   // + alignment slack
  uint32_t cols = 32;
  while (cols < 2u * (uint32_t)p.BN) cols <<= 1;
  p.tmem_cols = cols;                                               // <= 512
  CUtensorMap mx, mw;
  if ((rc = tc_make_map(&mx, x, (uint64_t)D, (uint64_t)M, (uint64_t)D, TC_BM))) return rc;
  if ((rc = tc_make_map(&mw, what, (uint64_t)w.Dp, (uint64_t)w.Kp, (uint64_t)w.Dp, (uint32_t)p.BN))) return rc;
  const bool regrow = p.nchunks == 1 && p.BN <= 96;
  auto kern = mil ? (regrow ? align_tc_kernel<true, true> : align_tc_kernel<true, false>)
                  : (regrow ? align_tc_kernel<false, true> : align_tc_kernel<false, false>);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const int grid = std::min(p.ntiles, kNumSMs);
  kern<<<grid, TC_THREADS, smem, st>>>(mx, mw, p);
  if ((rc = after_launch())) return rc;
  if (p.rowstat) {   // in place when the caller did not ask for logits (p.logits aliases probs)
    normalize_rows_kernel<<<kNumSMs * 16, 256, 0, st>>>(p.logits, p.rowstat, 1, M, (int)KO, probs);
    if ((rc = after_launch())) return rc;
  }
  return 0;
}

}  // namespace wsovod
