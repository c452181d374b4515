// align_tc2.cu -- the alignment contraction for large vocabularies (K + 1 > 256: LVIS, c4) on CTA PAIRS:
// tcgen05.mma.cta_group::2 kind::tf32, one 256 x 256 accumulator step per SM pair.
//
// Why a second kernel: at K = 1203 the one-CTA-per-tile kernel (align_tc.cu) ran the tensor pipe 40 % of the time with
// nothing saturated.  A CTA pair shares the text tile -- each CTA stages its own 128 rows of x and HALF of the 256 text
// rows (32 KB per stage instead of 48 KB, six stages instead of four), the leader CTA's single thread issues the MMAs
// for both SMs, each CTA's TMEM receives the accumulator rows of its own x rows -- but what actually bounded that
// kernel was its epilogue (four warps, three TMEM sweeps, accurate expf, scalar row-segment stores): with the epilogue
// below the main loop runs at the TF32 peak, and the row softmax, which cannot be completed before the last chunk of a
// row exists, is finished inside the kernel instead of by a second pass over the matrix.
//
//   cluster = 2 CTAs (one TPC), grid = 2 x min(#units, resident clusters), persistent, 384 threads per CTA.
//   unit    = (256-row tile, 256-column chunk of the vocabulary).  Unit u runs on pair u % #pairs: the chunks of a tile
//             run on neighbouring pairs at the same time (its x rows come from HBM once, its logits are complete
//             within one accumulator step), and the 148 SMs finish within one unit of each other (125 tiles x 5 chunks
//             over 74 pairs: 8 or 9 units each; whole tiles per pair would be 1 or 2 = 84 % busy).
//   warp 0  TMA producer (both CTAs): x tile [128 x 32 fp32] + text half tile [128 x 32] per stage, completing on the
//           LEADER's `full` barrier (cp.async.bulk.tensor ... .cta_group::2)
//   warp 1  MMA issuer (leader CTA only): 4 x (M = 256, N <= 256, K = 8) per stage; `tcgen05.commit ... multicast` arrives
//           on `cons[stage]` of BOTH CTAs, and on `tfull[buf]` of both when the accumulator is complete
//   warps 2-3  norm warps (both CTAs): wait for `cons[stage]` (the MMAs are done with the stage, the bytes are still
//           there), add up ||x_r||^2 as the unit streams by, then hand the stage back (`empty`)
//   warps 4-7  epilogue (both CTAs, thread == accumulator row): logits = acc * T / ||x_r|| (+ bias) into a 128B-swizzled
//           [32 x 32] box per warp and out with a TMA store; running (max, sum of exp) of the row over the unit;
//           releases the accumulator buffer on the leader's `tempty` (remote arrive for the peer CTA); hands in the
//           unit's TICKET (release) once its stores have completed
//   warps 8-11 finishers: probs = exp(logit - max_r) / sum_r for 4-row jobs taken from one global counter, each after
//           the tickets of all chunks of its rows are in; every other warp joins them when its own role is done.
// A row of 1204 logits never exists on chip (TMEM holds 512 columns), so its softmax needs the logits back from memory
// once all chunks are known; done inside the kernel they come back from L2 while the tensor pipe works on later units,
// instead of a second pass over 2 x 154 MB after it (what was measured on the way is in DESIGN.md, Kernel 2a).
// The last chunk of the vocabulary runs with N rounded up to 16 only (K = 1203: 192 instead of 256 columns).
// Row pitches that are not a multiple of 16 bytes (TMA) take plain stores and a separate softmax pass.
// Launched cooperatively: the finishing warps wait for tickets of other CTAs, so the whole grid must be resident.
#include "align.cuh"
#include "tc_ptx.cuh"

#include <algorithm>

namespace wsovod {

constexpr int TC2_BN = 256;          // accumulator columns per unit (UMMA N), both TMEM buffers = 512 columns
constexpr int TC2_HALF = TC2_BN / 2; // text rows staged per CTA and stage
constexpr int TC2_STAGING = 2 * 32 * 128;   // epilogue staging per warp: two [32 rows x 32 floats] boxes
constexpr int TC2_JOBS = 32;                // finishing jobs per half tile: 4 rows each
constexpr int TC2_THREADS = 384;            // 8 warps as in align_tc.cu + 4 warps that only finish row softmaxes

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory object in CTA `rank` of the cluster (shared::cluster window)
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done = 0;
  for (int spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(a), "r"(parity) : "memory");
    if (spin > TC_SPIN_LIMIT) __trap();
  }
}
// TMA tile load of a CTA pair: lands in the executing CTA's shared memory, completes on the leader CTA's barrier
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, void* smem, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs once the pair's MMAs issued so far have completed
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ float ex2_approx(float t) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t)); return r; }

struct Tc2Params {
  const float* bias;
  float* logits;       // [M, KO]
  int tma_out;         // logits leave through swizzled shared memory + TMA stores (row pitch a multiple of 16 bytes)
  float* probs;        // non-null: the row softmax is finished inside the kernel (needs tma_out), in place if == logits
  int* tickets;        // [ntiles * 8 + 1] zeroed: chunks of each 32-row block whose logits are in memory; + the job counter
  float2* rowstat;     // [M, nchunks] (chunk max, sum of exp(logit - chunk max)) written by the epilogue for the finishers
  int64_t M;
  int KO, nchunks, kblocks, stages, ntiles;   // ntiles: 256-row tiles
  int norm;
  float temperature;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC2_THREADS, 1)
align_tc2_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                 const __grid_constant__ CUtensorMap map_out, const Tc2Params p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t a_bytes = TC_BM * TC_BK * 4, b_bytes = TC2_HALF * TC_BK * 4;
  unsigned char* sa = smem;                                   // [stages][16 KB] own 128 rows of x
  unsigned char* sb = smem + (size_t)p.stages * a_bytes;      // [stages][16 KB] own half of the text tile
  unsigned char* stage_out = sb + (size_t)p.stages * b_bytes;  // [4 warps][2][32 rows x 128 B, 128B-swizzled] (or [32][33] floats)
  float* snorm = reinterpret_cast<float*>(stage_out + 4 * TC2_STAGING);   // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(snorm + 2 * TC_BM);
  uint64_t* full = bars;                    // leader's: all four tile loads of the stage have landed (both CTAs)
  uint64_t* cons = bars + p.stages;         // the pair's MMAs have read the stage
  uint64_t* empty = bars + 2 * p.stages;    // the norm warps are done with it too: the producer may refill
  uint64_t* tfull = bars + 3 * p.stages;    // accumulator buffer complete
  uint64_t* tempty = tfull + 2;             // leader's: both CTAs' epilogues have drained the buffer
  uint64_t* nfull = tempty + 2;
  uint64_t* nempty = nfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(nempty + 2);

  const uint32_t rank = cluster_ctarank();
  const int pair = (int)(blockIdx.x >> 1), npairs = (int)(gridDim.x >> 1);
  const int64_t U = (int64_t)p.ntiles * p.nchunks;
  // unit i of this pair is pair + i * npairs; npairs >= nchunks, so a pair never sees two units of one tile
  const int u_step = npairs, u_first = pair;
  const int u_count = (int)((U - pair + npairs - 1) / npairs);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_out) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&cons[s], 1); mbar_init(&empty[s], 2); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 256);
      mbar_init(&nfull[b], 2); mbar_init(&nempty[b], 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    // one warp of EACH CTA of the pair issues the paired allocation (all 512 columns: two 256-column buffers)
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();                // the peer's barriers exist before anything is signalled across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int i = 0; i < u_count; ++i) {
        const int u = u_first + i * u_step;
        const int tile = u / p.nchunks, c = u - tile * p.nchunks;
        const int64_t row0 = (int64_t)tile * (2 * TC_BM) + rank * TC_BM;
        const int xr = row0 < p.M ? (int)row0 : 0;                       // a half tile past the last row: any rows do
        const int nc = min(TC2_BN, (p.KO - c * TC2_BN + 15) & ~15);      // accumulator columns of this chunk
        const int wr = c * TC2_BN + (int)rank * (nc >> 1);               // this CTA's half of the chunk's text rows
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          const uint32_t lfull = mapa_u32(smem_u32(&full[stage]), 0);
          if (rank == 0) mbar_expect_tx(&full[stage], 2 * (a_bytes + b_bytes));
          tma_load_2d_pair(&map_x, sa + (size_t)stage * a_bytes, lfull, kb * TC_BK, xr);
          tma_load_2d_pair(&map_w, sb + (size_t)stage * b_bytes, lfull, kb * TC_BK, wr);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      int stage = 0; uint32_t phase = 0; uint32_t it = 0;
      for (int i = 0; i < u_count; ++i, ++it) {
        const int c = (u_first + i * u_step) % p.nchunks;
        const int nc = min(TC2_BN, (p.KO - c * TC2_BN + 15) & ~15);
        // instruction descriptor: D = F32, A = B = TF32, both K-major, N = nc, M = 256 (128 rows per CTA)
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(nc >> 3) << 17) | ((uint32_t)((2 * TC_BM) >> 4) << 24);
        const uint32_t buf = it & 1, aphase = (it >> 1) & 1;
        mbar_wait_cluster(&tempty[buf], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * (uint32_t)TC2_BN;
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(smem_u32(sa + (size_t)stage * a_bytes));
          const uint64_t bdesc = umma_desc_sw128(smem_u32(sb + (size_t)stage * b_bytes));
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k)
            tc_mma_tf32_pair(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
          tc_commit_pair(&cons[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        tc_commit_pair(&tfull[buf]);
      }
    }
  } else if (warp < 4) {
    // ===== norm warps =====
    const int t = (warp - 2) * 32 + lane;        // rows t and t + 64 of this CTA's half tile
    int stage = 0; uint32_t phase = 0;
    for (int i = 0; i < u_count; ++i) {
      float ss0 = 0.f, ss1 = 0.f;
      for (int kb = 0; kb < p.kblocks; ++kb) {
        mbar_wait(&cons[stage], phase);
        if (p.norm) {
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
      const uint32_t nb = i & 1, nphase = (i >> 1) & 1;
      mbar_wait(&nempty[nb], nphase ^ 1);        // the epilogue has read the previous use of this buffer
      snorm[nb * TC_BM + t] = ss0;
      snorm[nb * TC_BM + t + 64] = ss1;
      __syncwarp();
      if (lane == 0) mbar_arrive(&nfull[nb]);
    }
  } else if (warp < 8) {
    // ===== epilogue warps =====
    const int wq = warp - 4;
    unsigned char* stw = stage_out + wq * TC2_STAGING;
    float* st = reinterpret_cast<float*>(stw);
    const float bias = p.bias ? __ldg(p.bias) : 0.f;
    const uint32_t ltempty0 = mapa_u32(smem_u32(&tempty[0]), 0), ltempty1 = mapa_u32(smem_u32(&tempty[1]), 0);
    uint32_t it = 0, sbuf = 0;
    float scale = 1.f;
    for (int i = 0; i < u_count; ++i, ++it) {
      const int u = u_first + i * u_step;
      const int tile = u / p.nchunks, c = u - tile * p.nchunks;
      const int64_t wrow0 = (int64_t)tile * (2 * TC_BM) + rank * TC_BM + wq * 32;
      const int wrows = (int)max((int64_t)0, min((int64_t)32, p.M - wrow0));
      {
        const uint32_t nb = i & 1, nphase = (i >> 1) & 1;
        mbar_wait(&nfull[nb], nphase);
        scale = p.norm ? p.temperature / fmaxf(sqrtf(snorm[nb * TC_BM + wq * 32 + lane]), 1e-12f) : 1.f;
        mbar_arrive(&nempty[nb]);
      }
      const uint32_t buf = it & 1, aphase = (it >> 1) & 1;
      mbar_wait(&tfull[buf], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + buf * (uint32_t)TC2_BN;
      const int col0 = c * TC2_BN;
      const int ncols = min(TC2_BN, p.KO - col0);
      float v[32];
      float cm = -FLT_MAX, cs = 0.f;
      for (int j = 0; j < ncols; j += 32) {
        tmem_ld32(taddr + j, v);
        if (p.tma_out) {
          // thread == row: its 32 logits are 128 contiguous bytes of the output row.  They go into a [32 x 128 B] box
          // in the TMA 128B swizzle (16-byte chunk q of row r at chunk q ^ (r & 7): conflict-free for a quarter-warp),
          // and one lane sends the box; columns >= KO and rows >= M are clipped by the tensor map.
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the box two stores back has been read
          __syncwarp();
          unsigned char* box = stw + sbuf * (32 * 128);
          unsigned char* dst = box + lane * 128;
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaf(v[i], scale, bias);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(dst + ((q ^ (lane & 7)) << 4)) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          if (p.probs) {
            // running (max, sum of exp) of this row over the unit's columns, one rescale per 32-column slab; the sums use
            // the fast exponential (they are dominated by the terms next to the maximum, where its error is ~1e-7)
            float m4[4] = {-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
#pragma unroll
            for (int i = 0; i < 32; ++i) if (j + i < ncols) m4[i & 3] = fmaxf(m4[i & 3], v[i]);
            const float nm = fmaxf(cm, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
            const float nb = -nm * 1.4426950408889634f;
            float a4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 32; ++i) if (j + i < ncols) a4[i & 3] += ex2_approx(fmaf(v[i], 1.4426950408889634f, nb));
            cs = cs * ex2_approx(fmaf(cm, 1.4426950408889634f, nb)) + ((a4[0] + a4[1]) + (a4[2] + a4[3]));
            cm = nm;
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0 && wrows > 0) {
            // L2 policy: evict-last while the finishers still have to read the rows back, evict-first when nobody does
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
                         ::"l"(&map_out), "r"(smem_u32(box)), "r"(col0 + j), "r"((int)wrow0),
                           "l"(p.probs ? 0x14F0000000000000ull : 0x12F0000000000000ull) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          sbuf ^= 1;
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) st[lane * 33 + i] = fmaf(v[i], scale, bias);
          __syncwarp();
          const int cc = col0 + j + lane;
          if (cc < p.KO)
            for (int rr = 0; rr < 32; ++rr)
              if (rr < wrows) p.logits[(wrow0 + rr) * p.KO + cc] = st[rr * 33 + lane];
          __syncwarp();
        }
      }
      tc_fence_before();
      mbar_arrive_cluster(buf ? ltempty1 : ltempty0);
      if (p.probs) {
        // the pair's part of these 32 rows is in memory once the bulk stores (lane 0's) have completed: then the ticket
        // goes in as a release (the lanes' statistics are ordered before it by the warp barrier); the finishers poll
        // it with acquire loads
        if (lane < wrows) p.rowstat[(wrow0 + lane) * p.nchunks + c] = make_float2(cm, cs);
        __syncwarp();
        if (lane == 0) {
          asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
          asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p.tickets + (wrow0 >> 5)), "r"(1) : "memory");
        }
      }
    }
    if (lane == 0 && p.tma_out) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  // ===== finishing: probs[r, :] = exp(logits[r, :] - max_r) / sum_r =====
  // Jobs are 4-row blocks of the whole matrix in row order (= the order in which the interleaved units complete tiles),
  // taken from ONE global counter by whichever warp of whichever CTA is free: the four finisher warps of every CTA
  // from the start, every other warp once its own role is done (the tail of the kernel is all finishing, spread over
  // all SMs).  Tying a tile's finish to the pair that produced it -- "last one in finishes", or a designated pair --
  // left the slow pairs with the most work.  A job waits until the tickets of all chunks of its rows are in, merges the
  // per-chunk statistics of its rows and streams the rows back from L2 through independent 16-byte loads, two rows in
  // flight and the next two prefetched.
  // The exponentials of this (TF32) path are ex2.approx: relative error ~1e-6 on probabilities whose logits carry 1e-3.
  __syncwarp();
  if (p.probs) {
    const int n4 = p.KO >> 2;
    constexpr float L2E = 1.4426950408889634f;
    int* const global_next = p.tickets + p.ntiles * 8;      // zeroed with the tickets
    const int njobs = p.ntiles * 2 * TC2_JOBS;
    for (;;) {
      int j = 0;
      if (lane == 0) j = atomicAdd(global_next, 1);
      j = __shfl_sync(0xffffffffu, j, 0);
      if (j >= njobs) break;
      const int row0 = j * (TC_BM / TC2_JOBS);               // jobs in row order = the order in which tiles complete
      if (row0 >= p.M) continue;
      {
        const int* tk = p.tickets + (row0 >> 5);
        for (int spin = 0;; ++spin) {
          int got;
          asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(got) : "l"(tk) : "memory");
          if (got >= p.nchunks) break;
          __nanosleep(100);
          if (spin > (1 << 24)) __trap();
        }
      }
      const int nrows = (int)min((int64_t)(TC_BM / TC2_JOBS), p.M - row0);
      float nb_l = 0.f, inv_l = 0.f;               // lane rr: -max * log2(e) and 1 / sum of row rr
      if (lane < nrows) {
        const float2* rs = p.rowstat + ((int64_t)row0 + lane) * p.nchunks;
        float m = -FLT_MAX, sum = 0.f;
        for (int c = 0; c < p.nchunks; ++c) {
          const float2 st2 = __ldcg(rs + c);
          const float nm = fmaxf(m, st2.x);
          sum = sum * __expf(m - nm) + st2.y * __expf(st2.x - nm);
          m = nm;
        }
        nb_l = -m * L2E; inv_l = 1.f / sum;
      }
      auto ex2 = [](float t) { return ex2_approx(t); };
      if (n4 <= 320) {
        // rows of up to 1280 columns, two at a time = 20 independent 16-byte loads per lane.  Rolling prefetch: as soon
        // as an element of this pair of rows has been consumed its register is reloaded with the same element of the next
        // pair, so the next round trip runs under this one's exponentials and stores.
        float4 a[2][10];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const float4* s4 = reinterpret_cast<const float4*>(p.logits + (int64_t)(row0 + min(q, nrows - 1)) * p.KO);
#pragma unroll
          for (int i = 0; i < 10; ++i) a[q][i] = __ldcg(s4 + min(i * 32 + lane, n4 - 1));
        }
        for (int rr = 0; rr < nrows; rr += 2) {
          const bool more = rr + 2 < nrows;
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int r_this = min(rr + q, nrows - 1), r_next = min(rr + 2 + q, nrows - 1);
            const float nb = __shfl_sync(0xffffffffu, nb_l, r_this);
            const float inv = __shfl_sync(0xffffffffu, inv_l, r_this);
            float4* d4 = reinterpret_cast<float4*>(p.probs + (int64_t)(row0 + r_this) * p.KO);
            const float4* s4 = reinterpret_cast<const float4*>(p.logits + (int64_t)(row0 + r_next) * p.KO);
#pragma unroll
            for (int i = 0; i < 10; ++i) {
              float4 o;
              o.x = ex2(fmaf(a[q][i].x, L2E, nb)) * inv; o.y = ex2(fmaf(a[q][i].y, L2E, nb)) * inv;
              o.z = ex2(fmaf(a[q][i].z, L2E, nb)) * inv; o.w = ex2(fmaf(a[q][i].w, L2E, nb)) * inv;
              if (more) a[q][i] = __ldcg(s4 + min(i * 32 + lane, n4 - 1));
              if (i * 32 + lane < n4 && rr + q < nrows) d4[i * 32 + lane] = o;
            }
          }
        }
      } else {
        for (int rr = 0; rr < nrows; ++rr) {
          const float nb = __shfl_sync(0xffffffffu, nb_l, rr), inv = __shfl_sync(0xffffffffu, inv_l, rr);
          const float4* s4 = reinterpret_cast<const float4*>(p.logits + (int64_t)(row0 + rr) * p.KO);
          float4* d4 = reinterpret_cast<float4*>(p.probs + (int64_t)(row0 + rr) * p.KO);
          for (int k0 = 0; k0 < n4; k0 += 32 * 16) {
            float4 a[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = __ldcg(s4 + min(k0 + i * 32 + lane, n4 - 1));
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float4 o;
              o.x = ex2(fmaf(a[i].x, L2E, nb)) * inv; o.y = ex2(fmaf(a[i].y, L2E, nb)) * inv;
              o.z = ex2(fmaf(a[i].z, L2E, nb)) * inv; o.w = ex2(fmaf(a[i].w, L2E, nb)) * inv;
              if (k0 + i * 32 + lane < n4) d4[k0 + i * 32 + lane] = o;
            }
          }
        }
      }
    }
  }
  __syncwarp();
  tc_fence_before();
  cluster_sync_all();                // nobody leaves while the peer may still read its shared memory or signal its barriers
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// resident 2-CTA clusters of this kernel at `smem` bytes per CTA (74 on a full B200), queried once per size
static int pair_slots(size_t smem) {
  static size_t cached_smem = 0;
  static int cached = 0;
  if (cached && cached_smem == smem) return cached;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * (kNumSMs / 2)); cfg.blockDim = dim3(TC2_THREADS); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, align_tc2_kernel, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = kNumSMs / 2; }
  cached_smem = smem; cached = std::min(n, kNumSMs / 2);
  return cached;
}

// probs[r, :] = softmax(logits[r, :]) with the whole row in registers (KO <= 2048): one read and one write of the
// matrix, in place if the caller wants.  One warp per row; VEC: 16-byte accesses (row pitch a multiple of 16 bytes).
template <bool VEC>
__global__ void __launch_bounds__(256) softmax_rows_reg_kernel(const float* logits, int64_t M, int KO, float* probs) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < M; r += nwarps) {
    float mx = -FLT_MAX, sum = 0.f;
    if (VEC) {
      const float4* s4 = reinterpret_cast<const float4*>(logits + r * KO);
      float4* d4 = reinterpret_cast<float4*>(probs + r * KO);
      const int n4 = KO >> 2;
      float4 v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (i * 32 < n4) {
          const int k = i * 32 + lane;
          v[i] = k < n4 ? __ldcs(s4 + k) : make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
          mx = fmaxf(mx, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
        }
      mx = warp_max(mx);
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (i * 32 < n4) {
          v[i].x = expf(v[i].x - mx); v[i].y = expf(v[i].y - mx); v[i].z = expf(v[i].z - mx); v[i].w = expf(v[i].w - mx);
          if (i * 32 + lane < n4) sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
      const float inv = 1.f / warp_sum(sum);
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (i * 32 + lane < n4) d4[i * 32 + lane] = make_float4(v[i].x * inv, v[i].y * inv, v[i].z * inv, v[i].w * inv);
    } else {
      const float* src = logits + r * KO;
      float* dst = probs + r * KO;
      float v[64];
#pragma unroll
      for (int i = 0; i < 64; ++i)
        if (i * 32 < KO) {
          v[i] = i * 32 + lane < KO ? src[i * 32 + lane] : -FLT_MAX;
          mx = fmaxf(mx, v[i]);
        }
      mx = warp_max(mx);
#pragma unroll
      for (int i = 0; i < 64; ++i)
        if (i * 32 < KO) {
          v[i] = expf(v[i] - mx);
          if (i * 32 + lane < KO) sum += v[i];
        }
      const float inv = 1.f / warp_sum(sum);
#pragma unroll
      for (int i = 0; i < 64; ++i)
        if (i * 32 + lane < KO) dst[i * 32 + lane] = v[i] * inv;
    }
  }
}

int softmax_rows_launch(const float* logits, int64_t M, int64_t KO, float* probs, cudaStream_t st) {
  if (KO > 2048) {
    row_softmax_kernel<<<(unsigned)ceil_div(M, 8), 256, 0, st>>>(logits, M, (int)KO, probs);
  } else if ((KO & 3) == 0 && (((uintptr_t)logits | (uintptr_t)probs) & 15) == 0) {
    softmax_rows_reg_kernel<true><<<kNumSMs * 8, 256, 0, st>>>(logits, M, (int)KO, probs);
  } else {
    softmax_rows_reg_kernel<false><<<kNumSMs * 8, 256, 0, st>>>(logits, M, (int)KO, probs);
  }
  return after_launch();
}

// logits[M, KO] for KO > 256 and, if asked, their row softmax (`probs` may alias `logits`); `what` is the normalised,
// zero-padded text matrix, `tickets` ntiles * 8 zeroed ints (align_tc2_tickets)
int64_t align_tc2_tickets(int64_t M) { return ceil_div(M, 2 * TC_BM) * 8 + 1; }

int align_tc2_launch(const float* x, const float* what, int64_t M, int64_t D, int64_t KO, int64_t Kp, int64_t Dp,
                     float temperature, int norm, const float* bias, float* logits, float* probs, int* tickets,
                     float* rowstat, cudaStream_t st) {
  Tc2Params p;
  p.bias = bias; p.logits = logits; p.M = M; p.KO = (int)KO;
  p.nchunks = (int)ceil_div(KO, TC2_BN);
  p.kblocks = (int)ceil_div(D, TC_BK);
  p.ntiles = (int)ceil_div(M, 2 * TC_BM);
  p.norm = norm; p.temperature = temperature;
  p.tma_out = ((KO & 3) == 0 && ((uintptr_t)logits & 15) == 0) ? 1 : 0;
  const bool fuse = probs && p.tma_out && KO <= 2048 && ((uintptr_t)probs & 15) == 0 && tickets && rowstat;
  p.probs = fuse ? probs : nullptr;
  p.tickets = tickets;
  p.rowstat = reinterpret_cast<float2*>(rowstat);
  const size_t stage_bytes = (size_t)(TC_BM + TC2_HALF) * TC_BK * 4;
  const size_t extra = 4 * TC2_STAGING + 2 * TC_BM * sizeof(float) + 512;   // staging, norms, barriers + slots
  p.stages = (int)std::max<size_t>(2, std::min<size_t>(8, ((size_t)kMaxSmemOptin - 1024 - extra) / stage_bytes));
  const size_t smem = (size_t)p.stages * stage_bytes + extra + 1024;
  CUtensorMap mx, mw, mo;
  int rc;
  if ((rc = tc_make_map(&mx, x, (uint64_t)D, (uint64_t)M, (uint64_t)D, TC_BM))) return rc;
  if ((rc = tc_make_map(&mw, what, (uint64_t)Dp, (uint64_t)Kp, (uint64_t)Dp, TC2_HALF))) return rc;
  mo = mx;                            // unused without tma_out
  if (p.tma_out && (rc = tc_make_map(&mo, logits, (uint64_t)KO, (uint64_t)M, (uint64_t)KO, 32))) return rc;
  cudaError_t e = cudaFuncSetAttribute(align_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const int64_t units = (int64_t)p.ntiles * p.nchunks;
  const int pairs = (int)std::max<int64_t>(1, std::min<int64_t>(units, pair_slots(smem)));
  // The finishing warps wait for tickets of OTHER CTAs of this grid, so every CTA must be resident: a cooperative launch
  // is gang-scheduled (it starts when all 2 x pairs CTAs fit at once; with another kernel of the caller holding SMs -- a
  // second stream -- it waits instead of starting half of the grid, which could then wait for the other half forever)
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(TC2_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, align_tc2_kernel, mx, mw, mo, p);
  if (e != cudaSuccess) return (int)e;
  if ((rc = after_launch())) return rc;
  return probs && !fuse ? softmax_rows_launch(logits, M, KO, probs, st) : 0;
}


}  // namespace wsovod

