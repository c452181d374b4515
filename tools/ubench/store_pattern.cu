// Micro-benchmark: how fast can the pooled tensor's write pattern go by itself?  (R,C,7,7) fp32, a CTA owns
// (image, 4 channels) and writes 4 x 196-byte runs per proposal (stride C*196 B between proposals).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o store_pattern store_pattern.cu && ./store_pattern
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int POLICY>
__device__ __forceinline__ void st(float* p, float v) {
  if (POLICY == 0) __stcs(p, v);
  else if (POLICY == 1) *p = v;
  else if (POLICY == 2) __stwt(p, v);
  else __stcg(p, v);
}

template <int POLICY>
__global__ void __launch_bounds__(1024, 1) pattern(float* out, int C, int R, int CG, int skew) {
  const int cg = blockIdx.x % CG, n = blockIdx.x / CG;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int total = R * 49;
  const size_t c49 = (size_t)C * 49;
  float* outc = out + (size_t)cg * 4 * 49 + (size_t)n * R * c49;
  // skew: CTAs start at different proposals (emulates drift between channel groups)
  const int start = skew ? (int)(((unsigned)blockIdx.x * 2654435761u) % (unsigned)R) : 0;
  for (int f = wid * 32 + lane; f < total; f += nw * 32) {
    int r = f / 49;
    const int bin = f - r * 49;
    r += start; if (r >= R) r -= R;
    float* o = outc + (size_t)r * c49 + bin;
    const float v = (float)f;
#pragma unroll
    for (int k = 0; k < 4; ++k) st<POLICY>(o + k * 49, v + k);
  }
}

// (V1) warp-owned proposals: 49 bins as 32 + 17 lanes, the 8 stores of a proposal back to back
__global__ void __launch_bounds__(1024, 1) pattern_warp(float* out, int C, int R, int CG) {
  const int cg = blockIdx.x % CG, n = blockIdx.x / CG;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const size_t c49 = (size_t)C * 49;
  float* outc = out + (size_t)cg * 4 * 49 + (size_t)n * R * c49;
  for (int r = wid; r < R; r += nw) {
    float* o = outc + (size_t)r * c49;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      __stcs(o + k * 49 + lane, (float)r);
      if (lane < 17) __stcs(o + k * 49 + 32 + lane, (float)r);
    }
  }
}
// (V1b) as V1 but channel-major inside each half: k0 A, k1 A, k2 A, k3 A, k0 B, ... (adjacent pieces 4 stores apart)
__global__ void __launch_bounds__(1024, 1) pattern_warp_b(float* out, int C, int R, int CG) {
  const int cg = blockIdx.x % CG, n = blockIdx.x / CG;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const size_t c49 = (size_t)C * 49;
  float* outc = out + (size_t)cg * 4 * 49 + (size_t)n * R * c49;
  for (int r = wid; r < R; r += nw) {
    float* o = outc + (size_t)r * c49;
#pragma unroll
    for (int k = 0; k < 4; ++k) __stcs(o + k * 49 + lane, (float)r);
#pragma unroll
    for (int k = 0; k < 4; ++k) if (lane < 17) __stcs(o + k * 49 + 32 + lane, (float)r);
  }
}
// (V4) every warp walks a contiguous run of slots, 64 per pass (slot f and f + 32 per lane, any alignment
// against the 49-slot proposals); ORDER 0: k-major per half (k0 A..k3 A, k0 B..k3 B), 1: halves adjacent
template <int ORDER>
__global__ void __launch_bounds__(1024, 1) pattern_contig(float* out, int C, int R, int CG) {
  const int cg = blockIdx.x % CG, n = blockIdx.x / CG;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const size_t c49 = (size_t)C * 49;
  float* outc = out + (size_t)cg * 4 * 49 + (size_t)n * R * c49;
  const int per_w = (R + nw - 1) / nw;
  const int r0 = min(R, wid * per_w), r1 = min(R, r0 + per_w);
  const int total = (r1 - r0) * 49;
  for (int g = lane; g < total; g += 64) {
    const int ra = g / 49, ba = g - ra * 49;
    const int g2 = g + 32, rb = g2 / 49, bb = g2 - rb * 49;
    float* oa = outc + (size_t)(r0 + ra) * c49 + ba;
    float* ob = outc + (size_t)(r0 + rb) * c49 + bb;
    const bool hb = g2 < total;
    if (ORDER == 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k) __stcs(oa + k * 49, (float)g);
#pragma unroll
      for (int k = 0; k < 4; ++k) if (hb) __stcs(ob + k * 49, (float)g);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) { __stcs(oa + k * 49, (float)g); if (hb) __stcs(ob + k * 49, (float)g); }
    }
  }
}
// (V5) contiguous run, ONE slot per lane per pass (32 slots per pass)
__global__ void __launch_bounds__(1024, 1) pattern_contig32(float* out, int C, int R, int CG) {
  const int cg = blockIdx.x % CG, n = blockIdx.x / CG;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const size_t c49 = (size_t)C * 49;
  float* outc = out + (size_t)cg * 4 * 49 + (size_t)n * R * c49;
  const int per_w = (R + nw - 1) / nw;
  const int r0 = min(R, wid * per_w), r1 = min(R, r0 + per_w);
  const int total = (r1 - r0) * 49;
  for (int g = lane; g < total; g += 32) {
    const int ra = g / 49, ba = g - ra * 49;
    float* oa = outc + (size_t)(r0 + ra) * c49 + ba;
#pragma unroll
    for (int k = 0; k < 4; ++k) __stcs(oa + k * 49, (float)g);
  }
}
// (V6) strided passes over a slot stream padded to 64 slots per proposal (49 real + 15 idle): every warp
// store stays inside one proposal (bins 0..31 or 32..48), the two halves come from different warps
__global__ void __launch_bounds__(1024, 1) pattern_pad64(float* out, int C, int R, int CG) {
  const int cg = blockIdx.x % CG, n = blockIdx.x / CG;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int total = R * 64;
  const size_t c49 = (size_t)C * 49;
  float* outc = out + (size_t)cg * 4 * 49 + (size_t)n * R * c49;
  for (int f = wid * 32 + lane; f < total; f += nw * 32) {
    const int r = f >> 6, bin = f & 63;
    if (bin >= 49) continue;
    float* o = outc + (size_t)r * c49 + bin;
#pragma unroll
    for (int k = 0; k < 4; ++k) __stcs(o + k * 49, (float)f);
  }
}
// (V7) as V6 but a lane handles slot f and f + 32 of the SAME proposal (one warp = one proposal per pass)
__global__ void __launch_bounds__(1024, 1) pattern_pad64_pair(float* out, int C, int R, int CG) {
  const int cg = blockIdx.x % CG, n = blockIdx.x / CG;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const size_t c49 = (size_t)C * 49;
  float* outc = out + (size_t)cg * 4 * 49 + (size_t)n * R * c49;
  for (int r = wid; r < R; r += nw) {
    float* o = outc + (size_t)r * c49 + lane;
#pragma unroll
    for (int k = 0; k < 4; ++k) __stcs(o + k * 49, (float)r);
    if (lane < 17) {
#pragma unroll
      for (int k = 0; k < 4; ++k) __stcs(o + 32 + k * 49, (float)r);
    }
  }
}
// (V8) three proposals per 160 slots (147 bins + 13 idle): passes 2 and 4 of every five hold two proposals.
// SPLIT=0: one store instruction per channel (straddles), SPLIT=1: two, one per proposal
template <int SPLIT>
__global__ void __launch_bounds__(1024, 1) pattern_3in5(float* out, int C, int R, int CG) {
  const int cg = blockIdx.x % CG, n = blockIdx.x / CG;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int groups = R / 3;
  const int total = groups * 160;
  const size_t c49 = (size_t)C * 49;
  float* outc = out + (size_t)cg * 4 * 49 + (size_t)n * R * c49;
  for (int f = wid * 32 + lane; f < total; f += nw * 32) {
    const int grp = f / 160, s = f - grp * 160;
    if (s >= 147) continue;
    const int which = s >= 98 ? 2 : (s >= 49 ? 1 : 0);
    const int bin = s - which * 49;
    const int first = (f - lane) - grp * 160;                  // slot of lane 0
    const int wfirst = first >= 98 ? 2 : (first >= 49 ? 1 : 0);
    float* o = outc + (size_t)(grp * 3 + which) * c49 + bin;
    const bool a = which == wfirst;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (SPLIT) {
        asm volatile("{ .reg .pred p; setp.ne.u32 p, %2, 0; @p st.global.cs.f32 [%0], %1; }" ::"l"(o + k * 49), "f"((float)f), "r"((unsigned)a) : "memory");
        asm volatile("{ .reg .pred p; setp.eq.u32 p, %2, 0; @p st.global.cs.f32 [%0], %1; }" ::"l"(o + k * 49), "f"((float)f), "r"((unsigned)a) : "memory");
      } else {
        __stcs(o + k * 49, (float)f);
      }
    }
  }
}
// (V2) same shape of accesses on a fake layout with 48 floats per (proposal, channel): every warp store is
// 64-byte aligned, i.e. only whole 32-byte sectors are written
__global__ void __launch_bounds__(1024, 1) pattern48(float* out, int C, int R, int CG) {
  const int cg = blockIdx.x % CG, n = blockIdx.x / CG;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int total = R * 48;
  const size_t c48 = (size_t)C * 48;
  float* outc = out + (size_t)cg * 4 * 48 + (size_t)n * R * c48;
  for (int f = wid * 32 + lane; f < total; f += nw * 32) {
    const int r = f / 48, bin = f - r * 48;
    float* o = outc + (size_t)r * c48 + bin;
#pragma unroll
    for (int k = 0; k < 4; ++k) __stcs(o + k * 48, (float)f);
  }
}
// (V3) real layout, each lane stores its 4 channels for bin AND the warp only ever writes whole proposals:
// lanes walk 196 consecutive floats of the chunk but in (bin-major) order: lane l, step t -> run index
__global__ void __launch_bounds__(1024, 1) pattern_shuffled(float* out, int C, int R, int CG) {
  const int cg = blockIdx.x % CG, n = blockIdx.x / CG;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const size_t c49 = (size_t)C * 49;
  float* outc = out + (size_t)cg * 4 * 49 + (size_t)n * R * c49;
  for (int r = wid; r < R; r += nw) {
    float* o = outc + (size_t)r * c49;
    for (int v = lane; v < 196; v += 32) __stcs(o + v, (float)r);
  }
}

// (1) 784-byte chunk per (proposal, channel group) written as 49 float4 (same bytes, same chunk addresses)
__global__ void __launch_bounds__(1024, 1) chunk128(float* out, int C, int R, int CG) {
  const int cg = blockIdx.x % CG, n = blockIdx.x / CG;
  const int total = R * 49;
  const size_t c49 = (size_t)C * 49;
  float* outc = out + (size_t)cg * 4 * 49 + (size_t)n * R * c49;
  for (int f = threadIdx.x; f < total; f += blockDim.x) {
    const int r = f / 49, v = f - r * 49;
    __stcs(reinterpret_cast<float4*>(outc + (size_t)r * c49) + v, make_float4(1.f, 2.f, 3.f, (float)f));
  }
}
// (3) the same chunk written as 196 consecutive scalars (lanes walk the run)
__global__ void __launch_bounds__(1024, 1) chunk32(float* out, int C, int R, int CG) {
  const int cg = blockIdx.x % CG, n = blockIdx.x / CG;
  const int total = R * 196;
  const size_t c49 = (size_t)C * 49;
  float* outc = out + (size_t)cg * 4 * 49 + (size_t)n * R * c49;
  for (int f = threadIdx.x; f < total; f += blockDim.x) {
    const int r = f / 196, v = f - r * 196;
    __stcs(outc + (size_t)r * c49 + v, (float)f);
  }
}
// (4) 784-byte chunks through shared memory and cp.async.bulk (TMA) stores: one warp stages 32 chunks
__global__ void __launch_bounds__(1024, 1) chunk_tma(float* out, int C, int R, int CG) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int cg = blockIdx.x % CG, n = blockIdx.x / CG;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const size_t c49 = (size_t)C * 49;
  float* outc = out + (size_t)cg * 4 * 49 + (size_t)n * R * c49;
  float* stage = reinterpret_cast<float*>(smem) + wid * 2 * 196;   // two chunks per warp (double buffer)
  int buf = 0;
  for (int r = wid; r < R; r += nw) {
    float* sb = stage + buf * 196;
    // wait until the bulk store that last read this buffer is done (<= 1 group may stay in flight)
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncwarp();
    for (int v = lane; v < 196; v += 32) sb[v] = (float)(r + v);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      const unsigned sa = (unsigned)__cvta_generic_to_shared(sb);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 784;" ::"l"(outc + (size_t)r * c49), "r"(sa) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    buf ^= 1;
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void linear(float4* out, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
    __stcs(out + i, make_float4(1.f, 2.f, 3.f, 4.f));
}

int main() {
  const int N = 8, C = 512, R = 4000;
  const size_t elems = (size_t)N * R * C * 49;
  float* out;
  cudaMalloc(&out, elems * 4);
  char* flush;
  cudaMalloc(&flush, 256 << 20);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  auto run = [&](const char* name, auto fn) {
    float best = 1e9f;
    for (int it = 0; it < 6; ++it) {
      cudaMemsetAsync(flush, it, 256 << 20);
      cudaEventRecord(a);
      fn();
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      if (it > 0 && ms < best) best = ms;
    }
    printf("%-28s %.3f ms  %.0f GB/s  (%s)\n", name, best, elems * 4 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
  };
  run("linear float4 .cs", [&] { linear<<<148 * 8, 512>>>((float4*)out, elems / 4); });
  run("pattern .cs", [&] { pattern<0><<<N * 128, 1024>>>(out, C, R, 128, 0); });
  run("pattern default", [&] { pattern<1><<<N * 128, 1024>>>(out, C, R, 128, 0); });
  run("pattern .wt", [&] { pattern<2><<<N * 128, 1024>>>(out, C, R, 128, 0); });
  run("pattern .cg", [&] { pattern<3><<<N * 128, 1024>>>(out, C, R, 128, 0); });
  run("pattern .cs skewed", [&] { pattern<0><<<N * 128, 1024>>>(out, C, R, 128, 1); });
  run("pattern default skewed", [&] { pattern<1><<<N * 128, 1024>>>(out, C, R, 128, 1); });
  run("V1 warp-owned 32+17 lanes", [&] { pattern_warp<<<N * 128, 1024>>>(out, C, R, 128); });
  run("V1b warp-owned, k-major halves", [&] { pattern_warp_b<<<N * 128, 1024>>>(out, C, R, 128); });
  run("V4 contiguous 64/pass k-major", [&] { pattern_contig<0><<<N * 128, 1024>>>(out, C, R, 128); });
  run("V4 contiguous 64/pass halves adjacent", [&] { pattern_contig<1><<<N * 128, 1024>>>(out, C, R, 128); });
  run("V5 contiguous 32/pass", [&] { pattern_contig32<<<N * 128, 1024>>>(out, C, R, 128); });
  run("V6 padded 64 slots, strided passes", [&] { pattern_pad64<<<N * 128, 1024>>>(out, C, R, 128); });
  run("V7 padded, one proposal per warp pass", [&] { pattern_pad64_pair<<<N * 128, 1024>>>(out, C, R, 128); });
  run("V8 3 proposals / 5 passes, straddling", [&] { pattern_3in5<0><<<N * 128, 1024>>>(out, C, R - R % 3, 128); });
  run("V8 3 proposals / 5 passes, split stores", [&] { pattern_3in5<1><<<N * 128, 1024>>>(out, C, R - R % 3, 128); });
  run("V2 fake 48-float runs (whole sectors)", [&] { pattern48<<<N * 128, 1024>>>(out, C, R, 128); });
  run("V3 warp-owned chunk, consecutive", [&] { pattern_shuffled<<<N * 128, 1024>>>(out, C, R, 128); });
  run("chunk 784B as float4", [&] { chunk128<<<N * 128, 1024>>>(out, C, R, 128); });
  run("chunk 784B as float4 512thr", [&] { chunk128<<<N * 128, 512>>>(out, C, R, 128); });
  run("chunk 784B as scalars", [&] { chunk32<<<N * 128, 1024>>>(out, C, R, 128); });
  run("chunk 784B TMA bulk", [&] { chunk_tma<<<N * 128, 1024, 32 * 2 * 784>>>(out, C, R, 128); });
  cudaFuncSetAttribute(chunk_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  run("chunk 784B TMA bulk +150KB smem", [&] { chunk_tma<<<N * 128, 1024, 200 * 1024>>>(out, C, R, 128); });
  run("pattern .cs 512thr", [&] { pattern<0><<<N * 128, 512>>>(out, C, R, 128, 0); });
  return 0;
}
