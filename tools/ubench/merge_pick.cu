// merge_pick.cu -- latency of one pick of the single-warp K-way merge (det_topk), several formulations.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o merge_pick merge_pick.cu && ./merge_pick
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <algorithm>
#include <vector>
typedef unsigned long long ull;
constexpr ull kDead = ~0ull;
constexpr int RW = 3, K = 80, LEN = 100, TOPK = 100;

__global__ void redux_chain(unsigned* out, long long* cyc) {
  unsigned v = threadIdx.x * 2654435761u;
  const long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) v = __reduce_min_sync(0xffffffffu, v + threadIdx.x + i) ^ (threadIdx.x * 7u);
  const long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) cyc[0] = (t1 - t0) / 256;
}
__global__ void shfl_chain(unsigned* out, long long* cyc) {
  unsigned v = threadIdx.x * 2654435761u;
  const long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) v = __shfl_xor_sync(0xffffffffu, v + i, 1) ^ (threadIdx.x * 7u);
  const long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) cyc[0] = (t1 - t0) / 256;
}
__global__ void ballot_chain(unsigned* out, long long* cyc) {
  unsigned v = threadIdx.x * 2654435761u;
  const long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) v = __ballot_sync(0xffffffffu, (v + i) & 1) ^ (threadIdx.x * 7u);
  const long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) cyc[0] = (t1 - t0) / 256;
}

template <int MODE>
__global__ void merge(const ull* __restrict__ keys, ull* __restrict__ outk, long long* cyc) {
  __shared__ ull sel[TOPK];
  const int lane = threadIdx.x;
  int pos[RW], len[RW], off[RW];
  ull head[RW], next[RW];
#pragma unroll
  for (int j = 0; j < RW; ++j) {
    const int c = lane + j * 32;
    off[j] = c * LEN; len[j] = c < K ? LEN : 0; pos[j] = 0;
    head[j] = len[j] > 0 ? keys[off[j]] : kDead;
    next[j] = len[j] > 1 ? keys[off[j] + 1] : kDead;
  }
  ull acc = 0;
#pragma unroll
  for (int j = 0; j < RW; ++j) acc += head[j] + next[j];
  if (acc == 12345ull) printf("x");
  const long long t0 = clock64();
  for (int i = 0; i < TOPK; ++i) {
    ull best = head[0];
    int bj = 0;
#pragma unroll
    for (int j = 1; j < RW; ++j)
      if (head[j] < best) { best = head[j]; bj = j; }
    bool win;
    if (MODE == 0) {            // two REDUX
      const uint32_t hi = (uint32_t)(best >> 32), lo = (uint32_t)best;
      const uint32_t hmin = __reduce_min_sync(0xffffffffu, hi);
      const uint32_t lmin = __reduce_min_sync(0xffffffffu, hi == hmin ? lo : 0xffffffffu);
      if (hmin == 0xffffffffu && lmin == 0xffffffffu) break;
      win = hi == hmin && lo == lmin;
    } else if (MODE == 1) {     // one REDUX + ballot; second REDUX only on score ties
      const uint32_t hi = (uint32_t)(best >> 32), lo = (uint32_t)best;
      const uint32_t hmin = __reduce_min_sync(0xffffffffu, hi);
      if (hmin == 0xffffffffu) { if (__all_sync(0xffffffffu, best == kDead)) break; }
      const unsigned m = __ballot_sync(0xffffffffu, hi == hmin);
      if (m & (m - 1)) {
        const uint32_t lmin = __reduce_min_sync(0xffffffffu, hi == hmin ? lo : 0xffffffffu);
        win = hi == hmin && lo == lmin;
      } else win = hi == hmin;
    } else if (MODE == 3) {
      const uint32_t hi = (uint32_t)(best >> 32), lo = (uint32_t)best;
      const uint32_t hmin = __reduce_min_sync(0xffffffffu, hi);
      const uint32_t lmin = __reduce_min_sync(0xffffffffu, hi == hmin ? lo : 0xffffffffu);
      if (hmin == 0xffffffffu && lmin == 0xffffffffu) break;
      win = hi == hmin && lo == lmin;
    } else {                    // 64-bit butterfly with shuffles
      ull m = best;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { const ull t = __shfl_xor_sync(0xffffffffu, m, o); m = t < m ? t : m; }
      if (m == kDead) break;
      win = best == m;
    }
    if (win) {
#pragma unroll
      for (int j = 0; j < RW; ++j)
        if (j == bj) {
          ++pos[j];
          head[j] = next[j];
          if (MODE != 3) next[j] = pos[j] + 1 < len[j] ? keys[off[j] + pos[j] + 1] : kDead;
        }
      sel[i] = best;
    }
  }
  const long long t1 = clock64();
  __syncwarp();
  for (int i = lane; i < TOPK; i += 32) outk[i] = sel[i];
  if (lane == 0) cyc[0] = (t1 - t0) / TOPK;
}

__global__ void mergew(const ull* __restrict__ keys, ull* __restrict__ outk, long long* cyc) {
  __shared__ ull sel[TOPK];
  const int lane = threadIdx.x;
  int pos[RW], len[RW], off[RW];
  ull h0[RW], h1[RW], h2[RW], h3[RW];
#pragma unroll
  for (int j = 0; j < RW; ++j) {
    const int c = lane + j * 32;
    off[j] = c * LEN; len[j] = c < K ? LEN : 0; pos[j] = 4;
    h0[j] = len[j] > 0 ? keys[off[j]] : kDead;
    h1[j] = len[j] > 1 ? keys[off[j] + 1] : kDead;
    h2[j] = len[j] > 2 ? keys[off[j] + 2] : kDead;
    h3[j] = len[j] > 3 ? keys[off[j] + 3] : kDead;
  }
  ull acc = 0;
#pragma unroll
  for (int j = 0; j < RW; ++j) acc += h0[j] + h1[j] + h2[j] + h3[j];
  if (acc == 12345ull) printf("x");
  const long long t0 = clock64();
  for (int i = 0; i < TOPK; ++i) {
    ull best = h0[0];
    int bj = 0;
#pragma unroll
    for (int j = 1; j < RW; ++j)
      if (h0[j] < best) { best = h0[j]; bj = j; }
    const uint32_t hi = (uint32_t)(best >> 32), lo = (uint32_t)best;
    const uint32_t hmin = __reduce_min_sync(0xffffffffu, hi);
    const uint32_t lmin = __reduce_min_sync(0xffffffffu, hi == hmin ? lo : 0xffffffffu);
    if (hmin == 0xffffffffu && lmin == 0xffffffffu) break;
    if (hi == hmin && lo == lmin) {
#pragma unroll
      for (int j = 0; j < RW; ++j)
        if (j == bj) {
          h0[j] = h1[j]; h1[j] = h2[j]; h2[j] = h3[j]; h3[j] = kDead;
          if (h0[j] == kDead && pos[j] < len[j]) {
            const ull* src = keys + off[j] + pos[j];
            const int left = len[j] - pos[j];
            h0[j] = src[0];
            h1[j] = left > 1 ? src[1] : kDead;
            h2[j] = left > 2 ? src[2] : kDead;
            h3[j] = left > 3 ? src[3] : kDead;
            pos[j] += 4;
          }
        }
      sel[i] = best;
    }
  }
  const long long t1 = clock64();
  __syncwarp();
  for (int i = lane; i < TOPK; i += 32) outk[i] = sel[i];
  if (lane == 0) cyc[0] = (t1 - t0) / TOPK;
}
// runs' first 16 keys staged in shared memory: the pick loop only touches registers and shared memory
__global__ void mergel(const ull* __restrict__ keys, ull* __restrict__ outk, long long* cyc) {
  __shared__ ull sel[TOPK];
  __shared__ ull st[96][17];
  const int lane = threadIdx.x;
  for (int i = lane; i < 96 * 16; i += 32) { const int c = i / 16, q = i % 16; st[c][q] = c < K ? keys[c * LEN + q] : kDead; }
  __syncwarp();
  int pos[RW];
  ull head[RW];
#pragma unroll
  for (int j = 0; j < RW; ++j) { pos[j] = 0; head[j] = st[lane + j * 32][0]; }
  const long long t0 = clock64();
  for (int i = 0; i < TOPK; ++i) {
    ull best = head[0];
    int bj = 0;
#pragma unroll
    for (int j = 1; j < RW; ++j)
      if (head[j] < best) { best = head[j]; bj = j; }
    const uint32_t hi = (uint32_t)(best >> 32), lo = (uint32_t)best;
    const uint32_t hmin = __reduce_min_sync(0xffffffffu, hi);
    const uint32_t lmin = __reduce_min_sync(0xffffffffu, hi == hmin ? lo : 0xffffffffu);
    if (hmin == 0xffffffffu && lmin == 0xffffffffu) break;
    if (hi == hmin && lo == lmin) {
#pragma unroll
      for (int j = 0; j < RW; ++j)
        if (j == bj) { ++pos[j]; head[j] = pos[j] < 16 ? st[lane + j * 32][pos[j]] : kDead; }
      sel[i] = best;
    }
  }
  const long long t1 = clock64();
  __syncwarp();
  for (int i = lane; i < TOPK; i += 32) outk[i] = sel[i];
  if (lane == 0) cyc[0] = (t1 - t0) / TOPK;
}

int main() {
  std::vector<ull> h((size_t)K * LEN);
  srand(1);
  for (int c = 0; c < K; ++c) {
    std::vector<ull> r(LEN);
    for (auto& v : r) v = ((ull)(rand() & 0x7fffffff) << 32) | (unsigned)(rand() * 80 + c);
    std::sort(r.begin(), r.end());
    for (int i = 0; i < LEN; ++i) h[(size_t)c * LEN + i] = r[i];
  }
  ull *d, *o; long long* cyc; unsigned* uo;
  cudaMalloc(&d, h.size() * 8); cudaMalloc(&o, TOPK * 8); cudaMalloc(&cyc, 8); cudaMalloc(&uo, 128);
  cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  long long c;
  std::vector<ull> ref(h); std::sort(ref.begin(), ref.end());
  auto report = [&](const char* name, bool check) {
    cudaDeviceSynchronize();
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    bool ok = true;
    if (check) { std::vector<ull> got(TOPK); cudaMemcpy(got.data(), o, TOPK * 8, cudaMemcpyDeviceToHost); for (int i = 0; i < TOPK; ++i) ok &= got[i] == ref[i]; }
    printf("%-28s %lld cycles/iter %s\n", name, c, check ? (ok ? "ok" : "WRONG") : "");
  };
  for (int rep = 0; rep < 2; ++rep) {
    redux_chain<<<1, 32>>>(uo, cyc); report("REDUX.MIN dependent chain", false);
    shfl_chain<<<1, 32>>>(uo, cyc); report("SHFL dependent chain", false);
    ballot_chain<<<1, 32>>>(uo, cyc); report("VOTE dependent chain", false);
    merge<0><<<1, 32>>>(d, o, cyc); report("merge: 2 REDUX", true);
    merge<1><<<1, 32>>>(d, o, cyc); report("merge: REDUX + ballot", true);
    merge<2><<<1, 32>>>(d, o, cyc); report("merge: 64-bit butterfly", true);
    merge<3><<<1, 32>>>(d, o, cyc); report("merge: no reload (wrong)", false);
    mergew<<<1, 32>>>(d, o, cyc); report("merge: 4-key windows", true);
    mergel<<<1, 32>>>(d, o, cyc); report("merge: smem-staged runs", true);
  }
  return 0;
}
