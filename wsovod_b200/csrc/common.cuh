// common.cuh -- shared host/device helpers of libwsovod_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include <atomic>

#include "../../include/wsovod_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libwsovod_b200 is written for sm_100a (B200) only"
#endif

#define WSOVOD_API extern "C" __attribute__((visibility("default")))

namespace wsovod {

constexpr int kNumSMs = 148;             // B200: 2 dies x 74 SMs
constexpr int kMaxSmemOptin = 232448;    // 227 KB dynamic shared memory per CTA

extern std::atomic<uint64_t> g_launches; // kernels launched by this library (bench `gpu_launches`)

// process-wide tuning / test switches (wsovod_b200_tune): never change results, only which kernel runs
enum { TUNE_POOL_PATH = 0,    // 0 = library's choice, 1 = scan kernels, 2 = block-max planes wherever they apply
       TUNE_POOL_GROUP = 1,   // 1 = deal a pass's bins into bank-conflict-free quarter-warps (default), 0 = row-major lanes
       TUNE_ALIGN_PAIR = 2,   // 1 = CTA-pair contraction kernel for K + 1 > 256 (default), 0 = one CTA per tile
       TUNE_COUNT = 16 };
extern std::atomic<int> g_tune[TUNE_COUNT];
inline int tune(int key) { return g_tune[key].load(std::memory_order_relaxed); }

inline int after_launch() {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace wsovod
