// tc_ptx.cuh -- PTX wrappers shared by the tcgen05 kernels (align_tc.cu: one CTA per tile; align_tc2.cu: CTA pairs):
// mbarriers with bounded waits, TMA tile loads, tcgen05 fences / commit / MMA / TMEM loads, the UMMA shared-memory
// descriptor of a 128B-swizzled K-major operand.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace wsovod {

constexpr int TC_BM = 128;          // rows per tile (UMMA M)
constexpr int TC_BK = 32;           // fp32 elements per stage row = 128 B = one swizzle atom
constexpr int TC_THREADS = 256;
constexpr int TC_SPIN_LIMIT = 1 << 26;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done = 0;
  for (int spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(a), "r"(parity) : "memory");
    if (spin > TC_SPIN_LIMIT) __trap();   // a protocol bug must abort the launch, never hang the GPU
  }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, void* smem, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// UMMA shared-memory descriptor: K-major operand, 128B swizzle, rows of 128 B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3ffff) >> 4);        // start address, bits [0,14)
  d |= (uint64_t)0 << 16;                              // leading byte offset (unused: one atom along K)
  d |= (uint64_t)(1024 >> 4) << 32;                    // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                              // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                              // layout type: SWIZZLE_128B
  return d;
}

// host: 2-D tensor map over a row-major fp32 matrix, box = [box_rows x 32 floats], 128B swizzle (align_tc.cu)
int tc_make_map(CUtensorMap* m, const float* base, uint64_t inner, uint64_t rows, uint64_t row_stride_elems,
                uint32_t box_rows);
// CTA-pair contraction for K + 1 > 256 (align_tc2.cu)
int align_tc2_launch(const float* x, const float* what, int64_t M, int64_t D, int64_t KO, int64_t Kp, int64_t Dp,
                     float temperature, int norm, const float* bias, float* logits, float* probs, int* tickets,
                     float* rowstat, cudaStream_t st);
int64_t align_tc2_tickets(int64_t M);
// probs = row softmax of logits (in place allowed), rows held in registers up to 2048 columns (align_tc2.cu)
int softmax_rows_launch(const float* logits, int64_t M, int64_t KO, float* probs, cudaStream_t st);

}  // namespace wsovod
