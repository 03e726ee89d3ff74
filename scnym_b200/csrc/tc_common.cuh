// tcgen05 / TMEM / mbarrier / bulk-copy helpers shared by the tensor-core kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

#define TC_M 128                    // rows per MMA (TMEM lanes)
// Two tile geometries (template parameters HALVES, KT of the kernel):
//   <1, 64>: 128 rows per CTA, 64 genes per tile (128-byte rows, SWIZZLE_128B)
//   <2, 32>: 256 rows per CTA (two accumulators in TMEM), 32 genes per tile (64-byte rows, SWIZZLE_64B).
//            Every W1T tile feeds two MMA row blocks, which halves the W1T bytes each SM pulls through L2
//            per MMA cycle -- at <1,64> the 148 SMs together ask for ~12 TB/s, the L2 limit.
#define TC_W 8                      // entries per prefetch window of a densify thread
#define TC_SPIN_LIMIT (1u << 28)    // watchdog: trap instead of hanging the GPU if a barrier never completes

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!ok && ++spins > TC_SPIN_LIMIT) __trap();
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major operand tile with dense rows of ROWB = 2*KT bytes (128 -> SWIZZLE_128B, 64 -> SWIZZLE_64B):
// LBO = 16 B (unused for swizzled K-major), SBO = 8 rows, descriptor version 1 (Blackwell)
template <int KT>
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t smem_addr) {
  constexpr uint32_t ROWB = 2 * KT;
  constexpr uint32_t LAYOUT = (ROWB == 128) ? 2u : 4u;
  const uint32_t lo = ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16);
  const uint32_t hi = ((8u * ROWB) >> 4) | (1u << 14) | (LAYOUT << 29);
  return (uint64_t)lo | ((uint64_t)hi << 32);
}
// byte offset of 16-byte chunk `c` of row `r` inside a swizzled tile (relative to the row start)
template <int KT>
__device__ __forceinline__ uint32_t swz_chunk(uint32_t c, uint32_t r) {
  return (KT == 64) ? ((c ^ (r & 7u)) << 4) : ((c ^ ((r >> 1) & 3u)) << 4);
}
__device__ __forceinline__ void st_shared_zero16(uint32_t addr) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "r"(0u) : "memory");
}
__device__ __forceinline__ void st_shared_u16(uint32_t addr, uint16_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}


// hi = bf16(x), lo = bf16(x - hi) for two values at once, packed as {x0 in the low half, x1 in the high half}.  The packed
// conversion (F2FP.BF16.F32.PACK_AB) issues at ALU rate; the scalar F2F.BF16.F32 runs on the quarter-rate conversion pipe
// and made the operand staging of the tcgen05 GEMM conversion-bound.
__device__ __forceinline__ void split_bf16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(x0 - h0, x1 - h1);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
