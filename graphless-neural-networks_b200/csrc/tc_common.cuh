// Shared device helpers of the tcgen05 projection kernels (gemm_tc.cu, gemm_tall.cu): mbarriers,
// tcgen05 fences / commit / MMA, the SWIZZLE_128B shared-memory descriptor and the bf16 hi/lo split.
#pragma once
#include <cuda_bf16.h>

#include "gemm.cuh"

namespace glnn {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded spin: a protocol bug must trap instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spin = 0;; ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if (spin > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Swizzled shared-memory matrix descriptor (version 1 = Blackwell).  layout: 2 = SWIZZLE_128B
// (128-byte rows, 8-row groups 1024 bytes apart), 4 = SWIZZLE_64B (64-byte rows, groups 512 apart).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout = 2) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout) << 61;
  return d;
}

__device__ __forceinline__ void split4(const float4 v, uint2& hi, uint2& lo) {
  const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y);
  const __nv_bfloat162 h23 = __floats2bfloat162_rn(v.z, v.w);
  const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
  const __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - f01.x, v.y - f01.y);
  const __nv_bfloat162 l23 = __floats2bfloat162_rn(v.z - f23.x, v.w - f23.y);
  hi.x = *reinterpret_cast<const uint32_t*>(&h01);
  hi.y = *reinterpret_cast<const uint32_t*>(&h23);
  lo.x = *reinterpret_cast<const uint32_t*>(&l01);
  lo.y = *reinterpret_cast<const uint32_t*>(&l23);
}


// One 4-column piece of an output row in every format a consumer may want: fp32 C, bf16 hi/lo planes
// (operand of a following GEMM) and q24 (operand of a following gather).  n is a multiple of 4.
__device__ __forceinline__ void emit4(const GemmArgs& g, int64_t mr, int64_t n, const float4 v) {
  if (g.C) {
    float* cp = g.C + mr * g.ldc + n;
    if (g.vecC && n + 3 < g.N) {
      *reinterpret_cast<float4*>(cp) = v;
    } else {
      cp[0] = v.x;
      if (n + 1 < g.N) cp[1] = v.y;
      if (n + 2 < g.N) cp[2] = v.z;
      if (n + 3 < g.N) cp[3] = v.w;
    }
  }
  if (g.Ch && n + 3 < g.ldcp) {  // bf16 hi / lo planes of C (pad columns hold 0)
    uint2 hi, lo;
    split4(v, hi, lo);
    *reinterpret_cast<uint2*>(g.Ch + mr * g.ldcp + n) = hi;
    *reinterpret_cast<uint2*>(g.Cl + mr * g.ldcp + n) = lo;
  }
  if (g.Cq && n + 3 < g.N) {  // q24: round to 24 bits, hi16 block then mid8 block of the row
    const uint32_t b0 = __float_as_uint(v.x) + 0x80u, b1 = __float_as_uint(v.y) + 0x80u,
                   b2 = __float_as_uint(v.z) + 0x80u, b3 = __float_as_uint(v.w) + 0x80u;
    uint8_t* row = g.Cq + mr * g.ldcq;
    uint2 hi;
    hi.x = (b0 >> 16) | (b1 & 0xFFFF0000u);
    hi.y = (b2 >> 16) | (b3 & 0xFFFF0000u);
    *reinterpret_cast<uint2*>(row + 2 * n) = hi;
    *reinterpret_cast<uint32_t*>(row + 2 * g.N + n) =
        ((b0 >> 8) & 0xFFu) | (((b1 >> 8) & 0xFFu) << 8) | (((b2 >> 8) & 0xFFu) << 16) |
        (((b3 >> 8) & 0xFFu) << 24);
  }
}

}  // namespace tc
}  // namespace glnn
