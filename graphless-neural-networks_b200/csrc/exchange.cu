// Layer-embedding exchange of the dst-row sharded teacher: every rank pushes
// its slab of a replica chunk straight into the same offset of every peer's replica over
// NVLink 5 / NVSwitch (peer-mapped symmetric memory), from a small SM-driven kernel: one 16-byte load
// of the local slab feeds up to 7 posted 16-byte peer stores.  Round 1 used the copy engines for this
// (G-1 cudaMemcpyPeerAsync per chunk): 0.37-0.45 TB/s per rank at N=8 while the HBM-saturating gather
// ran, against ~0.75 TB/s per direction that kernel stores reach on this fabric -- and the exchange
// was the scaling limit (3.9 of 9.75 ms exposed).  A few CTAs suffice: the kernel is bound by the
// links, not by the SMs, and the aggregation of the next chunk keeps the remaining SMs busy.
#include <algorithm>

#include "common.cuh"

namespace glnn {

constexpr int kMaxPush = 8;
struct PushArgs {
  const uint4* src;
  uint4* dst[kMaxPush];
  int n_dst;
  int64_t n16;  // 16-byte units
};

__device__ __forceinline__ uint4 ld_stream16(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

__global__ void __launch_bounds__(512) peer_push_kernel(const PushArgs a) {
  constexpr int U = 4;  // loads in flight per thread
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < a.n16; i += U * stride) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = ld_stream16(a.src + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u)
      for (int p = 0; p < a.n_dst; ++p) a.dst[p][i + u * stride] = v[u];
  }
  for (; i < a.n16; i += stride) {
    const uint4 v = ld_stream16(a.src + i);
    for (int p = 0; p < a.n_dst; ++p) a.dst[p][i] = v;
  }
}

}  // namespace glnn

extern "C" int glnn_peer_push(const void* src, void* const* dst, int n_dst, int64_t bytes, int ctas,
                              glnn_stream_t stream) {
  using namespace glnn;
  GLNN_REQUIRE(n_dst >= 0 && n_dst <= kMaxPush && bytes >= 0, GLNN_ERR_ARG,
               "peer_push: 0..%d destinations, bytes >= 0", kMaxPush);
  if (n_dst == 0 || bytes == 0) return 0;
  GLNN_REQUIRE(src && dst, GLNN_ERR_ARG, "peer_push: null pointer");
  GLNN_REQUIRE(bytes % 16 == 0 && aligned16(src), GLNN_ERR_ALIGN, "peer_push: 16-byte granularity");
  PushArgs a{};
  a.src = static_cast<const uint4*>(src);
  a.n_dst = n_dst;
  a.n16 = bytes / 16;
  for (int p = 0; p < n_dst; ++p) {
    GLNN_REQUIRE(dst[p] && aligned16(dst[p]), GLNN_ERR_ALIGN, "peer_push: destination %d", p);
    a.dst[p] = static_cast<uint4*>(dst[p]);
  }
  const int64_t want = (a.n16 + 512 * 4 - 1) / (512 * 4);
  const unsigned grid = static_cast<unsigned>(std::max<int64_t>(1, std::min<int64_t>(want, ctas > 0 ? ctas : 32)));
  peer_push_kernel<<<grid, 512, 0, static_cast<cudaStream_t>(stream)>>>(a);
  GLNN_LAUNCH_OK("peer_push_kernel");
  return 0;
}
