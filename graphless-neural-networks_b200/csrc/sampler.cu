// Device-side neighbour sampling and block construction: what
// dgl.dataloading.MultiLayerNeighborSampler + NodeDataLoader do on the host for `train_sage`
// (train_and_eval.py:179-190), as kernels over the CSR that already lives in HBM.
//
//   glnn_sample_neighbors   per seed: min(in_degree, fanout) of its in-edges, uniformly WITHOUT
//                           replacement (dgl sample_neighbors(replace=False)); fanout < 0 = all
//   glnn_block_relabel      global source ids of the sampled edges -> local ids of the block
//                           (dst nodes first, models.py:105-109), through a node -> local-id map
#include <algorithm>

#include "common.cuh"

namespace glnn {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__device__ __forceinline__ int64_t ptr_at(const void* indptr, int indptr64, int64_t i) {
  return indptr64 ? reinterpret_cast<const int64_t*>(indptr)[i]
                  : static_cast<int64_t>(reinterpret_cast<const int32_t*>(indptr)[i]);
}

// counts[i] = min(deg(seed_i), fanout)  (fanout < 0: deg)
__global__ void sample_count_kernel(const void* __restrict__ indptr, int indptr64,
                                    const int64_t* __restrict__ seeds, int64_t m, int fanout,
                                    int64_t* __restrict__ counts) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int64_t v = seeds[i];
  const int64_t deg = ptr_at(indptr, indptr64, v + 1) - ptr_at(indptr, indptr64, v);
  counts[i] = (fanout < 0 || deg <= fanout) ? deg : fanout;
}

// One warp per seed.  deg <= fanout (or fanout < 0): the whole row is copied (coalesced).  Otherwise
// `fanout` DISTINCT edge positions are drawn with Floyd's algorithm (exactly uniform over the
// fanout-subsets; lane 0 draws, <= 32 draws per seed on the reference's fan-outs) and emitted in
// increasing position order, so a block row is a subsequence of the CSR row.
constexpr int kMaxFanout = 64;
__global__ void __launch_bounds__(256) sample_fill_kernel(const void* __restrict__ indptr, int indptr64,
                                                          const int32_t* __restrict__ indices,
                                                          const int64_t* __restrict__ seeds, int64_t m,
                                                          int fanout, uint64_t rng_seed,
                                                          const int64_t* __restrict__ out_ptr,
                                                          int32_t* __restrict__ out_src) {
  __shared__ int64_t s_pick[8][kMaxFanout];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 8 + warp;
  if (i >= m) return;
  const int64_t v = seeds[i];
  const int64_t beg = ptr_at(indptr, indptr64, v), deg = ptr_at(indptr, indptr64, v + 1) - beg;
  const int64_t o = out_ptr[i];
  if (fanout < 0 || deg <= fanout) {
    for (int64_t j = lane; j < deg; j += 32) out_src[o + j] = indices[beg + j];
    return;
  }
  if (lane == 0) {
    int64_t* pick = s_pick[warp];
    int cnt = 0;
    // Floyd: for j = deg - fanout .. deg - 1: t = U[0, j]; insert t if new, else j
    for (int64_t j = deg - fanout; j < deg; ++j) {
      const uint64_t r = mix64(rng_seed ^ mix64(static_cast<uint64_t>(v) * 0x9E3779B97F4A7C15ull +
                                                static_cast<uint64_t>(j - (deg - fanout)) + 1));
      const int64_t t = static_cast<int64_t>(r % static_cast<uint64_t>(j + 1));
      bool seen = false;
      for (int k = 0; k < cnt; ++k) seen |= (pick[k] == t);
      const int64_t val = seen ? j : t;
      int k = cnt++;               // insertion sort: keep the positions increasing
      while (k > 0 && pick[k - 1] > val) { pick[k] = pick[k - 1]; --k; }
      pick[k] = val;
    }
  }
  __syncwarp();
  for (int j = lane; j < fanout; j += 32) out_src[o + j] = indices[beg + s_pick[warp][j]];
}

// map[node] = local id (prepared by the caller: seeds 0..m-1, the other touched nodes m..); in place
__global__ void relabel_kernel(int32_t* __restrict__ src, int64_t total, const int32_t* __restrict__ map) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < total) src[i] = map[src[i]];
}

// flag[src[i]] = 1
__global__ void mark_kernel(const int32_t* __restrict__ src, int64_t total, uint8_t* __restrict__ flag) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < total) flag[src[i]] = 1;
}

}  // namespace glnn

extern "C" int glnn_sample_count(const void* indptr, int indptr64, const int64_t* seeds, int64_t m,
                                 int fanout, int64_t* counts, glnn_stream_t stream) {
  using namespace glnn;
  GLNN_REQUIRE(m >= 0, GLNN_ERR_ARG, "sample_count: negative m");
  if (m == 0) return 0;
  GLNN_REQUIRE(indptr && seeds && counts, GLNN_ERR_ARG, "sample_count: null pointer");
  GLNN_REQUIRE(fanout <= kMaxFanout, GLNN_ERR_SHAPE, "sample: fanout must be <= %d (or < 0 for all)", kMaxFanout);
  sample_count_kernel<<<static_cast<unsigned>((m + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      indptr, indptr64, seeds, m, fanout, counts);
  GLNN_LAUNCH_OK("sample_count_kernel");
  return 0;
}

extern "C" int glnn_sample_neighbors(const void* indptr, int indptr64, const int32_t* indices,
                                     const int64_t* seeds, int64_t m, int fanout, uint64_t rng_seed,
                                     const int64_t* out_ptr, int32_t* out_src, glnn_stream_t stream) {
  using namespace glnn;
  GLNN_REQUIRE(m >= 0, GLNN_ERR_ARG, "sample_neighbors: negative m");
  if (m == 0) return 0;
  GLNN_REQUIRE(indptr && seeds && out_ptr && (indices || true), GLNN_ERR_ARG, "sample_neighbors: null pointer");
  GLNN_REQUIRE(fanout <= kMaxFanout, GLNN_ERR_SHAPE, "sample: fanout must be <= %d (or < 0 for all)", kMaxFanout);
  const int64_t blocks = (m + 7) / 8;
  GLNN_REQUIRE(blocks < (1LL << 31), GLNN_ERR_SHAPE, "sample_neighbors: too many seeds");
  sample_fill_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      indptr, indptr64, indices, seeds, m, fanout, rng_seed, out_ptr, out_src);
  GLNN_LAUNCH_OK("sample_fill_kernel");
  return 0;
}

extern "C" int glnn_block_mark(const int32_t* src, int64_t total, uint8_t* flag, glnn_stream_t stream) {
  using namespace glnn;
  if (total <= 0) return 0;
  GLNN_REQUIRE(src && flag, GLNN_ERR_ARG, "block_mark: null pointer");
  mark_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, total, flag);
  GLNN_LAUNCH_OK("mark_kernel");
  return 0;
}

extern "C" int glnn_block_relabel(int32_t* src, int64_t total, const int32_t* map, glnn_stream_t stream) {
  using namespace glnn;
  if (total <= 0) return 0;
  GLNN_REQUIRE(src && map, GLNN_ERR_ARG, "block_relabel: null pointer");
  relabel_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, total, map);
  GLNN_LAUNCH_OK("relabel_kernel");
  return 0;
}
