// On-device graph construction ("on-device CSR build (create_formats_,
// COO -> CSR sort/scan, subgraph(idx_obs) for the inductive split)").  What DGL does on the host when
// the reference calls dgl.graph((src, dst)) / g.create_formats_() (dataloader.py:78,105,
// train_and_eval.py:178) and g.subgraph(idx_obs) (train_and_eval.py:324), as kernels over edge lists
// that already live in HBM.  Pure integer work, HBM-bound: no tensor cores here.
//
//   glnn_csr_from_coo   CSR over destinations, STABLE (the edges of a row keep their input order,
//                       multi-edges kept): one pass that narrows the ids to int32 and counts in / out
//                       degrees, an exclusive scan for indptr, then ceil(bits(N-1) / 8) passes of a
//                       least-significant-digit radix sort of (dst key, src payload) pairs.  A pass is
//                       per-tile digit histogram -> scan over (digit, tile) -> stable scatter; the
//                       rank of an item inside its tile comes from warp match groups in item order
//                       plus warp-private digit counters, so equal keys never swap.
//   glnn_csr_subgraph   node-induced subgraph with relabelled nodes.  The new destination is a
//                       function of the old one, so no sort is needed: a warp per old row counts,
//                       then compacts, its kept sources in order (ballot ranks).
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace glnn {
namespace csrb {

constexpr int kThreads = 256;
// items per thread of a radix tile (tile = 256 threads x items): a template parameter, chosen per call
// (GLNN_CSR_ITEMS = 4 | 8 | 16, default 8 -> 2048-item tiles)
constexpr int kMinItems = 4;
constexpr int kScanItems = 16;
constexpr int kScanTile = kThreads * kScanItems;  // 4096 elements per scan block
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ int64_t ptr_at(const void* indptr, int indptr64, int64_t i) {
  return indptr64 ? reinterpret_cast<const int64_t*>(indptr)[i]
                  : static_cast<int64_t>(reinterpret_cast<const int32_t*>(indptr)[i]);
}

// ---------------------------------------------------------------------------------------------
// exclusive scan of int32 (totals stay below 2^31: they count edges)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(kFull, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// exclusive prefix of v over the 256 threads of the block; *total = block sum.  s_warp: 8 ints.
__device__ __forceinline__ int block_excl_scan(int v, int* s_warp, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int inc = warp_incl_scan(v, lane);
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  int before = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) {
    const int c = s_warp[w];
    if (w < warp) before += c;
    tot += c;
  }
  __syncthreads();  // s_warp may be reused by the caller
  *total = tot;
  return before + inc - v;
}

__global__ void __launch_bounds__(kThreads) scan_reduce_kernel(const int32_t* __restrict__ in, int64_t n,
                                                               int32_t* __restrict__ bsum) {
  __shared__ int s_warp[kThreads / 32];
  const int64_t base = static_cast<int64_t>(blockIdx.x) * kScanTile;
  int v = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    const int64_t j = base + i * kThreads + threadIdx.x;
    if (j < n) v += in[j];
  }
  int tot;
  block_excl_scan(v, s_warp, &tot);
  if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}

// one block: in-place exclusive scan of the block sums (a carry runs over 256-element pieces)
__global__ void __launch_bounds__(kThreads) scan_bsums_kernel(int32_t* __restrict__ bsum, int64_t nb) {
  __shared__ int s_warp[kThreads / 32];
  int carry = 0;
  for (int64_t base = 0; base < nb; base += kThreads) {
    const int64_t j = base + threadIdx.x;
    const int v = j < nb ? bsum[j] : 0;
    int tot;
    const int ex = block_excl_scan(v, s_warp, &tot);
    if (j < nb) bsum[j] = carry + ex;
    carry += tot;
  }
}

// out may alias in: every thread holds its 16 inputs in registers before anything is written
__global__ void __launch_bounds__(kThreads) scan_apply_kernel(const int32_t* in, int64_t n,
                                                              const int32_t* __restrict__ bsum,
                                                              int32_t* out) {
  __shared__ int s_warp[kThreads / 32];
  const int64_t base = static_cast<int64_t>(blockIdx.x) * kScanTile +
                       static_cast<int64_t>(threadIdx.x) * kScanItems;
  int x[kScanItems];
  int v = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    x[i] = base + i < n ? in[base + i] : 0;
    v += x[i];
  }
  int tot;
  int run = bsum[blockIdx.x] + block_excl_scan(v, s_warp, &tot);
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    if (base + i < n) out[base + i] = run;
    run += x[i];
  }
}

// exclusive scan of in[0..n) into out (may alias); bsum: scratch of ceil(n / 4096) ints
static int exclusive_scan(const int32_t* in, int32_t* out, int64_t n, int32_t* bsum, cudaStream_t st) {
  if (n <= 0) return 0;
  const int64_t nb = (n + kScanTile - 1) / kScanTile;
  GLNN_REQUIRE(nb < (1LL << 31), GLNN_ERR_SHAPE, "csr build: scan too long");
  scan_reduce_kernel<<<static_cast<unsigned>(nb), kThreads, 0, st>>>(in, n, bsum);
  GLNN_LAUNCH_OK("scan_reduce_kernel");
  scan_bsums_kernel<<<1, kThreads, 0, st>>>(bsum, nb);
  GLNN_LAUNCH_OK("scan_bsums_kernel");
  scan_apply_kernel<<<static_cast<unsigned>(nb), kThreads, 0, st>>>(in, n, bsum, out);
  GLNN_LAUNCH_OK("scan_apply_kernel");
  return 0;
}

// ---------------------------------------------------------------------------------------------
// COO -> CSR
// ---------------------------------------------------------------------------------------------
// key = dst, val = src as int32; in / out degree counters; ids outside [0, N) are counted in *status
// and replaced by node 0 so that every later kernel stays in bounds (the caller raises on status).
template <typename T>
__global__ void __launch_bounds__(kThreads) coo_prepare_kernel(const T* __restrict__ src,
                                                               const T* __restrict__ dst, int64_t E,
                                                               int64_t N, int32_t* __restrict__ key,
                                                               int32_t* __restrict__ val,
                                                               int32_t* __restrict__ in_cnt,
                                                               int32_t* __restrict__ out_cnt,
                                                               int32_t* __restrict__ status) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; e < E; e += stride) {
    int64_t s = static_cast<int64_t>(src[e]), d = static_cast<int64_t>(dst[e]);
    if (s < 0 || s >= N || d < 0 || d >= N) {
      atomicAdd(status, 1);
      s = 0;
      d = 0;
    }
    key[e] = static_cast<int32_t>(d);
    val[e] = static_cast<int32_t>(s);
    atomicAdd(in_cnt + d, 1);
    if (out_cnt) atomicAdd(out_cnt + s, 1);
  }
}

__global__ void widen_kernel(const int32_t* __restrict__ in, int64_t n, int64_t* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}

// digit histogram of one tile, stored digit-major: hist[digit * nb + tile]
template <int kItems>
__global__ void __launch_bounds__(kThreads) radix_hist_kernel(const int32_t* __restrict__ key, int64_t E,
                                                              int shift, int32_t* __restrict__ hist,
                                                              int64_t nb) {
  __shared__ int h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = static_cast<int64_t>(blockIdx.x) * (kThreads * kItems);
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    const int64_t j = base + i * kThreads + threadIdx.x;
    if (j < E) atomicAdd(&h[(key[j] >> shift) & 255], 1);
  }
  __syncthreads();
  hist[static_cast<int64_t>(threadIdx.x) * nb + blockIdx.x] = h[threadIdx.x];
}

// Stable scatter of one tile.  Item order inside the tile: warp w owns 32 * kItems consecutive items,
// iteration i covers 32 consecutive ones, lane = position.  Rank of an item
// among the tile's items with the same digit = (count in earlier warps) + (count in this warp's earlier
// iterations) + (lower lanes of its match group).  goff = exclusive scan of hist (digit-major), i.e.
// the first output slot of (digit, tile).  out_key == nullptr on the last pass.
template <int kItems, bool HW_MATCH>
__global__ void __launch_bounds__(kThreads) radix_scatter_kernel(
    const int32_t* __restrict__ key, const int32_t* __restrict__ val, int64_t E, int shift,
    const int32_t* __restrict__ goff, int64_t nb, int32_t* __restrict__ out_key,
    int32_t* __restrict__ out_val) {
  __shared__ int cnt[kThreads / 32][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) cnt[w][threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = static_cast<int64_t>(blockIdx.x) * (kThreads * kItems) + warp * (32 * kItems);
  int k[kItems], v[kItems], rank[kItems];
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    const int64_t j = base + i * 32 + lane;
    const bool valid = j < E;
    k[i] = valid ? key[j] : 0;
    v[i] = valid ? val[j] : 0;
    const unsigned d = valid ? static_cast<unsigned>((k[i] >> shift) & 255) : 0xffffffffu;
    unsigned grp;  // lanes holding the same digit (invalid lanes form their own group)
    if constexpr (HW_MATCH) {
      grp = __match_any_sync(kFull, d);
    } else {
      // eight ballots, one per digit bit, refine the peer mask; the MATCH instruction serialises over
      // the ~28 distinct digits of a warp in the MIO pipe (ncu: mio_throttle + short_scoreboard were
      // the top stalls of this kernel), votes do not
      grp = __ballot_sync(kFull, valid);
      if (!valid) grp = ~grp;
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const bool bit = (d >> b) & 1u;
        const unsigned m = __ballot_sync(kFull, bit);
        grp &= bit ? m : ~m;
      }
    }
    const unsigned lower = grp & ((1u << lane) - 1u);
    int old = 0;
    if (valid && lower == 0) {  // lowest lane of the group advances the warp's counter
      old = cnt[warp][d];
      cnt[warp][d] = old + __popc(grp);
    }
    old = __shfl_sync(kFull, old, __ffs(grp) - 1);
    rank[i] = old + __popc(lower);
    __syncwarp();  // the counter update is visible to the next iteration's group leaders
  }
  __syncthreads();
  {  // thread t = digit t: first slot of (digit, tile), then running offsets per warp
    int run = goff[static_cast<int64_t>(threadIdx.x) * nb + blockIdx.x];
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) {
      const int c = cnt[w][threadIdx.x];
      cnt[w][threadIdx.x] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    const int64_t j = base + i * 32 + lane;
    if (j < E) {
      const int pos = cnt[warp][(k[i] >> shift) & 255] + rank[i];
      if (out_key) out_key[pos] = k[i];
      out_val[pos] = v[i];
    }
  }
}

static int radix_passes(int64_t N) {
  int bits = 0;
  while (bits < 31 && (1LL << bits) < N) ++bits;  // ids are < N <= 2^bits
  return (bits + 7) / 8;
}

struct Workspace {
  int32_t *key_a, *key_b, *val_x, *in_cnt, *out_cnt, *hist, *bsum;
  size_t bytes;
};

static size_t up256(size_t b) { return (b + 255) & ~static_cast<size_t>(255); }

static int items_per_thread() {
  static const int items = []() {
    const char* e = getenv("GLNN_CSR_ITEMS");
    const int v = e ? atoi(e) : 8;
    return (v == 4 || v == 16) ? v : 8;
  }();
  return items;
}

// sized for the smallest tile, so that every tile choice fits
static Workspace carve(void* base, int64_t N, int64_t E) {
  Workspace w;
  const int64_t min_tile = kThreads * kMinItems;
  const int64_t nb = (E + min_tile - 1) / min_tile;
  const int64_t scan_n = std::max<int64_t>(256 * nb, N + 1);
  uint8_t* p = static_cast<uint8_t*>(base);
  size_t off = 0;
  auto take = [&](size_t b) {
    uint8_t* r = p ? p + off : nullptr;
    off += up256(b);
    return reinterpret_cast<int32_t*>(r);
  };
  w.key_a = take(sizeof(int32_t) * E);
  w.key_b = take(sizeof(int32_t) * E);
  w.val_x = take(sizeof(int32_t) * E);
  w.in_cnt = take(sizeof(int32_t) * (N + 1));
  w.out_cnt = take(sizeof(int32_t) * N);
  w.hist = take(sizeof(int32_t) * 256 * nb);
  w.bsum = take(sizeof(int32_t) * ((scan_n + kScanTile - 1) / kScanTile + 1));
  w.bytes = off;
  return w;
}

// ---------------------------------------------------------------------------------------------
// node-induced subgraph
// ---------------------------------------------------------------------------------------------
// warp per old row v with relabel[v] >= 0: new_cnt[relabel[v]] = kept sources of the row
__global__ void __launch_bounds__(kThreads) subgraph_count_kernel(const void* __restrict__ indptr,
                                                                  int indptr64,
                                                                  const int32_t* __restrict__ indices,
                                                                  int64_t N,
                                                                  const int32_t* __restrict__ relabel,
                                                                  int32_t* __restrict__ new_cnt) {
  const int lane = threadIdx.x & 31;
  const int64_t v = static_cast<int64_t>(blockIdx.x) * (kThreads / 32) + (threadIdx.x >> 5);
  if (v >= N) return;
  const int r = relabel[v];
  if (r < 0) return;
  const int64_t beg = ptr_at(indptr, indptr64, v), end = ptr_at(indptr, indptr64, v + 1);
  int c = 0;
  for (int64_t j = beg + lane; j < end; j += 32) c += relabel[indices[j]] >= 0 ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(kFull, c, o);
  if (lane == 0) new_cnt[r] = c;
}

__global__ void __launch_bounds__(kThreads) subgraph_fill_kernel(const void* __restrict__ indptr,
                                                                 int indptr64,
                                                                 const int32_t* __restrict__ indices,
                                                                 int64_t N,
                                                                 const int32_t* __restrict__ relabel,
                                                                 const int32_t* __restrict__ new_indptr,
                                                                 int32_t* __restrict__ new_indices,
                                                                 int32_t* __restrict__ out_cnt) {
  // new_indices == nullptr: out-degrees only
  const int lane = threadIdx.x & 31;
  const int64_t v = static_cast<int64_t>(blockIdx.x) * (kThreads / 32) + (threadIdx.x >> 5);
  if (v >= N) return;
  const int r = relabel[v];
  if (r < 0) return;
  const int64_t beg = ptr_at(indptr, indptr64, v), end = ptr_at(indptr, indptr64, v + 1);
  int o = new_indptr[r];
  for (int64_t b = beg; b < end; b += 32) {  // whole warp iterates together: ballots stay converged
    const int64_t j = b + lane;
    const int s = j < end ? relabel[indices[j]] : -1;
    const unsigned kept = __ballot_sync(kFull, s >= 0);
    if (s >= 0) {
      if (new_indices) new_indices[o + __popc(kept & ((1u << lane) - 1u))] = s;
      if (out_cnt) atomicAdd(out_cnt + s, 1);
    }
    o += __popc(kept);
  }
}

}  // namespace csrb
}  // namespace glnn

extern "C" size_t glnn_csr_build_workspace_bytes(int64_t n_nodes, int64_t n_edges) {
  if (n_nodes < 0 || n_edges < 0) return 0;
  return glnn::csrb::carve(nullptr, n_nodes, n_edges).bytes;
}

extern "C" int glnn_csr_from_coo(const void* src, const void* dst, int idx64, int64_t n_edges,
                                 int64_t n_nodes, int32_t* indptr, int32_t* indices, int64_t* out_deg,
                                 int32_t* status, void* workspace, size_t ws_bytes,
                                 glnn_stream_t stream) {
  using namespace glnn;
  using namespace glnn::csrb;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GLNN_REQUIRE(n_nodes >= 0 && n_edges >= 0, GLNN_ERR_ARG, "csr_from_coo: negative size");
  GLNN_REQUIRE(n_edges < (1LL << 31) - 4096 && n_nodes < (1LL << 31) - 1, GLNN_ERR_SHAPE,
               "csr_from_coo: int32 CSR needs fewer than 2^31 edges and nodes");
  GLNN_REQUIRE(indptr && status, GLNN_ERR_ARG, "csr_from_coo: null output");
  GLNN_REQUIRE(n_edges == 0 || n_nodes > 0, GLNN_ERR_SHAPE, "csr_from_coo: edges without nodes");
  GLNN_CUDA_OK(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
  if (n_edges == 0) {
    GLNN_CUDA_OK(cudaMemsetAsync(indptr, 0, sizeof(int32_t) * (n_nodes + 1), st));
    if (out_deg && n_nodes) GLNN_CUDA_OK(cudaMemsetAsync(out_deg, 0, sizeof(int64_t) * n_nodes, st));
    return 0;
  }
  GLNN_REQUIRE(src && dst && indices && workspace, GLNN_ERR_ARG, "csr_from_coo: null pointer");
  const Workspace w = carve(workspace, n_nodes, n_edges);
  GLNN_REQUIRE(ws_bytes >= w.bytes, GLNN_ERR_SHAPE, "csr_from_coo: workspace too small (%zu < %zu)",
               ws_bytes, w.bytes);
  GLNN_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, GLNN_ERR_ALIGN,
               "csr_from_coo: workspace must be 256-byte aligned");
  const int passes = radix_passes(n_nodes);
  const int items = items_per_thread();
  static const bool hw_match = getenv("GLNN_CSR_MATCH") != nullptr && atoi(getenv("GLNN_CSR_MATCH")) == 1;
  const int64_t tile = static_cast<int64_t>(kThreads) * items;
  const int64_t nb = (n_edges + tile - 1) / tile;
  // payload ping-pong between the workspace buffer X and the output `indices` (Y), arranged so that
  // the last pass writes Y: an odd number of passes starts in X, an even number in Y
  int32_t* val_cur = (passes & 1) ? w.val_x : indices;
  int32_t* val_nxt = (passes & 1) ? indices : w.val_x;
  int32_t* key_cur = w.key_a;
  int32_t* key_nxt = w.key_b;

  GLNN_CUDA_OK(cudaMemsetAsync(w.in_cnt, 0, sizeof(int32_t) * (n_nodes + 1), st));
  if (out_deg) GLNN_CUDA_OK(cudaMemsetAsync(w.out_cnt, 0, sizeof(int32_t) * n_nodes, st));
  const unsigned pgrid =
      static_cast<unsigned>(std::min<int64_t>((n_edges + kThreads - 1) / kThreads, 16LL * sm_count()));
  if (idx64)
    coo_prepare_kernel<int64_t><<<pgrid, kThreads, 0, st>>>(
        static_cast<const int64_t*>(src), static_cast<const int64_t*>(dst), n_edges, n_nodes, key_cur,
        val_cur, w.in_cnt, out_deg ? w.out_cnt : nullptr, status);
  else
    coo_prepare_kernel<int32_t><<<pgrid, kThreads, 0, st>>>(
        static_cast<const int32_t*>(src), static_cast<const int32_t*>(dst), n_edges, n_nodes, key_cur,
        val_cur, w.in_cnt, out_deg ? w.out_cnt : nullptr, status);
  GLNN_LAUNCH_OK("coo_prepare_kernel");
  int rc = exclusive_scan(w.in_cnt, indptr, n_nodes + 1, w.bsum, st);  // in_cnt[N] = 0 -> indptr[N] = E
  if (rc != 0) return rc;
  if (out_deg) {
    widen_kernel<<<static_cast<unsigned>((n_nodes + 255) / 256), 256, 0, st>>>(w.out_cnt, n_nodes, out_deg);
    GLNN_LAUNCH_OK("widen_kernel");
  }
  for (int p = 0; p < passes; ++p) {
    const int shift = 8 * p;
    const bool last = p == passes - 1;
    const unsigned grid = static_cast<unsigned>(nb);
    if (items == 4) radix_hist_kernel<4><<<grid, kThreads, 0, st>>>(key_cur, n_edges, shift, w.hist, nb);
    else if (items == 16) radix_hist_kernel<16><<<grid, kThreads, 0, st>>>(key_cur, n_edges, shift, w.hist, nb);
    else radix_hist_kernel<8><<<grid, kThreads, 0, st>>>(key_cur, n_edges, shift, w.hist, nb);
    GLNN_LAUNCH_OK("radix_hist_kernel");
    rc = exclusive_scan(w.hist, w.hist, 256 * nb, w.bsum, st);
    if (rc != 0) return rc;
    int32_t* ok = last ? nullptr : key_nxt;
#define GLNN_SCATTER(IT, HW) \
  radix_scatter_kernel<IT, HW><<<grid, kThreads, 0, st>>>(key_cur, val_cur, n_edges, shift, w.hist, nb, ok, val_nxt)
    if (hw_match) {
      if (items == 4) GLNN_SCATTER(4, true);
      else if (items == 16) GLNN_SCATTER(16, true);
      else GLNN_SCATTER(8, true);
    } else {
      if (items == 4) GLNN_SCATTER(4, false);
      else if (items == 16) GLNN_SCATTER(16, false);
      else GLNN_SCATTER(8, false);
    }
#undef GLNN_SCATTER
    GLNN_LAUNCH_OK("radix_scatter_kernel");
    std::swap(key_cur, key_nxt);
    std::swap(val_cur, val_nxt);
  }
  // passes == 0 (a single node): every edge is 0 -> 0 and val_cur == indices already holds them
  return 0;
}

extern "C" int glnn_csr_subgraph(const void* indptr, int indptr64, const int32_t* indices,
                                 int64_t n_nodes, const int32_t* relabel, int64_t n_new,
                                 int32_t* new_indptr, int32_t* new_indices, int64_t* new_out_deg,
                                 void* workspace, size_t ws_bytes, glnn_stream_t stream) {
  using namespace glnn;
  using namespace glnn::csrb;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GLNN_REQUIRE(n_nodes >= 0 && n_new >= 0 && n_new <= n_nodes, GLNN_ERR_ARG, "csr_subgraph: bad sizes");
  GLNN_REQUIRE(n_nodes < (1LL << 31) - 1, GLNN_ERR_SHAPE, "csr_subgraph: too many nodes");
  GLNN_REQUIRE(new_indptr, GLNN_ERR_ARG, "csr_subgraph: null output");
  // workspace: new_cnt [n_new + 1], out_cnt [n_new], scan block sums
  const size_t need = up256(sizeof(int32_t) * (n_new + 1)) + up256(sizeof(int32_t) * n_new) +
                      up256(sizeof(int32_t) * ((n_new + 1 + kScanTile - 1) / kScanTile + 1));
  if (n_new == 0 || n_nodes == 0) {
    GLNN_CUDA_OK(cudaMemsetAsync(new_indptr, 0, sizeof(int32_t) * (n_new + 1), st));
    if (new_out_deg && n_new) GLNN_CUDA_OK(cudaMemsetAsync(new_out_deg, 0, sizeof(int64_t) * n_new, st));
    return 0;
  }
  GLNN_REQUIRE(indptr && relabel && workspace, GLNN_ERR_ARG, "csr_subgraph: null pointer");
  GLNN_REQUIRE(ws_bytes >= need, GLNN_ERR_SHAPE, "csr_subgraph: workspace too small (%zu < %zu)", ws_bytes, need);
  GLNN_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, GLNN_ERR_ALIGN,
               "csr_subgraph: workspace must be 256-byte aligned");
  uint8_t* p = static_cast<uint8_t*>(workspace);
  int32_t* new_cnt = reinterpret_cast<int32_t*>(p);
  int32_t* out_cnt = reinterpret_cast<int32_t*>(p + up256(sizeof(int32_t) * (n_new + 1)));
  int32_t* bsum = reinterpret_cast<int32_t*>(p + up256(sizeof(int32_t) * (n_new + 1)) +
                                             up256(sizeof(int32_t) * n_new));
  GLNN_CUDA_OK(cudaMemsetAsync(new_cnt, 0, sizeof(int32_t) * (n_new + 1), st));
  GLNN_CUDA_OK(cudaMemsetAsync(out_cnt, 0, sizeof(int32_t) * n_new, st));
  const int64_t blocks = (n_nodes + kThreads / 32 - 1) / (kThreads / 32);
  subgraph_count_kernel<<<static_cast<unsigned>(blocks), kThreads, 0, st>>>(indptr, indptr64, indices,
                                                                            n_nodes, relabel, new_cnt);
  GLNN_LAUNCH_OK("subgraph_count_kernel");
  const int rc = exclusive_scan(new_cnt, new_indptr, n_new + 1, bsum, st);
  if (rc != 0) return rc;
  if (new_indices || new_out_deg) {  // second phase (a subgraph without edges has no index array)
    subgraph_fill_kernel<<<static_cast<unsigned>(blocks), kThreads, 0, st>>>(
        indptr, indptr64, indices, n_nodes, relabel, new_indptr, new_indices,
        new_out_deg ? out_cnt : nullptr);
    GLNN_LAUNCH_OK("subgraph_fill_kernel");
    if (new_out_deg) {
      widen_kernel<<<static_cast<unsigned>((n_new + 255) / 256), 256, 0, st>>>(out_cnt, n_new, new_out_deg);
      GLNN_LAUNCH_OK("widen_kernel");
    }
  }
  return 0;
}
