// CSR neighbour aggregation (K1/K2/K3 of SURVEY.md section 2.3) for sm_100a.
//
// HBM/L2-bound gather: every destination row is owned by a group of G lanes (G = 2..32, so that
// narrow rows -- 47-class logits, 7-class cora -- pack several rows per warp instead of idling
// lanes), each lane holds VPL 16-byte column chunks.  A group reads G neighbour ids with one
// coalesced load, broadcasts them with warp shuffles and keeps U independent 16-byte row loads per
// lane in flight (memory-level parallelism is what hides the ~1 us DRAM latency of a random row).
// Rows whose in-degree exceeds HUB_T (power-law hubs) are deferred and then processed by the whole
// CTA: the neighbour range is cut into one segment per group, partials are combined through shared
// memory in a fixed order, so results are deterministic.
//
// The epilogue fuses everything the reference does between the aggregation and the next dense op:
// self term, 1/(deg+1) (SAGEConv "gcn"), per-destination scale (GraphConv), bias, eval-BatchNorm
// affine and ReLU -- so a project-first layer needs no further pass over Y.
#include "common.cuh"

namespace glnn {

struct SpmmArgs {
  const void* indptr;
  const int32_t* indices;
  const float* X;
  int64_t ldx;
  float* Y;
  int64_t ldy;
  int64_t n_dst;
  int d;
  int indptr64;
  int self_add;
  int mean_plus_one;
  const float* src_scale;
  const float* dst_scale;
  const float* bias;
  const float* col_scale;
  const float* col_shift;
  int relu;
};

constexpr int kWarps = 8;
constexpr int kHubT = 1024;

__device__ __forceinline__ int64_t load_ptr(const SpmmArgs& a, int64_t i) {
  return a.indptr64 ? __ldg(reinterpret_cast<const int64_t*>(a.indptr) + i)
                    : static_cast<int64_t>(__ldg(reinterpret_cast<const int32_t*>(a.indptr) + i));
}

template <int W>
__device__ __forceinline__ void load_chunk(const float* p, float (&v)[W]) {
  if constexpr (W == 4) {
    float4 t = ldg4(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
    v[0] = __ldg(p);
  }
}

template <int W>
__device__ __forceinline__ void store_chunk(float* p, const float (&v)[W]) {
  if constexpr (W == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    p[0] = v[0];
  }
}

// acc += sum over edges [beg, end) of (scale *) X[indices[e], my columns]
template <int G, int VPL, int W, bool HAS_SS>
__device__ __forceinline__ void gather_range(const SpmmArgs& a, int64_t beg, int64_t end, int gl,
                                             int lane_base, unsigned gmask, float (&acc)[VPL][W]) {
  constexpr int U = (G >= 4) ? 4 : G;
  for (int64_t base = beg; base < end; base += G) {
    const int64_t e = base + gl;
    int my = -1;
    float mys = 1.f;
    if (e < end) {
      my = __ldg(a.indices + e);
      if constexpr (HAS_SS) mys = __ldg(a.src_scale + my);
    }
    const int cnt = static_cast<int>(min(static_cast<int64_t>(G), end - base));
    for (int j = 0; j < cnt; j += U) {
      int u[U];
      float s[U];
      float v[U][VPL][W];
#pragma unroll
      for (int t = 0; t < U; ++t) {
        u[t] = __shfl_sync(gmask, my, lane_base + j + t);
        if constexpr (HAS_SS) s[t] = __shfl_sync(gmask, mys, lane_base + j + t);
      }
#pragma unroll
      for (int t = 0; t < U; ++t) {
#pragma unroll
        for (int p = 0; p < VPL; ++p) {
          const int col = (gl + p * G) * W;
          if (u[t] >= 0 && col < a.d) {
            load_chunk<W>(a.X + static_cast<int64_t>(u[t]) * a.ldx + col, v[t][p]);
          } else {
#pragma unroll
            for (int w = 0; w < W; ++w) v[t][p][w] = 0.f;
          }
        }
      }
#pragma unroll
      for (int t = 0; t < U; ++t) {
#pragma unroll
        for (int p = 0; p < VPL; ++p) {
#pragma unroll
          for (int w = 0; w < W; ++w) {
            if constexpr (HAS_SS) acc[p][w] = fmaf(v[t][p][w], s[t], acc[p][w]);
            else acc[p][w] += v[t][p][w];
          }
        }
      }
    }
  }
}

template <int G, int VPL, int W>
__device__ __forceinline__ void epilogue_store(const SpmmArgs& a, int64_t row, int64_t deg, int gl,
                                               float (&acc)[VPL][W]) {
  const float inv_den = static_cast<float>(deg + 1);
  const float ds = a.dst_scale ? __ldg(a.dst_scale + row) : 1.f;
#pragma unroll
  for (int p = 0; p < VPL; ++p) {
    const int col = (gl + p * G) * W;
    if (col >= a.d) continue;
    float self[W];
    if (a.self_add) load_chunk<W>(a.X + row * a.ldx + col, self);
    float out[W];
#pragma unroll
    for (int w = 0; w < W; ++w) {
      float x = acc[p][w];
      if (a.self_add) x += self[w];
      if (a.mean_plus_one) x = x / inv_den;
      x *= ds;
      if (a.bias) x += __ldg(a.bias + col + w);
      if (a.relu == 2) x = fmaxf(x, 0.f);
      if (a.col_scale) x = fmaf(x, __ldg(a.col_scale + col + w), __ldg(a.col_shift + col + w));
      if (a.relu == 1) x = fmaxf(x, 0.f);
      out[w] = x;
    }
    store_chunk<W>(a.Y + row * a.ldy + col, out);
  }
}

template <int G, int VPL, bool VEC, bool HAS_SS>
__global__ void __launch_bounds__(kWarps * 32) spmm_csr_kernel(const SpmmArgs a) {
  constexpr int W = VEC ? 4 : 1;
  constexpr int RPW = 32 / G;          // rows per warp
  constexpr int NG = kWarps * RPW;     // groups (= rows) per CTA
  constexpr int PW = G * VPL * W;      // padded row width held by a group
  __shared__ float s_part[NG][PW];
  __shared__ int s_hub[NG];
  __shared__ int s_nhub;

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int sub = lane / G;
  const int gl = lane % G;
  const int lane_base = sub * G;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << lane_base);
  const int gidx = warp * RPW + sub;
  const int64_t row0 = static_cast<int64_t>(blockIdx.x) * NG;

  if (threadIdx.x == 0) s_nhub = 0;
  __syncthreads();

  {
    const int64_t row = row0 + gidx;
    if (row < a.n_dst) {
      const int64_t beg = load_ptr(a, row), end = load_ptr(a, row + 1);
      if (end - beg > kHubT) {
        if (gl == 0) s_hub[atomicAdd(&s_nhub, 1)] = gidx;
      } else {
        float acc[VPL][W];
#pragma unroll
        for (int p = 0; p < VPL; ++p)
#pragma unroll
          for (int w = 0; w < W; ++w) acc[p][w] = 0.f;
        gather_range<G, VPL, W, HAS_SS>(a, beg, end, gl, lane_base, gmask, acc);
        epilogue_store<G, VPL, W>(a, row, end - beg, gl, acc);
      }
    }
  }
  __syncthreads();

  const int nhub = s_nhub;
  for (int h = 0; h < nhub; ++h) {
    const int64_t row = row0 + s_hub[h];
    const int64_t beg = load_ptr(a, row), end = load_ptr(a, row + 1);
    const int64_t deg = end - beg;
    const int64_t seg = ((deg + NG - 1) / NG + G - 1) / G * G;
    const int64_t sb = min(end, beg + gidx * seg), se = min(end, sb + seg);
    float acc[VPL][W];
#pragma unroll
    for (int p = 0; p < VPL; ++p)
#pragma unroll
      for (int w = 0; w < W; ++w) acc[p][w] = 0.f;
    gather_range<G, VPL, W, HAS_SS>(a, sb, se, gl, lane_base, gmask, acc);
#pragma unroll
    for (int p = 0; p < VPL; ++p)
#pragma unroll
      for (int w = 0; w < W; ++w) s_part[gidx][(gl + p * G) * W + w] = acc[p][w];
    __syncthreads();
    if (gidx == 0) {
#pragma unroll
      for (int p = 0; p < VPL; ++p)
#pragma unroll
        for (int w = 0; w < W; ++w) {
          float t = 0.f;
          for (int g = 0; g < NG; ++g) t += s_part[g][(gl + p * G) * W + w];
          acc[p][w] = t;
        }
      epilogue_store<G, VPL, W>(a, row, deg, gl, acc);
    }
    __syncthreads();
  }
}

template <int G, int VPL, bool VEC>
static int launch_cfg(const SpmmArgs& a, cudaStream_t st) {
  constexpr int NG = kWarps * (32 / G);
  const int64_t blocks = (a.n_dst + NG - 1) / NG;
  if (blocks == 0) return 0;
  GLNN_REQUIRE(blocks < (1LL << 31), GLNN_ERR_SHAPE, "spmm: too many rows (%lld)", (long long)a.n_dst);
  if (a.src_scale)
    spmm_csr_kernel<G, VPL, VEC, true><<<static_cast<unsigned>(blocks), kWarps * 32, 0, st>>>(a);
  else
    spmm_csr_kernel<G, VPL, VEC, false><<<static_cast<unsigned>(blocks), kWarps * 32, 0, st>>>(a);
  GLNN_LAUNCH_OK("spmm_csr_kernel");
  return 0;
}

template <bool VEC>
static int launch_width(const SpmmArgs& a, cudaStream_t st) {
  const int lanes = VEC ? (a.d + 3) / 4 : a.d;  // column chunks per row
  if (lanes <= 2) return launch_cfg<2, 1, VEC>(a, st);
  if (lanes <= 4) return launch_cfg<4, 1, VEC>(a, st);
  if (lanes <= 8) return launch_cfg<8, 1, VEC>(a, st);
  if (lanes <= 16) return launch_cfg<16, 1, VEC>(a, st);
  if (lanes <= 32) return launch_cfg<32, 1, VEC>(a, st);
  if (lanes <= 64) return launch_cfg<32, 2, VEC>(a, st);
  if (lanes <= 96) return launch_cfg<32, 3, VEC>(a, st);
  return launch_cfg<32, 4, VEC>(a, st);
}

}  // namespace glnn

extern "C" int glnn_spmm_csr_f32(const void* indptr, int indptr64, const int32_t* indices,
                                 const float* X, int64_t ldx, float* Y, int64_t ldy, int64_t n_dst,
                                 int64_t n_src, int d, int self_add, int mean_plus_one,
                                 const float* src_scale, const float* dst_scale, const float* bias,
                                 const float* col_scale, const float* col_shift, int relu,
                                 glnn_stream_t stream) {
  using namespace glnn;
  GLNN_REQUIRE(n_dst >= 0 && n_src >= 0 && d >= 0, GLNN_ERR_ARG, "spmm: negative size");
  if (n_dst == 0 || d == 0) return 0;
  GLNN_REQUIRE(indptr && X && Y, GLNN_ERR_ARG, "spmm: null indptr/X/Y");
  GLNN_REQUIRE(ldx >= d && ldy >= d, GLNN_ERR_SHAPE, "spmm: leading dimension smaller than d=%d", d);
  GLNN_REQUIRE(!self_add || n_src >= n_dst, GLNN_ERR_SHAPE,
               "spmm: self_add needs dst nodes to be a prefix of src nodes (n_src=%lld < n_dst=%lld)",
               (long long)n_src, (long long)n_dst);
  GLNN_REQUIRE((col_scale == nullptr) == (col_shift == nullptr), GLNN_ERR_ARG,
               "spmm: col_scale and col_shift must be given together");
  GLNN_REQUIRE(relu >= 0 && relu <= 2, GLNN_ERR_ARG, "spmm: relu must be 0, 1 or 2");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool vec = (d % 4 == 0) && (ldx % 4 == 0) && (ldy % 4 == 0) && aligned16(X) && aligned16(Y);
  const int chunk = vec ? 512 : 128;
  for (int c0 = 0; c0 < d; c0 += chunk) {
    SpmmArgs a;
    a.indptr = indptr;
    a.indices = indices;
    a.X = X + c0;
    a.ldx = ldx;
    a.Y = Y + c0;
    a.ldy = ldy;
    a.n_dst = n_dst;
    a.d = min(chunk, d - c0);
    a.indptr64 = indptr64;
    a.self_add = self_add;
    a.mean_plus_one = mean_plus_one;
    a.src_scale = src_scale;
    a.dst_scale = dst_scale;
    a.bias = bias ? bias + c0 : nullptr;
    a.col_scale = col_scale ? col_scale + c0 : nullptr;
    a.col_shift = col_shift ? col_shift + c0 : nullptr;
    a.relu = relu;
    int rc = vec ? launch_width<true>(a, st) : launch_width<false>(a, st);
    if (rc != 0) return rc;
  }
  return 0;
}
