// CSR neighbour aggregation (K1/K2/K3 of SURVEY.md section 2.3) for sm_100a.
//
// HBM/L2-bound gather: every destination row is owned by a group of G lanes (G = 2..32, so that
// narrow rows -- 47-class logits, 7-class cora -- pack several rows per warp instead of idling
// lanes), each lane holds VPL 16-byte column chunks.  A group reads G neighbour ids with one
// coalesced load, broadcasts them with warp shuffles and keeps U independent 16-byte row loads per
// lane in flight (memory-level parallelism is what hides the ~1 us DRAM latency of a random row).
//
// Power-law hubs: a row above kHubT in-edges would pin one warp (or one CTA) for milliseconds --
// invisible when the whole kernel runs 18 ms on one GPU, but the critical path once the rows are
// sharded over 4-8 GPUs (the shard that owns the hubs ran 2.6x longer than the others).  The main
// kernel therefore only REGISTERS such rows: it cuts them into kHubSeg-edge tasks in a device task
// list; a second, persistent kernel drains the list with every SM (each task is split over the
// groups of a CTA, reduced through shared memory and atomically added into an fp32 scratch row);
// a third, tiny kernel applies the epilogue to the hub rows.  If the scratch capacity is ever
// exceeded the CTA processes the row itself, so the result is always complete.
//
// The epilogue fuses everything the reference does between the aggregation and the next dense op:
// self term, 1/(deg+1) (SAGEConv "gcn"), per-destination scale (GraphConv), bias, eval-BatchNorm
// affine and ReLU -- and can emit the row as bf16 hi/lo planes, the operand format of the
// tensor-core projection that follows an aggregate-first layer.
#include <cuda_bf16.h>

#include <algorithm>
#include <map>
#include <mutex>

#include "common.cuh"

// Resident CTAs per SM of the narrow (<= 8 values per lane) kernels: 3 (85 registers, no spills, 24
// warps) measured 3-5 % faster than 4 (64 registers, ~200 B of spills in the gather loop) and 8 %
// faster than 2 on the q24 rows of the products forward.
#ifndef GLNN_SPMM_MINB
#define GLNN_SPMM_MINB 3
#endif

namespace glnn {

struct HubTask {
  int64_t beg;
  int32_t len;
  int32_t slot;
};

struct SpmmArgs {
  const void* indptr;
  const int32_t* indices;
  const float* X;
  int64_t ldx;
  const uint8_t* Xq;  // optional q24 input instead of X: row = dq x hi16 then dq x mid8 (dq = d
                      // rounded up to 8), rows ldq bytes apart
  int64_t ldq;
  int dq;
  float* Y;        // fp32 output (may be null when planes are written)
  int64_t ldy;
  uint16_t* Yh;    // optional bf16 hi / lo plane output
  uint16_t* Yl;
  int64_t ldyp;
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
  int log_softmax;       // epilogue ends with log_softmax over the first d_valid columns (scalar stores)
  int d_valid;
  // L2 residency hints: gathered rows of source ids < hot_below are loaded evict_last, all other
  // gathered rows, the index stream and the output stream evict_first (0 = no hints)
  int hot_below;
  // optional initial value of the accumulators: acc[v, :] starts from Yinit[v, :] (fp32, ldyi) instead
  // of zero -- the second pass of an aggregation that was split by source block (dist_teacher.py)
  const float* Yinit;
  int64_t ldyi;
  // optional sparse copy of the gathered matrix ("s24", see glnn_compact_s24): row = `cap` words
  // [fp32 bits 31..8 | column 7..0] sorted by column, zero-padded, rows lds words apart; *cap_dev =
  // max non-zeros per row.  Used by spmm_csr_s24_kernel only; Xq stays the self / hub / dense source.
  const uint32_t* Xs;
  int64_t lds;
  const int* cap_dev;
  int cap_limit;         // sparse path only when *cap_dev <= cap_limit
  // hub scratch (library-owned, per device and stream)
  int* hub_ctr;          // [0] tasks registered, [1] hub rows registered, [2] next task to run
  HubTask* hub_tasks;
  int64_t* hub_rows;     // row id per slot
  float* hub_acc;        // [cap_rows][kHubAccLd]
  int cap_tasks, cap_rows;
};

constexpr int kWarps = 8;
constexpr int kHubT = 1024;      // rows above this many in-edges go through the task list
constexpr int kHubSeg = 2048;    // edges per task
constexpr int kHubAccLd = 512;   // widest column chunk of one launch
constexpr int kCapTasks = 1 << 16;
constexpr int kCapRows = 1 << 13;

__device__ __forceinline__ int64_t load_ptr(const SpmmArgs& a, int64_t i) {
  return a.indptr64 ? __ldg(reinterpret_cast<const int64_t*>(a.indptr) + i)
                    : static_cast<int64_t>(__ldg(reinterpret_cast<const int32_t*>(a.indptr) + i));
}

// L2 eviction policies.  The gather of a power-law graph re-reads a small set of hub rows many times
// while most rows are touched once per pass; marking the hub rows evict_last and every streaming
// access evict_first keeps the hubs resident in the 126 MB L2 instead of letting the stream of cold
// rows push them out.
struct Policies {
  uint64_t hot, cold;
};
__device__ __forceinline__ Policies make_policies(bool hints) {
  Policies p;
  if (hints) {
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p.hot));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p.cold));
  } else {
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p.hot));
    p.cold = p.hot;
  }
  return p;
}
__device__ __forceinline__ uint4 ld16(const void* p, uint64_t pol) {
  uint4 r;
  asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ uint2 ld8(const void* p, uint64_t pol) {
  uint2 r;
  asm volatile("ld.global.nc.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;"
               : "=r"(r.x), "=r"(r.y)
               : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ uint32_t ld4(const void* p, uint64_t pol) {
  uint32_t r;
  asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ void st16(void* p, const uint4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void st8(void* p, const uint2 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v2.u32 [%0], {%1,%2}, %3;" ::"l"(p), "r"(v.x), "r"(v.y),
               "l"(pol)
               : "memory");
}

// q24 chunk: 8 values from 16 bytes of hi16 + 8 bytes of mid8, value = (hi16 << 16 | mid8 << 8).
__device__ __forceinline__ void decode_q24(const uint4 h, const uint2 m, float (&v)[8]) {
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
  const uint32_t mw[2] = {m.x, m.y};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t hi = (i & 1) ? (hw[i >> 1] & 0xFFFF0000u) : (hw[i >> 1] << 16);
    const uint32_t mid = ((mw[i >> 2] >> ((i & 3) * 8)) & 0xFFu) << 8;
    v[i] = __uint_as_float(hi | mid);
  }
}

// W consecutive columns of source row `r`: W = 4 / 1 from the fp32 matrix, W = 8 from a q24 matrix.
template <int W>
__device__ __forceinline__ void load_chunk(const SpmmArgs& a, int64_t r, int col, float (&v)[W],
                                           uint64_t pol) {
  if constexpr (W == 8) {
    const uint8_t* row = a.Xq + r * a.ldq;
    decode_q24(ld16(row + 2 * col, pol), ld8(row + 2 * a.dq + col, pol), v);
  } else if constexpr (W == 4) {
    const uint4 t = ld16(a.X + r * a.ldx + col, pol);
    v[0] = __uint_as_float(t.x); v[1] = __uint_as_float(t.y);
    v[2] = __uint_as_float(t.z); v[3] = __uint_as_float(t.w);
  } else {
    v[0] = __uint_as_float(ld4(a.X + r * a.ldx + col, pol));
  }
}

template <bool PLANES = true>
__device__ __forceinline__ void store4(const SpmmArgs& a, int64_t row, int col, const float* v,
                                       uint64_t pol) {
  if (a.Y)
    st16(a.Y + row * a.ldy + col,
         make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]),
                    __float_as_uint(v[3])),
         pol);
  if (PLANES && a.Yh) {
    const __nv_bfloat162 h01 = __floats2bfloat162_rn(v[0], v[1]), h23 = __floats2bfloat162_rn(v[2], v[3]);
    const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
    const __nv_bfloat162 l01 = __floats2bfloat162_rn(v[0] - f01.x, v[1] - f01.y);
    const __nv_bfloat162 l23 = __floats2bfloat162_rn(v[2] - f23.x, v[3] - f23.y);
    uint2 uh, ul;
    uh.x = *reinterpret_cast<const uint32_t*>(&h01); uh.y = *reinterpret_cast<const uint32_t*>(&h23);
    ul.x = *reinterpret_cast<const uint32_t*>(&l01); ul.y = *reinterpret_cast<const uint32_t*>(&l23);
    st8(a.Yh + row * a.ldyp + col, uh, pol);
    st8(a.Yl + row * a.ldyp + col, ul, pol);
  }
}

template <int W>
__device__ __forceinline__ void store_chunk(const SpmmArgs& a, int64_t row, int col,
                                            const float (&v)[W], uint64_t pol) {
  if constexpr (W == 8) {
    // one 16-byte store per plane (two 8-byte halves of a sector issued by different instructions
    // reached DRAM as partial writes: ncu showed 4.5 GB written for 2.5 GB of planes)
    if (a.Y) {
      store4<false>(a, row, col, v, pol);
      store4<false>(a, row, col + 4, v + 4, pol);
    }
    if (a.Yh) {
      uint32_t h[4], l[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        const float2 f = __bfloat1622float2(hh);
        const __nv_bfloat162 ll = __floats2bfloat162_rn(v[2 * i] - f.x, v[2 * i + 1] - f.y);
        h[i] = *reinterpret_cast<const uint32_t*>(&hh);
        l[i] = *reinterpret_cast<const uint32_t*>(&ll);
      }
      st16(a.Yh + row * a.ldyp + col, make_uint4(h[0], h[1], h[2], h[3]), pol);
      st16(a.Yl + row * a.ldyp + col, make_uint4(l[0], l[1], l[2], l[3]), pol);
    }
  } else if constexpr (W == 4) {
    store4(a, row, col, v, pol);
  } else {
    if (a.Y) a.Y[row * a.ldy + col] = v[0];
    if (a.Yh) {
      const __nv_bfloat16 h = __float2bfloat16_rn(v[0]);
      const __nv_bfloat16 l = __float2bfloat16_rn(v[0] - __bfloat162float(h));
      a.Yh[row * a.ldyp + col] = *reinterpret_cast<const uint16_t*>(&h);
      a.Yl[row * a.ldyp + col] = *reinterpret_cast<const uint16_t*>(&l);
    }
  }
}

// acc += sum over edges [beg, end) of (scale *) X[indices[e], my columns]
template <int G, int VPL, int W, bool HAS_SS>
__device__ __forceinline__ void gather_range(const SpmmArgs& a, int64_t beg, int64_t end, int gl,
                                             int lane_base, unsigned gmask, const Policies& pol,
                                             float (&acc)[VPL][W]) {
  // neighbours in flight per group iteration: q24 rows are 25 % smaller and their decode costs
  // registers, so keep 8 of them in flight as RAW words (6 registers each) and decode afterwards
  constexpr int U = (W == 8) ? (G >= 8 ? 8 : G) : ((G >= 4) ? 4 : G);
  // (requesting the ids of the next batch ahead of the current rows was measured and does not help:
  // 14.8 / 8.4 / 4.4 ms with vs 14.9 / 8.3 / 4.4 ms without at d = 256 / 100 / 48 -- the kernel is
  // DRAM-bound, not latency-bound)
  for (int64_t base = beg; base < end; base += G) {
    const int my = (base + gl < end) ? static_cast<int>(ld4(a.indices + base + gl, pol.cold)) : -1;
    float mys = 1.f;
    if constexpr (HAS_SS) {
      if (my >= 0) mys = __ldg(a.src_scale + my);
    }
    const int cnt = static_cast<int>(min(static_cast<int64_t>(G), end - base));
    for (int j = 0; j < cnt; j += U) {
      int u[U];
      float s[U];
#pragma unroll
      for (int t = 0; t < U; ++t) {
        u[t] = __shfl_sync(gmask, my, lane_base + j + t);
        if constexpr (HAS_SS) s[t] = __shfl_sync(gmask, mys, lane_base + j + t);
      }
      if constexpr (W == 8) {
        uint4 hw[U][VPL];
        uint2 mw[U][VPL];
#pragma unroll
        for (int t = 0; t < U; ++t) {
          const uint64_t pl = (u[t] < a.hot_below) ? pol.hot : pol.cold;
#pragma unroll
          for (int p = 0; p < VPL; ++p) {
            const int col = (gl + p * G) * 8;
            if (u[t] >= 0 && col < a.d) {
              const uint8_t* row = a.Xq + static_cast<int64_t>(u[t]) * a.ldq;
              hw[t][p] = ld16(row + 2 * col, pl);
              mw[t][p] = ld8(row + 2 * a.dq + col, pl);
            } else {
              hw[t][p] = make_uint4(0u, 0u, 0u, 0u);
              mw[t][p] = make_uint2(0u, 0u);
            }
          }
        }
#pragma unroll
        for (int t = 0; t < U; ++t) {
#pragma unroll
          for (int p = 0; p < VPL; ++p) {
            float x[8];
            decode_q24(hw[t][p], mw[t][p], x);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if constexpr (HAS_SS) acc[p][i] = fmaf(x[i], s[t], acc[p][i]);
              else acc[p][i] += x[i];
            }
          }
        }
      } else {
        float v[U][VPL][W];
#pragma unroll
        for (int t = 0; t < U; ++t) {
          const uint64_t pl = (u[t] < a.hot_below) ? pol.hot : pol.cold;
#pragma unroll
          for (int p = 0; p < VPL; ++p) {
            const int col = (gl + p * G) * W;
            if (u[t] >= 0 && col < a.d) {
              load_chunk<W>(a, static_cast<int64_t>(u[t]), col, v[t][p], pl);
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
}

template <int G>
__device__ __forceinline__ float group_max(float v, unsigned gmask) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(gmask, v, o));
  return v;
}
template <int G>
__device__ __forceinline__ float group_sum(float v, unsigned gmask) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
  return v;
}

// Must be called by every lane of the group that owns `row` (the log-softmax tail reduces over it).
template <int G, int VPL, int W>
__device__ __forceinline__ void epilogue_store(const SpmmArgs& a, int64_t row, int64_t deg, int gl,
                                               unsigned gmask, const Policies& pol,
                                               float (&acc)[VPL][W]) {
  const float inv_den = static_cast<float>(deg + 1);
  const float ds = a.dst_scale ? __ldg(a.dst_scale + row) : 1.f;
#pragma unroll
  for (int p = 0; p < VPL; ++p) {
    const int col = (gl + p * G) * W;
    if (col >= a.d) continue;
    float self[W];
    if (a.self_add) load_chunk<W>(a, row, col, self, pol.cold);
    if (a.Yinit) {
#pragma unroll
      for (int w = 0; w < W; ++w)
        if (col + w < a.d) acc[p][w] += __ldg(a.Yinit + row * a.ldyi + col + w);
    }
#pragma unroll
    for (int w = 0; w < W; ++w) {
      float x = acc[p][w];
      if (a.self_add) x += self[w];
      if (a.mean_plus_one) x = x / inv_den;
      x *= ds;
      if (col + w < a.d) {  // a q24 row ends in up to 7 zero pad columns: they stay zero
        if (a.bias) x += __ldg(a.bias + col + w);
        if (a.relu == 2) x = fmaxf(x, 0.f);
        if (a.col_scale) x = fmaf(x, __ldg(a.col_scale + col + w), __ldg(a.col_shift + col + w));
        if (a.relu == 1) x = fmaxf(x, 0.f);
      } else {
        x = 0.f;
      }
      acc[p][w] = x;
    }
  }
  if (a.log_softmax) {
    // evaluate()'s log_softmax (train_and_eval.py:98) over the first d_valid columns, fused: the
    // whole row lives in this group's registers
    float m = -INFINITY;
#pragma unroll
    for (int p = 0; p < VPL; ++p)
#pragma unroll
      for (int w = 0; w < W; ++w)
        if ((gl + p * G) * W + w < a.d_valid) m = fmaxf(m, acc[p][w]);
    m = group_max<G>(m, gmask);
    float sum = 0.f;
#pragma unroll
    for (int p = 0; p < VPL; ++p)
#pragma unroll
      for (int w = 0; w < W; ++w)
        if ((gl + p * G) * W + w < a.d_valid) sum += expf(acc[p][w] - m);
    sum = group_sum<G>(sum, gmask);
    const float lse = m + logf(sum);
    float* y = a.Y + row * a.ldy;
#pragma unroll
    for (int p = 0; p < VPL; ++p)
#pragma unroll
      for (int w = 0; w < W; ++w) {
        const int c = (gl + p * G) * W + w;
        if (c < a.d_valid) y[c] = acc[p][w] - lse;
      }
    return;
  }
#pragma unroll
  for (int p = 0; p < VPL; ++p) {
    const int col = (gl + p * G) * W;
    if (col >= a.d) continue;
    store_chunk<W>(a, row, col, acc[p], pol.cold);
  }
}

template <int VPL, int W>
__device__ __forceinline__ void zero_acc(float (&acc)[VPL][W]) {
#pragma unroll
  for (int p = 0; p < VPL; ++p)
#pragma unroll
    for (int w = 0; w < W; ++w) acc[p][w] = 0.f;
}

// Whole-CTA processing of one edge range: every group takes a fixed sub-range, partials are combined
// through shared memory in group order; group 0 ends up with the sum in `acc`.
template <int G, int VPL, int W, bool HAS_SS, int NG>
__device__ __forceinline__ void cta_gather(const SpmmArgs& a, int64_t beg, int64_t end, int gidx, int gl,
                                           int lane_base, unsigned gmask, const Policies& pol,
                                           float (*s_part)[G * VPL * W], float (&acc)[VPL][W]) {
  const int64_t len = end - beg;
  const int64_t seg = ((len + NG - 1) / NG + G - 1) / G * G;
  const int64_t sb = min(end, beg + gidx * seg), se = min(end, sb + seg);
  zero_acc<VPL, W>(acc);
  gather_range<G, VPL, W, HAS_SS>(a, sb, se, gl, lane_base, gmask, pol, acc);
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
  }
}

template <int G, int VPL, int W, bool HAS_SS>
__global__ void __launch_bounds__(kWarps * 32, (VPL * W <= 8 ? GLNN_SPMM_MINB : 1)) spmm_csr_kernel(const SpmmArgs a) {
  constexpr int RPW = 32 / G;          // rows per warp
  constexpr int NG = kWarps * RPW;     // groups (= rows) per CTA
  __shared__ float s_part[NG][G * VPL * W];
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
  const Policies pol = make_policies(a.hot_below > 0);

  if (threadIdx.x == 0) s_nhub = 0;
  __syncthreads();

  {
    const int64_t row = row0 + gidx;
    if (row < a.n_dst) {
      const int64_t beg = load_ptr(a, row), end = load_ptr(a, row + 1);
      const int64_t deg = end - beg;
      if (deg > kHubT) {
        // register the row in the device task list (one lane), fall back to this CTA when full
        int slot = -1;
        if (gl == 0) {
          const int ntask = static_cast<int>((deg + kHubSeg - 1) / kHubSeg);
          const int s = atomicAdd(a.hub_ctr + 1, 1);
          if (s < a.cap_rows) {
            const int t0 = atomicAdd(a.hub_ctr + 0, ntask);
            if (t0 + ntask <= a.cap_tasks) {
              slot = s;
              a.hub_rows[s] = row;
              for (int t = 0; t < ntask; ++t) {
                HubTask tk;
                tk.beg = beg + static_cast<int64_t>(t) * kHubSeg;
                tk.len = static_cast<int32_t>(min(static_cast<int64_t>(kHubSeg), end - tk.beg));
                tk.slot = s;
                a.hub_tasks[t0 + t] = tk;
              }
            } else {
              a.hub_rows[s] = -1;  // slot burnt: the finish kernel skips it
              for (int t = t0; t < min(t0 + ntask, a.cap_tasks); ++t) {
                HubTask tk;
                tk.beg = 0; tk.len = 0; tk.slot = s;  // empty filler so the drain sees valid entries
                a.hub_tasks[t] = tk;
              }
            }
          }
        }
        slot = __shfl_sync(gmask, slot, lane_base);
        if (slot >= 0) {
          for (int c = gl; c < kHubAccLd; c += G)
            a.hub_acc[static_cast<int64_t>(slot) * kHubAccLd + c] = 0.f;
        } else if (gl == 0) {
          s_hub[atomicAdd(&s_nhub, 1)] = gidx;
        }
      } else {
        float acc[VPL][W];
        zero_acc<VPL, W>(acc);
        gather_range<G, VPL, W, HAS_SS>(a, beg, end, gl, lane_base, gmask, pol, acc);
        epilogue_store<G, VPL, W>(a, row, deg, gl, gmask, pol, acc);
      }
    }
  }
  __syncthreads();

  const int nhub = s_nhub;  // only when the task list overflowed
  for (int h = 0; h < nhub; ++h) {
    const int64_t row = row0 + s_hub[h];
    const int64_t beg = load_ptr(a, row), end = load_ptr(a, row + 1);
    float acc[VPL][W];
    cta_gather<G, VPL, W, HAS_SS, NG>(a, beg, end, gidx, gl, lane_base, gmask, pol, s_part, acc);
    if (gidx == 0) epilogue_store<G, VPL, W>(a, row, end - beg, gl, gmask, pol, acc);
    __syncthreads();
  }
}

// Persistent drain of the hub task list: one task per CTA iteration.
template <int G, int VPL, int W, bool HAS_SS>
__global__ void __launch_bounds__(kWarps * 32) spmm_hub_kernel(const SpmmArgs a) {
  constexpr int RPW = 32 / G;
  constexpr int NG = kWarps * RPW;
  __shared__ float s_part[NG][G * VPL * W];
  __shared__ int s_task;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / G, gl = lane % G, lane_base = sub * G;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << lane_base);
  const int gidx = warp * RPW + sub;
  const Policies pol = make_policies(a.hot_below > 0);
  const int ntasks = min(a.hub_ctr[0], a.cap_tasks);
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_task = atomicAdd(a.hub_ctr + 2, 1);
    __syncthreads();
    const int t = s_task;
    if (t >= ntasks) break;
    const HubTask tk = a.hub_tasks[t];
    float acc[VPL][W];
    cta_gather<G, VPL, W, HAS_SS, NG>(a, tk.beg, tk.beg + tk.len, gidx, gl, lane_base, gmask, pol,
                                      s_part, acc);
    if (gidx == 0 && tk.len > 0) {
#pragma unroll
      for (int p = 0; p < VPL; ++p)
#pragma unroll
        for (int w = 0; w < W; ++w) {
          const int col = (gl + p * G) * W + w;
          if (col < a.d) atomicAdd(a.hub_acc + static_cast<int64_t>(tk.slot) * kHubAccLd + col, acc[p][w]);
        }
    }
  }
}

// Epilogue of the registered hub rows (one group per row).
template <int G, int VPL, int W>
__global__ void __launch_bounds__(kWarps * 32) spmm_hub_finish_kernel(const SpmmArgs a) {
  constexpr int RPW = 32 / G;
  constexpr int NG = kWarps * RPW;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / G, gl = lane % G, lane_base = sub * G;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << lane_base);
  const int gidx = warp * RPW + sub;
  const Policies pol = make_policies(a.hot_below > 0);
  const int nrows = min(a.hub_ctr[1], a.cap_rows);
  for (int s = blockIdx.x * NG + gidx; s < nrows; s += gridDim.x * NG) {
    const int64_t row = a.hub_rows[s];
    if (row < 0) continue;
    const int64_t beg = load_ptr(a, row), end = load_ptr(a, row + 1);
    float acc[VPL][W];
#pragma unroll
    for (int p = 0; p < VPL; ++p)
#pragma unroll
      for (int w = 0; w < W; ++w) {
        const int col = (gl + p * G) * W + w;
        acc[p][w] = col < a.d ? a.hub_acc[static_cast<int64_t>(s) * kHubAccLd + col] : 0.f;
      }
    epilogue_store<G, VPL, W>(a, row, end - beg, gl, gmask, pol, acc);
  }
}

// ---- sparse rows ("s24") for post-ReLU embeddings: EXPERIMENTAL, opt-in ---------------------------
// Written at the end of round 1 without a GPU run (tests/test_zz_next_rows_gpu.py holds its parity
// test as a non-strict xfail); nothing on the default path calls it.  DESIGN.md section 8 item 3.
//
// q24 -> s24: one warp per row.  Lane l decodes its 8 columns, keeps the non-zeros as
// [bits 31..8 | column], the warp packs them (prefix sum over the lanes) through shared memory and
// writes the whole row -- entries, then zeros up to dq words -- with 16-byte stores.  The per-matrix
// capacity (max non-zeros of a row) is reduced per CTA and published with one atomicMax.
constexpr int kS24MaxD = 256;  // the column id has 8 bits

__global__ void __launch_bounds__(kWarps * 32) compact_s24_kernel(const uint8_t* __restrict__ Q,
                                                                  int64_t ldq, int64_t rows, int dq,
                                                                  uint32_t* __restrict__ S,
                                                                  int64_t lds, int* __restrict__ cap) {
  __shared__ __align__(16) uint32_t s_row[kWarps][kS24MaxD];
  __shared__ int s_max;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lds_row = (dq + 31) / 32 * 32;  // words written per row (entries, then zeros)
  if (threadIdx.x == 0) s_max = 0;
  __syncthreads();
  int local_max = 0;
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * kWarps + warp; row < rows;
       row += static_cast<int64_t>(gridDim.x) * kWarps) {
    const int col = lane * 8;
    uint32_t wd[8];
    uint32_t mask = 0u;  // bit i: column col + i is non-zero
    if (col < dq) {
      const uint8_t* r = Q + row * ldq;
      float x[8];
      decode_q24(*reinterpret_cast<const uint4*>(r + 2 * col),
                 *reinterpret_cast<const uint2*>(r + 2 * dq + col), x);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t b = __float_as_uint(x[i]);
        wd[i] = b | static_cast<uint32_t>(col + i);
        if ((b << 1) != 0u) mask |= 1u << i;  // skips +0.0 and -0.0
      }
    }
    const int cnt = __popc(mask);
    int pre = cnt;  // inclusive prefix sum over the lanes
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, pre, o);
      if (lane >= o) pre += t;
    }
    const int total = __shfl_sync(0xffffffffu, pre, 31);
    const int base = pre - cnt;
    for (int i = lane; i < lds_row; i += 32) s_row[warp][i] = 0u;
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if ((mask >> i) & 1u) s_row[warp][base + __popc(mask & ((1u << i) - 1u))] = wd[i];
    __syncwarp();
    uint32_t* out = S + row * lds;
    for (int i = lane * 4; i < lds_row; i += 128)
      *reinterpret_cast<uint4*>(out + i) = *reinterpret_cast<const uint4*>(&s_row[warp][i]);
    __syncwarp();
    local_max = max(local_max, total);
  }
  if (lane == 0 && local_max > 0) atomicMax(&s_max, local_max);
  __syncthreads();
  if (threadIdx.x == 0 && s_max > 0) atomicMax(cap, s_max);
}

// acc_s[0..255] += sum over edges [beg, end) of the sparse rows.  Lanes take the entries
// lane + 32 k of a row (columns of one instruction are nearly consecutive: few bank conflicts);
// U rows are loaded before any is accumulated; zero words (padding) are skipped, so no two lanes of
// an instruction ever touch the same accumulator (the columns of a row are distinct).
__device__ __forceinline__ void gather_range_s24(const SpmmArgs& a, int64_t beg, int64_t end, int lane,
                                                 int cap8, const Policies& pol,
                                                 float* __restrict__ acc_s) {
  constexpr int U = 4, KMAX = kS24MaxD / 32;
  for (int64_t base = beg; base < end; base += 32) {
    const int my = (base + lane < end) ? static_cast<int>(ld4(a.indices + base + lane, pol.cold)) : -1;
    const int cnt = static_cast<int>(min(static_cast<int64_t>(32), end - base));
    for (int j = 0; j < cnt; j += U) {
      uint32_t w[U][KMAX];
#pragma unroll
      for (int t = 0; t < U; ++t) {
        const int u = __shfl_sync(0xffffffffu, my, (j + t) & 31);
        const bool live = (j + t < cnt) && u >= 0;
        const uint32_t* r = a.Xs + static_cast<int64_t>(live ? u : 0) * a.lds + lane;
#pragma unroll
        for (int k = 0; k < KMAX; ++k)  // whole 32-byte sectors up to the capacity, nothing beyond
          w[t][k] = (live && lane + 32 * k < cap8) ? ld4(r + 32 * k, pol.cold) : 0u;
      }
#pragma unroll
      for (int t = 0; t < U; ++t) {
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
          const uint32_t v = w[t][k];
          if (v != 0u) acc_s[v & 0xffu] += __uint_as_float(v & 0xffffff00u);
        }
        __syncwarp();  // the next row's entries of a column may sit in another lane
      }
    }
  }
}

// spmm_csr_kernel<32, 1, 8, false> with the non-hub rows gathered from the sparse copy when the
// matrix is sparse enough (*cap_dev <= cap_limit); hubs, the self term and dense matrices use q24.
__global__ void __launch_bounds__(kWarps * 32, GLNN_SPMM_MINB) spmm_csr_s24_kernel(const SpmmArgs a) {
  constexpr int G = 32, VPL = 1, W = 8, NG = kWarps;
  __shared__ float s_part[NG][G * VPL * W];
  __shared__ int s_hub[NG];
  __shared__ int s_nhub;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gl = lane, lane_base = 0, gidx = warp;
  const unsigned gmask = 0xffffffffu;
  const int64_t row0 = static_cast<int64_t>(blockIdx.x) * NG;
  const Policies pol = make_policies(a.hot_below > 0);
  const int cap = *a.cap_dev;
  const bool sparse = cap <= a.cap_limit;
  const int cap8 = (cap + 7) & ~7;  // entries read per row: the capacity rounded up to a 32-byte sector
  if (threadIdx.x == 0) s_nhub = 0;
  __syncthreads();
  {
    const int64_t row = row0 + gidx;
    if (row < a.n_dst) {
      const int64_t beg = load_ptr(a, row), end = load_ptr(a, row + 1);
      const int64_t deg = end - beg;
      if (deg > kHubT) {  // identical to spmm_csr_kernel: register the row, the q24 hub kernels finish it
        int slot = -1;
        if (gl == 0) {
          const int ntask = static_cast<int>((deg + kHubSeg - 1) / kHubSeg);
          const int s = atomicAdd(a.hub_ctr + 1, 1);
          if (s < a.cap_rows) {
            const int t0 = atomicAdd(a.hub_ctr + 0, ntask);
            if (t0 + ntask <= a.cap_tasks) {
              slot = s;
              a.hub_rows[s] = row;
              for (int t = 0; t < ntask; ++t) {
                HubTask tk;
                tk.beg = beg + static_cast<int64_t>(t) * kHubSeg;
                tk.len = static_cast<int32_t>(min(static_cast<int64_t>(kHubSeg), end - tk.beg));
                tk.slot = s;
                a.hub_tasks[t0 + t] = tk;
              }
            } else {
              a.hub_rows[s] = -1;
              for (int t = t0; t < min(t0 + ntask, a.cap_tasks); ++t) {
                HubTask tk;
                tk.beg = 0; tk.len = 0; tk.slot = s;
                a.hub_tasks[t] = tk;
              }
            }
          }
        }
        slot = __shfl_sync(gmask, slot, lane_base);
        if (slot >= 0) {
          for (int c = gl; c < kHubAccLd; c += G)
            a.hub_acc[static_cast<int64_t>(slot) * kHubAccLd + c] = 0.f;
        } else if (gl == 0) {
          s_hub[atomicAdd(&s_nhub, 1)] = gidx;
        }
      } else {
        float acc[VPL][W];
        if (sparse) {
          float* acc_s = s_part[gidx];  // this warp's 256 accumulators
#pragma unroll
          for (int i = 0; i < W; ++i) acc_s[lane * W + i] = 0.f;
          __syncwarp();
          gather_range_s24(a, beg, end, lane, cap8, pol, acc_s);
#pragma unroll
          for (int i = 0; i < W; ++i) acc[0][i] = acc_s[lane * W + i];
          __syncwarp();
        } else {
          zero_acc<VPL, W>(acc);
          gather_range<G, VPL, W, false>(a, beg, end, gl, lane_base, gmask, pol, acc);
        }
        epilogue_store<G, VPL, W>(a, row, deg, gl, gmask, pol, acc);
      }
    }
  }
  __syncthreads();
  const int nhub = s_nhub;  // only when the task list overflowed
  for (int h = 0; h < nhub; ++h) {
    const int64_t row = row0 + s_hub[h];
    const int64_t beg = load_ptr(a, row), end = load_ptr(a, row + 1);
    float acc[VPL][W];
    cta_gather<G, VPL, W, false, NG>(a, beg, end, gidx, gl, lane_base, gmask, pol, s_part, acc);
    if (gidx == 0) epilogue_store<G, VPL, W>(a, row, end - beg, gl, gmask, pol, acc);
    __syncthreads();
  }
}

// ---- EXPERIMENT (tools/exp_spmm_tma.py): neighbour rows pulled by TMA bulk copies -----------------
// north_star names TMA for the neighbour pull.  This kernel is the A/B partner of spmm_csr_kernel
// <32,1,8,false> on 256-wide q24 rows: one warp per destination row, lane 0 issues one
// cp.async.bulk (global -> shared, mbarrier complete_tx) per neighbour ROW into a per-warp ring of
// STAGES row buffers, the warp waits on the row's mbarrier, decodes its 8 columns from shared memory
// and accumulates.  No hub path: it is only run on graphs without hubs.  Measured result and verdict:
// profiles/README.md (r2 "TMA gather A/B") and DESIGN.md section 4.1.
__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
template <int STAGES>
__global__ void __launch_bounds__(kWarps * 32) spmm_tma_q24_kernel(const SpmmArgs a) {
  extern __shared__ __align__(128) uint8_t s_ring[];   // [kWarps][STAGES][ldq]
  __shared__ __align__(8) uint64_t s_bar[kWarps][STAGES];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * kWarps + warp;
  const uint32_t row_bytes = static_cast<uint32_t>(a.ldq);
  uint8_t* ring = s_ring + static_cast<size_t>(warp) * STAGES * row_bytes;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&s_bar[warp][s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (row >= a.n_dst) return;
  const Policies pol = make_policies(false);
  const int64_t beg = load_ptr(a, row), end = load_ptr(a, row + 1);
  const int64_t deg = end - beg;
  float acc[1][8];
  zero_acc<1, 8>(acc);
  auto issue = [&](int64_t src, int stage) {  // lane 0 only
    const uint32_t bar = smem_addr(&s_bar[warp][stage]);
    const uint32_t dst = smem_addr(ring + static_cast<size_t>(stage) * row_bytes);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(row_bytes) : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(dst), "l"(a.Xq + src * a.ldq), "r"(row_bytes), "r"(bar)
        : "memory");
  };
  // ids are fetched 32 at a time (one coalesced load); the edge counter picks the ring stage and the
  // barrier parity
  int64_t issued = 0, done = 0;
  // the id batch of the ISSUE pointer (it runs at most STAGES edges ahead of the consume pointer)
  int my_i = -1;
  int64_t my_i_base = -32;
  auto id_issue = [&](int64_t j) -> int {
    const int64_t b = j & ~int64_t(31);
    if (b != my_i_base) {
      my_i_base = b;
      my_i = (beg + b + lane < end) ? static_cast<int>(ld4(a.indices + beg + b + lane, pol.cold)) : -1;
    }
    return __shfl_sync(0xffffffffu, my_i, static_cast<int>(j & 31));
  };
  for (; issued < deg && issued < STAGES; ++issued) {
    const int u = id_issue(issued);
    if (lane == 0) issue(u, static_cast<int>(issued % STAGES));
  }
  for (; done < deg; ++done) {
    const int stage = static_cast<int>(done % STAGES);
    const uint32_t parity = static_cast<uint32_t>((done / STAGES) & 1);
    const uint32_t bar = smem_addr(&s_bar[warp][stage]);
    uint32_t ok = 0;
    while (!ok) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok)
          : "r"(bar), "r"(parity)
          : "memory");
    }
    const uint8_t* r = ring + static_cast<size_t>(stage) * row_bytes;
    const uint4 h = *reinterpret_cast<const uint4*>(r + 16 * lane);
    const uint2 m = *reinterpret_cast<const uint2*>(r + 2 * a.dq + 8 * lane);
    float x[8];
    decode_q24(h, m, x);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[0][i] += x[i];
    __syncwarp();   // every lane has read the stage before it is refilled
    if (issued < deg) {
      const int u = id_issue(issued);
      if (lane == 0) issue(u, stage);
      ++issued;
    }
  }
  epilogue_store<32, 1, 8>(a, row, deg, lane, 0xffffffffu, pol, acc);
}

int spmm_run_tma_exp(const glnn_spmm_desc& q, int stages, cudaStream_t st) {
  const int d = q.d, dq = (d + 7) / 8 * 8;
  GLNN_REQUIRE(q.X_q24 && !q.X && dq == 256 && q.ldq == 768 && aligned16(q.X_q24), GLNN_ERR_SHAPE,
               "spmm_tma_exp: 256-wide q24 rows of 768 bytes only");
  GLNN_REQUIRE(q.indptr && q.indices && (q.Y || q.Y_hi) && !q.src_scale && !q.log_softmax, GLNN_ERR_ARG,
               "spmm_tma_exp: unsupported arguments");
  GLNN_REQUIRE(stages == 4 || stages == 8, GLNN_ERR_ARG, "spmm_tma_exp: stages must be 4 or 8");
  if (q.n_dst == 0) return 0;
  SpmmArgs a{};
  a.indptr = q.indptr; a.indices = q.indices; a.Xq = q.X_q24; a.ldq = q.ldq; a.dq = dq;
  a.Y = q.Y; a.ldy = q.ldy; a.Yh = q.Y_hi; a.Yl = q.Y_lo; a.ldyp = q.ldyp;
  a.n_dst = q.n_dst; a.d = d; a.indptr64 = q.indptr64;
  a.self_add = q.self_add; a.mean_plus_one = q.mean_plus_one; a.dst_scale = q.dst_scale;
  a.bias = q.bias; a.col_scale = q.col_scale; a.col_shift = q.col_shift; a.relu = q.relu;
  const unsigned blocks = static_cast<unsigned>((q.n_dst + kWarps - 1) / kWarps);
  const size_t smem = static_cast<size_t>(kWarps) * stages * q.ldq;
  if (stages == 8) {
    GLNN_CUDA_OK(cudaFuncSetAttribute(spmm_tma_q24_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
    spmm_tma_q24_kernel<8><<<blocks, kWarps * 32, smem, st>>>(a);
  } else {
    GLNN_CUDA_OK(cudaFuncSetAttribute(spmm_tma_q24_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
    spmm_tma_q24_kernel<4><<<blocks, kWarps * 32, smem, st>>>(a);
  }
  GLNN_LAUNCH_OK("spmm_tma_q24_kernel");
  return 0;
}

template <int G, int VPL, int W>
static int launch_cfg(const SpmmArgs& a, cudaStream_t st) {
  constexpr int NG = kWarps * (32 / G);
  const int64_t blocks = (a.n_dst + NG - 1) / NG;
  if (blocks == 0) return 0;
  GLNN_REQUIRE(blocks < (1LL << 31), GLNN_ERR_SHAPE, "spmm: too many rows (%lld)", (long long)a.n_dst);
  GLNN_CUDA_OK(cudaMemsetAsync(a.hub_ctr, 0, 4 * sizeof(int), st));
  const unsigned nb = static_cast<unsigned>(blocks), hub_grid = static_cast<unsigned>(4 * sm_count());
  if (a.src_scale) {
    spmm_csr_kernel<G, VPL, W, true><<<nb, kWarps * 32, 0, st>>>(a);
    spmm_hub_kernel<G, VPL, W, true><<<hub_grid, kWarps * 32, 0, st>>>(a);
  } else {
    spmm_csr_kernel<G, VPL, W, false><<<nb, kWarps * 32, 0, st>>>(a);
    spmm_hub_kernel<G, VPL, W, false><<<hub_grid, kWarps * 32, 0, st>>>(a);
  }
  spmm_hub_finish_kernel<G, VPL, W><<<32, kWarps * 32, 0, st>>>(a);
  GLNN_LAUNCH_OK("spmm_csr_kernel");
  return 0;
}

template <int W>
static int launch_width(const SpmmArgs& a, cudaStream_t st) {
  const int lanes = (a.d + W - 1) / W;  // column chunks per row
  if (lanes <= 2) return launch_cfg<2, 1, W>(a, st);
  if (lanes <= 4) return launch_cfg<4, 1, W>(a, st);
  if (lanes <= 8) return launch_cfg<8, 1, W>(a, st);
  if (lanes <= 16) return launch_cfg<16, 1, W>(a, st);
  if (lanes <= 32) return launch_cfg<32, 1, W>(a, st);
  if (lanes <= 64) return launch_cfg<32, 2, W>(a, st);
  if constexpr (W == 8) {
    return GLNN_ERR_SHAPE;  // q24 rows wider than 512 are not produced
  } else {
    if (lanes <= 96) return launch_cfg<32, 3, W>(a, st);
    return launch_cfg<32, 4, W>(a, st);
  }
}

// Hub scratch: one per (device, stream) so that concurrent streams never share counters.
struct HubScratch {
  int* ctr = nullptr;
  HubTask* tasks = nullptr;
  int64_t* rows = nullptr;
  float* acc = nullptr;
};
static std::mutex g_hub_mu;
static std::map<std::pair<int, cudaStream_t>, HubScratch> g_hub;

static int hub_scratch(cudaStream_t st, HubScratch* out) {
  int dev = 0;
  GLNN_CUDA_OK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_hub_mu);
  auto key = std::make_pair(dev, st);
  auto it = g_hub.find(key);
  if (it == g_hub.end()) {
    HubScratch h;
    GLNN_CUDA_OK(cudaMalloc(&h.ctr, 4 * sizeof(int)));
    GLNN_CUDA_OK(cudaMalloc(&h.tasks, sizeof(HubTask) * kCapTasks));
    GLNN_CUDA_OK(cudaMalloc(&h.rows, sizeof(int64_t) * kCapRows));
    GLNN_CUDA_OK(cudaMalloc(&h.acc, sizeof(float) * kCapRows * kHubAccLd));
    it = g_hub.emplace(key, h).first;
  }
  *out = it->second;
  return 0;
}

int spmm_run(const glnn_spmm_desc& d0, cudaStream_t st) {
  const glnn_spmm_desc& q = d0;
  const int d = q.d;
  GLNN_REQUIRE(q.n_dst >= 0 && q.n_src >= 0 && d >= 0, GLNN_ERR_ARG, "spmm: negative size");
  if (q.n_dst == 0 || d == 0) return 0;
  GLNN_REQUIRE(q.indptr && (q.X || q.X_q24) && (q.Y || q.Y_hi), GLNN_ERR_ARG, "spmm: null indptr/X/Y");
  GLNN_REQUIRE(!(q.X && q.X_q24), GLNN_ERR_ARG, "spmm: give X or X_q24, not both");
  const bool q24 = q.X_q24 != nullptr;
  const int dq = (d + 7) / 8 * 8;
  GLNN_REQUIRE(!q24 || (d <= 512 && aligned16(q.X_q24) && q.ldq % 16 == 0 && q.ldq >= 3 * dq),
               GLNN_ERR_SHAPE,
               "spmm: q24 input needs d <= 512, a 16-byte aligned buffer and ldq %% 16 == 0, ldq >= "
               "3 * roundup(d, 8)");
  GLNN_REQUIRE((q.Y_hi == nullptr) == (q.Y_lo == nullptr), GLNN_ERR_ARG, "spmm: output planes come in pairs");
  const int dout = q24 ? dq : d;  // a q24 row is processed in whole 8-column chunks
  GLNN_REQUIRE((q24 || q.ldx >= d) && (!q.Y || q.log_softmax || q.ldy >= dout) &&
                   (!q.Y_hi || q.ldyp >= dout),
               GLNN_ERR_SHAPE, "spmm: leading dimension smaller than d=%d", d);
  GLNN_REQUIRE(!q.self_add || q.n_src >= q.n_dst, GLNN_ERR_SHAPE,
               "spmm: self_add needs dst nodes to be a prefix of src nodes (n_src=%lld < n_dst=%lld)",
               (long long)q.n_src, (long long)q.n_dst);
  GLNN_REQUIRE((q.col_scale == nullptr) == (q.col_shift == nullptr), GLNN_ERR_ARG,
               "spmm: col_scale and col_shift must be given together");
  GLNN_REQUIRE(q.relu >= 0 && q.relu <= 2, GLNN_ERR_ARG, "spmm: relu must be 0, 1 or 2");
  GLNN_REQUIRE(!q.Y_hi || (q.ldyp % 4 == 0 && (reinterpret_cast<uintptr_t>(q.Y_hi) & 7) == 0 &&
                           (reinterpret_cast<uintptr_t>(q.Y_lo) & 7) == 0),
               GLNN_ERR_ALIGN, "spmm: output planes need ldyp %% 4 == 0 and 8-byte alignment");
  GLNN_REQUIRE(q.log_softmax >= 0 && q.log_softmax <= d, GLNN_ERR_ARG,
               "spmm: log_softmax must be 0 or the number of leading columns (<= d)");
  GLNN_REQUIRE(!q.log_softmax || (q.Y && !q.Y_hi && d <= 512 && q.ldy >= q.log_softmax), GLNN_ERR_ARG,
               "spmm: the log_softmax epilogue writes fp32 Y only, d <= 512, ldy >= log_softmax");
  GLNN_REQUIRE(q.hot_below >= 0, GLNN_ERR_ARG, "spmm: hot_below must be >= 0");
  GLNN_REQUIRE(!q.Y_init || q.ldyi >= d, GLNN_ERR_SHAPE, "spmm: ldyi < d");
  HubScratch hs;
  int rc = hub_scratch(st, &hs);
  if (rc != 0) return rc;
  const bool vec = q24 || ((d % 4 == 0) && (q.ldx % 4 == 0) && aligned16(q.X));
  GLNN_REQUIRE(!q24 || ((!q.Y || q.log_softmax || (q.ldy % 4 == 0 && aligned16(q.Y))) &&
                        (!q.Y_hi || (q.ldyp % 8 == 0 && aligned16(q.Y_hi) && aligned16(q.Y_lo)))),
               GLNN_ERR_ALIGN, "spmm: q24 input needs 16-byte aligned output rows");
  const bool vec_out = (!q.Y || q.log_softmax || (q.ldy % 4 == 0 && aligned16(q.Y)));
  const int chunk = (vec && vec_out) ? 512 : 128;
  GLNN_REQUIRE(!q.log_softmax || d <= chunk, GLNN_ERR_SHAPE,
               "spmm: the log_softmax epilogue needs the row in one launch (d <= %d here)", chunk);
  for (int c0 = 0; c0 < d; c0 += chunk) {
    SpmmArgs a;
    a.indptr = q.indptr;
    a.indices = q.indices;
    a.X = q.X ? q.X + c0 : nullptr;
    a.ldx = q.ldx;
    a.Xq = q.X_q24;
    a.ldq = q.ldq;
    a.dq = dq;
    a.Y = q.Y ? q.Y + c0 : nullptr;
    a.ldy = q.ldy;
    a.Yh = q.Y_hi ? q.Y_hi + c0 : nullptr;
    a.Yl = q.Y_lo ? q.Y_lo + c0 : nullptr;
    a.ldyp = q.ldyp;
    a.n_dst = q.n_dst;
    a.d = min(chunk, d - c0);
    a.indptr64 = q.indptr64;
    a.self_add = q.self_add;
    a.mean_plus_one = q.mean_plus_one;
    a.src_scale = q.src_scale;
    a.dst_scale = q.dst_scale;
    a.bias = q.bias ? q.bias + c0 : nullptr;
    a.col_scale = q.col_scale ? q.col_scale + c0 : nullptr;
    a.col_shift = q.col_shift ? q.col_shift + c0 : nullptr;
    a.relu = q.relu;
    a.log_softmax = q.log_softmax > 0;
    a.d_valid = q.log_softmax;
    a.hot_below = q.hot_below;
    a.Yinit = q.Y_init ? q.Y_init + c0 : nullptr;
    a.ldyi = q.ldyi;
    a.Xs = nullptr;
    a.lds = 0;
    a.cap_dev = nullptr;
    a.cap_limit = 0;
    a.hub_ctr = hs.ctr;
    a.hub_tasks = hs.tasks;
    a.hub_rows = hs.rows;
    a.hub_acc = hs.acc;
    a.cap_tasks = kCapTasks;
    a.cap_rows = kCapRows;
    rc = q24 ? launch_width<8>(a, st) : ((vec && vec_out) ? launch_width<4>(a, st) : launch_width<1>(a, st));
    if (rc != 0) return rc;
  }
  return 0;
}

// EXPERIMENTAL (see compact_s24_kernel): q24 -> s24 and the aggregation that reads it.
int compact_s24(const uint8_t* Q, int64_t ldq, int64_t rows, int d, uint32_t* S, int64_t lds, int* cap_dev,
                cudaStream_t st) {
  const int dq = (d + 7) / 8 * 8;
  GLNN_REQUIRE(Q && S && cap_dev, GLNN_ERR_ARG, "compact_s24: null pointer");
  GLNN_REQUIRE(d > 0 && dq <= kS24MaxD, GLNN_ERR_SHAPE, "compact_s24: d must be in [1, %d]", kS24MaxD);
  GLNN_REQUIRE(aligned16(Q) && ldq % 16 == 0 && ldq >= 3 * dq, GLNN_ERR_ALIGN, "compact_s24: bad q24 layout");
  GLNN_REQUIRE(aligned16(S) && lds % 32 == 0 && lds >= (dq + 31) / 32 * 32, GLNN_ERR_ALIGN,
               "compact_s24: lds must be a multiple of 32 words and >= roundup(d, 32)");
  GLNN_CUDA_OK(cudaMemsetAsync(cap_dev, 0, sizeof(int), st));
  if (rows == 0) return 0;
  const unsigned blocks = static_cast<unsigned>(std::min<int64_t>((rows + kWarps - 1) / kWarps, 16LL * sm_count()));
  compact_s24_kernel<<<blocks, kWarps * 32, 0, st>>>(Q, ldq, rows, dq, S, lds, cap_dev);
  GLNN_LAUNCH_OK("compact_s24_kernel");
  return 0;
}

int spmm_run_s24(const glnn_spmm_desc& q, const uint32_t* S, int64_t lds, const int* cap_dev,
                 cudaStream_t st) {
  const int d = q.d, dq = (d + 7) / 8 * 8;
  GLNN_REQUIRE(S && cap_dev && q.X_q24 && !q.X, GLNN_ERR_ARG,
               "spmm_s24: needs the q24 matrix (self term, hubs, dense fallback) and its s24 copy");
  GLNN_REQUIRE(dq > 128 && dq <= kS24MaxD, GLNN_ERR_SHAPE, "spmm_s24: 128 < roundup(d, 8) <= 256 only");
  GLNN_REQUIRE(!q.src_scale && !q.log_softmax, GLNN_ERR_ARG, "spmm_s24: no src_scale / log_softmax");
  GLNN_REQUIRE(aligned16(S) && lds % 32 == 0 && lds >= (dq + 31) / 32 * 32, GLNN_ERR_ALIGN,
               "spmm_s24: lds must be a multiple of 32 words and >= roundup(d, 32)");
  GLNN_REQUIRE(q.n_dst >= 0 && q.indptr && (q.Y || q.Y_hi), GLNN_ERR_ARG, "spmm_s24: null indptr / output");
  if (q.n_dst == 0) return 0;
  GLNN_REQUIRE(aligned16(q.X_q24) && q.ldq % 16 == 0 && q.ldq >= 3 * dq, GLNN_ERR_ALIGN, "spmm_s24: bad q24 layout");
  GLNN_REQUIRE((q.Y_hi == nullptr) == (q.Y_lo == nullptr), GLNN_ERR_ARG, "spmm_s24: output planes come in pairs");
  GLNN_REQUIRE((!q.Y || (q.ldy % 4 == 0 && q.ldy >= dq && aligned16(q.Y))) &&
                   (!q.Y_hi || (q.ldyp % 8 == 0 && q.ldyp >= dq && aligned16(q.Y_hi) && aligned16(q.Y_lo))),
               GLNN_ERR_ALIGN, "spmm_s24: output rows must be 16-byte aligned and >= roundup(d, 8) wide");
  GLNN_REQUIRE(!q.self_add || q.n_src >= q.n_dst, GLNN_ERR_SHAPE, "spmm_s24: self_add needs n_src >= n_dst");
  GLNN_REQUIRE((q.col_scale == nullptr) == (q.col_shift == nullptr) && q.relu >= 0 && q.relu <= 2,
               GLNN_ERR_ARG, "spmm_s24: bad epilogue arguments");
  HubScratch hs;
  int rc = hub_scratch(st, &hs);
  if (rc != 0) return rc;
  SpmmArgs a;
  a.indptr = q.indptr; a.indices = q.indices; a.X = nullptr; a.ldx = 0;
  a.Xq = q.X_q24; a.ldq = q.ldq; a.dq = dq;
  a.Y = q.Y; a.ldy = q.ldy; a.Yh = q.Y_hi; a.Yl = q.Y_lo; a.ldyp = q.ldyp;
  a.n_dst = q.n_dst; a.d = d; a.indptr64 = q.indptr64;
  a.self_add = q.self_add; a.mean_plus_one = q.mean_plus_one;
  a.src_scale = nullptr; a.dst_scale = q.dst_scale; a.bias = q.bias;
  a.col_scale = q.col_scale; a.col_shift = q.col_shift; a.relu = q.relu;
  a.log_softmax = 0; a.d_valid = 0; a.hot_below = q.hot_below;
  a.Yinit = q.Y_init; a.ldyi = q.ldyi;
  a.Xs = S; a.lds = lds; a.cap_dev = cap_dev;
  a.cap_limit = (q.ldq * 4 / 5) / 4;  // sparse rows must be at least 20 % shorter than the q24 rows
  a.hub_ctr = hs.ctr; a.hub_tasks = hs.tasks; a.hub_rows = hs.rows; a.hub_acc = hs.acc;
  a.cap_tasks = kCapTasks; a.cap_rows = kCapRows;
  const int64_t blocks = (a.n_dst + kWarps - 1) / kWarps;
  GLNN_REQUIRE(blocks < (1LL << 31), GLNN_ERR_SHAPE, "spmm_s24: too many rows");
  GLNN_CUDA_OK(cudaMemsetAsync(a.hub_ctr, 0, 4 * sizeof(int), st));
  spmm_csr_s24_kernel<<<static_cast<unsigned>(blocks), kWarps * 32, 0, st>>>(a);
  spmm_hub_kernel<32, 1, 8, false><<<static_cast<unsigned>(4 * sm_count()), kWarps * 32, 0, st>>>(a);
  spmm_hub_finish_kernel<32, 1, 8><<<32, kWarps * 32, 0, st>>>(a);
  GLNN_LAUNCH_OK("spmm_csr_s24_kernel");
  return 0;
}

// fp32 -> q24 (top 24 bits of each value, rounded to nearest; pad columns zero).
__global__ void __launch_bounds__(256) quantize_q24_kernel(const float* __restrict__ X, int64_t ldx,
                                                           int64_t rows, int d, int dq,
                                                           uint8_t* __restrict__ Q, int64_t ldq) {
  const int cpr = dq / 4;  // 4-column pieces per row
  const int64_t total = rows * cpr;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / cpr;
    const int c = static_cast<int>(i % cpr) * 4;
    uint32_t b[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      b[j] = (c + j < d) ? __float_as_uint(__ldg(X + r * ldx + c + j)) + 0x80u : 0u;
    uint8_t* row = Q + r * ldq;
    uint2 hi;
    hi.x = (b[0] >> 16) | (b[1] & 0xFFFF0000u);
    hi.y = (b[2] >> 16) | (b[3] & 0xFFFF0000u);
    *reinterpret_cast<uint2*>(row + 2 * c) = hi;
    *reinterpret_cast<uint32_t*>(row + 2 * dq + c) = ((b[0] >> 8) & 0xFFu) | (((b[1] >> 8) & 0xFFu) << 8) |
                                                     (((b[2] >> 8) & 0xFFu) << 16) |
                                                     (((b[3] >> 8) & 0xFFu) << 24);
  }
}

int quantize_q24(const float* X, int64_t ldx, int64_t rows, int d, uint8_t* Q, int64_t ldq,
                 cudaStream_t st) {
  if (rows == 0 || d == 0) return 0;
  const int dq = (d + 7) / 8 * 8;
  const int64_t total = rows * (dq / 4);
  const unsigned blocks = static_cast<unsigned>(std::min<int64_t>((total + 255) / 256, 16LL * sm_count()));
  quantize_q24_kernel<<<blocks, 256, 0, st>>>(X, ldx, rows, d, dq, Q, ldq);
  GLNN_LAUNCH_OK("quantize_q24_kernel");
  return 0;
}

}  // namespace glnn

extern "C" int glnn_spmm_csr(const glnn_spmm_desc* desc, glnn_stream_t stream) {
  GLNN_REQUIRE(desc != nullptr, GLNN_ERR_ARG, "spmm: null descriptor");
  return glnn::spmm_run(*desc, static_cast<cudaStream_t>(stream));
}

extern "C" int glnn_exp_spmm_tma_q24(const glnn_spmm_desc* desc, int stages, glnn_stream_t stream) {
  GLNN_REQUIRE(desc != nullptr, GLNN_ERR_ARG, "spmm_tma_exp: null descriptor");
  return glnn::spmm_run_tma_exp(*desc, stages, static_cast<cudaStream_t>(stream));
}

extern "C" int64_t glnn_s24_row_words(int d) {
  if (d <= 0) return 0;
  return (static_cast<int64_t>(d) + 31) / 32 * 32;
}

extern "C" int glnn_compact_s24(const uint8_t* X_q24, int64_t ldq, int64_t rows, int d, uint32_t* X_s24,
                                int64_t lds, int32_t* cap_dev, glnn_stream_t stream) {
  return glnn::compact_s24(X_q24, ldq, rows, d, X_s24, lds, cap_dev, static_cast<cudaStream_t>(stream));
}

extern "C" int glnn_spmm_csr_s24(const glnn_spmm_desc* desc, const uint32_t* X_s24, int64_t lds,
                                 const int32_t* cap_dev, glnn_stream_t stream) {
  GLNN_REQUIRE(desc != nullptr, GLNN_ERR_ARG, "spmm_s24: null descriptor");
  return glnn::spmm_run_s24(*desc, X_s24, lds, cap_dev, static_cast<cudaStream_t>(stream));
}

extern "C" int64_t glnn_q24_row_bytes(int d) {
  if (d <= 0) return 0;
  const int64_t dq = (d + 7) / 8 * 8;
  return (3 * dq + 31) / 32 * 32;
}

extern "C" int glnn_quantize_q24_f32(const float* X, int64_t ldx, int64_t rows, int d, uint8_t* X_q24,
                                     int64_t ldq, glnn_stream_t stream) {
  using namespace glnn;
  GLNN_REQUIRE(rows >= 0 && d >= 0, GLNN_ERR_ARG, "quantize_q24: negative size");
  if (rows == 0 || d == 0) return 0;
  GLNN_REQUIRE(X && X_q24, GLNN_ERR_ARG, "quantize_q24: null pointer");
  const int dq = (d + 7) / 8 * 8;
  GLNN_REQUIRE(ldx >= d && ldq >= 3 * dq && ldq % 16 == 0 && aligned16(X_q24), GLNN_ERR_SHAPE,
               "quantize_q24: need ldx >= d, ldq >= 3 * roundup(d, 8), ldq %% 16 == 0, 16-byte aligned output");
  return quantize_q24(X, ldx, rows, d, X_q24, ldq, static_cast<cudaStream_t>(stream));
}

static glnn_spmm_desc spmm_desc_basic(const void* indptr, int indptr64, const int32_t* indices,
                                      int64_t n_dst, int64_t n_src, int d, int self_add,
                                      int mean_plus_one, const float* src_scale,
                                      const float* dst_scale, const float* bias,
                                      const float* col_scale, const float* col_shift, int relu) {
  glnn_spmm_desc q{};
  q.indptr = indptr; q.indptr64 = indptr64; q.indices = indices;
  q.n_dst = n_dst; q.n_src = n_src; q.d = d;
  q.self_add = self_add; q.mean_plus_one = mean_plus_one;
  q.src_scale = src_scale; q.dst_scale = dst_scale; q.bias = bias;
  q.col_scale = col_scale; q.col_shift = col_shift; q.relu = relu;
  return q;
}

extern "C" int glnn_spmm_csr_f32(const void* indptr, int indptr64, const int32_t* indices,
                                 const float* X, int64_t ldx, float* Y, int64_t ldy, int64_t n_dst,
                                 int64_t n_src, int d, int self_add, int mean_plus_one,
                                 const float* src_scale, const float* dst_scale, const float* bias,
                                 const float* col_scale, const float* col_shift, int relu,
                                 glnn_stream_t stream) {
  GLNN_REQUIRE(Y != nullptr || n_dst <= 0 || d <= 0, GLNN_ERR_ARG, "spmm: null indptr/X/Y");
  glnn_spmm_desc q = spmm_desc_basic(indptr, indptr64, indices, n_dst, n_src, d, self_add, mean_plus_one,
                                     src_scale, dst_scale, bias, col_scale, col_shift, relu);
  q.X = X; q.ldx = ldx; q.Y = Y; q.ldy = ldy;
  return glnn::spmm_run(q, static_cast<cudaStream_t>(stream));
}

extern "C" int glnn_spmm_csr_planes(const void* indptr, int indptr64, const int32_t* indices,
                                    const float* X, int64_t ldx, uint16_t* Y_hi, uint16_t* Y_lo,
                                    int64_t ldyp, int64_t n_dst, int64_t n_src, int d, int self_add,
                                    int mean_plus_one, const float* src_scale, const float* dst_scale,
                                    const float* bias, const float* col_scale, const float* col_shift,
                                    int relu, glnn_stream_t stream) {
  GLNN_REQUIRE((Y_hi && Y_lo) || n_dst <= 0 || d <= 0, GLNN_ERR_ARG, "spmm_planes: null output plane");
  glnn_spmm_desc q = spmm_desc_basic(indptr, indptr64, indices, n_dst, n_src, d, self_add, mean_plus_one,
                                     src_scale, dst_scale, bias, col_scale, col_shift, relu);
  q.X = X; q.ldx = ldx; q.Y_hi = Y_hi; q.Y_lo = Y_lo; q.ldyp = ldyp;
  return glnn::spmm_run(q, static_cast<cudaStream_t>(stream));
}

extern "C" int glnn_spmm_csr_q24_planes(const void* indptr, int indptr64, const int32_t* indices,
                                        const uint8_t* X_q24, uint16_t* Y_hi, uint16_t* Y_lo,
                                        int64_t ldyp, int64_t n_dst, int64_t n_src, int d, int self_add,
                                        int mean_plus_one, const float* src_scale,
                                        const float* dst_scale, glnn_stream_t stream) {
  GLNN_REQUIRE((X_q24 && Y_hi && Y_lo) || n_dst <= 0 || d <= 0, GLNN_ERR_ARG, "spmm_q24: null pointer");
  glnn_spmm_desc q = spmm_desc_basic(indptr, indptr64, indices, n_dst, n_src, d, self_add, mean_plus_one,
                                     src_scale, dst_scale, nullptr, nullptr, nullptr, 0);
  q.X_q24 = X_q24; q.ldq = 3 * static_cast<int64_t>((d + 7) / 8 * 8);
  q.Y_hi = Y_hi; q.Y_lo = Y_lo; q.ldyp = ldyp;
  return glnn::spmm_run(q, static_cast<cudaStream_t>(stream));
}
