// Student path (B): one pass of train_mini_batch (train_and_eval.py:59-86) and the eval forward of
// evaluate_mini_batch (:108-136) for models.MLP, with no autograd anywhere.
//
// A step is a fixed kernel sequence
//   gather -> [GEMM(+bias) -> BN batch stats -> BN apply + ReLU + dropout] x (L-1) -> GEMM(+bias)
//   -> log-softmax + NLL|KL loss + d(lamb*loss)/dlogits (+ last-layer bias grad)
//   -> [dW GEMM, dX GEMM, BN backward (2 kernels, also yields dgamma/dbeta/dbias)] x (L-1) -> dW_0
//   -> flat Adam -> advance
// over fixed workspace addresses; everything that changes from step to step (batch offset, Adam
// step number, dropout stream) is derived on the device from a step counter, and everything that
// changes from pass to pass (input matrix, targets, lamb, learning rate) lives in a small device
// struct.  The sequence is therefore captured ONCE into a CUDA graph and replayed nb times per
// pass with a single host read of the loss at the end of the pass (the reference syncs per step).
//
// Data parallel (glnn_mlp_train_pass_dp): the global batch of the reference is
// split by rows over the GPUs of one box; every exchange of the step is FUSED into the kernel that
// produces or consumes the data, over peer-mapped (NVLink) symmetric memory -- no NCCL call, no extra
// copy, the whole step still one CUDA graph:
//   * BatchNorm statistics: bn_stats / bn_bwd_stats write their per-split partials straight into
//     every peer's buffer, the last CTA raises a flag on every peer; bn_apply / bn_bwd_apply wait for
//     the flags and combine all partials in rank order (Chan), so every rank normalises with the
//     statistics of the GLOBAL batch, exactly as the single-process reference does;
//   * gradients + Adam: each rank owns 1/G of the flat parameter buffer; adam_dp_kernel sums that
//     slice of the gradients over all peers (P2P loads, fixed order), applies Adam with its slice of
//     the moments and stores the new parameters into every peer's buffer (P2P stores) --
//     reduce-scatter, optimizer and all-gather in one kernel.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include <cuda_bf16.h>

#include "common.cuh"
#include "gemm.cuh"

namespace glnn {

int split_planes(const float* X, int64_t ldx, int64_t rows, int cols, uint16_t* hi, uint16_t* lo,
                 int64_t ldp, cudaStream_t st);  // planes.cu

// torch.optim.Adam hyper-parameters as the update kernels take them.  The host keeps them in DOUBLE,
// as torch does: 1 - beta is formed in double and only then rounded to fp32 (fp32(1 - 0.999) differs
// from 1.f - fp32(0.999) by 1.3e-5 relative -- found by the injected-gradient test), and the bias
// corrections 1 - beta^t are evaluated in double from the double betas.
struct AdamHp {
  double lr, beta1, beta2;
  float b1, b2, omb1, omb2, eps, wd;
};
static inline AdamHp make_adam_hp(const glnn_adam_hparams& h) {
  AdamHp a;
  a.lr = h.lr; a.beta1 = h.beta1; a.beta2 = h.beta2;
  a.b1 = static_cast<float>(h.beta1); a.b2 = static_cast<float>(h.beta2);
  a.omb1 = static_cast<float>(1.0 - h.beta1); a.omb2 = static_cast<float>(1.0 - h.beta2);
  a.eps = static_cast<float>(h.eps); a.wd = static_cast<float>(h.weight_decay);
  return a;
}

struct PassParams {       // device-resident, rewritten once per pass
  const float* X;
  int64_t ldx;
  const void* target;
  const int64_t* perm;
  const uint8_t* masks;
  float* loss_sum;
  int64_t step0;
  uint64_t seed;
  int32_t kind;
  float lamb;
  AdamHp hp;
};

struct Dims {
  int L, F, H, C, norm;
  float p_drop, bn_eps, bn_mom;
  int64_t R;   // batch rows processed by this GPU
  int64_t Rg;  // rows of the (global) batch: R * world
  int world, rank;
};

// Data-parallel exchange state, passed by value to the kernels that communicate.  Every rank owns a
// symmetric region laid out identically; base[r] is rank r's region as mapped into this process.
constexpr int kMaxPeers = 8;
constexpr int kSlots = 40;  // 2 per hidden layer (<= 15) + gradients + parameters
struct DpDev {
  int world, rank;
  unsigned char* base[kMaxPeers];
  int64_t off_flags;   // uint32 [kSlots][kMaxPeers]: flag[slot][src] = last epoch signalled by src
  int64_t off_done;    // uint32 [kSlots]: CTA completion counters (local use)
  int64_t off_epoch;   // uint32: step epoch of this rank, starts at 1, never reset
  int64_t off_part;    // float [2 (L-1)][world * rs][2][H]
  int64_t off_params, off_grads;  // flat parameter / gradient buffers inside the region
};
__host__ __device__ __forceinline__ uint32_t* dp_flags(const DpDev& dp, int r) {
  return reinterpret_cast<uint32_t*>(dp.base[r] + dp.off_flags);
}

// ---- cross-GPU signalling over peer-mapped memory ----
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Raise flag[slot][my rank] = epoch on every rank (called by ONE thread after a system fence).
__device__ __forceinline__ void dp_signal_all(const DpDev& dp, int slot, uint32_t epoch) {
  __threadfence_system();
  for (int r = 0; r < dp.world; ++r) st_release_sys(dp_flags(dp, r) + slot * kMaxPeers + dp.rank, epoch);
}
// Block-wide wait until every rank has signalled `epoch` on `slot`.  A protocol bug must trap
// instead of hanging the GPU: bounded by ~4 s of clock64.
__device__ __forceinline__ void dp_wait_all(const DpDev& dp, int slot, uint32_t epoch) {
  if (threadIdx.x < dp.world && threadIdx.y == 0 && threadIdx.z == 0) {
    const uint32_t* f = dp_flags(dp, dp.rank) + slot * kMaxPeers + threadIdx.x;
    const long long t0 = clock64();
    while (static_cast<int32_t>(ld_acquire_sys(f) - epoch) < 0) {
      if (clock64() - t0 > 8000000000LL) __trap();
    }
  }
  __syncthreads();
}
// Called by every thread of a CTA after its peer stores: the LAST CTA of the grid signals `slot`.
__device__ __forceinline__ void dp_last_cta_signals(const DpDev& dp, int slot, uint32_t epoch,
                                                    unsigned total_ctas) {
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) {
    uint32_t* done = reinterpret_cast<uint32_t*>(dp.base[dp.rank] + dp.off_done) + slot;
    const unsigned prev = atomicInc(done, total_ctas - 1);  // wraps to 0 for the next step
    if (prev == total_ctas - 1) dp_signal_all(dp, slot, epoch);
  }
}
__device__ __forceinline__ uint32_t dp_epoch(const DpDev& dp) {
  return *reinterpret_cast<const uint32_t*>(dp.base[dp.rank] + dp.off_epoch);
}

__host__ __device__ __forceinline__ int64_t imax64(int64_t a, int64_t b) { return a > b ? a : b; }
__host__ __device__ __forceinline__ int64_t imin64(int64_t a, int64_t b) { return a < b ? a : b; }
static inline int64_t up(int64_t x, int64_t a = 64) { return (x + a - 1) / a * a; }
static inline int in_dim(const Dims& d, int l) { return l == 0 ? d.F : d.H; }
static inline int out_dim(const Dims& d, int l) { return l == d.L - 1 ? d.C : d.H; }

// flat parameter layout: [W_0 | b_0 | W_1 | b_1 | ... | gamma_0 | beta_0 | ...]
struct ParamLayout {
  int64_t w[16], b[16], gamma[16], beta[16], total;
};
static ParamLayout param_layout(const Dims& d) {
  ParamLayout p{};
  int64_t o = 0;
  for (int l = 0; l < d.L; ++l) {
    p.w[l] = o; o += static_cast<int64_t>(in_dim(d, l)) * out_dim(d, l);
    p.b[l] = o; o += out_dim(d, l);
  }
  if (d.norm == 1)
    for (int l = 0; l < d.L - 1; ++l) {
      p.gamma[l] = o; o += d.H;
      p.beta[l] = o; o += d.H;
    }
  p.total = o;
  return p;
}

// workspace carve-up (offsets in floats).  GEMM operands live as bf16 hi/lo plane pairs ("P" fields:
// hi plane at the offset, lo plane right behind it, leading dimension pad8(width)); everything a
// non-GEMM kernel reads stays fp32.
struct WsLayout {
  int64_t pp, ctr, xbP, tgt, z[16], aP[16], mean[16], invstd[16], logits, dlogP, dh[2], dzP, wP[16],
      part, perm, fold, total_bytes;
  int rs;  // row splits of the column reductions
};
static inline int pad8(int d) { return (d + 7) / 8 * 8; }
static inline int64_t plane_pair_floats(int64_t rows, int width) { return rows * pad8(width); }
// columns per thread of the column-wise BatchNorm kernels
static inline int bn_vec(const Dims& d) { return d.H % 4 == 0 ? 4 : 1; }
static int row_splits(const Dims& d) {
  const int col_tiles = std::max(1, (d.H + 32 * bn_vec(d) - 1) / (32 * bn_vec(d)));
  int rs = std::max(1, 2 * 148 / col_tiles);  // one wave of 2 resident CTAs per SM (B200: 148 SMs)
  rs = static_cast<int>(std::min<int64_t>(rs, std::max<int64_t>(1, d.R / 32)));
  return std::max(1, std::min(rs, 64));
}
constexpr int64_t kPermCap = 1LL << 22;  // rows of one train_pass call (the host splits longer passes)

static WsLayout ws_layout(const Dims& d, bool train) {
  WsLayout w{};
  int64_t o = 0;
  auto take = [&](int64_t floats) { int64_t r = o; o += up(floats); return r; };
  w.pp = take((sizeof(PassParams) + 3) / 4);
  w.ctr = take(4);
  w.xbP = take(plane_pair_floats(d.R, d.F));
  w.tgt = take(std::max<int64_t>(d.R * d.C, 2 * d.R));
  for (int l = 0; l < d.L - 1; ++l) {
    w.z[l] = take(d.R * d.H);
    w.aP[l] = take(plane_pair_floats(d.R, d.H));
    w.mean[l] = take(d.H);
    w.invstd[l] = take(d.H);
  }
  w.logits = take(d.R * d.C);
  w.dlogP = take(plane_pair_floats(d.R, d.C));
  w.dh[0] = take(d.R * d.H);
  w.dh[1] = take(d.R * d.H);
  w.dzP = take(plane_pair_floats(d.R, d.H));
  for (int l = 0; l < d.L; ++l) w.wP[l] = take(plane_pair_floats(out_dim(d, l), in_dim(d, l)));
  w.rs = row_splits(d);
  w.part = take(static_cast<int64_t>(w.rs) * 2 * d.H);
  w.fold = take(2LL * d.H * std::max(1, d.L - 1));
  w.perm = take(train ? 2 * kPermCap : 0);  // int64 copy of the pass permutation
  w.total_bytes = o * 4;
  return w;
}

struct PlaneRef {
  uint16_t *hi, *lo;
  int64_t ld;
};
static PlaneRef plane_ref(float* ws, int64_t off, int64_t rows, int width) {
  PlaneRef p;
  p.ld = pad8(width);
  p.hi = reinterpret_cast<uint16_t*>(ws + off);
  p.lo = p.hi + rows * p.ld;
  return p;
}

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_planes(uint16_t* __restrict__ hi, uint16_t* __restrict__ lo,
                                             int64_t idx, float v) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
  hi[idx] = *reinterpret_cast<const uint16_t*>(&h);
  lo[idx] = *reinterpret_cast<const uint16_t*>(&l);
}

__global__ void __launch_bounds__(256) gather_kernel(const PassParams* __restrict__ pp,
                                                     const int* __restrict__ ctr, int64_t R,
                                                     int64_t Rg, int64_t row0, int F,
                                                     int C, uint16_t* __restrict__ xb_hi,
                                                     uint16_t* __restrict__ xb_lo, int64_t ldxb,
                                                     float* __restrict__ tgt) {
  const int lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= R) return;
  const int64_t src = pp->perm[static_cast<int64_t>(*ctr) * Rg + row0 + r];
  const float* x = pp->X + src * pp->ldx;
  for (int j = lane; j < F; j += 32) store_planes(xb_hi, xb_lo, r * ldxb + j, __ldg(x + j));
  if (pp->kind == 0) {
    if (lane == 0) reinterpret_cast<int64_t*>(tgt)[r] = static_cast<const int64_t*>(pp->target)[src];
  } else {
    const float* t = static_cast<const float*>(pp->target) + src * C;
    for (int j = lane; j < C; j += 32) tgt[r * C + j] = __ldg(t + j);
  }
}

// ---- column-wise BatchNorm kernels -------------------------------------------------------------
// block (32, 8): a thread owns V adjacent columns (V = 4 when H % 4 == 0: 16-byte loads of the fp32
// operands, 8-byte stores per bf16 plane; V = 1 otherwise), threadIdx.y strides over the rows of the
// CTA's row split, four row loads in flight per thread.  All of Z / dA (33.5 MB at bs 4096, H 2048)
// stays L2-resident between the kernels of a step.
template <int V>
struct ColVec {
  float v[V];
};
template <int V>
__device__ __forceinline__ ColVec<V> ld_cols(const float* __restrict__ p) {
  ColVec<V> r;
  if constexpr (V == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else {
#pragma unroll
    for (int i = 0; i < V; ++i) r.v[i] = p[i];
  }
  return r;
}
template <int V>
__device__ __forceinline__ void store_planes_cols(uint16_t* __restrict__ hi, uint16_t* __restrict__ lo,
                                                  int64_t idx, const ColVec<V>& y) {
  if constexpr (V == 4) {
    uint32_t h[2], l[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const __nv_bfloat162 hh = __floats2bfloat162_rn(y.v[2 * i], y.v[2 * i + 1]);
      h[i] = *reinterpret_cast<const uint32_t*>(&hh);
      const float f0 = __uint_as_float(h[i] << 16), f1 = __uint_as_float(h[i] & 0xffff0000u);
      const __nv_bfloat162 ll = __floats2bfloat162_rn(y.v[2 * i] - f0, y.v[2 * i + 1] - f1);
      l[i] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    *reinterpret_cast<uint2*>(hi + idx) = make_uint2(h[0], h[1]);
    *reinterpret_cast<uint2*>(lo + idx) = make_uint2(l[0], l[1]);
  } else {
#pragma unroll
    for (int i = 0; i < V; ++i) store_planes(hi, lo, idx + i, y.v[i]);
  }
}
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}
// Dropout of one (step, layer): everything that does not change along a thread's row loop is read
// ONCE (the PassParams fields live in global memory; re-reading them per row put a dependent load
// and a branch between the row loads and serialised them).
struct DropCtx {
  const uint8_t* masks;  // host-injected keep masks of this (step, layer), row 0; null -> device stream
  uint64_t key;          // device stream: (seed, step, layer)
  float p, scale;        // p <= 0: no dropout
  uint32_t thr;          // device stream: drop when the element's 16 random bits are < thr
};
__device__ __forceinline__ DropCtx drop_ctx(const PassParams* __restrict__ pp, int step, int layer,
                                            int nlay, int64_t Rg, int H, float p_drop) {
  DropCtx d;
  d.p = p_drop;
  d.scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  d.masks = nullptr;
  d.key = 0;
  d.thr = static_cast<uint32_t>(ceilf(fminf(fmaxf(p_drop, 0.f), 1.f) * 65536.f));
  if (p_drop > 0.f) {
    const uint8_t* m = pp->masks;
    if (m) d.masks = m + (static_cast<int64_t>(step) * nlay + layer) * Rg * H;
    d.key = pp->seed + 0x9E3779B97F4A7C15ull * (static_cast<uint64_t>(pp->step0 + step) + 1) +
            0xD1B54A32D192ED03ull * (static_cast<uint64_t>(layer) + 1);
  }
  return d;
}
// keep/(1-p) factors of V adjacent elements of global row r (the decision of an element does not
// depend on V).  mword: the V mask bytes of the elements, preloaded by the caller in mask mode.
template <int V>
__device__ __forceinline__ ColVec<V> keep_cols(const DropCtx& d, int H, int64_t r, int c, uint32_t mword) {
  ColVec<V> k;
  if (d.p <= 0.f) {
#pragma unroll
    for (int i = 0; i < V; ++i) k.v[i] = 1.f;
    return k;
  }
  if (d.masks) {
#pragma unroll
    for (int i = 0; i < V; ++i) k.v[i] = ((mword >> (8 * i)) & 0xffu) ? d.scale : 0.f;
    return k;
  }
  // counter-based stream: one splitmix64 per aligned GROUP OF FOUR adjacent elements, 16 bits each,
  // compared as integers against thr = ceil(p * 2^16) (keep probability exact to 2^-16), keyed by
  // (seed, step, layer, group index)
  const uint64_t e0 = static_cast<uint64_t>(r) * H + c;
  if constexpr (V == 4) {  // c and H multiples of 4: one whole group
    const uint64_t x = splitmix64(d.key + (e0 >> 2));
    const uint32_t lo = static_cast<uint32_t>(x), hi = static_cast<uint32_t>(x >> 32);
    k.v[0] = (lo & 0xffffu) >= d.thr ? d.scale : 0.f;
    k.v[1] = (lo >> 16) >= d.thr ? d.scale : 0.f;
    k.v[2] = (hi & 0xffffu) >= d.thr ? d.scale : 0.f;
    k.v[3] = (hi >> 16) >= d.thr ? d.scale : 0.f;
  } else {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const uint64_t e = e0 + i;
      const uint64_t x = splitmix64(d.key + (e >> 2));
      const uint32_t u = static_cast<uint32_t>(x >> (16 * (e & 3))) & 0xffffu;
      k.v[i] = u >= d.thr ? d.scale : 0.f;
    }
  }
  return k;
}
template <int V>
__device__ __forceinline__ uint32_t ld_mask(const DropCtx& d, int H, int64_t r, int c) {
  if constexpr (V == 4) return *reinterpret_cast<const uint32_t*>(d.masks + r * H + c);
  uint32_t m = 0;
#pragma unroll
  for (int i = 0; i < V; ++i) m |= static_cast<uint32_t>(d.masks[r * H + c + i]) << (8 * i);
  return m;
}

// Row loop of the column-wise kernels: threadIdx.y strides over rows [r0, r1); UB rows are LOADED
// first (A, optionally B, optionally the mask bytes: UB or 2 UB independent 16-byte loads in flight
// per thread) and only then handed to f(local row, a, b, keep) -- written out explicitly because the
// compiler does not move loads across the per-row dropout code.
template <int V, int UB, bool HAS_B, bool DROP, class F>
__device__ __forceinline__ void for_rows(const float* __restrict__ A, const float* __restrict__ B, int H,
                                         int c, int64_t r0, int64_t r1, int ty, const DropCtx& dc,
                                         int64_t grow0, F&& f) {
  for (int64_t rb = r0 + ty; rb < r1; rb += 8 * UB) {
    ColVec<V> a[UB], b[UB];
    uint32_t m[UB];
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const int64_t r = rb + 8 * u;
      m[u] = 0;
      if (r < r1) {
        a[u] = ld_cols<V>(A + r * H + c);
        if constexpr (HAS_B) b[u] = ld_cols<V>(B + r * H + c);
        if constexpr (DROP) {
          if (dc.masks) m[u] = ld_mask<V>(dc, H, grow0 + r, c);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const int64_t r = rb + 8 * u;
      if (r < r1) {
        ColVec<V> k;
        if constexpr (DROP) {
          k = keep_cols<V>(dc, H, grow0 + r, c, m[u]);
        } else {
#pragma unroll
          for (int i = 0; i < V; ++i) k.v[i] = 1.f;
        }
        f(r, a[u], b[u], k);
      }
    }
  }
}

// Column statistics of Z[R,H] over a row split: chunk mean and chunk M2 in ONE pass with the chunk's
// first row as the shift K (sum (z-K), sum (z-K)^2: no cancellation, K is a sample of the column),
// combined later with Chan's formula.
template <int V>
__global__ void __launch_bounds__(256, 2) bn_stats_kernel(const float* __restrict__ Z, int64_t R, int H,
                                                       int rs, float* __restrict__ part, const DpDev dp,
                                                       int slot) {
  __shared__ float s1[8][32 * V + 1], s2[8][32 * V + 1];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int c = (blockIdx.x * 32 + tx) * V;
  const int64_t rows = (R + rs - 1) / rs;
  const int64_t r0 = blockIdx.y * rows, r1 = min(R, r0 + rows);
  const float cnt = static_cast<float>(imax64(r1 - r0, 1));
  float s[V], q[V];
  ColVec<V> K;
#pragma unroll
  for (int i = 0; i < V; ++i) { s[i] = 0.f; q[i] = 0.f; K.v[i] = 0.f; }
  if (c < H && r0 < r1) {
    K = ld_cols<V>(Z + r0 * H + c);
    for_rows<V, 8, false, false>(Z, nullptr, H, c, r0, r1, ty, DropCtx{}, 0,
                                 [&](int64_t, const ColVec<V>& z, const ColVec<V>&, const ColVec<V>&) {
#pragma unroll
                                   for (int i = 0; i < V; ++i) {
                                     const float dz = z.v[i] - K.v[i];
                                     s[i] += dz;
                                     q[i] = fmaf(dz, dz, q[i]);
                                   }
                                 });
  }
#pragma unroll
  for (int i = 0; i < V; ++i) { s1[ty][tx * V + i] = s[i]; s2[ty][tx * V + i] = q[i]; }
  __syncthreads();
  if (ty == 0 && c < H) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float S = 0.f, Q = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) { S += s1[j][tx * V + i]; Q += s2[j][tx * V + i]; }
      const float mu = K.v[i] + S / cnt;
      const float m2 = fmaxf(Q - S * (S / cnt), 0.f);
      if (dp.world > 1) {  // this split's partials go to every rank's exchange buffer
        const int64_t o =
            ((static_cast<int64_t>(slot) * dp.world + dp.rank) * rs + blockIdx.y) * 2 * H + c + i;
        for (int r = 0; r < dp.world; ++r) {
          float* pr = reinterpret_cast<float*>(dp.base[r] + dp.off_part);
          pr[o] = mu;
          pr[o + H] = m2;
        }
      } else {
        part[(static_cast<int64_t>(blockIdx.y) * 2 + 0) * H + c + i] = mu;
        part[(static_cast<int64_t>(blockIdx.y) * 2 + 1) * H + c + i] = m2;
      }
    }
  }
  if (dp.world > 1) dp_last_cta_signals(dp, slot, dp_epoch(dp), gridDim.x * gridDim.y);
}

// Combines the split statistics, normalises, applies gamma/beta, ReLU and dropout; the first row
// tile also stores mean / invstd for the backward pass and updates the running statistics
// (momentum update with the UNBIASED variance, as nn.BatchNorm1d does).  norm == 0: y = z.
template <int V>
__global__ void __launch_bounds__(256, 2) bn_apply_kernel(
    const float* __restrict__ Z, uint16_t* __restrict__ A_hi, uint16_t* __restrict__ A_lo, int64_t lda,
    int64_t R, int H, int rs,
    const float* __restrict__ part, const float* __restrict__ gamma, const float* __restrict__ beta,
    float* __restrict__ run_mean, float* __restrict__ run_var, float* __restrict__ save_mean,
    float* __restrict__ save_invstd, int norm, float eps, float mom, float p_drop,
    const PassParams* __restrict__ pp, const int* __restrict__ ctr, int layer, int nlay,
    int rows_per_block, const DpDev dp, int slot, int64_t Rg) {
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int c = (blockIdx.x * 32 + tx) * V;
  if (dp.world > 1 && norm) {  // statistics of every rank's rows must have arrived
    dp_wait_all(dp, slot, dp_epoch(dp));
    part = reinterpret_cast<const float*>(dp.base[dp.rank] + dp.off_part) +
           static_cast<int64_t>(slot) * dp.world * rs * 2 * H;
  }
  const bool active = c < H;
  float mu[V], inv[V], g[V], b[V];
#pragma unroll
  for (int i = 0; i < V; ++i) { mu[i] = 0.f; inv[i] = 1.f; g[i] = 1.f; b[i] = 0.f; }
  if (norm) {
    // Chan combination of the split statistics: threadIdx.y = j combines the splits s = j (mod 8) in
    // ascending order, then all threads combine the 8 partials in the order j = 0..7 -- a fixed tree,
    // so every CTA (and, data parallel, every rank: s is rank-major) gets identical bits.
    __shared__ float sn[8], smu[8][32 * V + 1], sq[8][32 * V + 1];
    const int64_t rows = (R + rs - 1) / rs;
    float n = 0.f, m2[V];
#pragma unroll
    for (int i = 0; i < V; ++i) m2[i] = 0.f;
    if (active) {
      for (int s = ty; s < rs * dp.world; s += 8) {
        const int sl = s % rs;
        const float nb = static_cast<float>(imax64(imin64(R, (sl + 1) * rows) - sl * rows, 0));
        if (nb <= 0.f) continue;
        const ColVec<V> mb = ld_cols<V>(part + (static_cast<int64_t>(s) * 2 + 0) * H + c);
        const ColVec<V> qb = ld_cols<V>(part + (static_cast<int64_t>(s) * 2 + 1) * H + c);
        const float tot = n + nb;
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const float delta = mb.v[i] - mu[i];
          mu[i] += delta * (nb / tot);
          m2[i] += qb.v[i] + delta * delta * (n * nb / tot);
        }
        n = tot;
      }
    }
    if (tx == 0) sn[ty] = n;
#pragma unroll
    for (int i = 0; i < V; ++i) { smu[ty][tx * V + i] = mu[i]; sq[ty][tx * V + i] = m2[i]; }
    __syncthreads();
    n = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) { mu[i] = 0.f; m2[i] = 0.f; }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float nb = sn[j];
      if (nb <= 0.f) continue;
      const float tot = n + nb;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float delta = smu[j][tx * V + i] - mu[i];
        mu[i] += delta * (nb / tot);
        m2[i] += sq[j][tx * V + i] + delta * delta * (n * nb / tot);
      }
      n = tot;
    }
    if (active) {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float var = m2[i] / static_cast<float>(Rg);
        inv[i] = 1.f / sqrtf(var + eps);
        g[i] = gamma[c + i];
        b[i] = beta[c + i];
        if (blockIdx.y == 0 && ty == 0) {
          save_mean[c + i] = mu[i];
          save_invstd[c + i] = inv[i];
          const float unb = m2[i] / static_cast<float>(imax64(Rg - 1, 1));
          run_mean[c + i] = (1.f - mom) * run_mean[c + i] + mom * mu[i];
          run_var[c + i] = (1.f - mom) * run_var[c + i] + mom * unb;
        }
      }
    }
  }
  if (!active) return;
  const int step = *ctr;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * rows_per_block;
  const int64_t r1 = min(R, r0 + rows_per_block);
  const DropCtx dc = drop_ctx(pp, step, layer, nlay, Rg, H, p_drop);
  for_rows<V, 4, false, true>(Z, nullptr, H, c, r0, r1, ty, dc, dp.rank * R,
                              [&](int64_t r, const ColVec<V>& z, const ColVec<V>&, const ColVec<V>& k) {
                                ColVec<V> y;
#pragma unroll
                                for (int i = 0; i < V; ++i) {
                                  const float t = norm ? (z.v[i] - mu[i]) * inv[i] * g[i] + b[i] : z.v[i];
                                  y.v[i] = fmaxf(t, 0.f) * k.v[i];
                                }
                                store_planes_cols<V>(A_hi, A_lo, r * lda + c, y);
                              });
}

// Backward through dropout, ReLU and BatchNorm.  Pass 1: per split  S1 = sum g, S2 = sum g*xhat with
// g = dA * keep/(1-p) * [y > 0].
template <int V>
__global__ void __launch_bounds__(256, 2) bn_bwd_stats_kernel(
    const float* __restrict__ dA, const float* __restrict__ Z, int64_t R, int H, int rs,
    const float* __restrict__ gamma, const float* __restrict__ beta,
    const float* __restrict__ save_mean, const float* __restrict__ save_invstd, int norm, float p_drop,
    const PassParams* __restrict__ pp, const int* __restrict__ ctr, int layer, int nlay,
    float* __restrict__ part, const DpDev dp, int slot, int64_t Rg) {
  __shared__ float s1[8][32 * V + 1], s2[8][32 * V + 1];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int c = (blockIdx.x * 32 + tx) * V;
  const int64_t rows = (R + rs - 1) / rs;
  const int64_t r0 = blockIdx.y * rows, r1 = min(R, r0 + rows);
  float a1[V], a2[V];
#pragma unroll
  for (int i = 0; i < V; ++i) { a1[i] = 0.f; a2[i] = 0.f; }
  if (c < H) {
    float mu[V], inv[V], g[V], b[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      mu[i] = norm ? save_mean[c + i] : 0.f;
      inv[i] = norm ? save_invstd[c + i] : 1.f;
      g[i] = norm ? gamma[c + i] : 1.f;
      b[i] = norm ? beta[c + i] : 0.f;
    }
    const int step = *ctr;
    const DropCtx dc = drop_ctx(pp, step, layer, nlay, Rg, H, p_drop);
    for_rows<V, 4, true, true>(Z, dA, H, c, r0, r1, ty, dc, dp.rank * R,
                               [&](int64_t, const ColVec<V>& z, const ColVec<V>& da, const ColVec<V>& k) {
#pragma unroll
                                 for (int i = 0; i < V; ++i) {
                                   const float xh = (z.v[i] - mu[i]) * inv[i];
                                   const float y = norm ? xh * g[i] + b[i] : z.v[i];
                                   float gr = da.v[i] * k.v[i];
                                   gr = y > 0.f ? gr : 0.f;
                                   a1[i] += gr;
                                   a2[i] = fmaf(gr, xh, a2[i]);
                                 }
                               });
  }
#pragma unroll
  for (int i = 0; i < V; ++i) { s1[ty][tx * V + i] = a1[i]; s2[ty][tx * V + i] = a2[i]; }
  __syncthreads();
  if (ty == 0 && c < H) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float t1 = 0.f, t2 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) { t1 += s1[j][tx * V + i]; t2 += s2[j][tx * V + i]; }
      if (dp.world > 1) {
        const int64_t o =
            ((static_cast<int64_t>(slot) * dp.world + dp.rank) * rs + blockIdx.y) * 2 * H + c + i;
        for (int r = 0; r < dp.world; ++r) {
          float* pr = reinterpret_cast<float*>(dp.base[r] + dp.off_part);
          pr[o] = t1;
          pr[o + H] = t2;
        }
      } else {
        part[(static_cast<int64_t>(blockIdx.y) * 2 + 0) * H + c + i] = t1;
        part[(static_cast<int64_t>(blockIdx.y) * 2 + 1) * H + c + i] = t2;
      }
    }
  }
  if (dp.world > 1) dp_last_cta_signals(dp, slot, dp_epoch(dp), gridDim.x * gridDim.y);
}

// Pass 2: dZ = gamma*invstd/R * (R*g - S1 - xhat*S2) written as bf16 planes (it only feeds the dW and
// dX projections); dgamma = S2,
// dbeta = S1; the Linear bias gradient is the column sum of dZ (mathematically 0 in front of a
// BatchNorm; the reference computes it the same way and Adam still sees its rounding noise).
template <int V>
__global__ void __launch_bounds__(256, 2) bn_bwd_apply_kernel(
    const float* __restrict__ dA, uint16_t* __restrict__ dZ_hi, uint16_t* __restrict__ dZ_lo,
    int64_t lddz, const float* __restrict__ Z, int64_t R, int H, int rs,
    const float* __restrict__ part, const float* __restrict__ gamma, const float* __restrict__ beta,
    const float* __restrict__ save_mean, const float* __restrict__ save_invstd, int norm, float p_drop,
    const PassParams* __restrict__ pp, const int* __restrict__ ctr, int layer, int nlay,
    float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias,
    int rows_per_block, const DpDev dp, int slot, int64_t Rg) {
  __shared__ float sb[8][32 * V + 1];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int c = (blockIdx.x * 32 + tx) * V;
  if (dp.world > 1 && norm) {
    dp_wait_all(dp, slot, dp_epoch(dp));
    part = reinterpret_cast<const float*>(dp.base[dp.rank] + dp.off_part) +
           static_cast<int64_t>(slot) * dp.world * rs * 2 * H;
  }
  float colsum[V];
#pragma unroll
  for (int i = 0; i < V; ++i) colsum[i] = 0.f;
  if (c < H) {
    float S1[V], S2[V], mu[V], inv[V], g[V], b[V], k[V];
#pragma unroll
    for (int i = 0; i < V; ++i) { S1[i] = 0.f; S2[i] = 0.f; mu[i] = 0.f; inv[i] = 1.f; g[i] = 1.f; b[i] = 0.f; }
    const float fR = static_cast<float>(Rg);
    if (norm) {
#pragma unroll 8
      for (int s = 0; s < rs * dp.world; ++s) {
        const ColVec<V> p1 = ld_cols<V>(part + (static_cast<int64_t>(s) * 2 + 0) * H + c);
        const ColVec<V> p2 = ld_cols<V>(part + (static_cast<int64_t>(s) * 2 + 1) * H + c);
#pragma unroll
        for (int i = 0; i < V; ++i) { S1[i] += p1.v[i]; S2[i] += p2.v[i]; }
      }
#pragma unroll
      for (int i = 0; i < V; ++i) {
        mu[i] = save_mean[c + i]; inv[i] = save_invstd[c + i]; g[i] = gamma[c + i]; b[i] = beta[c + i];
        // S1 / S2 are already sums over the GLOBAL batch: only rank 0 contributes them to the
        // gradient reduction
        if (blockIdx.y == 0 && ty == 0) {
          dgamma[c + i] = dp.rank == 0 ? S2[i] : 0.f;
          dbeta[c + i] = dp.rank == 0 ? S1[i] : 0.f;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) k[i] = g[i] * inv[i] / fR;
    const int step = *ctr;
    const int64_t r0 = static_cast<int64_t>(blockIdx.y) * rows_per_block;
    const int64_t r1 = min(R, r0 + rows_per_block);
    const DropCtx dc = drop_ctx(pp, step, layer, nlay, Rg, H, p_drop);
    for_rows<V, 4, true, true>(Z, dA, H, c, r0, r1, ty, dc, dp.rank * R,
                               [&](int64_t r, const ColVec<V>& z, const ColVec<V>& da, const ColVec<V>& ks) {
                                 ColVec<V> dz;
#pragma unroll
                                 for (int i = 0; i < V; ++i) {
                                   const float xh = (z.v[i] - mu[i]) * inv[i];
                                   const float y = norm ? xh * g[i] + b[i] : z.v[i];
                                   float gr = da.v[i] * ks.v[i];
                                   gr = y > 0.f ? gr : 0.f;
                                   dz.v[i] = norm ? k[i] * (fR * gr - S1[i] - xh * S2[i]) : gr;
                                   colsum[i] += dz.v[i];
                                 }
                                 store_planes_cols<V>(dZ_hi, dZ_lo, r * lddz + c, dz);
                               });
  }
#pragma unroll
  for (int i = 0; i < V; ++i) sb[ty][tx * V + i] = colsum[i];
  __syncthreads();
  if (ty == 0 && c < H) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float t = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) t += sb[j][tx * V + i];
      atomicAdd(dbias + c + i, t);
    }
  }
}

// log-softmax + loss + gradient of lamb*loss w.r.t. the logits; one warp per row.  Also accumulates
// the last layer's bias gradient (column sums of dlogits).
__global__ void __launch_bounds__(256) loss_kernel(const float* __restrict__ logits,
                                                   const float* __restrict__ tgt, int64_t R,
                                                   int64_t Rg, int C,
                                                   const PassParams* __restrict__ pp,
                                                   uint16_t* __restrict__ dl_hi,
                                                   uint16_t* __restrict__ dl_lo, int64_t lddl,
                                                   float* __restrict__ dbias) {
  extern __shared__ float s_col[];  // [C] column sums + [8] row losses
  float* s_loss = s_col + C;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int j = threadIdx.x; j < C; j += blockDim.x) s_col[j] = 0.f;
  __syncthreads();
  const int64_t r = static_cast<int64_t>(blockIdx.x) * 8 + warp;
  float loss = 0.f;
  if (r < R) {
    const float* x = logits + r * C;
    float m = -INFINITY;
    for (int j = lane; j < C; j += 32) m = fmaxf(m, x[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float se = 0.f;
    for (int j = lane; j < C; j += 32) se += expf(x[j] - m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
    const float lse = m + logf(se);
    const float sc = pp->lamb / static_cast<float>(Rg);
    if (pp->kind == 0) {
      const int y = static_cast<int>(reinterpret_cast<const int64_t*>(tgt)[r]);
      for (int j = lane; j < C; j += 32) {
        const float s = x[j] - lse;
        const float d = (expf(s) - (j == y ? 1.f : 0.f)) * sc;
        store_planes(dl_hi, dl_lo, r * lddl + j, d);
        atomicAdd(&s_col[j], d);
        if (j == y) loss = -s;
      }
    } else {
      const float* t = tgt + r * C;
      float st = 0.f;
      for (int j = lane; j < C; j += 32) st += expf(t[j]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) st += __shfl_xor_sync(0xffffffffu, st, o);
      for (int j = lane; j < C; j += 32) {
        const float s = x[j] - lse, tj = t[j], et = expf(tj);
        const float d = (expf(s) * st - et) * sc;
        store_planes(dl_hi, dl_lo, r * lddl + j, d);
        atomicAdd(&s_col[j], d);
        loss = fmaf(et, tj - s, loss);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
  }
  if (lane == 0) s_loss[warp] = loss;
  __syncthreads();
  for (int j = threadIdx.x; j < C; j += blockDim.x) atomicAdd(dbias + j, s_col[j]);
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_loss[w];
    atomicAdd(pp->loss_sum, t / static_cast<float>(Rg));
  }
}

// torch.optim.Adam (amsgrad=False, L2 weight decay) over the flat parameter buffer.  Hyper-parameters
// and the step number come from the pass's device state (pp, ctr: the fused student step) or, when pp
// is null, by value (glnn_adam_step_f32: teacher training, the injected-gradient parity test).
__device__ __forceinline__ void adam_scalars(const AdamHp& h, double t, float* step_size, float* bc2_sqrt) {
  const double bc1 = 1.0 - pow(h.beta1, t);
  const double bc2 = 1.0 - pow(h.beta2, t);
  *step_size = static_cast<float>(h.lr / bc1);
  *bc2_sqrt = static_cast<float>(sqrt(bc2));
}
__device__ __forceinline__ float adam_update(const AdamHp& h, float step_size, float bc2s, float p,
                                             float g, float* m, float* v) {
  if (h.wd != 0.f) g = fmaf(h.wd, p, g);
  const float mi = *m * h.b1 + g * h.omb1;          // exp_avg.mul_(beta1).add_(grad, alpha=1 - beta1)
  const float vi = *v * h.b2 + g * g * h.omb2;      // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  *m = mi;
  *v = vi;
  const float denom = sqrtf(vi) / bc2s + h.eps;
  return p - step_size * (mi / denom);
}
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v,
                                                   int64_t n, const PassParams* __restrict__ pp,
                                                   const int* __restrict__ ctr, const AdamHp hv,
                                                   int64_t step_by_value) {
  __shared__ float s_step, s_bc2;
  const AdamHp h = pp ? pp->hp : hv;
  if (threadIdx.x == 0)
    adam_scalars(h, static_cast<double>(pp ? pp->step0 + *ctr + 1 : step_by_value), &s_step, &s_bc2);
  __syncthreads();
  const float step_size = s_step, bc2s = s_bc2;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float mi = m[i], vi = v[i];
    p[i] = adam_update(h, step_size, bc2s, p[i], g[i], &mi, &vi);
    m[i] = mi;
    v[i] = vi;
  }
}

// Data-parallel optimizer step: reduce-scatter of the gradients, Adam and all-gather of the new
// parameters in ONE kernel over peer memory.  This rank owns elements [lo, hi) of the flat buffers:
// it waits until every rank has published its gradients, sums that slice over the ranks in rank
// order (P2P loads, 16 bytes per lane), applies torch.optim.Adam with its slice of the moments and
// stores the new parameters into every rank's parameter buffer (P2P stores).  The last CTA then
// signals "parameters updated".  Moments outside [lo, hi) are not touched (the host gathers them
// once per pass).
__global__ void __launch_bounds__(256) adam_dp_kernel(const DpDev dp, float* __restrict__ m,
                                                      float* __restrict__ v, int64_t lo, int64_t hi,
                                                      const PassParams* __restrict__ pp,
                                                      const int* __restrict__ ctr, int slot_grads,
                                                      int slot_params) {
  __shared__ float s_step, s_bc2;
  const uint32_t epoch = dp_epoch(dp);
  const AdamHp h = pp->hp;
  if (threadIdx.x == 0) adam_scalars(h, static_cast<double>(pp->step0 + *ctr + 1), &s_step, &s_bc2);
  dp_wait_all(dp, slot_grads, epoch);  // includes __syncthreads
  const float step_size = s_step, bc2s = s_bc2;
  float* p_loc = reinterpret_cast<float*>(dp.base[dp.rank] + dp.off_params);
  // lo, hi are multiples of 4 and the buffers 16-byte aligned
  for (int64_t i = lo + 4 * (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x); i < hi;
       i += 4LL * gridDim.x * blockDim.x) {
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < dp.world; ++r) {
      const float4* gp = reinterpret_cast<const float4*>(dp.base[r] + dp.off_grads) + (i >> 2);
      float4 t;
      asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];"
                   : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
                   : "l"(gp));
      g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
    }
    const float4 p4 = *reinterpret_cast<const float4*>(p_loc + i);
    float4 m4 = *reinterpret_cast<const float4*>(m + i), v4 = *reinterpret_cast<const float4*>(v + i);
    float gi[4] = {g.x, g.y, g.z, g.w}, pi[4] = {p4.x, p4.y, p4.z, p4.w};
    float mi[4] = {m4.x, m4.y, m4.z, m4.w}, vi[4] = {v4.x, v4.y, v4.z, v4.w}, po[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) po[j] = adam_update(h, step_size, bc2s, pi[j], gi[j], &mi[j], &vi[j]);
    *reinterpret_cast<float4*>(m + i) = make_float4(mi[0], mi[1], mi[2], mi[3]);
    *reinterpret_cast<float4*>(v + i) = make_float4(vi[0], vi[1], vi[2], vi[3]);
    const float4 out = make_float4(po[0], po[1], po[2], po[3]);
    for (int r = 0; r < dp.world; ++r)
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(dp.base[r] + dp.off_params) + i) = out;
  }
  dp_last_cta_signals(dp, slot_params, epoch, gridDim.x);
}

// "My gradients are complete": one thread, after every backward kernel of the step in stream order.
__global__ void dp_signal_kernel(const DpDev dp, int slot) {
  if (threadIdx.x == 0) dp_signal_all(dp, slot, dp_epoch(dp));
}
// Start of a step: the parameters written by every rank's optimizer kernel of the PREVIOUS step have
// landed in this rank's buffer (epoch - 1; trivially true for the first step after allocation).
__global__ void dp_wait_kernel(const DpDev dp, int slot) { dp_wait_all(dp, slot, dp_epoch(dp) - 1); }

__global__ void advance_kernel(int* ctr, int64_t* nbt, int n_norm, uint32_t* epoch) {
  if (threadIdx.x == 0) {
    *ctr += 1;
    if (epoch) *epoch += 1;
  }
  if (nbt && threadIdx.x < n_norm) nbt[threadIdx.x] += 1;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// Tensor-core projection on plane operands: C = op(A) op(B) (+ bias, BN affine, ReLU), fp32 and/or
// plane output.
static int gemm_p(const PlaneRef& A, int tA, const PlaneRef& B, int tB, float* C, int64_t ldc,
                  const PlaneRef* Cp, int64_t M, int64_t N, int64_t K, const float* bias,
                  const float* col_scale, const float* col_shift, int relu, cudaStream_t st) {
  return glnn_gemm_bf16x3_planes(A.hi, A.lo, A.ld, tA, B.hi, B.lo, B.ld, tB, C, ldc,
                                 Cp ? Cp->hi : nullptr, Cp ? Cp->lo : nullptr, Cp ? Cp->ld : 0, M, N, K,
                                 nullptr, bias, col_scale, col_shift, relu, st);
}

// Data-parallel ownership of the flat buffers: equal slices of a multiple of 4 elements; the flat
// buffers are padded to P4(P) = slice * world elements by the host.
static inline int64_t dp_slice(int64_t P, int world) { return (P + 4LL * world - 1) / (4LL * world) * 4; }

struct StepCtx {
  Dims d;
  ParamLayout pl;
  WsLayout wl;
  float *params, *grads, *m, *v, *bn_stats;
  int64_t* nbt;
  float* ws;
  DpDev dp;  // dp.world == 1: single GPU
};
constexpr int kSlotGrads = kSlots - 2, kSlotParams = kSlots - 1;

#define BN_LAUNCH(name) (bv == 4 ? name<4> : name<1>)
// Side stream + fork / join events of the calling thread (weight-gradient GEMMs run beside the
// activation-gradient chain; works identically under stream capture and in direct launches).
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  bool ok = false;
};
static SideStream& side_stream() {
  static thread_local SideStream s;
  static const bool disabled = getenv("GLNN_MLP_NO_FORK") != nullptr;
  if (!s.stream && !disabled) {
    if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) == cudaSuccess &&
        cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) == cudaSuccess &&
        cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) == cudaSuccess)
      s.ok = true;
  }
  return s;
}
static int enqueue_step(const StepCtx& c, cudaStream_t st) {
  SideStream& side = side_stream();
  const Dims& d = c.d;
  const int64_t R = d.R;
  const PassParams* pp = reinterpret_cast<const PassParams*>(c.ws + c.wl.pp);
  int* ctr = reinterpret_cast<int*>(c.ws + c.wl.ctr);
  float* tgt = c.ws + c.wl.tgt;
  float* part = c.ws + c.wl.part;
  const int rs = c.wl.rs;
  const int nlay = d.L - 1;
  const dim3 blk(32, 8);
  const int bv = bn_vec(d);
  const int col_tiles = (d.H + 32 * bv - 1) / (32 * bv);
  // the apply kernels run as ONE wave of 2 resident CTAs per SM (their register budget), each CTA
  // looping over its rows: a second partial wave cost more than the longer loop
  const int64_t want_tiles = std::max<int64_t>(1, std::min<int64_t>(2LL * 148 / col_tiles, R / 32));
  const int rows_per_block = static_cast<int>((R + want_tiles - 1) / want_tiles);
  const int row_tiles = static_cast<int>((R + rows_per_block - 1) / rows_per_block);
  int rc;

  const PlaneRef xb = plane_ref(c.ws, c.wl.xbP, R, d.F);
  const PlaneRef dlog = plane_ref(c.ws, c.wl.dlogP, R, d.C);
  const PlaneRef dzp = plane_ref(c.ws, c.wl.dzP, R, d.H);
  const DpDev& dp = c.dp;
  const bool is_dp = dp.world > 1;
  if (is_dp) {  // every rank's optimizer kernel of the previous step has written our parameters
    dp_wait_kernel<<<1, 32, 0, st>>>(dp, kSlotParams);
    GLNN_LAUNCH_OK("dp_wait_kernel");
  }
  // Top of the step, two independent branches: the side stream turns the weights into planes (they
  // changed in the previous Adam update, or were loaded from a state_dict between passes) and zeroes
  // the bias-gradient accumulators while the main stream gathers the batch.
  PlaneRef wp[16], ap[16];
  cudaStream_t ws_st = st;
  if (side.ok) {
    GLNN_CUDA_OK(cudaEventRecord(side.fork, st));
    GLNN_CUDA_OK(cudaStreamWaitEvent(side.stream, side.fork, 0));
    ws_st = side.stream;
  }
  for (int l = 0; l < d.L; ++l) {
    wp[l] = plane_ref(c.ws, c.wl.wP[l], out_dim(d, l), in_dim(d, l));
    if (l < d.L - 1) ap[l] = plane_ref(c.ws, c.wl.aP[l], R, d.H);
    rc = split_planes(c.params + c.pl.w[l], in_dim(d, l), out_dim(d, l), in_dim(d, l), wp[l].hi,
                      wp[l].lo, wp[l].ld, ws_st);
    if (rc != 0) return rc;
    GLNN_CUDA_OK(cudaMemsetAsync(c.grads + c.pl.b[l], 0, sizeof(float) * out_dim(d, l), ws_st));
  }
  if (side.ok) GLNN_CUDA_OK(cudaEventRecord(side.join, side.stream));

  gather_kernel<<<static_cast<unsigned>((R + 7) / 8), 256, 0, st>>>(pp, ctr, R, d.Rg, d.rank * R, d.F,
                                                                    d.C, xb.hi, xb.lo, xb.ld, tgt);
  GLNN_LAUNCH_OK("gather_kernel");
  if (side.ok) GLNN_CUDA_OK(cudaStreamWaitEvent(st, side.join, 0));

  // forward
  const PlaneRef* h = &xb;
  for (int l = 0; l < d.L; ++l) {
    const int din = in_dim(d, l), dout = out_dim(d, l);
    float* z = (l == d.L - 1) ? c.ws + c.wl.logits : c.ws + c.wl.z[l];
    rc = gemm_p(*h, 0, wp[l], 1, z, dout, nullptr, R, dout, din, c.params + c.pl.b[l], nullptr, nullptr,
                0, st);
    if (rc != 0) return rc;
    if (l == d.L - 1) break;
    if (d.norm) {
      BN_LAUNCH(bn_stats_kernel)<<<dim3(col_tiles, rs), blk, 0, st>>>(z, R, d.H, rs, part, dp, l);
      GLNN_LAUNCH_OK("bn_stats_kernel");
    }
    BN_LAUNCH(bn_apply_kernel)<<<dim3(col_tiles, row_tiles), blk, 0, st>>>(
        z, ap[l].hi, ap[l].lo, ap[l].ld, R, d.H, rs, part,
        d.norm ? c.params + c.pl.gamma[l] : nullptr, d.norm ? c.params + c.pl.beta[l] : nullptr,
        d.norm ? c.bn_stats + 2LL * l * d.H : nullptr,
        d.norm ? c.bn_stats + (2LL * l + 1) * d.H : nullptr, c.ws + c.wl.mean[l],
        c.ws + c.wl.invstd[l], d.norm, d.bn_eps, d.bn_mom, d.p_drop, pp, ctr, l, nlay,
        rows_per_block, dp, l, d.Rg);
    GLNN_LAUNCH_OK("bn_apply_kernel");
    h = &ap[l];
  }

  // loss + dlogits (+ last bias grad)
  loss_kernel<<<static_cast<unsigned>((R + 7) / 8), 256, sizeof(float) * (d.C + 8), st>>>(
      c.ws + c.wl.logits, tgt, R, d.Rg, d.C, pp, dlog.hi, dlog.lo, dlog.ld, c.grads + c.pl.b[d.L - 1]);
  GLNN_LAUNCH_OK("loss_kernel");

  // backward
  const PlaneRef* dz = &dlog;  // gradient w.r.t. the output of layer l's Linear
  for (int l = d.L - 1; l >= 0; --l) {
    const int din = in_dim(d, l), dout = out_dim(d, l);
    const PlaneRef& hin = (l == 0) ? xb : ap[l - 1];
    // dW_l [dout, din] = dz^T [dout, R] * hin [R, din]   (both operands MN-major).  Nothing on the
    // way down to the input depends on it, so for l > 0 it runs on a forked side stream next to
    // dA_{l-1} and the BatchNorm-backward statistics (it fills the SMs their partial waves leave
    // idle) and is joined before bn_bwd_apply overwrites the dz planes it reads.
    const bool forked = l > 0 && side.ok;
    if (forked) {
      GLNN_CUDA_OK(cudaEventRecord(side.fork, st));
      GLNN_CUDA_OK(cudaStreamWaitEvent(side.stream, side.fork, 0));
    }
    rc = gemm_p(*dz, 1, hin, 0, c.grads + c.pl.w[l], din, nullptr, dout, din, R, nullptr, nullptr,
                nullptr, 0, forked ? side.stream : st);
    if (rc != 0) return rc;
    if (forked) GLNN_CUDA_OK(cudaEventRecord(side.join, side.stream));
    if (l == 0) break;
    // dA_{l-1} [R, din] = dz [R, dout] * W_l [dout, din]  (W as MN-major B operand)
    float* da = c.ws + c.wl.dh[l & 1];
    rc = gemm_p(*dz, 0, wp[l], 0, da, din, nullptr, R, din, dout, nullptr, nullptr, nullptr, 0, st);
    if (rc != 0) return rc;
    const int k = l - 1;  // hidden layer whose activation we go back through
    const float* gam = d.norm ? c.params + c.pl.gamma[k] : nullptr;
    const float* bet = d.norm ? c.params + c.pl.beta[k] : nullptr;
    if (d.norm) {
      BN_LAUNCH(bn_bwd_stats_kernel)<<<dim3(col_tiles, rs), blk, 0, st>>>(
          da, c.ws + c.wl.z[k], R, d.H, rs, gam, bet, c.ws + c.wl.mean[k], c.ws + c.wl.invstd[k],
          d.norm, d.p_drop, pp, ctr, k, nlay, part, dp, nlay + k, d.Rg);
      GLNN_LAUNCH_OK("bn_bwd_stats_kernel");
    }
    if (forked) GLNN_CUDA_OK(cudaStreamWaitEvent(st, side.join, 0));
    BN_LAUNCH(bn_bwd_apply_kernel)<<<dim3(col_tiles, row_tiles), blk, 0, st>>>(
        da, dzp.hi, dzp.lo, dzp.ld, c.ws + c.wl.z[k], R, d.H, rs, part, gam, bet, c.ws + c.wl.mean[k],
        c.ws + c.wl.invstd[k], d.norm, d.p_drop, pp, ctr, k, nlay,
        d.norm ? c.grads + c.pl.gamma[k] : nullptr, d.norm ? c.grads + c.pl.beta[k] : nullptr,
        c.grads + c.pl.b[k], rows_per_block, dp, nlay + k, d.Rg);
    GLNN_LAUNCH_OK("bn_bwd_apply_kernel");
    dz = &dzp;
  }

  const int64_t P = c.pl.total;
  if (is_dp) {
    dp_signal_kernel<<<1, 32, 0, st>>>(dp, kSlotGrads);
    GLNN_LAUNCH_OK("dp_signal_kernel");
    const int64_t slice = dp_slice(P, dp.world);
    const int64_t lo = slice * dp.rank, hi = lo + slice;  // buffers hold slice * world elements
    const int64_t n4 = std::max<int64_t>((hi - lo) / 4, 1);
    const unsigned ablocks = static_cast<unsigned>(std::min<int64_t>((n4 + 255) / 256, 2LL * sm_count()));
    adam_dp_kernel<<<ablocks, 256, 0, st>>>(dp, c.m, c.v, lo, hi, pp, ctr, kSlotGrads, kSlotParams);
    GLNN_LAUNCH_OK("adam_dp_kernel");
  } else {
    const unsigned ablocks = static_cast<unsigned>(std::min<int64_t>((P + 255) / 256, 8LL * sm_count()));
    adam_kernel<<<ablocks, 256, 0, st>>>(c.params, c.grads, c.m, c.v, P, pp, ctr, AdamHp{}, 0);
    GLNN_LAUNCH_OK("adam_kernel");
  }
  advance_kernel<<<1, 32, 0, st>>>(ctr, d.norm ? c.nbt : nullptr, d.norm ? d.L - 1 : 0,
                                   is_dp ? reinterpret_cast<uint32_t*>(dp.base[dp.rank] + dp.off_epoch)
                                         : nullptr);
  GLNN_LAUNCH_OK("advance_kernel");
  return 0;
}

// graph cache: one executable graph per distinct (shape, buffer set)
struct GraphKey {
  Dims d;
  const void *params, *grads, *m, *v, *bn, *nbt, *ws;
  DpDev dp;
  bool operator<(const GraphKey& o) const {
    return memcmp(this, &o, sizeof(GraphKey)) < 0;
  }
};
static std::mutex g_mu;
static std::map<GraphKey, cudaGraphExec_t> g_graphs;
static cudaStream_t g_cap_stream = nullptr;

static int get_graph(const StepCtx& c, cudaGraphExec_t* out) {
  GraphKey k;
  memset(&k, 0, sizeof(k));
  k.d = c.d; k.params = c.params; k.grads = c.grads; k.m = c.m; k.v = c.v; k.bn = c.bn_stats;
  k.nbt = c.nbt; k.ws = c.ws; k.dp = c.dp;
  std::lock_guard<std::mutex> lock(g_mu);
  auto it = g_graphs.find(k);
  if (it != g_graphs.end()) { *out = it->second; return 0; }
  if (!g_cap_stream) GLNN_CUDA_OK(cudaStreamCreateWithFlags(&g_cap_stream, cudaStreamNonBlocking));
  if (g_graphs.size() >= 64) {  // bounded cache
    for (auto& kv : g_graphs) cudaGraphExecDestroy(kv.second);
    g_graphs.clear();
  }
  GLNN_CUDA_OK(cudaStreamBeginCapture(g_cap_stream, cudaStreamCaptureModeThreadLocal));
  int rc = enqueue_step(c, g_cap_stream);
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamEndCapture(g_cap_stream, &graph);
  if (rc != 0) { if (graph) cudaGraphDestroy(graph); return rc; }
  if (e != cudaSuccess) { set_error("graph capture failed: %s", cudaGetErrorString(e)); return (int)e; }
  cudaGraphExec_t exec = nullptr;
  e = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) { set_error("graph instantiate failed: %s", cudaGetErrorString(e)); return (int)e; }
  g_graphs[k] = exec;
  *out = exec;
  return 0;
}

static int make_dims(const glnn_mlp_desc* desc, int64_t rows, Dims* d) {
  GLNN_REQUIRE(desc, GLNN_ERR_ARG, "mlp: null desc");
  GLNN_REQUIRE(desc->num_layers >= 1 && desc->num_layers <= 16, GLNN_ERR_SHAPE,
               "mlp: num_layers must be in [1,16]");
  GLNN_REQUIRE(desc->feat_dim > 0 && desc->label_dim > 0 && (desc->num_layers == 1 || desc->hidden_dim > 0),
               GLNN_ERR_SHAPE, "mlp: non-positive dimension");
  GLNN_REQUIRE(desc->norm == 0 || desc->norm == 1, GLNN_ERR_ARG, "mlp: norm must be 0 or 1");
  GLNN_REQUIRE(desc->dropout >= 0.f && desc->dropout < 1.f, GLNN_ERR_ARG, "mlp: dropout in [0,1)");
  memset(d, 0, sizeof(Dims));
  d->L = desc->num_layers; d->F = desc->feat_dim; d->H = desc->num_layers == 1 ? 1 : desc->hidden_dim;
  d->C = desc->label_dim; d->norm = desc->num_layers == 1 ? 0 : desc->norm;
  d->p_drop = desc->dropout; d->bn_eps = desc->bn_eps; d->bn_mom = desc->bn_momentum; d->R = rows;
  d->Rg = rows; d->world = 1; d->rank = 0;
  return 0;
}

}  // namespace glnn

extern "C" int64_t glnn_mlp_param_count(const glnn_mlp_desc* desc) {
  glnn::Dims d;
  if (glnn::make_dims(desc, 1, &d) != 0) return -1;
  return glnn::param_layout(d).total;
}

extern "C" int64_t glnn_mlp_bn_stat_count(const glnn_mlp_desc* desc) {
  glnn::Dims d;
  if (glnn::make_dims(desc, 1, &d) != 0) return -1;
  return d.norm ? 2LL * d.H * (d.L - 1) : 0;
}

extern "C" int64_t glnn_mlp_workspace_bytes(const glnn_mlp_desc* desc, int64_t rows) {
  glnn::Dims d;
  if (rows < 1 || glnn::make_dims(desc, rows, &d) != 0) return -1;
  return glnn::ws_layout(d, true).total_bytes;
}

namespace glnn {

// Control area at the start of every rank's symmetric region (see DpDev).
constexpr int64_t kDpOffFlags = 0, kDpOffDone = 1280, kDpOffEpoch = 1536, kDpOffPart = 2048;
static int64_t dp_control_bytes(const Dims& d_local, int world) {
  const int rs = row_splits(d_local);
  const int64_t part = 2LL * std::max(1, d_local.L - 1) * world * rs * 2 * d_local.H * sizeof(float);
  return (kDpOffPart + part + 255) / 256 * 256;
}

// bs = rows of the GLOBAL batch; grp == nullptr: single GPU.
static int train_pass_impl(const glnn_dp_group* grp, const glnn_mlp_desc* desc, float* params,
                           float* grads, float* exp_avg, float* exp_avg_sq, float* bn_stats,
                           int64_t* num_batches_tracked, int64_t adam_step0,
                           const glnn_adam_hparams* hp, const float* X, int64_t ldx,
                           const void* target, int target_kind, const int64_t* perm, int64_t nb,
                           int64_t bs, const uint8_t* drop_masks, uint64_t seed, float lamb,
                           float* loss_sum, void* workspace, int64_t workspace_bytes,
                           glnn_stream_t stream) {
  Dims d;
  GLNN_REQUIRE(bs >= 1 && nb >= 0, GLNN_ERR_ARG, "mlp_train_pass: bad batch geometry");
  const int world = grp ? grp->world : 1;
  GLNN_REQUIRE(world >= 1 && world <= kMaxPeers && (!grp || (grp->rank >= 0 && grp->rank < world)),
               GLNN_ERR_ARG, "mlp_train_pass_dp: world must be 1..%d and rank inside it", kMaxPeers);
  GLNN_REQUIRE(bs % world == 0, GLNN_ERR_SHAPE,
               "mlp_train_pass_dp: the global batch (%lld rows) must split evenly over %d ranks",
               (long long)bs, world);
  int rc = make_dims(desc, bs / world, &d);
  if (rc != 0) return rc;
  d.Rg = bs; d.world = world; d.rank = grp ? grp->rank : 0;
  if (nb == 0) return 0;
  GLNN_REQUIRE(params && grads && exp_avg && exp_avg_sq && hp && X && target && perm && loss_sum &&
                   workspace, GLNN_ERR_ARG, "mlp_train_pass: null pointer");
  GLNN_REQUIRE(!d.norm || bn_stats, GLNN_ERR_ARG, "mlp_train_pass: bn_stats required with norm=1");
  GLNN_REQUIRE(d.L - 1 <= 15, GLNN_ERR_SHAPE, "mlp_train_pass: at most 16 layers");
  GLNN_REQUIRE(!d.norm || bs >= 2, GLNN_ERR_SHAPE,
               "mlp_train_pass: BatchNorm needs more than 1 row per batch (torch raises too)");
  GLNN_REQUIRE(target_kind == 0 || target_kind == 1, GLNN_ERR_ARG, "mlp_train_pass: target_kind");
  GLNN_REQUIRE(ldx >= d.F, GLNN_ERR_SHAPE, "mlp_train_pass: ldx < feat_dim");
  GLNN_REQUIRE(nb * bs <= kPermCap, GLNN_ERR_SHAPE,
               "mlp_train_pass: at most %lld rows per call (split the pass)", (long long)kPermCap);
  (void)sm_count();  // resolve device attributes before any stream capture
  GLNN_REQUIRE(aligned16(workspace), GLNN_ERR_ALIGN, "mlp_train_pass: workspace alignment");
  StepCtx c;
  c.d = d;
  c.pl = param_layout(d);
  c.wl = ws_layout(d, true);
  GLNN_REQUIRE(workspace_bytes >= c.wl.total_bytes, GLNN_ERR_WORKSPACE,
               "mlp_train_pass: workspace %lld < %lld bytes", (long long)workspace_bytes,
               (long long)c.wl.total_bytes);
  c.params = params; c.grads = grads; c.m = exp_avg; c.v = exp_avg_sq; c.bn_stats = bn_stats;
  c.nbt = num_batches_tracked; c.ws = static_cast<float*>(workspace);
  memset(&c.dp, 0, sizeof(c.dp));
  c.dp.world = 1;
  if (world > 1) {
    c.dp.world = world; c.dp.rank = grp->rank;
    for (int r = 0; r < world; ++r) {
      GLNN_REQUIRE(grp->base[r] != nullptr, GLNN_ERR_ARG, "mlp_train_pass_dp: null peer base %d", r);
      c.dp.base[r] = static_cast<unsigned char*>(grp->base[r]);
    }
    const int64_t ctrl = dp_control_bytes(d, world);
    const int64_t flat = dp_slice(c.pl.total, world) * world * static_cast<int64_t>(sizeof(float));
    unsigned char* base = c.dp.base[grp->rank];
    const int64_t op = reinterpret_cast<unsigned char*>(params) - base;
    const int64_t og = reinterpret_cast<unsigned char*>(grads) - base;
    GLNN_REQUIRE(op >= ctrl && og >= ctrl && op + flat <= grp->bytes && og + flat <= grp->bytes &&
                     (op + flat <= og || og + flat <= op) && op % 16 == 0 && og % 16 == 0,
                 GLNN_ERR_ARG,
                 "mlp_train_pass_dp: params / grads must be 16-byte aligned, glnn_mlp_dp_flat_count() "
                 "elements long and inside this rank's symmetric region behind its control area");
    c.dp.off_flags = kDpOffFlags; c.dp.off_done = kDpOffDone; c.dp.off_epoch = kDpOffEpoch;
    c.dp.off_part = kDpOffPart; c.dp.off_params = op; c.dp.off_grads = og;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  // per-pass state: permutation copy, pass parameters, step counter
  int64_t* perm_dev = reinterpret_cast<int64_t*>(c.ws + c.wl.perm);
  cudaPointerAttributes attr;
  cudaMemcpyKind kind = cudaMemcpyDefault;
  if (cudaPointerGetAttributes(&attr, perm) != cudaSuccess) cudaGetLastError();
  GLNN_CUDA_OK(cudaMemcpyAsync(perm_dev, perm, sizeof(int64_t) * nb * bs, kind, st));
  PassParams pp;
  memset(&pp, 0, sizeof(pp));
  pp.X = X; pp.ldx = ldx; pp.target = target; pp.perm = perm_dev; pp.masks = drop_masks;
  pp.loss_sum = loss_sum; pp.step0 = adam_step0; pp.seed = seed; pp.kind = target_kind;
  pp.lamb = lamb; pp.hp = make_adam_hp(*hp);
  GLNN_CUDA_OK(cudaMemcpyAsync(c.ws + c.wl.pp, &pp, sizeof(pp), cudaMemcpyHostToDevice, st));
  GLNN_CUDA_OK(cudaMemsetAsync(c.ws + c.wl.ctr, 0, sizeof(int), st));
  // `pp` lives on this stack frame: the copy above must have been staged before we return
  // (cudaMemcpyAsync from pageable memory returns after staging, so this is already guaranteed).

  static const bool no_graph = getenv("GLNN_NO_GRAPH") != nullptr;
  if (!no_graph && nb >= 2) {
    cudaGraphExec_t exec = nullptr;
    rc = get_graph(c, &exec);
    if (rc != 0) return rc;
    for (int64_t i = 0; i < nb; ++i) GLNN_CUDA_OK(cudaGraphLaunch(exec, st));
    return 0;
  }
  for (int64_t i = 0; i < nb; ++i) {
    rc = enqueue_step(c, st);
    if (rc != 0) return rc;
  }
  return 0;
}

}  // namespace glnn

extern "C" int glnn_mlp_train_pass(const glnn_mlp_desc* desc, float* params, float* grads,
                                   float* exp_avg, float* exp_avg_sq, float* bn_stats,
                                   int64_t* num_batches_tracked, int64_t adam_step0,
                                   const glnn_adam_hparams* hp, const float* X, int64_t ldx,
                                   const void* target, int target_kind, const int64_t* perm,
                                   int64_t nb, int64_t bs, const uint8_t* drop_masks, uint64_t seed,
                                   float lamb, float* loss_sum, void* workspace,
                                   int64_t workspace_bytes, glnn_stream_t stream) {
  return glnn::train_pass_impl(nullptr, desc, params, grads, exp_avg, exp_avg_sq, bn_stats,
                               num_batches_tracked, adam_step0, hp, X, ldx, target, target_kind, perm,
                               nb, bs, drop_masks, seed, lamb, loss_sum, workspace, workspace_bytes,
                               stream);
}

extern "C" int glnn_adam_step_f32(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                                  int64_t n, int64_t step, const glnn_adam_hparams* hp,
                                  glnn_stream_t stream) {
  using namespace glnn;
  GLNN_REQUIRE(n >= 0 && step >= 1 && hp, GLNN_ERR_ARG, "adam_step: n >= 0, step >= 1, hp != NULL");
  if (n == 0) return 0;
  GLNN_REQUIRE(params && grads && exp_avg && exp_avg_sq, GLNN_ERR_ARG, "adam_step: null pointer");
  const AdamHp hv = make_adam_hp(*hp);
  const unsigned blocks = static_cast<unsigned>(std::min<int64_t>((n + 255) / 256, 8LL * sm_count()));
  adam_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(params, grads, exp_avg, exp_avg_sq, n,
                                                                     nullptr, nullptr, hv, step);
  GLNN_LAUNCH_OK("adam_kernel");
  return 0;
}

extern "C" int64_t glnn_mlp_dp_control_bytes(const glnn_mlp_desc* desc, int64_t bs_global, int world) {
  glnn::Dims d;
  if (world < 1 || world > glnn::kMaxPeers || bs_global < world || bs_global % world != 0) return -1;
  if (glnn::make_dims(desc, bs_global / world, &d) != 0) return -1;
  return glnn::dp_control_bytes(d, world);
}

extern "C" int64_t glnn_mlp_dp_flat_count(const glnn_mlp_desc* desc, int world) {
  glnn::Dims d;
  if (world < 1 || world > glnn::kMaxPeers || glnn::make_dims(desc, 1, &d) != 0) return -1;
  return glnn::dp_slice(glnn::param_layout(d).total, world) * world;
}

extern "C" int glnn_mlp_dp_init(void* region_local, int64_t control_bytes, glnn_stream_t stream) {
  using namespace glnn;
  GLNN_REQUIRE(region_local && control_bytes >= kDpOffPart, GLNN_ERR_ARG, "mlp_dp_init: bad region");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GLNN_CUDA_OK(cudaMemsetAsync(region_local, 0, static_cast<size_t>(control_bytes), st));
  const uint32_t one = 1;
  GLNN_CUDA_OK(cudaMemcpyAsync(static_cast<unsigned char*>(region_local) + kDpOffEpoch, &one, sizeof(one),
                               cudaMemcpyHostToDevice, st));
  GLNN_CUDA_OK(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int glnn_mlp_train_pass_dp(const glnn_dp_group* grp, const glnn_mlp_desc* desc, float* params,
                                      float* grads, float* exp_avg, float* exp_avg_sq, float* bn_stats,
                                      int64_t* num_batches_tracked, int64_t adam_step0,
                                      const glnn_adam_hparams* hp, const float* X, int64_t ldx,
                                      const void* target, int target_kind, const int64_t* perm,
                                      int64_t nb, int64_t bs_global, const uint8_t* drop_masks,
                                      uint64_t seed, float lamb, float* loss_sum, void* workspace,
                                      int64_t workspace_bytes, glnn_stream_t stream) {
  using namespace glnn;
  GLNN_REQUIRE(grp != nullptr, GLNN_ERR_ARG, "mlp_train_pass_dp: null group");
  return train_pass_impl(grp->world > 1 ? grp : nullptr, desc, params, grads, exp_avg, exp_avg_sq,
                         bn_stats, num_batches_tracked, adam_step0, hp, X, ldx, target, target_kind,
                         perm, nb, bs_global, drop_masks, seed, lamb, loss_sum, workspace,
                         workspace_bytes, stream);
}

extern "C" int glnn_mlp_eval(const glnn_mlp_desc* desc, const float* params, const float* bn_stats,
                             const float* X, int64_t ldx, int64_t n, float* out, int64_t ldo,
                             int log_softmax, int64_t rows_per_chunk, void* workspace,
                             int64_t workspace_bytes, glnn_stream_t stream) {
  using namespace glnn;
  GLNN_REQUIRE(n >= 0 && rows_per_chunk >= 1, GLNN_ERR_ARG, "mlp_eval: bad sizes");
  Dims d;
  int rc = make_dims(desc, rows_per_chunk, &d);
  if (rc != 0) return rc;
  if (n == 0) return 0;
  GLNN_REQUIRE(params && X && out && workspace, GLNN_ERR_ARG, "mlp_eval: null pointer");
  GLNN_REQUIRE(!d.norm || bn_stats, GLNN_ERR_ARG, "mlp_eval: bn_stats required with norm=1");
  GLNN_REQUIRE(ldx >= d.F && ldo >= d.C, GLNN_ERR_SHAPE, "mlp_eval: leading dimension too small");
  const ParamLayout pl = param_layout(d);
  const WsLayout wl = ws_layout(d, false);
  GLNN_REQUIRE(workspace_bytes >= wl.total_bytes, GLNN_ERR_WORKSPACE,
               "mlp_eval: workspace %lld < %lld bytes", (long long)workspace_bytes,
               (long long)wl.total_bytes);
  float* ws = static_cast<float*>(workspace);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* fold = ws + wl.fold;
  if (d.norm)
    for (int l = 0; l < d.L - 1; ++l) {
      rc = glnn_bn_fold_f32(params + pl.gamma[l], params + pl.beta[l], bn_stats + 2LL * l * d.H,
                            bn_stats + (2LL * l + 1) * d.H, d.bn_eps, fold + 2LL * l * d.H,
                            fold + (2LL * l + 1) * d.H, d.H, st);
      if (rc != 0) return rc;
    }
  PlaneRef wp[16];
  for (int l = 0; l < d.L; ++l) {
    wp[l] = plane_ref(ws, wl.wP[l], out_dim(d, l), in_dim(d, l));
    rc = split_planes(params + pl.w[l], in_dim(d, l), out_dim(d, l), in_dim(d, l), wp[l].hi, wp[l].lo,
                      wp[l].ld, st);
    if (rc != 0) return rc;
  }
  for (int64_t r0 = 0; r0 < n; r0 += rows_per_chunk) {
    const int64_t R = std::min(rows_per_chunk, n - r0);
    // layout offsets are computed for rows_per_chunk rows, so a shorter last chunk fits
    const PlaneRef xb = plane_ref(ws, wl.xbP, rows_per_chunk, d.F);
    rc = split_planes(X + r0 * ldx, ldx, R, d.F, xb.hi, xb.lo, xb.ld, st);
    if (rc != 0) return rc;
    PlaneRef hbuf[2];
    if (d.L > 1) {
      hbuf[0] = plane_ref(ws, wl.aP[0], rows_per_chunk, d.H);
      hbuf[1] = plane_ref(ws, wl.dzP, rows_per_chunk, d.H);
    }
    const PlaneRef* h = &xb;
    for (int l = 0; l < d.L; ++l) {
      const int din = in_dim(d, l), dout = out_dim(d, l);
      const bool last = (l == d.L - 1);
      if (!last) {  // hidden layer: bias + eval-BN affine + ReLU in the epilogue, planes out
        rc = gemm_p(*h, 0, wp[l], 1, nullptr, 0, &hbuf[l & 1], R, dout, din, params + pl.b[l],
                    d.norm ? fold + 2LL * l * d.H : nullptr,
                    d.norm ? fold + (2LL * l + 1) * d.H : nullptr, 1, st);
        if (rc != 0) return rc;
        h = &hbuf[l & 1];
      } else {
        float* z = log_softmax ? ws + wl.logits : out + r0 * ldo;
        const int64_t ldz = log_softmax ? dout : ldo;
        rc = gemm_p(*h, 0, wp[l], 1, z, ldz, nullptr, R, dout, din, params + pl.b[l], nullptr, nullptr,
                    0, st);
        if (rc != 0) return rc;
      }
    }
    if (log_softmax) {
      rc = glnn_log_softmax_f32(ws + wl.logits, d.C, out + r0 * ldo, ldo, R, d.C, st);
      if (rc != 0) return rc;
    }
  }
  return 0;
}
