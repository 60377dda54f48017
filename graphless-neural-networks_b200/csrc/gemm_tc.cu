// fp32-faithful projection GEMM on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
// Numerics: every fp32 operand x is split into two bf16 values hi = bf16(x), lo = bf16(x - hi)
// (|x - hi - lo| <= 2^-18 |x|) and the product is accumulated in fp32 as hi*hi + hi*lo + lo*hi
// ("bf16x3"); the dropped lo*lo term is <= 2^-18 relative, so a dot product carries ~1e-5 worst-case
// and ~2e-6 typical relative error -- inside the 1e-4 parity bound with an order of magnitude to
// spare, at half the tensor-pipe cost of a 3xTF32 split (3 MMAs at the bf16 rate instead of 3 at
// the tf32 rate).  glnn_gemm_f32(impl=1) keeps the exact fp32 SIMT kernel as the anchor.
//
// Structure (one 128 x BN output tile per CTA, BK = 64 per stage, ring of stages):
//   warps 0-15 producers: coalesced 16-byte global loads of the fp32 A/B tiles, split into hi/lo in
//              registers, written to shared memory directly in the UMMA canonical SWIZZLE_128B
//              layout (K-major or MN-major, whichever matches the operand's contiguous dimension, so
//              no transposition is ever needed), fence.proxy.async, mbarrier arrive;
//   warp 16    one elected thread issues tcgen05.mma (kind::f16, M=128, N=BN, K=16): 4 k-steps x 3
//              products per stage, accumulator in TMEM; tcgen05.commit releases the stage;
//   warps 0-15 epilogue: tcgen05.ld of the accumulator, row scale / bias / eval-BN affine / ReLU,
//              16-byte stores.
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>

#include "gemm.cuh"
#include "tc_common.cuh"

namespace glnn {
namespace tc {

constexpr int BM = 128;
constexpr int BK = 64;          // fp32 elements per stage along K (= one 128-byte swizzle atom of bf16)
constexpr int NPROD = 512;      // producer threads (16 warps)
constexpr int NPWARPS = NPROD / 32;
constexpr int NTHREADS = NPROD + 32;

__device__ __forceinline__ float4 load4_guard(const float* p, int64_t i, int64_t lim) {
  if (i + 3 < lim) return ldg4(p + i);
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < lim) r.x = __ldg(p + i);
  if (i + 1 < lim) r.y = __ldg(p + i + 1);
  if (i + 2 < lim) r.z = __ldg(p + i + 2);
  return r;
}

// Register-staged tile loader.  All 16-byte global loads of a stage are issued first (A and B
// together: 12 independent loads per thread at BN = 256), then converted and stored, so that the
// memory system sees the whole stage in flight at once.
//
// KMAJ = true : ROWS (m or n) x 64 k from a matrix whose k index is contiguous (P[row*ld + k]).
//   smem: row r at r*128 B, 16-byte chunk c stored at chunk (c ^ (r & 7))       [K-major SW128]
// KMAJ = false: 64 k x ROWS (m or n) from a matrix whose m/n index is contiguous (P[k*ld + col]).
//   smem: 64-wide column block j at j*8192 B, k-group g (8 k) at g*1024 B, k row kk at kk*128 B,
//   16-byte chunk c stored at chunk (c ^ kk)                                     [MN-major SW128]
template <int ROWS, bool KMAJ>
struct TileIO {
  static constexpr int N4 = ROWS * (BK / 4) / NPROD;  // float4 per thread
  static_assert(N4 >= 1 && (ROWS * (BK / 4)) % NPROD == 0, "tile / producer-count mismatch");
  float4 v[N4];

  __device__ __forceinline__ void load(const float* __restrict__ P, int64_t ld, int64_t r0,
                                       int64_t rows, int64_t k0, int64_t kend, int tid) {
#pragma unroll
    for (int i = 0; i < N4; ++i) {
      const int f = tid + i * NPROD;
      if constexpr (KMAJ) {
        const int r = f >> 4, l = f & 15;
        const int64_t row = r0 + r;
        v[i] = (row < rows) ? load4_guard(P + row * ld, k0 + l * 4, kend)
                            : make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        constexpr int F4 = ROWS / 4;
        const int kk = f / F4, l = f % F4;
        const int64_t k = k0 + kk;
        v[i] = (k < kend) ? load4_guard(P + k * ld, r0 + l * 4, rows)
                          : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }

  __device__ __forceinline__ void store(uint8_t* s_hi, uint8_t* s_lo, int tid) const {
#pragma unroll
    for (int i = 0; i < N4; ++i) {
      const int f = tid + i * NPROD;
      uint32_t off;
      if constexpr (KMAJ) {
        const int r = f >> 4, l = f & 15;
        off = r * 128 + (((l >> 1) ^ (r & 7)) << 4) + ((l & 1) << 3);
      } else {
        constexpr int F4 = ROWS / 4;
        const int kk = f / F4, l = f % F4;
        const int j = l >> 4, c = (l & 15) >> 1, half = l & 1;
        off = j * 8192 + (kk >> 3) * 1024 + (kk & 7) * 128 + ((c ^ (kk & 7)) << 4) + (half << 3);
      }
      uint2 hi, lo;
      split4(v[i], hi, lo);
      *reinterpret_cast<uint2*>(s_hi + off) = hi;
      *reinterpret_cast<uint2*>(s_lo + off) = lo;
    }
  }
};

__device__ __forceinline__ void cp_async16(uint8_t* smem_dst, const void* gsrc, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc),
               "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Planes mode: the operand already exists as two bf16 planes, so a stage is filled by 16-byte
// cp.async copies (zero-filled past the matrix edge) that land directly at their swizzled position
// -- no registers, no conversion, a whole stage in flight per thread group.
template <int ROWS, bool KMAJ>
__device__ __forceinline__ void copy_planes(const uint16_t* __restrict__ hi,
                                            const uint16_t* __restrict__ lo, int64_t ld, int64_t r0,
                                            int64_t rows, int64_t k0, int64_t kend, uint8_t* s_hi,
                                            uint8_t* s_lo, int tid) {
  constexpr int NC = ROWS * 8 / NPROD;  // 16-byte chunks per thread and plane
  static_assert(NC >= 1, "tile too small for the producer count");
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    const int f = tid + i * NPROD;
    int64_t src;
    int valid;
    uint32_t off;
    if constexpr (KMAJ) {
      const int r = f >> 3, c = f & 7;
      const int64_t row = r0 + r, k = k0 + c * 8;
      valid = (row < rows) ? static_cast<int>(max(static_cast<long long>(0),
                                 min(static_cast<long long>(8), static_cast<long long>(kend - k)))) : 0;
      src = row * ld + k;
      off = r * 128 + ((c ^ (r & 7)) << 4);
    } else {
      constexpr int CH = ROWS / 8;  // chunks per k row
      const int kk = f / CH, ci = f % CH;
      const int64_t k = k0 + kk, col = r0 + ci * 8;
      valid = (k < kend) ? static_cast<int>(max(static_cast<long long>(0),
                               min(static_cast<long long>(8), static_cast<long long>(rows - col)))) : 0;
      src = k * ld + col;
      const int j = ci >> 3, c = ci & 7;
      off = j * 8192 + (kk >> 3) * 1024 + (kk & 7) * 128 + ((c ^ (kk & 7)) << 4);
    }
    if (valid <= 0) src = 0;
    cp_async16(s_hi + off, hi + src, static_cast<uint32_t>(valid * 2));
    cp_async16(s_lo + off, lo + src, static_cast<uint32_t>(valid * 2));
  }
}

template <int BN>
struct Smem {
  static constexpr int A_BYTES = BM * BK * 2;   // one bf16 plane
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 2 : 3;
  static constexpr int TOTAL = STAGES * STAGE + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

template <int BN, bool A_K, bool B_K, bool PLANES, bool SPLIT>
__global__ void __launch_bounds__(NTHREADS, 1) gemm_bf16x3_kernel(const GemmArgs g) {
  using S = Smem<BN>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // pointer arithmetic on the array itself keeps the shared address space visible to the compiler
  // (STS / LDS instead of generic ST / LD for the register-staged tiles and the epilogue staging)
  uint8_t* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + S::STAGES * S::STAGE);
  uint64_t* full = bars;                 // [STAGES]
  uint64_t* empty = bars + S::STAGES;    // [STAGES]
  uint64_t* accum = bars + 2 * S::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S::STAGES + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t m0 = static_cast<int64_t>(blockIdx.x) * BM;
  const int64_t n0 = static_cast<int64_t>(blockIdx.y) * BN;
  const int nkb_total = static_cast<int>((g.K + BK - 1) / BK);
  const int kb0 = SPLIT ? static_cast<int>(blockIdx.z) * g.kb_per_split : 0;
  const int nkb = SPLIT ? min(nkb_total, kb0 + g.kb_per_split) - kb0 : nkb_total;

  if (tid == 0) {
    // one arrival per producer WARP (lane 0 after __syncwarp): 512 per-thread arrivals on one
    // mbarrier serialise on its shared-memory word
    for (int s = 0; s < S::STAGES; ++s) {
      mbar_init(&full[s], NPWARPS);
      mbar_init(&empty[s], 1);
    }
    mbar_init(accum, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == NPWARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(static_cast<uint32_t>(BN))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < NPWARPS) {
    // ------------------------------- producers -------------------------------
    if constexpr (PLANES) {
      for (int i = 0; i < nkb; ++i) {
        if (i >= S::STAGES - 1) {  // the oldest stage in flight has landed: hand it to the MMA warp
          cp_async_wait<S::STAGES - 2>();
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&full[(i - (S::STAGES - 1)) % S::STAGES]);
        }
        const int s = i % S::STAGES, it = i / S::STAGES;
        if (it > 0) mbar_wait(&empty[s], (it - 1) & 1);
        uint8_t* st = tiles + s * S::STAGE;
        const int64_t k0 = static_cast<int64_t>(kb0 + i) * BK;
        copy_planes<BM, A_K>(g.Ah, g.Al, g.lda, m0, g.M, k0, g.K, st, st + S::A_BYTES, tid);
        copy_planes<BN, B_K>(g.Bh, g.Bl, g.ldb, n0, g.N, k0, g.K, st + 2 * S::A_BYTES,
                             st + 2 * S::A_BYTES + S::B_BYTES, tid);
        cp_async_commit();
      }
      cp_async_wait<0>();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0)
        for (int i = max(0, nkb - (S::STAGES - 1)); i < nkb; ++i) mbar_arrive(&full[i % S::STAGES]);
    } else {
    TileIO<BM, A_K> ta;
    TileIO<BN, B_K> tb;
    ta.load(g.A, g.lda, m0, g.M, 0, g.K, tid);
    tb.load(g.B, g.ldb, n0, g.N, 0, g.K, tid);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % S::STAGES, it = kb / S::STAGES;
      if (it > 0) mbar_wait(&empty[s], (it - 1) & 1);
      uint8_t* st = tiles + s * S::STAGE;
      uint8_t *a_hi = st, *a_lo = st + S::A_BYTES, *b_hi = st + 2 * S::A_BYTES,
              *b_lo = st + 2 * S::A_BYTES + S::B_BYTES;
      const int64_t k0 = static_cast<int64_t>(kb) * BK;
      ta.store(a_hi, a_lo, tid);
      tb.store(b_hi, b_lo, tid);
      if (kb + 1 < nkb) {  // next stage's loads go out before we signal this one
        ta.load(g.A, g.lda, m0, g.M, k0 + BK, g.K, tid);
        tb.load(g.B, g.ldb, n0, g.N, k0 + BK, g.K, tid);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
    }
    }
  } else if (lane == 0) {
    // ------------------------------- MMA issuer -------------------------------
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_K ? 0u : 1u) << 15) |
                           ((B_K ? 0u : 1u) << 16) | (static_cast<uint32_t>(BN >> 3) << 17) |
                           (static_cast<uint32_t>(BM >> 4) << 24);
    // K-major: 8-row groups 1024 B apart (SBO), LBO unused; advancing K by 16 bf16 = +32 B.
    // MN-major: 64-wide blocks 8192 B apart (LBO), 8-k groups 1024 B apart (SBO); K by 16 = +2048 B.
    const uint32_t a_lbo = A_K ? 16u : 8192u, b_lbo = B_K ? 16u : 8192u;
    const uint32_t a_step = A_K ? 2u : 128u, b_step = B_K ? 2u : 128u;  // in 16-byte units
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % S::STAGES, it = kb / S::STAGES;
      mbar_wait(&full[s], it & 1);
      tc_fence_after();
      const uint32_t st = smem_u32(tiles + s * S::STAGE);
      const uint64_t a_hi = make_desc(st, a_lbo, 1024), a_lo = make_desc(st + S::A_BYTES, a_lbo, 1024);
      const uint64_t b_hi = make_desc(st + 2 * S::A_BYTES, b_lbo, 1024),
                     b_lo = make_desc(st + 2 * S::A_BYTES + S::B_BYTES, b_lbo, 1024);
#pragma unroll
      for (int k = 0; k < BK / 16; ++k) {
        const uint64_t da = static_cast<uint64_t>(k * a_step), db = static_cast<uint64_t>(k * b_step);
        umma_bf16(tmem_base, a_hi + da, b_hi + db, idesc, (kb | k) != 0 ? 1u : 0u);  // kb is slice-local
        umma_bf16(tmem_base, a_hi + da, b_lo + db, idesc, 1u);
        umma_bf16(tmem_base, a_lo + da, b_hi + db, idesc, 1u);
      }
      umma_commit(&empty[s]);
    }
    umma_commit(accum);
  }

  if (warp < NPWARPS) {
    // ------------------------------- epilogue -------------------------------
    mbar_wait(accum, 0);
    tc_fence_after();
    const int q = warp & 3, part = warp >> 2;  // TMEM lane quarter, column part (4 parts)
    const int64_t m = m0 + q * 32 + lane;
    const float rs = (g.row_scale && m < g.M) ? __ldg(g.row_scale + m) : 1.f;
    constexpr int COLS_PER_WARP = BN / (NPWARPS / 4);
#pragma unroll 1
    for (int c0 = 0; c0 < COLS_PER_WARP; c0 += 16) {
      const int col = part * COLS_PER_WARP + c0;
      uint32_t r[16];
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + col;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,"
          "%15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
            "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
            "=r"(r[14]), "=r"(r[15])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      // epilogue math in registers: row scale, bias, (ReLU) affine (ReLU)
      float o[16];
      const bool lead = !SPLIT || blockIdx.z == 0;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int64_t n = n0 + col + j;
        float x = __uint_as_float(r[j]) * rs;
        if (n < g.N) {
          if (g.bias && lead) x += __ldg(g.bias + n);
          if constexpr (!SPLIT) {
            if (g.relu == 2) x = fmaxf(x, 0.f);
            if (g.col_scale) x = fmaf(x, __ldg(g.col_scale + n), __ldg(g.col_shift + n));
            if (g.relu == 1) x = fmaxf(x, 0.f);
          }
        }
        o[j] = x;
      }
      if constexpr (SPLIT) {
        if (m < g.M) {
          float* cp = g.C + m * g.ldc + n0 + col;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (n0 + col + j < g.N) atomicAdd(cp + j, o[j]);
        }
      } else {
        // stage the tile in shared memory (the operand ring is free: every MMA has completed) so
        // that global stores are full coalesced rows instead of one 16-byte piece per row
        float* cs = reinterpret_cast<float*>(tiles) + (q * 32 + lane) * (BN + 4) + col;
#pragma unroll
        for (int j4 = 0; j4 < 16; j4 += 4)
          *reinterpret_cast<float4*>(cs + j4) = make_float4(o[j4], o[j4 + 1], o[j4 + 2], o[j4 + 3]);
      }
    }
    if constexpr (!SPLIT) {
      asm volatile("bar.sync 1, %0;" ::"n"(NPROD) : "memory");
      const float* cs = reinterpret_cast<const float*>(tiles);
      for (int rr = warp; rr < BM; rr += NPWARPS) {
        const int64_t mr = m0 + rr;
        if (mr >= g.M) break;
#pragma unroll
        for (int c4 = lane; c4 < BN / 4; c4 += 32) {
          const int64_t n = n0 + c4 * 4;
          if (n >= g.N) continue;
          const float4 v = *reinterpret_cast<const float4*>(cs + rr * (BN + 4) + c4 * 4);
          emit4(g, mr, n, v);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NPWARPS) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(static_cast<uint32_t>(BN))
                 : "memory");
  }
}

template <int BN, bool A_K, bool B_K, bool PLANES, bool SPLIT>
static int launch(const GemmArgs& g, int splits, cudaStream_t st) {
  using S = Smem<BN>;
  static bool configured = false;
  auto kern = gemm_bf16x3_kernel<BN, A_K, B_K, PLANES, SPLIT>;
  if (!configured) {
    GLNN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    configured = true;
  }
  dim3 grid(static_cast<unsigned>((g.M + BM - 1) / BM), static_cast<unsigned>((g.N + BN - 1) / BN),
            static_cast<unsigned>(splits));
  kern<<<grid, NTHREADS, S::TOTAL, st>>>(g);
  GLNN_LAUNCH_OK("gemm_bf16x3_kernel");
  return 0;
}

template <int BN, bool PLANES, bool SPLIT>
static int launch_major(const GemmArgs& g, int splits, cudaStream_t st) {
  const bool a_k = !g.transA, b_k = g.transB;
  if (a_k && b_k) return launch<BN, true, true, PLANES, SPLIT>(g, splits, st);
  if (a_k && !b_k) return launch<BN, true, false, PLANES, SPLIT>(g, splits, st);
  if (!a_k && b_k) return launch<BN, false, true, PLANES, SPLIT>(g, splits, st);
  return launch<BN, false, false, PLANES, SPLIT>(g, splits, st);
}

}  // namespace tc

// Takes the shape when both operands can be read with aligned 16-byte loads along their contiguous
// dimension and the problem is big enough for a 128-row tile to pay off.
int gemm_tc(const GemmArgs& g0, cudaStream_t st, bool force, bool* taken) {
  *taken = false;
  static const bool disabled = getenv("GLNN_NO_TC") != nullptr;
  if (disabled && !force) return 0;
  GemmArgs g = g0;
  if (g.K < 1 || g.M < 1 || g.N < 1) return 0;
  if (!force && (g.M < 64 || g.N < 8)) return 0;
  if (!aligned16(g.A) || !aligned16(g.B) || (g.lda % 4) != 0 || (g.ldb % 4) != 0) return 0;
  // the guarded tail loader needs the contiguous extent to be a multiple of 4 only for alignment of
  // the NEXT row, which lda/ldb % 4 already guarantees
  if (!force && 2.0 * g.M * g.N * g.K < 2.0e8) return 0;  // tiny problems: launch-bound either way
  g.vecC = (g.ldc % 4 == 0) && aligned16(g.C);
  int rc;
  if (g.N > 128) rc = tc::launch_major<256, false, false>(g, 1, st);
  else rc = tc::launch_major<128, false, false>(g, 1, st);
  if (rc != 0) return rc;
  *taken = true;
  return 0;
}

// Operands given as bf16 hi / lo planes.  Split-K (atomic accumulation into a zeroed fp32 C) is used
// for skinny outputs with a long K (weight gradients), where one tile per CTA would leave most SMs
// idle.
int gemm_tc_planes(GemmArgs g, cudaStream_t st) {
  GLNN_REQUIRE(g.Ah && g.Al && g.Bh && g.Bl, GLNN_ERR_ARG, "gemm_planes: null operand plane");
  GLNN_REQUIRE(g.C || g.Ch || g.Cq, GLNN_ERR_ARG, "gemm_planes: no output");
  GLNN_REQUIRE(!g.Cq || (g.N % 8 == 0 && aligned16(g.Cq) && g.ldcq % 16 == 0 && g.ldcq >= 3 * g.N),
               GLNN_ERR_ALIGN,
               "gemm_planes: q24 output needs N %% 8 == 0, a 16-byte aligned buffer, ldq %% 16 == 0 and "
               "ldq >= 3 N");
  GLNN_REQUIRE((g.Ch == nullptr) == (g.Cl == nullptr), GLNN_ERR_ARG, "gemm_planes: C planes come in pairs");
  GLNN_REQUIRE(g.lda % 8 == 0 && g.ldb % 8 == 0 && aligned16(g.Ah) && aligned16(g.Al) &&
                   aligned16(g.Bh) && aligned16(g.Bl),
               GLNN_ERR_ALIGN, "gemm_planes: operand planes need 16-byte aligned rows (ld %% 8 == 0)");
  GLNN_REQUIRE(!g.Ch || (g.ldcp % 4 == 0 && g.ldcp >= g.N), GLNN_ERR_ALIGN, "gemm_planes: bad ldcp");
  g.vecC = g.C && (g.ldc % 4 == 0) && aligned16(g.C);
  {  // tall operands (the teacher's per-layer projection): persistent TMA-fed kernel
    g.kb_per_split = 0;
    bool taken = false;
    const int rc = gemm_tall_planes(g, st, &taken);
    if (rc != 0 || taken) return rc;
  }
  // 128 x 256 tiles (two-stage ring) when there are enough of them to fill the machine; otherwise
  // 128 x 128 tiles: twice the CTAs and a three-stage ring (two loads in flight) -- the small students
  // (bs 512, H 256 / 1024) run 4-32 CTAs per projection and are bound by the per-CTA timeline
  // (GLNN_TC_SMALL_TILES: 0 = round-1 rule, 1 = down to 128 columns, 2 = down to 64 columns)
  static const int small_tiles = getenv("GLNN_TC_SMALL_TILES") ? atoi(getenv("GLNN_TC_SMALL_TILES")) : 2;
  const int64_t mt = (g.M + tc::BM - 1) / tc::BM;
  int bn = g.N > 128 ? 256 : (g.N > 64 || small_tiles < 2 ? 128 : 64);
  if (small_tiles >= 1 && bn == 256 && mt * ((g.N + 255) / 256) * 2 <= sm_count()) bn = 128;
  if (small_tiles >= 2 && bn == 128 && g.N > 64 && mt * ((g.N + 127) / 128) * 2 <= sm_count()) bn = 64;
  const int64_t tiles = mt * ((g.N + bn - 1) / bn);
  const int nkb = static_cast<int>((g.K + tc::BK - 1) / tc::BK);
  int splits = 1;
  const bool linear = !g.relu && !g.col_scale && !g.Ch && !g.Cq && g.C;
  if (linear && tiles * 2 <= sm_count() && nkb >= 16) {
    splits = static_cast<int>(std::min<int64_t>((sm_count() + tiles - 1) / tiles, nkb / 4));
    if (splits < 1) splits = 1;
  }
  if (splits > 1) {
    g.kb_per_split = (nkb + splits - 1) / splits;
    splits = (nkb + g.kb_per_split - 1) / g.kb_per_split;
    if (g.ldc == g.N) {
      GLNN_CUDA_OK(cudaMemsetAsync(g.C, 0, sizeof(float) * g.M * g.N, st));
    } else {
      GLNN_CUDA_OK(cudaMemset2DAsync(g.C, sizeof(float) * g.ldc, 0, sizeof(float) * g.N, g.M, st));
    }
    return bn == 256 ? tc::launch_major<256, true, true>(g, splits, st)
                     : (bn == 128 ? tc::launch_major<128, true, true>(g, splits, st)
                                  : tc::launch_major<64, true, true>(g, splits, st));
  }
  g.kb_per_split = 0;
  return bn == 256 ? tc::launch_major<256, true, false>(g, 1, st)
                   : (bn == 128 ? tc::launch_major<128, true, false>(g, 1, st)
                                : tc::launch_major<64, true, false>(g, 1, st));
}

}  // namespace glnn
