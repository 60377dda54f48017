// Persistent, warp-specialised tcgen05 projection for TALL operands (the teacher's per-layer
// nn.Linear over all N nodes: M = 10^5..10^7 rows, K <= a few hundred, N <= 256), sm_100a only.
//
// Why a second kernel: gemm_tc.cu runs one output tile per CTA -- load, MMA and epilogue in
// sequence, which is right for the student's K = 2048 tiles but leaves a K = 256 tile latency-bound
// (measured 2.6 ms for 2.45M x 256 x 256 against 0.8 ms of compulsory HBM traffic).  Here a CTA per
// SM walks over the row tiles and the three phases of neighbouring tiles overlap:
//   warp 0      TMA producer: one lane issues cp.async.bulk.tensor (SWIZZLE_128B) for the bf16 hi/lo
//               planes of A (128 rows x 64 k) and W (BN x 64 k) into a shared-memory ring; the data
//               lands in the UMMA canonical K-major layout, mbarrier complete_tx signals the stage;
//   warp 1      MMA issuer: one lane issues tcgen05.mma kind::f16 (M = 128, N = BN, K = 16), three
//               products per k-step (hi*hi + hi*lo + lo*hi, see gemm_tc.cu for the numerics) into
//               one of TWO TMEM accumulators; tcgen05.commit frees the stage / publishes the tile;
//   warps 2-9   epilogue: tcgen05.ld of the finished accumulator while the next tile is being
//               multiplied into the other one; row scale, bias, eval-BN affine, ReLU in registers;
//               each warp stages its own 32 x 32 chunks through swizzled shared memory (no CTA-wide
//               barrier) so that global stores are whole 128-byte row segments; output as fp32,
//               bf16 hi/lo planes or q24.
// The weight planes are re-read from L2 for every tile (K = 256, N = 256 is 256 KB -- more than a
// CTA can hold next to the A ring); A is read from HBM exactly once.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "gemm.cuh"
#include "tc_common.cuh"

namespace glnn {
namespace tall {

using namespace tc;

constexpr int BM = 128;
// k-block of one pipeline stage: a template parameter.  64 (SWIZZLE_128B rows) gives 2 / 3 / 4 stages
// at N = 256 / 128 / 64; 32 (SWIZZLE_64B rows) halves the stage and doubles the ring depth.  Round 2
// measured both on the products forward (GLNN_TALL_BK): 1.17 / 1.27 / 0.60 ms (64) against 1.23 /
// 1.31 / 0.71 ms (32) for the three projections -- the deeper ring does not help, so the ring depth
// is not what holds the kernel at 41 % tensor-pipe activity; 64 stays the default.
constexpr int SLAB = 32;  // columns per epilogue slab
// Epilogue warps per CTA: a template parameter, 8 (default) or 16 (GLNN_TALL_EW=16: 4 warps per TMEM
// lane quarter, 64 columns each, single-buffered TMEM reads to stay within 113 registers).  Round 2
// measured both on the products forward after ncu's source page had spread the stall samples evenly
// over the ~5000 epilogue instructions: 1.18 / 1.26 ms (8 warps) against 1.30 / 1.29 ms (16 warps) for
// the two 256-wide projections -- like the ring depth, the epilogue's issue rate is not what holds the
// kernel at 4 TB/s of algorithmic traffic.
template <int BN, int BK, int EW>
struct Cfg {
  static constexpr int NTHREADS = 64 + EW * 32;
  static constexpr int STAGE_ROWS = EW == 8 ? 16 : 8;   // rows of a warp's staging buffer
  static constexpr int A_BYTES = BM * BK * 2;  // one bf16 plane of one stage
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int STAGES = (BN == 256 ? 2 : (BN == 128 ? 3 : 4)) * (64 / BK);
  static constexpr uint32_t SBO = 8 * BK * 2;            // bytes between 8-row groups of a plane
  static constexpr uint32_t LAYOUT = BK == 64 ? 2 : 4;   // SWIZZLE_128B / SWIZZLE_64B
  static constexpr int STAGING = EW * STAGE_ROWS * SLAB * 4;  // STAGE_ROWS x 32 fp32 columns per epilogue warp
  static constexpr int EPI_VEC = 3 * BN * 4;                 // bias, BN scale, BN shift (padded to BN)
  static constexpr int TOTAL = STAGES * STAGE + STAGING + EPI_VEC + 1024 /*alignment slack*/ +
                               256 /*barriers*/;
  static_assert(TOTAL <= 232448, "exceeds the 227 KB of shared memory a CTA can opt into");
  static constexpr int TMEM_COLS = 2 * BN;  // two accumulators; 128 / 256 / 512 are powers of two
};

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint8_t* dst, const CUtensorMap* map, int c0, int c1,
                                            uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1, {%2, %3}], [%4], %5;" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// 32 lanes x 32 consecutive fp32 columns of TMEM -> 32 registers per thread (row = lane).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}

template <int BN, int BK, int EW>
__global__ void __launch_bounds__(64 + EW * 32, 1)
gemm_tall_kernel(const GemmArgs g, const __grid_constant__ CUtensorMap map_ah,
                 const __grid_constant__ CUtensorMap map_al, const __grid_constant__ CUtensorMap map_bh,
                 const __grid_constant__ CUtensorMap map_bl) {
  using S = Cfg<BN, BK, EW>;
  constexpr int NTHREADS = S::NTHREADS;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // SWIZZLE_128B atoms
  float* epi_vec = reinterpret_cast<float*>(tiles + S::STAGES * S::STAGE + S::STAGING);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + S::STAGES * S::STAGE + S::STAGING + S::EPI_VEC);
  uint64_t* full = bars;                       // [STAGES]  TMA -> MMA
  uint64_t* empty = bars + S::STAGES;          // [STAGES]  MMA -> TMA
  uint64_t* acc_full = bars + 2 * S::STAGES;   // [2]       MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;          // [2]       epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ntiles = static_cast<int>((g.M + BM - 1) / BM);
  const int nkb = static_cast<int>((g.K + BK - 1) / BK);

  if (tid == 0) {
    for (int s = 0; s < S::STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], EW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const bool has_affine = g.col_scale != nullptr;
  for (int n = tid; n < BN; n += NTHREADS) {  // epilogue vectors, padded so that the math is branch-free
    const bool in = n < g.N;
    epi_vec[n] = (in && g.bias) ? __ldg(g.bias + n) : 0.f;
    epi_vec[BN + n] = (in && has_affine) ? __ldg(g.col_scale + n) : 1.f;
    epi_vec[2 * BN + n] = (in && has_affine) ? __ldg(g.col_shift + n) : 0.f;
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(static_cast<uint32_t>(S::TMEM_COLS))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_ah)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_al)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_bh)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_bl)) : "memory");
      uint64_t pol_stream, pol_keep;  // A is read once, W by every CTA for every tile
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
      asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
      uint32_t it = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int m0 = t * BM;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % S::STAGES;
          mbar_wait(&empty[s], ((it / S::STAGES) & 1) ^ 1);
          uint8_t* st = tiles + s * S::STAGE;
          mbar_expect_tx(&full[s], S::STAGE);
          tma_load_2d(st, &map_ah, kb * BK, m0, &full[s], pol_stream);
          tma_load_2d(st + S::A_BYTES, &map_al, kb * BK, m0, &full[s], pol_stream);
          tma_load_2d(st + 2 * S::A_BYTES, &map_bh, kb * BK, 0, &full[s], pol_keep);
          tma_load_2d(st + 2 * S::A_BYTES + S::B_BYTES, &map_bl, kb * BK, 0, &full[s], pol_keep);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer -------------------------------
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(BN >> 3) << 17) |
                             (static_cast<uint32_t>(BM >> 4) << 24);  // f32 accum, bf16 x bf16, K-major
      uint32_t it = 0, tcount = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++tcount) {
        const uint32_t a = tcount & 1;
        mbar_wait(&acc_empty[a], ((tcount >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + a * BN;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % S::STAGES;
          mbar_wait(&full[s], (it / S::STAGES) & 1);
          tc_fence_after();
          const uint32_t st = smem_u32(tiles + s * S::STAGE);
          const uint64_t a_hi = make_desc(st, 16, S::SBO, S::LAYOUT),
                         a_lo = make_desc(st + S::A_BYTES, 16, S::SBO, S::LAYOUT);
          const uint64_t b_hi = make_desc(st + 2 * S::A_BYTES, 16, S::SBO, S::LAYOUT),
                         b_lo = make_desc(st + 2 * S::A_BYTES + S::B_BYTES, 16, S::SBO, S::LAYOUT);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t dk = static_cast<uint64_t>(k * 2);  // 16 bf16 = 32 bytes, in 16-byte units
            umma_bf16(tmem_d, a_hi + dk, b_hi + dk, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_bf16(tmem_d, a_hi + dk, b_lo + dk, idesc, 1u);
            umma_bf16(tmem_d, a_lo + dk, b_hi + dk, idesc, 1u);
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&acc_full[a]);
      }
    }
  } else {
    // ------------------------------- epilogue -------------------------------
    // Each warp owns 32 rows (its TMEM lane quarter) x BN / (EW/4) columns of the tile and works
    // through them in 32-column chunks on its own: TMEM -> registers -> epilogue math (bias / BN
    // vectors are broadcast reads from shared memory) -> private swizzled staging (STAGE_ROWS rows x 32
    // columns per pass) -> coalesced 128-byte row segments.  No CTA-wide barrier: the warps drift apart
    // and hide each other's latencies.  EW = 8: the next chunk's tcgen05.ld is in flight while the
    // current chunk is stored (two register buffers); EW = 16: one buffer, twice the warps.
    const int ew = warp - 2;          // 0..EW-1
    const int q = warp & 3;           // TMEM lane quarter this warp may access
    const int part = ew >> 2;         // which slice of the tile's columns
    constexpr int PARTS = EW / 4;
    constexpr int WCOLS = BN / PARTS;           // columns per warp
    constexpr int CHUNKS = WCOLS / SLAB;        // 32-column chunks per warp and tile
    static_assert(CHUNKS >= 1, "tile too narrow for this many epilogue warps");
    constexpr int SR = S::STAGE_ROWS;           // 16 (two passes) or 8 (four passes)
    constexpr int NBUF = EW == 8 ? 2 : 1;
    const uint32_t sb = smem_u32(tiles) + S::STAGES * S::STAGE + ew * (SR * SLAB * 4);
    const uint32_t ev = smem_u32(tiles) + S::STAGES * S::STAGE + S::STAGING;
    uint32_t tcount = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++tcount) {
      const uint32_t a = tcount & 1;
      const int64_t m0 = static_cast<int64_t>(t) * BM + q * 32;  // first row of this warp
      mbar_wait(&acc_full[a], (tcount >> 1) & 1);
      tc_fence_after();
      const float rs = (g.row_scale && m0 + lane < g.M) ? __ldg(g.row_scale + m0 + lane) : 1.f;
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * BN + part * WCOLS;
      uint32_t r[NBUF][32];
      tmem_ld32(tbase, r[0]);
#pragma unroll
      for (int c = 0; c < CHUNKS; ++c) {
        const int col0 = part * WCOLS + c * SLAB;
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (NBUF == 2 && c + 1 < CHUNKS) tmem_ld32(tbase + (c + 1) * SLAB, r[(c + 1) & 1]);
        float o[32];
        if (col0 < g.N) {
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 vb = lds128(ev + (col0 + j4 * 4) * 4);
            const float4 vs = lds128(ev + (BN + col0 + j4 * 4) * 4);
            const float4 vt = lds128(ev + (2 * BN + col0 + j4 * 4) * 4);
            const float bb[4] = {vb.x, vb.y, vb.z, vb.w}, ss[4] = {vs.x, vs.y, vs.z, vs.w},
                        tt[4] = {vt.x, vt.y, vt.z, vt.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float x = __uint_as_float(r[c % NBUF][j4 * 4 + j]) * rs + bb[j];
              if (g.relu == 2) x = fmaxf(x, 0.f);
              if (has_affine) x = fmaf(x, ss[j], tt[j]);
              if (g.relu == 1) x = fmaxf(x, 0.f);
              o[j4 * 4 + j] = x;
            }
          }
        }
        // the accumulator values of this chunk are consumed: fetch the next chunk (single buffer) or,
        // after the last one, hand the accumulator back to the MMA warp
        if (c + 1 < CHUNKS) {
          if (NBUF == 1) tmem_ld32(tbase + (c + 1) * SLAB, r[0]);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[a]);
        }
        if (col0 >= g.N) continue;
#pragma unroll
        for (int pass = 0; pass < 32 / SR; ++pass) {
          __syncwarp();  // the previous pass's staging reads are done
          if (lane / SR == pass) {
            const int rl = lane % SR;
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4)  // 16-byte piece p of row r lives at piece p ^ (r & 7)
              sts128(sb + (rl * SLAB + ((j4 ^ (rl & 7)) * 4)) * 4,
                     make_float4(o[4 * j4], o[4 * j4 + 1], o[4 * j4 + 2], o[4 * j4 + 3]));
          }
          __syncwarp();
          // coalesced stores: 8 lanes cover one 128-byte row segment, 4 rows per instruction
#pragma unroll
          for (int rr = 0; rr < SR; rr += 4) {
            const int r_out = rr + (lane >> 3), p = lane & 7;
            const int64_t mr = m0 + pass * SR + r_out;
            const int64_t n = col0 + p * 4;
            if (mr < g.M && n < g.N) {
              const float4 v = lds128(sb + (r_out * SLAB + ((p ^ (r_out & 7)) * 4)) * 4);
              emit4(g, mr, n, v);
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(static_cast<uint32_t>(S::TMEM_COLS))
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// host side: tensor maps (driver entry point resolved at run time, no link against libcuda)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// bf16 matrix [rows, cols] with row stride ld elements; box = bk columns x box_rows rows, swizzle
// over the box row (128 bytes at bk = 64, 64 bytes at bk = 32), out-of-bounds elements read as zero
// (ragged K, ragged last row tile, N < BN).
static int make_map(CUtensorMap* map, const uint16_t* base, int64_t rows, int64_t cols, int64_t ld,
                    int box_rows, int bk) {
  EncodeTiledFn fn = encode_fn();
  GLNN_REQUIRE(fn != nullptr, GLNN_ERR_DEVICE, "gemm_tall: cuTensorMapEncodeTiled not available");
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(bk), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<uint16_t*>(base), dims,
                        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GLNN_REQUIRE(r == CUDA_SUCCESS, GLNN_ERR_ARG, "gemm_tall: cuTensorMapEncodeTiled failed (%d)",
               static_cast<int>(r));
  return 0;
}

template <int BN, int BK, int EW>
static int launch(const GemmArgs& g, cudaStream_t st) {
  using S = Cfg<BN, BK, EW>;
  static bool configured = false;
  auto kern = gemm_tall_kernel<BN, BK, EW>;
  if (!configured) {
    GLNN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    configured = true;
  }
  CUtensorMap mah, mal, mbh, mbl;
  int rc;
  if ((rc = make_map(&mah, g.Ah, g.M, g.K, g.lda, BM, BK))) return rc;
  if ((rc = make_map(&mal, g.Al, g.M, g.K, g.lda, BM, BK))) return rc;
  if ((rc = make_map(&mbh, g.Bh, g.N, g.K, g.ldb, BN, BK))) return rc;
  if ((rc = make_map(&mbl, g.Bl, g.N, g.K, g.ldb, BN, BK))) return rc;
  const int ntiles = static_cast<int>((g.M + BM - 1) / BM);
  const int grid = std::min(ntiles, sm_count());
  kern<<<grid, S::NTHREADS, S::TOTAL, st>>>(g, mah, mal, mbh, mbl);
  GLNN_LAUNCH_OK("gemm_tall_kernel");
  return 0;
}

}  // namespace tall

// Takes planes x planes^T problems (A [M,K] K-major, B = nn.Linear weight [N,K]) with N <= 256 and at
// least one row tile per SM.  *taken = false leaves the problem to gemm_tc_planes.
int gemm_tall_planes(const GemmArgs& g, cudaStream_t st, bool* taken) {
  *taken = false;
  static const bool disabled = getenv("GLNN_NO_TALL") != nullptr;
  if (disabled) return 0;
  if (g.transA || !g.transB || g.N > 256 || g.N < 8 || g.K < 1) return 0;
  if (g.M < static_cast<int64_t>(tall::BM) * sm_count()) return 0;
  if (g.M >= (1LL << 31) - tall::BM) return 0;  // TMA coordinates are int32
  int rc;
  static const bool bk64 = !(getenv("GLNN_TALL_BK") != nullptr && atoi(getenv("GLNN_TALL_BK")) == 32);
  static const bool ew8 = !(getenv("GLNN_TALL_EW") != nullptr && atoi(getenv("GLNN_TALL_EW")) == 16);
  if (bk64) {  // default
    if (g.N > 128) rc = ew8 ? tall::launch<256, 64, 8>(g, st) : tall::launch<256, 64, 16>(g, st);
    else if (g.N > 64) rc = ew8 ? tall::launch<128, 64, 8>(g, st) : tall::launch<128, 64, 16>(g, st);
    else rc = tall::launch<64, 64, 8>(g, st);
  } else {
    if (g.N > 128) rc = tall::launch<256, 32, 8>(g, st);
    else if (g.N > 64) rc = tall::launch<128, 32, 8>(g, st);
    else rc = tall::launch<64, 32, 8>(g, st);
  }
  if (rc != 0) return rc;
  *taken = true;
  return 0;
}

}  // namespace glnn
