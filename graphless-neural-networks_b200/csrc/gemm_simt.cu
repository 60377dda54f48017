// fp32 SIMT GEMM with fused epilogue (K4 of SURVEY.md section 2.3): the exact-fp32 anchor for the
// dense projections, used for shapes the tcgen05 3xTF32 path does not take (odd K, tiny N, strided
// operands).  128 x BN x 16 tiles, 256 threads, 8 x (BN/16) register micro-tiles, k-major shared
// tiles so the inner product reads are conflict-free float4 broadcasts, register-staged double
// buffering (one __syncthreads per k-tile).  Split-K (gridDim.z) is used when a skinny output
// (weight gradients: M,N = layer dims, K = batch) would otherwise leave most SMs idle.
#include <algorithm>

#include "common.cuh"
#include "gemm.cuh"

namespace glnn {


constexpr int BM = 128, BK = 16, PAD = 4;

// 4 consecutive elements along the contiguous dimension starting at `i` (limit `lim`), zero filled.
__device__ __forceinline__ float4 load4(const float* base, int64_t i, int64_t lim, bool vec) {
  if (vec && i + 3 < lim) return ldg4(base + i);
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < lim) r.x = __ldg(base + i);
  if (i + 1 < lim) r.y = __ldg(base + i + 1);
  if (i + 2 < lim) r.z = __ldg(base + i + 2);
  if (i + 3 < lim) r.w = __ldg(base + i + 3);
  return r;
}

// Loads this thread's share of a [ROWS x BK] operand tile (ROWS = BM or BN) into registers.
// KC = the k index is the contiguous one in memory (A with transA=0, B with transB=1).
template <int ROWS, bool KC>
struct TileLoader {
  static constexpr int NV = ROWS * BK / 4 / 256;  // float4 per thread (2 for 128 rows)
  static_assert(NV >= 1, "tile too small");
  float4 r[NV];
  __device__ __forceinline__ void load(const float* P, int64_t ld, int64_t row0, int64_t rows,
                                       int64_t k0, int64_t kend, bool vec, int tid) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int f = tid + i * 256;
      if constexpr (KC) {
        const int rr = f / (BK / 4), k4 = f % (BK / 4);
        const int64_t row = row0 + rr;
        r[i] = (row < rows) ? load4(P + row * ld, k0 + k4 * 4, kend, vec)
                            : make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        const int kk = f / (ROWS / 4), r4 = f % (ROWS / 4);
        const int64_t k = k0 + kk;
        r[i] = (k < kend) ? load4(P + k * ld, row0 + r4 * 4, rows, vec)
                          : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  __device__ __forceinline__ void store(float (*S)[ROWS + PAD], int tid) const {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int f = tid + i * 256;
      if constexpr (KC) {
        const int rr = f / (BK / 4), k4 = f % (BK / 4);
        S[k4 * 4 + 0][rr] = r[i].x;
        S[k4 * 4 + 1][rr] = r[i].y;
        S[k4 * 4 + 2][rr] = r[i].z;
        S[k4 * 4 + 3][rr] = r[i].w;
      } else {
        const int kk = f / (ROWS / 4), r4 = f % (ROWS / 4);
        *reinterpret_cast<float4*>(&S[kk][r4 * 4]) = r[i];
      }
    }
  }
};

template <int BN, bool AKC, bool BKC, bool SPLIT>
__global__ void __launch_bounds__(256) gemm_f32_kernel(const GemmArgs g) {
  constexpr int TN = BN / 16;               // columns per thread
  constexpr int NV = TN >= 4 ? 4 : TN;      // contiguous columns per group
  constexpr int NGRP = TN / NV;             // column groups per thread (2 for BN=128)
  constexpr int BROWS = BN < 64 ? 64 : BN;  // loader granularity (>= 1 float4 per thread)
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BROWS + PAD];

  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int64_t m0 = static_cast<int64_t>(blockIdx.x) * BM;
  const int64_t n0 = static_cast<int64_t>(blockIdx.y) * BN;
  const int64_t kbeg = SPLIT ? static_cast<int64_t>(blockIdx.z) * g.k_per_split : 0;
  const int64_t kend = SPLIT ? min(g.K, kbeg + g.k_per_split) : g.K;

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  TileLoader<BM, AKC> la;
  TileLoader<BROWS, BKC> lb;
  const int ntiles = static_cast<int>((kend - kbeg + BK - 1) / BK);
  if (ntiles > 0) {
    la.load(g.A, g.lda, m0, g.M, kbeg, kend, g.vecA, tid);
    lb.load(g.B, g.ldb, n0, g.N, kbeg, kend, g.vecB, tid);
    la.store(As[0], tid);
    lb.store(Bs[0], tid);
  }
  __syncthreads();
  for (int t = 0; t < ntiles; ++t) {
    const int cur = t & 1;
    if (t + 1 < ntiles) {
      la.load(g.A, g.lda, m0, g.M, kbeg + (t + 1) * BK, kend, g.vecA, tid);
      lb.load(g.B, g.ldb, n0, g.N, kbeg + (t + 1) * BK, kend, g.vecB, tid);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[8], b[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
      a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
      for (int gq = 0; gq < NGRP; ++gq)
#pragma unroll
        for (int j = 0; j < NV; ++j) b[gq * NV + j] = Bs[cur][k][gq * (BN / NGRP) + tx * NV + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (t + 1 < ntiles) {
      la.store(As[cur ^ 1], tid);
      lb.store(Bs[cur ^ 1], tid);
    }
    __syncthreads();
  }

  const bool lead = !SPLIT || blockIdx.z == 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= g.M) continue;
    const float rs = g.row_scale ? __ldg(g.row_scale + m) : 1.f;
#pragma unroll
    for (int gq = 0; gq < NGRP; ++gq) {
      const int64_t nb = n0 + gq * (BN / NGRP) + tx * NV;
      float o[NV];
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int64_t n = nb + j;
        float x = acc[i][gq * NV + j] * rs;
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
      float* cp = g.C + m * g.ldc + nb;
      if constexpr (SPLIT) {
#pragma unroll
        for (int j = 0; j < NV; ++j)
          if (nb + j < g.N) atomicAdd(cp + j, o[j]);
      } else {
        bool done = false;
        if constexpr (NV == 4) {
          if (g.vecC && nb + 3 < g.N) {
            *reinterpret_cast<float4*>(cp) = make_float4(o[0], o[1], o[2], o[3]);
            done = true;
          }
        }
        if (!done) {
#pragma unroll
          for (int j = 0; j < NV; ++j)
            if (nb + j < g.N) cp[j] = o[j];
        }
      }
    }
  }
}

template <int BN, bool SPLIT>
static int launch_bn(const GemmArgs& g, int splits, cudaStream_t st) {
  dim3 grid(static_cast<unsigned>((g.M + BM - 1) / BM), static_cast<unsigned>((g.N + BN - 1) / BN),
            static_cast<unsigned>(splits));
  const bool akc = !g.transA, bkc = g.transB;
  if (akc && bkc) gemm_f32_kernel<BN, true, true, SPLIT><<<grid, 256, 0, st>>>(g);
  else if (akc && !bkc) gemm_f32_kernel<BN, true, false, SPLIT><<<grid, 256, 0, st>>>(g);
  else if (!akc && bkc) gemm_f32_kernel<BN, false, true, SPLIT><<<grid, 256, 0, st>>>(g);
  else gemm_f32_kernel<BN, false, false, SPLIT><<<grid, 256, 0, st>>>(g);
  GLNN_LAUNCH_OK("gemm_f32_kernel");
  return 0;
}

int gemm_simt(GemmArgs g, cudaStream_t st) {
  g.vecA = (g.lda % 4 == 0) && aligned16(g.A);
  g.vecB = (g.ldb % 4 == 0) && aligned16(g.B);
  g.vecC = (g.ldc % 4 == 0) && aligned16(g.C);
  const int bn = g.N <= 32 ? 32 : (g.N <= 64 ? 64 : 128);
  const int64_t tiles = ((g.N + bn - 1) / bn) * ((g.M + BM - 1) / BM);
  // split-K only for linear epilogues (weight gradients) with few tiles and a long K
  int splits = 1;
  const bool linear = !g.relu && !g.col_scale;
  if (linear && tiles * 2 <= sm_count() && g.K >= 1024) {
    splits = static_cast<int>(std::min<int64_t>((2 * sm_count() + tiles - 1) / tiles, g.K / 256));
    if (splits < 1) splits = 1;
  }
  if (splits > 1) {
    g.k_per_split = static_cast<int>(((g.K + splits - 1) / splits + BK - 1) / BK * BK);
    splits = static_cast<int>((g.K + g.k_per_split - 1) / g.k_per_split);
    if (g.ldc == g.N) {
      GLNN_CUDA_OK(cudaMemsetAsync(g.C, 0, sizeof(float) * g.M * g.N, st));
    } else {
      GLNN_CUDA_OK(cudaMemset2DAsync(g.C, sizeof(float) * g.ldc, 0, sizeof(float) * g.N, g.M, st));
    }
    if (bn == 32) return launch_bn<32, true>(g, splits, st);
    if (bn == 64) return launch_bn<64, true>(g, splits, st);
    return launch_bn<128, true>(g, splits, st);
  }
  g.k_per_split = 0;
  if (bn == 32) return launch_bn<32, false>(g, 1, st);
  if (bn == 64) return launch_bn<64, false>(g, 1, st);
  return launch_bn<128, false>(g, 1, st);
}

}  // namespace glnn

extern "C" int glnn_gemm_f32(const float* A, int64_t lda, int transA, const float* B, int64_t ldb,
                             int transB, float* C, int64_t ldc, int64_t M, int64_t N, int64_t K,
                             const float* row_scale, const float* bias, const float* col_scale,
                             const float* col_shift, int relu, int impl, glnn_stream_t stream) {
  using namespace glnn;
  GLNN_REQUIRE(M >= 0 && N >= 0 && K >= 0, GLNN_ERR_ARG, "gemm: negative size");
  if (M == 0 || N == 0) return 0;
  GLNN_REQUIRE(A && B && C, GLNN_ERR_ARG, "gemm: null operand");
  GLNN_REQUIRE(lda >= (transA ? M : K) && ldb >= (transB ? K : N) && ldc >= N, GLNN_ERR_SHAPE,
               "gemm: leading dimension too small (lda=%lld ldb=%lld ldc=%lld M=%lld N=%lld K=%lld)",
               (long long)lda, (long long)ldb, (long long)ldc, (long long)M, (long long)N,
               (long long)K);
  GLNN_REQUIRE((col_scale == nullptr) == (col_shift == nullptr), GLNN_ERR_ARG,
               "gemm: col_scale and col_shift must be given together");
  GLNN_REQUIRE(impl >= 0 && impl <= 2, GLNN_ERR_ARG, "gemm: impl must be 0, 1 or 2");
  GLNN_REQUIRE(relu >= 0 && relu <= 2, GLNN_ERR_ARG, "gemm: relu must be 0, 1 or 2");
  GemmArgs g;
  g.A = A; g.lda = lda; g.transA = transA ? 1 : 0;
  g.B = B; g.ldb = ldb; g.transB = transB ? 1 : 0;
  g.C = C; g.ldc = ldc; g.M = M; g.N = N; g.K = K;
  g.row_scale = row_scale; g.bias = bias; g.col_scale = col_scale; g.col_shift = col_shift;
  g.relu = relu;
  g.vecA = g.vecB = g.vecC = 0;
  g.k_per_split = 0;
  g.Ah = g.Al = g.Bh = g.Bl = nullptr;
  g.Ch = g.Cl = nullptr;
  g.ldcp = 0;
  g.Cq = nullptr;
  g.kb_per_split = 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (impl != 1) {
    bool taken = false;
    int rc = gemm_tc(g, st, impl == 2, &taken);
    if (rc != 0) return rc;
    if (taken) return 0;
    GLNN_REQUIRE(impl != 2, GLNN_ERR_SHAPE, "gemm: shape not eligible for the tcgen05 path");
  }
  return gemm_simt(g, st);
}
