// bf16 hi/lo "planes": the operand format of the tensor-core projection.  An fp32 matrix X is kept as
// two bf16 matrices hi = bf16(X), lo = bf16(X - hi) with a common leading dimension that is a
// multiple of 8 elements (16-byte rows), zero in the pad columns.  Same bytes as fp32, but a GEMM
// stage can then be filled by plain 16-byte async copies with no conversion in the GEMM itself.
// Producers that feed only GEMMs (aggregation epilogue, BN apply, loss gradient, ...) write planes
// directly; this file has the stand-alone splitter (weights, external inputs) and the C ABI.
#include <cuda_bf16.h>

#include "gemm.cuh"

namespace glnn {

__global__ void __launch_bounds__(256) split_planes_kernel(const float* __restrict__ X, int64_t ldx,
                                                           int64_t rows, int cols,
                                                           uint16_t* __restrict__ hi,
                                                           uint16_t* __restrict__ lo, int64_t ldp) {
  const int64_t chunks_per_row = ldp / 4;
  const int64_t total = rows * chunks_per_row;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / chunks_per_row;
    const int c = static_cast<int>(i % chunks_per_row) * 4;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = (c + j < cols) ? __ldg(X + r * ldx + c + j) : 0.f;
    __nv_bfloat162 h01 = __floats2bfloat162_rn(v[0], v[1]), h23 = __floats2bfloat162_rn(v[2], v[3]);
    const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
    __nv_bfloat162 l01 = __floats2bfloat162_rn(v[0] - f01.x, v[1] - f01.y);
    __nv_bfloat162 l23 = __floats2bfloat162_rn(v[2] - f23.x, v[3] - f23.y);
    uint2 uh, ul;
    uh.x = *reinterpret_cast<uint32_t*>(&h01); uh.y = *reinterpret_cast<uint32_t*>(&h23);
    ul.x = *reinterpret_cast<uint32_t*>(&l01); ul.y = *reinterpret_cast<uint32_t*>(&l23);
    *reinterpret_cast<uint2*>(hi + r * ldp + c) = uh;
    *reinterpret_cast<uint2*>(lo + r * ldp + c) = ul;
  }
}

int split_planes(const float* X, int64_t ldx, int64_t rows, int cols, uint16_t* hi, uint16_t* lo,
                 int64_t ldp, cudaStream_t st) {
  if (rows == 0 || ldp == 0) return 0;
  const int64_t total = rows * (ldp / 4);
  const unsigned blocks = static_cast<unsigned>(std::min<int64_t>((total + 255) / 256, 16LL * sm_count()));
  split_planes_kernel<<<blocks, 256, 0, st>>>(X, ldx, rows, cols, hi, lo, ldp);
  GLNN_LAUNCH_OK("split_planes_kernel");
  return 0;
}

// Projection whose result is written only in the 24-bit row-packed format read by a following gather.
int gemm_planes_q24(const uint16_t* A_hi, const uint16_t* A_lo, int64_t lda, const uint16_t* B_hi,
                    const uint16_t* B_lo, int64_t ldb, int transB, uint8_t* Cq, int64_t ldq, int64_t M,
                    int64_t N, int64_t K, const float* row_scale, const float* bias, const float* col_scale,
                    const float* col_shift, int relu, cudaStream_t st) {
  GemmArgs g;
  g.A = nullptr; g.B = nullptr;
  g.lda = lda; g.transA = 0; g.ldb = ldb; g.transB = transB ? 1 : 0;
  g.C = nullptr; g.ldc = 0; g.M = M; g.N = N; g.K = K;
  g.row_scale = row_scale; g.bias = bias; g.col_scale = col_scale; g.col_shift = col_shift;
  g.relu = relu;
  g.vecA = g.vecB = g.vecC = 0;
  g.k_per_split = 0;
  g.Ah = A_hi; g.Al = A_lo; g.Bh = B_hi; g.Bl = B_lo; g.Ch = nullptr; g.Cl = nullptr; g.ldcp = 0;
  g.Cq = Cq;
  g.ldcq = ldq;
  g.kb_per_split = 0;
  return gemm_tc_planes(g, st);
}

}  // namespace glnn

extern "C" int glnn_split_planes_f32(const float* X, int64_t ldx, int64_t rows, int cols, uint16_t* hi,
                                     uint16_t* lo, int64_t ldp, glnn_stream_t stream) {
  using namespace glnn;
  GLNN_REQUIRE(rows >= 0 && cols >= 0, GLNN_ERR_ARG, "split_planes: negative size");
  if (rows == 0) return 0;
  GLNN_REQUIRE(X && hi && lo, GLNN_ERR_ARG, "split_planes: null pointer");
  GLNN_REQUIRE(ldx >= cols && ldp >= cols && ldp % 8 == 0, GLNN_ERR_SHAPE,
               "split_planes: need ldx >= cols and ldp >= cols with ldp %% 8 == 0");
  GLNN_REQUIRE(aligned16(hi) && aligned16(lo), GLNN_ERR_ALIGN, "split_planes: planes must be 16-byte aligned");
  return split_planes(X, ldx, rows, cols, hi, lo, ldp, static_cast<cudaStream_t>(stream));
}

extern "C" int glnn_gemm_bf16x3_planes(const uint16_t* A_hi, const uint16_t* A_lo, int64_t lda,
                                       int transA, const uint16_t* B_hi, const uint16_t* B_lo,
                                       int64_t ldb, int transB, float* C, int64_t ldc, uint16_t* C_hi,
                                       uint16_t* C_lo, int64_t ldcp, int64_t M, int64_t N, int64_t K,
                                       const float* row_scale, const float* bias,
                                       const float* col_scale, const float* col_shift, int relu,
                                       glnn_stream_t stream) {
  using namespace glnn;
  GLNN_REQUIRE(M >= 0 && N >= 0 && K >= 0, GLNN_ERR_ARG, "gemm_planes: negative size");
  if (M == 0 || N == 0) return 0;
  GLNN_REQUIRE(lda >= (transA ? M : K) && ldb >= (transB ? K : N) && (!C || ldc >= N), GLNN_ERR_SHAPE,
               "gemm_planes: leading dimension too small");
  GLNN_REQUIRE((col_scale == nullptr) == (col_shift == nullptr), GLNN_ERR_ARG,
               "gemm_planes: col_scale and col_shift must be given together");
  GLNN_REQUIRE(relu >= 0 && relu <= 2, GLNN_ERR_ARG, "gemm_planes: relu must be 0, 1 or 2");
  GemmArgs g;
  g.A = nullptr; g.B = nullptr;
  g.lda = lda; g.transA = transA ? 1 : 0; g.ldb = ldb; g.transB = transB ? 1 : 0;
  g.C = C; g.ldc = ldc; g.M = M; g.N = N; g.K = K;
  g.row_scale = row_scale; g.bias = bias; g.col_scale = col_scale; g.col_shift = col_shift;
  g.relu = relu;
  g.vecA = g.vecB = g.vecC = 0;
  g.k_per_split = 0;
  g.Ah = A_hi; g.Al = A_lo; g.Bh = B_hi; g.Bl = B_lo; g.Ch = C_hi; g.Cl = C_lo; g.ldcp = ldcp;
  g.Cq = nullptr;
  g.ldcq = 0;
  g.kb_per_split = 0;
  return gemm_tc_planes(g, static_cast<cudaStream_t>(stream));
}

extern "C" int glnn_gemm_bf16x3_planes_q24(const uint16_t* A_hi, const uint16_t* A_lo, int64_t lda,
                                           const uint16_t* B_hi, const uint16_t* B_lo, int64_t ldb,
                                           int transB, uint8_t* C_q24, int64_t ldq, int64_t M, int64_t N,
                                           int64_t K,
                                           const float* row_scale, const float* bias,
                                           const float* col_scale, const float* col_shift, int relu,
                                           glnn_stream_t stream) {
  using namespace glnn;
  GLNN_REQUIRE(M >= 0 && N >= 0 && K >= 0, GLNN_ERR_ARG, "gemm_planes_q24: negative size");
  if (M == 0 || N == 0) return 0;
  GLNN_REQUIRE(C_q24 != nullptr, GLNN_ERR_ARG, "gemm_planes_q24: null output");
  GLNN_REQUIRE(lda >= K && ldb >= (transB ? K : N), GLNN_ERR_SHAPE, "gemm_planes_q24: leading dimension");
  GLNN_REQUIRE((col_scale == nullptr) == (col_shift == nullptr), GLNN_ERR_ARG,
               "gemm_planes_q24: col_scale and col_shift must be given together");
  return gemm_planes_q24(A_hi, A_lo, lda, B_hi, B_lo, ldb, transB, C_q24, ldq, M, N, K, row_scale, bias,
                         col_scale, col_shift, relu, static_cast<cudaStream_t>(stream));
}
