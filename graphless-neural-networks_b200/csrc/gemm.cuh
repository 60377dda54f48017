// Argument block shared by the SIMT and tcgen05 projection kernels.
#pragma once
#include "common.cuh"

namespace glnn {

struct GemmArgs {
  const float* A;
  int64_t lda;
  int transA;
  const float* B;
  int64_t ldb;
  int transB;
  float* C;
  int64_t ldc;
  int64_t M, N, K;
  const float* row_scale;
  const float* bias;
  const float* col_scale;
  const float* col_shift;
  int relu;
  int vecA, vecB, vecC;
  int k_per_split;  // multiple of BK
  // --- tensor-core "planes" mode: operands already split into bf16 hi / lo planes (same leading
  // dimensions lda / ldb, in elements); optional plane output of C; split-K over k-blocks ---
  const uint16_t *Ah, *Al, *Bh, *Bl;
  uint16_t *Ch, *Cl;
  int64_t ldcp;
  // optional 24-bit row-packed output ("q24": per row N x hi16 then N x mid8, 3N bytes): the format
  // a following neighbour GATHER reads when the layer output feeds an aggregate-first layer
  uint8_t* Cq;
  int64_t ldcq;  // bytes between q24 rows (>= 3 N, multiple of 16); N % 8 == 0
  int kb_per_split;
};

int gemm_simt(GemmArgs g, cudaStream_t st);                    // gemm_simt.cu
int gemm_tc(const GemmArgs& g, cudaStream_t st, bool force, bool* taken);  // gemm_tc.cu
int gemm_tc_planes(GemmArgs g, cudaStream_t st);                             // gemm_tc.cu
int gemm_tall_planes(const GemmArgs& g, cudaStream_t st, bool* taken);       // gemm_tall.cu

}  // namespace glnn
