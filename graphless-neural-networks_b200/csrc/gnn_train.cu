// Teacher TRAINING kernels: what `train` (train_and_eval.py:12-29) and
// `train_sage` (:32-56) need around the aggregation / projection kernels so that a training step is a
// plain kernel sequence with a hand-written backward -- no autograd:
//
//   glnn_nll_loss_grad_f32     log_softmax + NLLLoss(mean) over a row subset + d(lamb*loss)/dlogits
//   glnn_act_train_fwd_f32     [BatchNorm1d train] -> [ReLU] -> [Dropout]      (models.py:113-118,194-198)
//   glnn_act_train_bwd_f32     its backward (+ dgamma, dbeta, column sums = bias gradient)
//   glnn_spmm_csr_scatter_f32  transposed aggregation of a sampled block without building its transpose
//
// GraphConv applies ReLU inside the conv (fused into the forward epilogues), so for GCN the block is
// norm -> dropout and the backward ends with the mask of the block's (post-ReLU) INPUT; for SAGE the
// block is norm -> ReLU -> dropout.
#include <algorithm>

#include "common.cuh"

namespace glnn {

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// one warp per selected row
__global__ void __launch_bounds__(256) nll_loss_grad_kernel(
    const float* __restrict__ logits, int64_t ld, int c, const int64_t* __restrict__ labels,
    const int64_t* __restrict__ rows, const int64_t* __restrict__ label_rows, int64_t m, float lamb,
    float* __restrict__ dlogits, int64_t lddl, float* __restrict__ loss_out) {
  __shared__ float s_loss[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 8 + warp;
  float loss = 0.f;
  if (i < m) {
    const int64_t r = rows ? rows[i] : i;
    const int y = static_cast<int>(labels[label_rows ? label_rows[i] : r]);
    const float* x = logits + r * ld;
    float mx = -INFINITY;
    for (int j = lane; j < c; j += 32) mx = fmaxf(mx, x[j]);
    mx = wmax(mx);
    float se = 0.f;
    for (int j = lane; j < c; j += 32) se += expf(x[j] - mx);
    se = wsum(se);
    const float lse = mx + logf(se);
    const float sc = lamb / static_cast<float>(m);
    float* dl = dlogits + r * lddl;
    for (int j = lane; j < c; j += 32) {
      const float s = x[j] - lse;
      dl[j] = (expf(s) - (j == y ? 1.f : 0.f)) * sc;
      if (j == y) loss = -s;
    }
    loss = wsum(loss);
  }
  if (lane == 0) s_loss[warp] = loss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_loss[w];
    atomicAdd(loss_out, t / static_cast<float>(m));
  }
}

// ---- norm / relu / dropout block ---------------------------------------------------------------
struct ActArgs {
  int64_t n;
  int d;
  const float* X; int64_t ldx;
  float* Y; int64_t ldy;
  const float* gamma; const float* beta;
  float* rmean; float* rvar;
  float* smean; float* sinv;
  float eps, mom;
  int relu_post, relu_input;
  float p_drop;
  uint64_t seed;
  const uint8_t* mask;
  // backward
  const float* dY; int64_t lddy;
  float* dX; int64_t lddx;
  float* dgamma; float* dbeta; float* dbias;
  float* scratch;  // [2][d] column sums
};

__device__ __forceinline__ bool keep_of(const ActArgs& a, int64_t row, int col) {
  if (a.mask) return a.mask[row * a.d + col] != 0;
  uint64_t z = a.seed + 0x9E3779B97F4A7C15ull * static_cast<uint64_t>(row * a.d + col + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return static_cast<float>(static_cast<uint32_t>(z >> 40)) * (1.0f / 16777216.0f) >= a.p_drop;
}

// Column sums of two per-element quantities, block (32, 8): threadIdx.x = column of the tile,
// threadIdx.y strides over the rows of the block's row split; partials combined with atomics into
// scratch[0][col], scratch[1][col] (zeroed by the host).
//   MODE 0 (forward):  q1 = x - shift, q2 = (x - shift)^2 with shift = X[0, col]
//   MODE 1 (backward): g = dY (after dropout and post-ReLU masks); q1 = g, q2 = g * xhat
template <int MODE>
__global__ void __launch_bounds__(256) act_stats_kernel(const ActArgs a, int64_t rows_per_block) {
  __shared__ float s1[8][33], s2[8][33];
  const int col = blockIdx.x * 32 + threadIdx.x;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * rows_per_block;
  const int64_t r1 = min(a.n, r0 + rows_per_block);
  float q1 = 0.f, q2 = 0.f;
  if (col < a.d) {
    float shift = 0.f, mean = 0.f, inv = 0.f, ga = 1.f, be = 0.f;
    if (MODE == 0) {
      shift = a.X[col];
    } else {
      mean = a.smean[col]; inv = a.sinv[col]; ga = a.gamma[col]; be = a.beta[col];
    }
    const float inv_keep = 1.f / (1.f - a.p_drop);
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) {
      const float x = a.X[r * a.ldx + col];
      if (MODE == 0) {
        const float t = x - shift;
        q1 += t;
        q2 = fmaf(t, t, q2);
      } else {
        float g = a.dY[r * a.lddy + col];
        if (a.p_drop > 0.f) g = keep_of(a, r, col) ? g * inv_keep : 0.f;
        const float xhat = (x - mean) * inv;
        if (a.relu_post && fmaf(xhat, ga, be) <= 0.f) g = 0.f;
        q1 += g;
        q2 = fmaf(g, xhat, q2);
      }
    }
  }
  s1[threadIdx.y][threadIdx.x] = q1;
  s2[threadIdx.y][threadIdx.x] = q2;
  __syncthreads();
  if (threadIdx.y == 0 && col < a.d) {
    float t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { t1 += s1[j][threadIdx.x]; t2 += s2[j][threadIdx.x]; }
    atomicAdd(a.scratch + col, t1);
    atomicAdd(a.scratch + a.d + col, t2);
  }
}

// mean / invstd of the batch from the shifted sums; running statistics (momentum, unbiased variance)
__global__ void act_fwd_finalize_kernel(const ActArgs a) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= a.d) return;
  const float nf = static_cast<float>(a.n);
  const float m1 = a.scratch[col] / nf;
  const float var = fmaxf(a.scratch[a.d + col] / nf - m1 * m1, 0.f);
  const float mean = a.X[col] + m1;
  a.smean[col] = mean;
  a.sinv[col] = 1.f / sqrtf(var + a.eps);
  if (a.rmean) {
    a.rmean[col] = (1.f - a.mom) * a.rmean[col] + a.mom * mean;
    const float unb = var * (nf / fmaxf(nf - 1.f, 1.f));
    a.rvar[col] = (1.f - a.mom) * a.rvar[col] + a.mom * unb;
  }
}

__global__ void __launch_bounds__(256) act_fwd_apply_kernel(const ActArgs a) {
  const int col = blockIdx.x * 32 + threadIdx.x;
  if (col >= a.d) return;
  const bool norm = a.gamma != nullptr;
  float mean = 0.f, inv = 1.f, ga = 1.f, be = 0.f;
  if (norm) { mean = a.smean[col]; inv = a.sinv[col]; ga = a.gamma[col]; be = a.beta[col]; }
  const float inv_keep = 1.f / (1.f - a.p_drop);
  for (int64_t r = static_cast<int64_t>(blockIdx.y) * 8 + threadIdx.y; r < a.n;
       r += static_cast<int64_t>(gridDim.y) * 8) {
    float t = a.X[r * a.ldx + col];
    if (norm) t = fmaf((t - mean) * inv, ga, be);
    if (a.relu_post) t = fmaxf(t, 0.f);
    if (a.p_drop > 0.f) t = keep_of(a, r, col) ? t * inv_keep : 0.f;
    a.Y[r * a.ldy + col] = t;
  }
}

// dX and the column sums of dX (= gradient of the bias added in front of the block)
__global__ void __launch_bounds__(256) act_bwd_apply_kernel(const ActArgs a, int64_t rows_per_block) {
  __shared__ float sb[8][33];
  const int col = blockIdx.x * 32 + threadIdx.x;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * rows_per_block;
  const int64_t r1 = min(a.n, r0 + rows_per_block);
  const bool norm = a.gamma != nullptr;
  float csum = 0.f;
  if (col < a.d) {
    float mean = 0.f, inv = 1.f, ga = 1.f, be = 0.f, S1 = 0.f, S2 = 0.f;
    if (norm) {
      mean = a.smean[col]; inv = a.sinv[col]; ga = a.gamma[col]; be = a.beta[col];
      S1 = a.scratch[col]; S2 = a.scratch[a.d + col];
    }
    const float inv_keep = 1.f / (1.f - a.p_drop);
    const float nf = static_cast<float>(a.n);
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) {
      const float x = a.X[r * a.ldx + col];
      float g = a.dY[r * a.lddy + col];
      if (a.p_drop > 0.f) g = keep_of(a, r, col) ? g * inv_keep : 0.f;
      if (norm) {
        const float xhat = (x - mean) * inv;
        if (a.relu_post && fmaf(xhat, ga, be) <= 0.f) g = 0.f;
        g = (ga * inv / nf) * (nf * g - S1 - xhat * S2);
      } else if (a.relu_post && x <= 0.f) {
        g = 0.f;
      }
      if (a.relu_input && x <= 0.f) g = 0.f;
      a.dX[r * a.lddx + col] = g;
      csum += g;
    }
  }
  sb[threadIdx.y][threadIdx.x] = csum;
  __syncthreads();
  if (threadIdx.y == 0 && col < a.d) {
    if (a.dbias) {
      float t = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) t += sb[j][threadIdx.x];
      atomicAdd(a.dbias + col, t);
    }
    if (norm && blockIdx.y == 0) {
      if (a.dgamma) a.dgamma[col] = a.scratch[a.d + col];
      if (a.dbeta) a.dbeta[col] = a.scratch[col];
    }
  }
}

static int act_check(const glnn_act_desc* q, const char* who) {
  GLNN_REQUIRE(q != nullptr, GLNN_ERR_ARG, "%s: null descriptor", who);
  GLNN_REQUIRE(q->n >= 0 && q->d >= 0, GLNN_ERR_ARG, "%s: negative size", who);
  if (q->n == 0 || q->d == 0) return 0;
  GLNN_REQUIRE(q->X && q->ldx >= q->d, GLNN_ERR_ARG, "%s: X / ldx", who);
  GLNN_REQUIRE((q->gamma == nullptr) == (q->beta == nullptr), GLNN_ERR_ARG, "%s: gamma and beta come together", who);
  GLNN_REQUIRE(!q->gamma || (q->save_mean && q->save_invstd), GLNN_ERR_ARG,
               "%s: BatchNorm needs save_mean / save_invstd", who);
  GLNN_REQUIRE((q->running_mean == nullptr) == (q->running_var == nullptr), GLNN_ERR_ARG,
               "%s: running_mean and running_var come together", who);
  GLNN_REQUIRE(q->p_drop >= 0.f && q->p_drop < 1.f, GLNN_ERR_ARG, "%s: dropout in [0, 1)", who);
  GLNN_REQUIRE(!q->gamma || q->n >= 2, GLNN_ERR_SHAPE, "%s: BatchNorm needs more than 1 row", who);
  return 0;
}

static ActArgs act_args(const glnn_act_desc* q, float* scratch) {
  ActArgs a{};
  a.n = q->n; a.d = q->d; a.X = q->X; a.ldx = q->ldx; a.Y = q->Y; a.ldy = q->ldy;
  a.gamma = q->gamma; a.beta = q->beta; a.rmean = q->running_mean; a.rvar = q->running_var;
  a.smean = q->save_mean; a.sinv = q->save_invstd; a.eps = q->eps; a.mom = q->momentum;
  a.relu_post = q->relu_post; a.relu_input = q->relu_input; a.p_drop = q->p_drop; a.seed = q->seed;
  a.mask = q->keep_mask; a.scratch = scratch;
  return a;
}

static inline void split_rows(int64_t n, int col_tiles, int64_t* rows_per_block, unsigned* row_blocks) {
  const int64_t want = std::max<int64_t>(1, std::min<int64_t>(4LL * sm_count() / col_tiles, (n + 63) / 64));
  *rows_per_block = (n + want - 1) / want;
  *row_blocks = static_cast<unsigned>((n + *rows_per_block - 1) / *rows_per_block);
}

}  // namespace glnn

extern "C" int glnn_nll_loss_grad_f32(const float* logits, int64_t ld, int c, const int64_t* labels,
                                      const int64_t* rows, const int64_t* label_rows, int64_t m,
                                      float lamb, float* dlogits, int64_t lddl, float* loss_out,
                                      glnn_stream_t stream) {
  using namespace glnn;
  GLNN_REQUIRE(m >= 0 && c > 0, GLNN_ERR_ARG, "nll_loss_grad: bad size");
  if (m == 0) return 0;
  GLNN_REQUIRE(logits && labels && dlogits && loss_out && ld >= c && lddl >= c, GLNN_ERR_ARG,
               "nll_loss_grad: null pointer or leading dimension < c");
  const int64_t blocks = (m + 7) / 8;
  GLNN_REQUIRE(blocks < (1LL << 31), GLNN_ERR_SHAPE, "nll_loss_grad: too many rows");
  nll_loss_grad_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, ld, c, labels, rows, label_rows, m, lamb, dlogits, lddl, loss_out);
  GLNN_LAUNCH_OK("nll_loss_grad_kernel");
  return 0;
}

extern "C" int glnn_act_train_fwd_f32(const glnn_act_desc* q, float* scratch, glnn_stream_t stream) {
  using namespace glnn;
  int rc = act_check(q, "act_train_fwd");
  if (rc != 0) return rc;
  if (q->n == 0 || q->d == 0) return 0;
  GLNN_REQUIRE(q->Y && q->ldy >= q->d, GLNN_ERR_ARG, "act_train_fwd: Y / ldy");
  GLNN_REQUIRE(!q->gamma || scratch, GLNN_ERR_ARG, "act_train_fwd: BatchNorm needs 2*d floats of scratch");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ActArgs a = act_args(q, scratch);
  const int col_tiles = (q->d + 31) / 32;
  if (q->gamma) {
    GLNN_CUDA_OK(cudaMemsetAsync(scratch, 0, sizeof(float) * 2 * q->d, st));
    int64_t rpb; unsigned rb;
    split_rows(q->n, col_tiles, &rpb, &rb);
    act_stats_kernel<0><<<dim3(col_tiles, rb), dim3(32, 8), 0, st>>>(a, rpb);
    act_fwd_finalize_kernel<<<(q->d + 127) / 128, 128, 0, st>>>(a);
  }
  const unsigned rb2 = static_cast<unsigned>(std::max<int64_t>(
      1, std::min<int64_t>((q->n + 7) / 8, 8LL * sm_count() / col_tiles + 1)));
  act_fwd_apply_kernel<<<dim3(col_tiles, rb2), dim3(32, 8), 0, st>>>(a);
  GLNN_LAUNCH_OK("act_fwd_apply_kernel");
  return 0;
}

extern "C" int glnn_act_train_bwd_f32(const glnn_act_desc* q, const float* dY, int64_t lddy, float* dX,
                                      int64_t lddx, float* dgamma, float* dbeta, float* dbias,
                                      float* scratch, glnn_stream_t stream) {
  using namespace glnn;
  int rc = act_check(q, "act_train_bwd");
  if (rc != 0) return rc;
  if (q->n == 0 || q->d == 0) return 0;
  GLNN_REQUIRE(dY && dX && lddy >= q->d && lddx >= q->d, GLNN_ERR_ARG, "act_train_bwd: dY / dX");
  GLNN_REQUIRE(!q->gamma || scratch, GLNN_ERR_ARG, "act_train_bwd: BatchNorm needs 2*d floats of scratch");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ActArgs a = act_args(q, scratch);
  a.dY = dY; a.lddy = lddy; a.dX = dX; a.lddx = lddx; a.dgamma = dgamma; a.dbeta = dbeta; a.dbias = dbias;
  const int col_tiles = (q->d + 31) / 32;
  int64_t rpb; unsigned rb;
  split_rows(q->n, col_tiles, &rpb, &rb);
  if (q->gamma) {
    GLNN_CUDA_OK(cudaMemsetAsync(scratch, 0, sizeof(float) * 2 * q->d, st));
    act_stats_kernel<1><<<dim3(col_tiles, rb), dim3(32, 8), 0, st>>>(a, rpb);
  }
  if (dbias) GLNN_CUDA_OK(cudaMemsetAsync(dbias, 0, sizeof(float) * q->d, st));
  act_bwd_apply_kernel<<<dim3(col_tiles, rb), dim3(32, 8), 0, st>>>(a, rpb);
  GLNN_LAUNCH_OK("act_bwd_apply_kernel");
  return 0;
}

// ---- transposed aggregation by scatter ---------------------------------------------------------
namespace glnn {
// dX[u, :] += scale[v] * dY[v, :] for every edge u -> v of the CSR (rows = v); with self_add also
// dX[v, :] += scale[v] * dY[v, :].  One warp per destination row, lanes over the columns; fp32 atomics
// (the summation order over a source's out-edges is not fixed: results agree to rounding).
__global__ void __launch_bounds__(256) spmm_scatter_kernel(const void* __restrict__ indptr, int indptr64,
                                                           const int32_t* __restrict__ indices,
                                                           const float* __restrict__ dY, int64_t lddy,
                                                           const float* __restrict__ scale,
                                                           float* __restrict__ dX, int64_t lddx,
                                                           int64_t n_dst, int d, int self_add) {
  const int lane = threadIdx.x & 31;
  const int64_t v = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (v >= n_dst) return;
  const int64_t beg = indptr64 ? reinterpret_cast<const int64_t*>(indptr)[v]
                               : static_cast<int64_t>(reinterpret_cast<const int32_t*>(indptr)[v]);
  const int64_t end = indptr64 ? reinterpret_cast<const int64_t*>(indptr)[v + 1]
                               : static_cast<int64_t>(reinterpret_cast<const int32_t*>(indptr)[v + 1]);
  const float s = scale ? scale[v] : 1.f;
  for (int c0 = 0; c0 < d; c0 += 32) {
    const int c = c0 + lane;
    const float g = (c < d) ? dY[v * lddy + c] * s : 0.f;
    if (c < d && self_add) atomicAdd(dX + v * lddx + c, g);
    for (int64_t e = beg; e < end; ++e) {
      const int64_t u = indices[e];
      if (c < d) atomicAdd(dX + u * lddx + c, g);
    }
  }
}
}  // namespace glnn

extern "C" int glnn_spmm_csr_scatter_f32(const void* indptr, int indptr64, const int32_t* indices,
                                         const float* dY, int64_t lddy, const float* scale, float* dX,
                                         int64_t lddx, int64_t n_dst, int d, int self_add,
                                         glnn_stream_t stream) {
  using namespace glnn;
  GLNN_REQUIRE(n_dst >= 0 && d >= 0, GLNN_ERR_ARG, "spmm_scatter: negative size");
  if (n_dst == 0 || d == 0) return 0;
  GLNN_REQUIRE(indptr && dY && dX && lddy >= d && lddx >= d, GLNN_ERR_ARG, "spmm_scatter: null pointer / ld");
  const int64_t blocks = (n_dst + 7) / 8;
  GLNN_REQUIRE(blocks < (1LL << 31), GLNN_ERR_SHAPE, "spmm_scatter: too many rows");
  spmm_scatter_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      indptr, indptr64, indices, dY, lddy, scale, dX, lddx, n_dst, d, self_add);
  GLNN_LAUNCH_OK("spmm_scatter_kernel");
  return 0;
}
