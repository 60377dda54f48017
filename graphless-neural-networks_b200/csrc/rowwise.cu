// Small row-wise / column-wise kernels around the two hot loops: eval-BatchNorm folding (K5),
// log-softmax (K7) and the evaluate() reductions (K8 + accuracy).
#include <algorithm>

#include "common.cuh"

namespace glnn {

__global__ void bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ mean, const float* __restrict__ var,
                               float eps, float* __restrict__ scale, float* __restrict__ shift,
                               int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // same operation order as torch's eval batch_norm: (x - mean) / sqrt(var + eps) * gamma + beta
  const float s = gamma[i] / sqrtf(var[i] + eps);
  scale[i] = s;
  shift[i] = beta[i] - mean[i] * s;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// one warp per row; c is small (7..47 on the configs) but any c works
__global__ void __launch_bounds__(256) log_softmax_kernel(const float* __restrict__ X, int64_t ldx,
                                                          float* __restrict__ Y, int64_t ldy,
                                                          int64_t n, int c) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const float* x = X + row * ldx;
  float m = -INFINITY;
  for (int j = lane; j < c; j += 32) m = fmaxf(m, x[j]);
  m = warp_max(m);
  float s = 0.f;
  for (int j = lane; j < c; j += 32) s += expf(x[j] - m);
  s = warp_sum(s);
  const float lse = m + logf(s);
  float* y = Y + row * ldy;
  for (int j = lane; j < c; j += 32) y[j] = x[j] - lse;
}

__global__ void __launch_bounds__(256) nll_acc_kernel(const float* __restrict__ LP, int64_t ld, int c,
                                                      const int64_t* __restrict__ labels,
                                                      const int64_t* __restrict__ idx, int64_t n,
                                                      float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarp = blockDim.x >> 5;
  float loss = 0.f, hit = 0.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * nwarp + warp; i < n;
       i += static_cast<int64_t>(gridDim.x) * nwarp) {
    const int64_t r = idx ? idx[i] : i;
    const float* x = LP + r * ld;
    const int64_t y = labels[r];
    // argmax with torch's tie rule (first maximal index)
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int j = lane; j < c; j += 32) {
      const float v = x[j];
      if (v > bv) { bv = v; bi = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) {
      loss -= x[y];
      hit += (bi == static_cast<int>(y)) ? 1.f : 0.f;
    }
  }
  __shared__ float s_loss[8], s_hit[8];
  if (lane == 0) { s_loss[warp] = loss; s_hit[warp] = hit; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float l = 0.f, h = 0.f;
    for (int w = 0; w < nwarp; ++w) { l += s_loss[w]; h += s_hit[w]; }
    atomicAdd(out, l);
    atomicAdd(out + 1, h);
  }
}

}  // namespace glnn

extern "C" int glnn_bn_fold_f32(const float* gamma, const float* beta, const float* mean,
                                const float* var, float eps, float* scale, float* shift, int n,
                                glnn_stream_t stream) {
  using namespace glnn;
  GLNN_REQUIRE(n >= 0, GLNN_ERR_ARG, "bn_fold: negative n");
  if (n == 0) return 0;
  GLNN_REQUIRE(gamma && beta && mean && var && scale && shift, GLNN_ERR_ARG, "bn_fold: null pointer");
  bn_fold_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      gamma, beta, mean, var, eps, scale, shift, n);
  GLNN_LAUNCH_OK("bn_fold_kernel");
  return 0;
}

extern "C" int glnn_log_softmax_f32(const float* X, int64_t ldx, float* Y, int64_t ldy, int64_t n,
                                    int c, glnn_stream_t stream) {
  using namespace glnn;
  GLNN_REQUIRE(n >= 0 && c >= 0, GLNN_ERR_ARG, "log_softmax: negative size");
  if (n == 0 || c == 0) return 0;
  GLNN_REQUIRE(X && Y, GLNN_ERR_ARG, "log_softmax: null pointer");
  GLNN_REQUIRE(ldx >= c && ldy >= c, GLNN_ERR_SHAPE, "log_softmax: leading dimension < c");
  const int64_t blocks = (n + 7) / 8;
  GLNN_REQUIRE(blocks < (1LL << 31), GLNN_ERR_SHAPE, "log_softmax: too many rows");
  log_softmax_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      X, ldx, Y, ldy, n, c);
  GLNN_LAUNCH_OK("log_softmax_kernel");
  return 0;
}

extern "C" int glnn_nll_acc_f32(const float* LP, int64_t ld, int c, const int64_t* labels,
                                const int64_t* idx, int64_t n, float* out, glnn_stream_t stream) {
  using namespace glnn;
  GLNN_REQUIRE(n >= 0 && c > 0, GLNN_ERR_ARG, "nll_acc: bad size");
  if (n == 0) return 0;
  GLNN_REQUIRE(LP && labels && out, GLNN_ERR_ARG, "nll_acc: null pointer");
  const int64_t want = (n + 7) / 8;
  const unsigned blocks = static_cast<unsigned>(std::min<int64_t>(want, 8LL * sm_count()));
  nll_acc_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(LP, ld, c, labels, idx, n,
                                                                        out);
  GLNN_LAUNCH_OK("nll_acc_kernel");
  return 0;
}
