// Shared plumbing for libglnn_b200.so: thread-local error text, argument checks, launch checks.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "glnn_b200.h"

namespace glnn {

void set_error(const char* fmt, ...);  // common.cu

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

#define GLNN_REQUIRE(cond, code, ...)  \
  do {                                 \
    if (!(cond)) {                     \
      ::glnn::set_error(__VA_ARGS__);  \
      return (code);                   \
    }                                  \
  } while (0)

#define GLNN_CUDA_OK(expr)                                                                  \
  do {                                                                                      \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess) {                                                               \
      ::glnn::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__,  \
                        __LINE__);                                                          \
      return static_cast<int>(e__);                                                         \
    }                                                                                       \
  } while (0)

#define GLNN_LAUNCH_OK(name)                                                             \
  do {                                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess) {                                                            \
      ::glnn::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));       \
      return static_cast<int>(e__);                                                      \
    }                                                                                    \
  } while (0)

inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

__device__ __forceinline__ float4 ldg4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}

}  // namespace glnn
