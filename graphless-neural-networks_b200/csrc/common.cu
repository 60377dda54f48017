#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace glnn {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace glnn

extern "C" int glnn_version(void) { return GLNN_ABI_VERSION; }

extern "C" const char* glnn_last_error(void) { return glnn::g_err; }

extern "C" int glnn_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  GLNN_CUDA_OK(cudaGetDevice(&dev));
  int sms = 0, maj = 0, min = 0;
  GLNN_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  GLNN_CUDA_OK(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
  GLNN_CUDA_OK(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count) *sm_count = sms;
  if (cc_major) *cc_major = maj;
  if (cc_minor) *cc_minor = min;
  GLNN_REQUIRE(maj == 10, GLNN_ERR_DEVICE,
               "libglnn_b200 is built for sm_100a only; current device is sm_%d%d", maj, min);
  return 0;
}
