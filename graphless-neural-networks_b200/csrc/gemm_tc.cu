// tcgen05 3xTF32 projection path -- placeholder until the tensor-core kernel lands; declines every
// shape so glnn_gemm_f32 falls through to the exact fp32 SIMT kernel.
#include "gemm.cuh"

namespace glnn {
int gemm_tc(const GemmArgs&, cudaStream_t, bool* taken) {
  *taken = false;
  return 0;
}
}  // namespace glnn
