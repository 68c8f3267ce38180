// TF32 tensor-core GEMM (tcgen05) for the fused denoiser -- placeholder until the tcgen05 kernel lands:
// reports PDR_ERR_UNSUPPORTED so that callers can never silently get a different arithmetic.
#include "common.cuh"
namespace pdr {
int launch_gemm_tf32(const PdrGemmArgs &a, cudaStream_t stream) {
  (void)a; (void)stream;
  set_error("gemm_fused: use_tf32=1 (tcgen05 path) is not built yet");
  return PDR_ERR_UNSUPPORTED;
}
}  // namespace pdr
