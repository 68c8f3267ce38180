// Column-tile width of the tcgen05 GEMM (kernel and launcher: gemm_tc.cuh; one translation unit per width).
#include "common.cuh"

namespace pdr {

int gemm_tf32_bn32(const PdrGemmArgs &a, cudaStream_t stream);
int gemm_tf32_bn64(const PdrGemmArgs &a, cudaStream_t stream);
int gemm_tf32_bn128(const PdrGemmArgs &a, cudaStream_t stream);
int gemm_tf32_bn256(const PdrGemmArgs &a, cudaStream_t stream);
constexpr int kPlanDoesNotFit = 12345;   // same value as in gemm_tc.cuh: the planner could not fit this width

int launch_gemm_tf32(const PdrGemmArgs &a, cudaStream_t stream) {
  if (a.N > 128) {
    const int rc = gemm_tf32_bn256(a, stream);
    if (rc != kPlanDoesNotFit) return rc;
  }
  if (a.N > 64) return gemm_tf32_bn128(a, stream);
  if (a.N > 32) return gemm_tf32_bn64(a, stream);
  return gemm_tf32_bn32(a, stream);
}

}  // namespace pdr
