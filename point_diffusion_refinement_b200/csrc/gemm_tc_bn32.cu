// 32-column tiles of the tcgen05 GEMM (gemm_tc.cuh): every epilogue flavour / experiment, resident and streamed W.
#include "gemm_tc.cuh"

namespace pdr {
int gemm_tf32_bn32(const PdrGemmArgs &a, cudaStream_t stream) { return dispatch_wres<32>(a, stream); }
}  // namespace pdr
