// Shared helpers for libpdr_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pdr_b200.h"

namespace pdr {

void set_error(const char *fmt, ...);
int check_launch(const char *what);

#define PDR_REQUIRE(cond, ...)             \
  do {                                     \
    if (!(cond)) {                         \
      pdr::set_error(__VA_ARGS__);         \
      return PDR_ERR_INVALID_ARGUMENT;     \
    }                                      \
  } while (0)

constexpr int kNumSMs = 148;  // B200

// Squared distance with the rounding sequence of the reference kernels as compiled by nvcc 12.9
// (SASS of sampling_gpu.cu / ball_query_gpu.cu / interpolate_gpu.cu / emd_kernel.cu / chamfer3D.cu):
//   t = RN(dy*dy); t = fma(dx,dx,t); t = fma(dz,dz,t).
// Explicit intrinsics so the result never depends on -fmad or on how ptxas contracts.
__device__ __forceinline__ float dist2_ref(float dx, float dy, float dz) {
  return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}
// pytorch3d-style accumulation over dimensions 0,1,2: fma(dz,dz, fma(dy,dy, dx*dx)).
__device__ __forceinline__ float dist2_xyz(float dx, float dy, float dz) {
  return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// Programmatic dependent launch (PDL).  The compiled step is ~270 small-to-medium kernels back to back on one stream:
// without PDL a kernel's grid is launched only after the previous grid has drained, so every launch pays the launch latency
// and the ramp of its prologue (barrier init, TMEM allocation, resident weights) on an idle GPU.  A kernel launched with
// launch_pdl() may start as soon as the kernel in front of it has called pdl_launch_dependents() (or exited) and resources
// free up; it must call pdl_wait() before it reads or writes anything the earlier kernels touch -- the wait returns once
// every earlier grid has completed and its memory operations are visible.  Kernels launched the classic way are fully
// ordered as before, so the two kinds mix freely on a stream (and in a captured graph: programmatic edges).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

int pdl_mode();         // PDR_PDL: 0 off, 1 the GEMM triggers its dependents at kernel start, 2 after its last MMA
bool pdl_enabled();     // PDR_PDL=0 launches everything the classic way (api.cu)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

}  // namespace pdr
