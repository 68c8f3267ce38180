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

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

}  // namespace pdr
