// HBM access-mix probe: measures the ceilings the store-dominated GEMMs of the step are held against.
// torch's fill_/zero_ write a constant (which the memory system may compress), so the write-only ceiling is taken
// here with incompressible data, once through ordinary vector stores and once through TMA bulk stores from shared
// memory -- the two store paths the GEMM epilogues use.  Diagnostic only; nothing on the hot path calls it.
#include "common.cuh"

namespace pdr {
namespace {

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
// finite, incompressible floats in [1, 2)
__device__ __forceinline__ float hash_float(uint32_t i) { return __uint_as_float(0x3f800000u | (mix32(i) >> 9)); }

// mode 0: dst[i] = hash(i), 16-byte stores, fully coalesced.  mode 1: the same plus src read at ratio 1 read : 3 writes.
__global__ void __launch_bounds__(256) probe_stg_kernel(float4 *dst, const float4 *src, size_t n4, int read_every) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  float acc = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const uint32_t h = (uint32_t)i * 4u;
    float4 v = make_float4(hash_float(h), hash_float(h + 1), hash_float(h + 2), hash_float(h + 3));
    if (read_every > 0 && (i / 32) % (size_t)read_every == 0) {      // one 512-byte segment in read_every is also read
      const float4 r = __ldg(src + i);
      acc += r.x + r.y + r.z + r.w;
    }
    dst[i] = v;
  }
  if (acc == 123.456f) dst[0].x = acc;                                // keep the loads
}

// mode 2: every warp fills a 4 KiB shared-memory tile with hash values and pushes it with one bulk store
__global__ void __launch_bounds__(128) probe_bulk_kernel(float *dst, size_t n_tiles) {
  __shared__ __align__(128) float tile[4][2][1024];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t wstride = (size_t)gridDim.x * 4;
  uint32_t buf = 0;
  for (size_t t = (size_t)blockIdx.x * 4 + warp; t < n_tiles; t += wstride) {
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncwarp();
    float *s = tile[warp][buf];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t h = (uint32_t)(t * 1024 + (size_t)(j * 32 + lane) * 4);
      *reinterpret_cast<float4 *>(s + (j * 32 + lane) * 4) =
          make_float4(hash_float(h), hash_float(h + 1), hash_float(h + 2), hash_float(h + 3));
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 4096;" ::"l"(dst + t * 1024),
                   "r"((uint32_t)__cvta_generic_to_shared(s))
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    buf ^= 1u;
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

}  // namespace
}  // namespace pdr

extern "C" int pdr_probe_hbm(int mode, float *dst, const float *src, size_t n_floats, void *stream) {
  using namespace pdr;
  PDR_REQUIRE(dst && n_floats >= 1024 && n_floats % 1024 == 0 && ((uintptr_t)dst % 16) == 0,
              "probe_hbm: dst must be 16-byte aligned, n a multiple of 1024");
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == 0 || mode == 1) {
    PDR_REQUIRE(mode == 0 || (src && ((uintptr_t)src % 16) == 0), "probe_hbm: mode 1 needs src");
    probe_stg_kernel<<<kNumSMs * 8, 256, 0, st>>>(reinterpret_cast<float4 *>(dst), reinterpret_cast<const float4 *>(src),
                                                 n_floats / 4, mode == 1 ? 3 : 0);
  } else if (mode == 2) {
    probe_bulk_kernel<<<kNumSMs * 6, 128, 0, st>>>(dst, n_floats / 1024);
  } else {
    set_error("probe_hbm: mode %d", mode);
    return PDR_ERR_INVALID_ARGUMENT;
  }
  return check_launch("probe_hbm");
}
