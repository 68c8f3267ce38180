// Elementwise part of the DDPM / FastDPM reverse step with device-side noise, for sm_100a.
//
// Reference: pointnet2/util.py:242-249 (x = (x - c*eps)/sqrt(alpha); x += sigma*z) and
// pointnet2/util_fastdpmv2.py:364-373 (x *= a; x += c*eps + sigma*z).  Both are the affine update
//     x <- x*scale_x + eps*scale_eps + sigma*z
// The reference draws z on the CPU and copies it to the GPU every step (util.py:118-123); here z comes
// from Philox4x32-10 evaluated in the kernel (counter = element/4 + offset, key = seed), so a
// 1000-step chain issues no host RNG work and no H2D traffic.  `noise` != NULL injects a given z
// (parity tests replay the oracle's noise).
#include "common.cuh"

namespace pdr {
namespace {

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  constexpr unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const unsigned hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0; key.y += W1;
  }
  return ctr;
}

__device__ __forceinline__ float2 box_muller(unsigned a, unsigned b) {
  // u in (0,1], v in [0,1)
  const float u = ((float)(a >> 8) + 1.0f) * (1.0f / 16777216.0f);
  const float v = (float)(b >> 8) * (1.0f / 16777216.0f);
  const float r = sqrtf(-2.0f * __logf(u));
  float s, c;
  __sincosf(6.283185307179586f * v, &s, &c);
  return make_float2(r * c, r * s);
}

__device__ __forceinline__ void normal4(uint64_t seed, uint64_t block, float out[4]) {
  const uint4 r = philox4x32_10(make_uint4((unsigned)block, (unsigned)(block >> 32), 0u, 0u),
                                make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
  const float2 a = box_muller(r.x, r.y), b = box_muller(r.z, r.w);
  out[0] = a.x; out[1] = a.y; out[2] = b.x; out[3] = b.y;
}

__global__ void __launch_bounds__(256)
affine_noise_kernel(size_t count, float *__restrict__ x, const float *__restrict__ eps, float scale_x,
                    float scale_eps, float sigma, const float *__restrict__ noise, uint64_t seed,
                    uint64_t offset) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // group of 4 elements
  const size_t i0 = q * 4;
  if (i0 >= count) return;
  float z[4] = {0.f, 0.f, 0.f, 0.f};
  if (sigma != 0.0f && !noise) normal4(seed, offset + q, z);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const size_t i = i0 + t;
    if (i < count) {
      const float zz = noise ? (sigma != 0.0f ? __ldg(noise + i) : 0.f) : z[t];
      float v = x ? x[i] * scale_x : 0.f;
      if (eps) v = __fmaf_rn(__ldg(eps + i), scale_eps, v);
      x[i] = __fmaf_rn(sigma, zz, v);
    }
  }
}

__global__ void __launch_bounds__(256)
normal_fill_kernel(size_t count, float *__restrict__ x, uint64_t seed, uint64_t offset) {
  const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t i0 = q * 4;
  if (i0 >= count) return;
  float z[4];
  normal4(seed, offset + q, z);
#pragma unroll
  for (int t = 0; t < 4; ++t)
    if (i0 + t < count) x[i0 + t] = z[t];
}

// point_upsample (pointnet2/models/point_upsample_module.py:4-27) as ONE pass: thread per output coordinate.
//   mid = coarse + centre_disp*s;  up[j] = mid + (grid_disp[j]*g)*s,  g = 1/sqrt(factor)
// Every product/sum is rounded separately (__fmul_rn/__fadd_rn) -- the reference is a chain of separate
// elementwise torch kernels, so no FMA contraction happens there; this keeps the result bit-identical.
__global__ void __launch_bounds__(256)
point_upsample_kernel(int b, int n, int reps, int with_centre, const float *__restrict__ coarse,
                      const float *__restrict__ disp, int ld_disp, float grid_scale, float out_scale,
                      float *__restrict__ refined, float *__restrict__ mid_out) {
  const long long total = (long long)b * n * 3;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (cloud, point, xyz)
  if (i >= total) return;
  const int c = (int)(i % 3);
  const long long bp = i / 3;                                              // cloud*n + point
  const int pt = (int)(bp % n);
  const long long cloud = bp / n;
  const float *d = disp + bp * ld_disp;
  const float mid = __fadd_rn(__ldg(coarse + i), __fmul_rn(__ldg(d + c), out_scale));
  if (mid_out) mid_out[i] = mid;
  const long long rows_out = (long long)n * (reps + (with_centre ? 1 : 0));
  float *o = refined + cloud * rows_out * 3;
  for (int j = 0; j < reps; ++j) {
    const float g = __fmul_rn(__ldg(d + 3 + 3 * j + c), grid_scale);
    o[((long long)pt * reps + j) * 3 + c] = __fadd_rn(mid, __fmul_rn(g, out_scale));
  }
  if (with_centre) o[((long long)n * reps + pt) * 3 + c] = mid;
}

}  // namespace
}  // namespace pdr

using namespace pdr;

extern "C" int pdr_point_upsample(int b, int n, int factor, int include_centre, const float *coarse,
                                  const float *displacement, float grid_scale, float out_scale, float *refined,
                                  float *intermediate, void *stream) {
  PDR_REQUIRE(b >= 0 && n >= 0 && factor >= 1, "point_upsample: bad sizes b=%d n=%d factor=%d", b, n, factor);
  if (b == 0 || n == 0) return PDR_OK;
  PDR_REQUIRE(coarse && displacement && refined, "point_upsample: null pointer");
  const int reps = include_centre ? factor - 1 : factor;
  const int ld = 3 * (reps + 1);
  const long long total = (long long)b * n * 3;
  point_upsample_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      b, n, reps, include_centre ? 1 : 0, coarse, displacement, ld, grid_scale, out_scale, refined, intermediate);
  return check_launch("point_upsample");
}

extern "C" int pdr_ddpm_update(size_t count, float *x, const float *eps, float c_eps, float inv_sqrt_alpha,
                               float sigma, const float *noise, uint64_t seed, uint64_t offset, void *stream) {
  if (count == 0) return PDR_OK;
  PDR_REQUIRE(x && eps, "ddpm_update: null pointer");
  // (x - c*eps) * inv = x*inv + eps*(-c*inv)
  const size_t groups = (count + 3) / 4;
  affine_noise_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      count, x, eps, inv_sqrt_alpha, -c_eps * inv_sqrt_alpha, sigma, noise, seed, offset);
  return check_launch("ddpm_update");
}

extern "C" int pdr_affine_noise_update(size_t count, float *x, const float *eps, float scale_x, float scale_eps,
                                       float sigma, const float *noise, uint64_t seed, uint64_t offset,
                                       void *stream) {
  if (count == 0) return PDR_OK;
  PDR_REQUIRE(x && eps, "affine_noise_update: null pointer");
  const size_t groups = (count + 3) / 4;
  affine_noise_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      count, x, eps, scale_x, scale_eps, sigma, noise, seed, offset);
  return check_launch("affine_noise_update");
}

extern "C" int pdr_normal_fill(size_t count, float *x, uint64_t seed, uint64_t offset, void *stream) {
  if (count == 0) return PDR_OK;
  PDR_REQUIRE(x, "normal_fill: null pointer");
  const size_t groups = (count + 3) / 4;
  normal_fill_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, (cudaStream_t)stream>>>(count, x, seed, offset);
  return check_launch("normal_fill");
}
