// Index-driven data movement for sm_100a: gather_points, group_points, three_interpolate (+ grads).
//
// Reference kernels: pointnet2_ops_lib/pointnet2_ops/_ext-src/src/sampling_gpu.cu:8-57,
// group_points_gpu.cu:8-75, interpolate_gpu.cu:72-154.  The reference launches `b` CTAs (one per
// cloud) and walks (channel, point) with a CTA-wide stride; here the grid covers
// (position tiles, channel tiles, b) so all 148 SMs are busy, indices are loaded once per position
// and reused across a tile of channels, and stores are coalesced along the innermost output axis.
#include "common.cuh"

namespace pdr {
namespace {

constexpr int kThreads = 256;
constexpr int kChanPerCta = 8;

// out[b, c, p] = points[b, c, idx[b, p]]   for p in [0, P); P = npoints*nsample (gather: nsample = 1)
__global__ void __launch_bounds__(kThreads)
gather_rows_kernel(int c, int n, int P, const float *__restrict__ points, const int *__restrict__ idx,
                   float *__restrict__ out) {
  const int bi = blockIdx.z;
  const int p = blockIdx.x * kThreads + threadIdx.x;
  if (p >= P) return;
  const int a = __ldg(idx + (size_t)bi * P + p);
  const int c0 = blockIdx.y * kChanPerCta;
  const int c1 = min(c0 + kChanPerCta, c);
  const float *src = points + ((size_t)bi * c + c0) * n + a;
  float *dst = out + ((size_t)bi * c + c0) * P + p;
#pragma unroll 4
  for (int l = c0; l < c1; ++l, src += n, dst += P) *dst = __ldg(src);
}

__global__ void __launch_bounds__(kThreads)
scatter_rows_add_kernel(int c, int n, int P, const float *__restrict__ grad_out,
                        const int *__restrict__ idx, float *__restrict__ grad_points) {
  const int bi = blockIdx.z;
  const int p = blockIdx.x * kThreads + threadIdx.x;
  if (p >= P) return;
  const int a = __ldg(idx + (size_t)bi * P + p);
  const int c0 = blockIdx.y * kChanPerCta;
  const int c1 = min(c0 + kChanPerCta, c);
  for (int l = c0; l < c1; ++l)
    atomicAdd(grad_points + ((size_t)bi * c + l) * n + a, __ldg(grad_out + ((size_t)bi * c + l) * P + p));
}

// out[b,c,j] = fma(p[i3], w3, fma(p[i1], w1, RN(p[i2]*w2)))  -- the rounding sequence nvcc emits for
// interpolate_gpu.cu:98-99 (checked in SASS), so results match the reference bit for bit.
__global__ void __launch_bounds__(kThreads)
three_interpolate_kernel(int c, int m, int n, const float *__restrict__ points,
                         const int *__restrict__ idx, const float *__restrict__ weight,
                         float *__restrict__ out) {
  const int bi = blockIdx.z;
  const int j = blockIdx.x * kThreads + threadIdx.x;
  if (j >= n) return;
  const size_t o = ((size_t)bi * n + j) * 3;
  const int i1 = __ldg(idx + o), i2 = __ldg(idx + o + 1), i3 = __ldg(idx + o + 2);
  const float w1 = __ldg(weight + o), w2 = __ldg(weight + o + 1), w3 = __ldg(weight + o + 2);
  const int c0 = blockIdx.y * kChanPerCta;
  const int c1 = min(c0 + kChanPerCta, c);
  for (int l = c0; l < c1; ++l) {
    const float *p = points + ((size_t)bi * c + l) * m;
    out[((size_t)bi * c + l) * n + j] =
        __fmaf_rn(__ldg(p + i3), w3, __fmaf_rn(__ldg(p + i1), w1, __fmul_rn(__ldg(p + i2), w2)));
  }
}

__global__ void __launch_bounds__(kThreads)
three_interpolate_grad_kernel(int c, int n, int m, const float *__restrict__ grad_out,
                              const int *__restrict__ idx, const float *__restrict__ weight,
                              float *__restrict__ grad_points) {
  const int bi = blockIdx.z;
  const int j = blockIdx.x * kThreads + threadIdx.x;
  if (j >= n) return;
  const size_t o = ((size_t)bi * n + j) * 3;
  const int i1 = __ldg(idx + o), i2 = __ldg(idx + o + 1), i3 = __ldg(idx + o + 2);
  const float w1 = __ldg(weight + o), w2 = __ldg(weight + o + 1), w3 = __ldg(weight + o + 2);
  const int c0 = blockIdx.y * kChanPerCta;
  const int c1 = min(c0 + kChanPerCta, c);
  for (int l = c0; l < c1; ++l) {
    const float g = __ldg(grad_out + ((size_t)bi * c + l) * n + j);
    float *gp = grad_points + ((size_t)bi * c + l) * m;
    atomicAdd(gp + i1, g * w1);
    atomicAdd(gp + i2, g * w2);
    atomicAdd(gp + i3, g * w3);
  }
}

int check_common(const char *op, int b, int c, int n, long long P) {
  PDR_REQUIRE(b >= 0 && c >= 0 && n >= 1 && P >= 0, "%s: bad sizes", op);
  PDR_REQUIRE(b <= 65535, "%s: b > 65535", op);
  PDR_REQUIRE(ceil_div(c > 0 ? c : 1, kChanPerCta) <= 65535, "%s: too many channels", op);
  PDR_REQUIRE(P < (1ll << 31), "%s: npoints*nsample overflows int32", op);
  return PDR_OK;
}

}  // namespace
}  // namespace pdr

using namespace pdr;

extern "C" int pdr_gather_points(int b, int c, int n, int m, const float *points, const int *idx,
                                 float *out, void *stream) {
  int rc = check_common("gather_points", b, c, n, m);
  if (rc) return rc;
  if (b == 0 || c == 0 || m == 0) return PDR_OK;
  PDR_REQUIRE(points && idx && out, "gather_points: null pointer");
  dim3 grid(ceil_div(m, kThreads), ceil_div(c, kChanPerCta), b);
  gather_rows_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(c, n, m, points, idx, out);
  return check_launch("gather_points");
}

extern "C" int pdr_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                                const int *idx, float *out, void *stream) {
  const long long P = (long long)npoints * nsample;
  int rc = check_common("group_points", b, c, n, P);
  if (rc) return rc;
  if (b == 0 || c == 0 || P == 0) return PDR_OK;
  PDR_REQUIRE(points && idx && out, "group_points: null pointer");
  dim3 grid(ceil_div((int)P, kThreads), ceil_div(c, kChanPerCta), b);
  gather_rows_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(c, n, (int)P, points, idx, out);
  return check_launch("group_points");
}

static int scatter_add(const char *op, int b, int c, int n, long long P, const float *grad_out,
                       const int *idx, float *grad_points, void *stream) {
  int rc = check_common(op, b, c, n, P);
  if (rc) return rc;
  if (b == 0 || c == 0) return PDR_OK;
  PDR_REQUIRE(grad_points, "%s: null pointer", op);
  cudaError_t e = cudaMemsetAsync(grad_points, 0, sizeof(float) * (size_t)b * c * n, (cudaStream_t)stream);
  if (e != cudaSuccess) { set_error("%s: memset: %s", op, cudaGetErrorString(e)); return PDR_ERR_CUDA; }
  if (P == 0) return PDR_OK;
  PDR_REQUIRE(grad_out && idx, "%s: null pointer", op);
  dim3 grid(ceil_div((int)P, kThreads), ceil_div(c, kChanPerCta), b);
  scatter_rows_add_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(c, n, (int)P, grad_out, idx, grad_points);
  return check_launch(op);
}

extern "C" int pdr_gather_points_grad(int b, int c, int n, int m, const float *grad_out, const int *idx,
                                      float *grad_points, void *stream) {
  return scatter_add("gather_points_grad", b, c, n, m, grad_out, idx, grad_points, stream);
}

extern "C" int pdr_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                                     const int *idx, float *grad_points, void *stream) {
  return scatter_add("group_points_grad", b, c, n, (long long)npoints * nsample, grad_out, idx, grad_points,
                     stream);
}

extern "C" int pdr_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx,
                                     const float *weight, float *out, void *stream) {
  int rc = check_common("three_interpolate", b, c, m, n);
  if (rc) return rc;
  if (b == 0 || c == 0 || n == 0) return PDR_OK;
  PDR_REQUIRE(points && idx && weight && out, "three_interpolate: null pointer");
  dim3 grid(ceil_div(n, kThreads), ceil_div(c, kChanPerCta), b);
  three_interpolate_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(c, m, n, points, idx, weight, out);
  return check_launch("three_interpolate");
}

extern "C" int pdr_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx,
                                          const float *weight, float *grad_points, void *stream) {
  int rc = check_common("three_interpolate_grad", b, c, m, n);
  if (rc) return rc;
  if (b == 0 || c == 0) return PDR_OK;
  PDR_REQUIRE(grad_points, "three_interpolate_grad: null pointer");
  cudaError_t e = cudaMemsetAsync(grad_points, 0, sizeof(float) * (size_t)b * c * m, (cudaStream_t)stream);
  if (e != cudaSuccess) { set_error("three_interpolate_grad: memset: %s", cudaGetErrorString(e)); return PDR_ERR_CUDA; }
  if (n == 0) return PDR_OK;
  PDR_REQUIRE(grad_out && idx && weight, "three_interpolate_grad: null pointer");
  dim3 grid(ceil_div(n, kThreads), ceil_div(c, kChanPerCta), b);
  three_interpolate_grad_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(c, n, m, grad_out, idx, weight, grad_points);
  return check_launch("three_interpolate_grad");
}
