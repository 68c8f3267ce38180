// Ball query for sm_100a.
//
// Replaces query_ball_point_kernel (reference: pointnet2_ops_lib/pointnet2_ops/_ext-src/src/
// ball_query_gpu.cu:9-47, binding ball_query.cpp:10-38 incl. the fork-specific `counts` output).
// Semantics kept bit for bit: for each centre, the first `nsample` point indices (in index order)
// with d^2 < r^2 (fp32, r^2 = r*r), unfilled slots repeat the first hit, counts = #found
// (<= nsample), rows without any hit are all zero with count 0.
//
// Mapping: the reference runs ONE THREAD per centre, each streaming all n points from global memory
// with only `b` CTAs.  Here one WARP owns a centre: the cloud is staged once per CTA in shared
// memory, the 32 lanes test 32 consecutive points per step (stride-3 word access is bank-conflict
// free), a ballot + popc prefix gives each hit its slot in index order, the warp leaves as soon as
// `nsample` hits exist, and the finished row goes out as one coalesced store.  Grid = (centre
// chunks, b): thousands of CTAs instead of b.
#include "common.cuh"

namespace pdr {
namespace {

constexpr int kBqThreads = 256;
constexpr int kBqWarps = kBqThreads / 32;
constexpr int kBqCentresPerWarp = 8;
constexpr int kBqCentresPerCta = kBqWarps * kBqCentresPerWarp;

template <bool STAGED>
__global__ void __launch_bounds__(kBqThreads)
ball_query_kernel(int n, int m, float radius2, int nsample, const float *__restrict__ new_xyz_all,
                  const float *__restrict__ xyz_all, int *__restrict__ idx_all,
                  int *__restrict__ counts_all) {
  extern __shared__ float smem[];
  int *s_found = reinterpret_cast<int *>(smem);          // [kBqWarps][nsample]
  float *sxyz = smem + kBqWarps * nsample;                // [n*3] when STAGED
  const int bi = blockIdx.y;
  const float *xyz = xyz_all + (size_t)bi * n * 3;
  const float *new_xyz = new_xyz_all + (size_t)bi * m * 3;
  int *idx = idx_all + (size_t)bi * m * nsample;
  int *counts = counts_all + (size_t)bi * m;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  if (STAGED) {
    for (int i = threadIdx.x; i < n * 3; i += kBqThreads) sxyz[i] = __ldg(xyz + i);
    __syncthreads();
  }
  const float *pts = STAGED ? sxyz : xyz;
  int *found = s_found + warp * nsample;
  const unsigned lt_mask = (1u << lane) - 1u;

  const int j0 = blockIdx.x * kBqCentresPerCta;
  for (int c = 0; c < kBqCentresPerWarp; ++c) {
    const int j = j0 + c * kBqWarps + warp;
    if (j >= m) break;
    const float cx = __ldg(new_xyz + j * 3 + 0), cy = __ldg(new_xyz + j * 3 + 1),
                cz = __ldg(new_xyz + j * 3 + 2);
    int cnt = 0;
    for (int k0 = 0; k0 < n && cnt < nsample; k0 += 128) {
      // four independent 32-point tests per trip for ILP; slots are assigned in index order
      bool hit[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = k0 + u * 32 + lane;
        const int kk = k < n ? k : n - 1;
        const float x = pts[kk * 3 + 0], y = pts[kk * 3 + 1], z = pts[kk * 3 + 2];
        const float d2 = dist2_ref(__fsub_rn(cx, x), __fsub_rn(cy, y), __fsub_rn(cz, z));
        hit[u] = (k < n) && (d2 < radius2);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const unsigned mask = __ballot_sync(0xffffffffu, hit[u]);
        if (mask) {
          const int pos = cnt + __popc(mask & lt_mask);
          if (hit[u] && pos < nsample) found[pos] = k0 + u * 32 + lane;
          cnt += __popc(mask);
        }
      }
    }
    cnt = cnt < nsample ? cnt : nsample;
    __syncwarp();
    const int first = cnt > 0 ? found[0] : 0;
    for (int l = lane; l < nsample; l += 32) idx[(size_t)j * nsample + l] = l < cnt ? found[l] : first;
    if (lane == 0) counts[j] = cnt;
    __syncwarp();
  }
}

}  // namespace
}  // namespace pdr

extern "C" int pdr_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                              const float *xyz, int *idx, int *counts, void *stream_) {
  using namespace pdr;
  cudaStream_t stream = (cudaStream_t)stream_;
  PDR_REQUIRE(b >= 0 && n >= 1 && m >= 0 && nsample >= 1, "ball_query: bad sizes b=%d n=%d m=%d ns=%d", b,
              n, m, nsample);
  PDR_REQUIRE(b <= 65535, "ball_query: b > 65535");
  PDR_REQUIRE(nsample <= 1024, "ball_query: nsample > 1024");
  if (b == 0 || m == 0) return PDR_OK;
  PDR_REQUIRE(new_xyz && xyz && idx && counts, "ball_query: null pointer");
  const float radius2 = radius * radius;  // fp32 product, ball_query_gpu.cu:24
  const dim3 grid(ceil_div(m, kBqCentresPerCta), b);
  const size_t found_bytes = (size_t)kBqWarps * nsample * sizeof(int);
  const size_t staged_bytes = found_bytes + (size_t)n * 3 * sizeof(float);
  if (staged_bytes <= 200 * 1024) {
    auto kern = ball_query_kernel<true>;
    if (staged_bytes + 2048 > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e != cudaSuccess) { set_error("ball_query: smem attr: %s", cudaGetErrorString(e)); return PDR_ERR_CUDA; }
    }
    kern<<<grid, kBqThreads, staged_bytes, stream>>>(n, m, radius2, nsample, new_xyz, xyz, idx, counts);
  } else {
    ball_query_kernel<false><<<grid, kBqThreads, found_bytes, stream>>>(n, m, radius2, nsample, new_xyz,
                                                                         xyz, idx, counts);
  }
  return check_launch("ball_query_kernel");
}
