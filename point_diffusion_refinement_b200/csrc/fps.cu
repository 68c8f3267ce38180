// Furthest point sampling for sm_100a.
//
// Replaces furthest_point_sampling_kernel (reference: pointnet2_ops_lib/pointnet2_ops/_ext-src/src/
// sampling_gpu.cu:69-173, launch :175-229).  Same result bit for bit, different machine mapping:
//
//   reference                                   here
//   ---------                                   ----
//   xyz + temp re-read from global every round  xyz staged once in shared memory (and registers when
//   (20 B/point/round through L1/L2)            <= 8 points/thread), running min distance kept in
//                                               registers for the whole kernel: HBM traffic is the
//                                               algorithmic 12n + 4m bytes per cloud
//   9-level shared-memory tree, 9 barriers      two REDUX.MAX/MIN warp reductions + ONE barrier per
//   per round                                   round (double-buffered partials)
//   argmax index carried through the scan       scan tracks the max VALUE only; the index is
//                                               recovered afterwards by the (few) lanes that hold it
//
// Tie-breaking contract (derived from the reference's strided scan + tree, sampling_gpu.cu:59-65,
// 108-109,115-168; bs = opt_n_threads(n), cuda_utils.h:15-19): among equal maxima the winner is the
// point k minimising (bitreverse_{log2 bs}(k mod bs), k div bs); no candidate at all -> index 0.
// The strict '>' scan of thread t over k = t, t+bs, ... finds the smallest k of its residue class,
// and the tree keeps the lower slot on equality, which yields exactly that order.
#include "common.cuh"

namespace pdr {
namespace {

constexpr int kFpsMaxThreads = 512;
constexpr int kFpsMaxPPT = 32;
constexpr int kFpsMaxOnchip = 16384;  // 16384 * 12 B = 192 KiB of shared memory

__host__ __device__ inline int fps_ref_block(int n) {  // cuda_utils.h:15-19 without the log() detour
  int p = 1;
  while (p * 2 <= n && p * 2 <= kFpsMaxThreads) p *= 2;
  return p;
}

struct WarpBest {
  unsigned val;  // float bits of the candidate distance + 1; 0 = no candidate
  unsigned tie;  // (bitrev(k mod bs) << 23) | (k div bs); smaller wins
};

__device__ __forceinline__ WarpBest warp_argmax(unsigned val, unsigned tie) {
  const unsigned vmax = __reduce_max_sync(0xffffffffu, val);
  const unsigned t = (val == vmax) ? tie : 0xffffffffu;
  const unsigned tmin = __reduce_min_sync(0xffffffffu, t);
  return {vmax, tmin};
}

template <int PPT, bool XYZ_REGS>
__global__ void __launch_bounds__(kFpsMaxThreads)
fps_onchip_kernel(int n, int m, int bs, int log2bs, const float *__restrict__ xyz_all,
                  int *__restrict__ idx_all) {
  extern __shared__ float sxyz[];  // n*3
  __shared__ uint2 s_red[2][16];

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int nwarps = blockDim.x >> 5;
  const float *xyz = xyz_all + (size_t)blockIdx.x * n * 3;
  int *idx = idx_all + (size_t)blockIdx.x * m;

  for (int i = tid; i < n * 3; i += blockDim.x) sxyz[i] = __ldg(xyz + i);
  __syncthreads();

  // Per-thread resident state: running min distance (the reference's `temp`), validity mask
  // (k < n and not skipped by the |p|^2 <= 1e-3 rule, sampling_gpu.cu:100-101), optionally xyz.
  float temp[PPT];
  float px[XYZ_REGS ? PPT : 1], py[XYZ_REGS ? PPT : 1], pz[XYZ_REGS ? PPT : 1];
  unsigned valid = 0;
  const bool active = tid < bs;
#pragma unroll
  for (int i = 0; i < PPT; ++i) {
    const int k = tid + i * bs;
    temp[i] = 1e10f;
    if (active && k < n) {
      const float x = sxyz[k * 3 + 0], y = sxyz[k * 3 + 1], z = sxyz[k * 3 + 2];
      if (XYZ_REGS) { px[i] = x; py[i] = y; pz[i] = z; }
      const float mag = __fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y)));
      if (!((double)mag <= 1e-3)) valid |= 1u << i;
    } else if (XYZ_REGS) {
      px[i] = py[i] = pz[i] = 0.f;
    }
  }
  const unsigned brev_t = log2bs ? (__brev((unsigned)tid) >> (32 - log2bs)) : 0u;

  int old = 0;
  if (tid == 0) idx[0] = 0;
  int buf = 0;
  for (int j = 1; j < m; ++j) {
    const float x1 = sxyz[old * 3 + 0], y1 = sxyz[old * 3 + 1], z1 = sxyz[old * 3 + 2];
    float lmax = -1.0f;
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
      float x2, y2, z2;
      if (XYZ_REGS) { x2 = px[i]; y2 = py[i]; z2 = pz[i]; }
      else {
        const int k = min(tid + i * bs, n - 1);
        x2 = sxyz[k * 3 + 0]; y2 = sxyz[k * 3 + 1]; z2 = sxyz[k * 3 + 2];
      }
      const float d = dist2_ref(__fsub_rn(x2, x1), __fsub_rn(y2, y1), __fsub_rn(z2, z1));
      const float t = fminf(d, temp[i]);
      if (valid & (1u << i)) {
        temp[i] = t;
        lmax = fmaxf(lmax, t);
      }
    }
    const unsigned val = lmax < 0.0f ? 0u : __float_as_uint(lmax) + 1u;
    const unsigned wmax = __reduce_max_sync(0xffffffffu, val);
    unsigned tie = 0xffffffffu;
    if (val == wmax && wmax != 0u) {
      int ibest = 0;
#pragma unroll
      for (int i = PPT - 1; i >= 0; --i)
        if ((valid & (1u << i)) && temp[i] == lmax) ibest = i;
      tie = (brev_t << 23) | (unsigned)ibest;
    }
    const unsigned wtie = __reduce_min_sync(0xffffffffu, tie);
    if (lane == 0) s_red[buf][warp] = make_uint2(wmax, wtie);
    __syncthreads();
    uint2 r = make_uint2(0u, 0xffffffffu);
    if (lane < nwarps) r = s_red[buf][lane];
    const WarpBest g = warp_argmax(r.x, r.y);
    if (g.val == 0u) {
      old = 0;
    } else {
      const unsigned t = log2bs ? (__brev(g.tie >> 23) >> (32 - log2bs)) : 0u;
      old = (int)((g.tie & 0x7fffffu) * (unsigned)bs + t);
    }
    if (tid == 0) idx[j] = old;
    buf ^= 1;
  }
}

// Any n: running min distance in a caller-provided global scratch (one CTA of 512 threads per cloud).
__global__ void __launch_bounds__(kFpsMaxThreads)
fps_global_kernel(int n, int m, int log2bs, const float *__restrict__ xyz_all, float *temp_all,
                  int *__restrict__ idx_all) {
  __shared__ uint2 s_red[2][16];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int bs = blockDim.x;
  const float *xyz = xyz_all + (size_t)blockIdx.x * n * 3;
  float *temp = temp_all + (size_t)blockIdx.x * n;
  int *idx = idx_all + (size_t)blockIdx.x * m;
  for (int k = tid; k < n; k += bs) temp[k] = 1e10f;
  const unsigned brev_t = log2bs ? (__brev((unsigned)tid) >> (32 - log2bs)) : 0u;
  int old = 0;
  if (tid == 0) idx[0] = 0;
  int buf = 0;
  for (int j = 1; j < m; ++j) {
    const float x1 = __ldg(xyz + old * 3 + 0), y1 = __ldg(xyz + old * 3 + 1), z1 = __ldg(xyz + old * 3 + 2);
    float best = -1.0f;
    int ibest = 0;
    for (int k = tid, i = 0; k < n; k += bs, ++i) {
      const float x2 = __ldg(xyz + k * 3 + 0), y2 = __ldg(xyz + k * 3 + 1), z2 = __ldg(xyz + k * 3 + 2);
      const float mag = __fmaf_rn(z2, z2, __fmaf_rn(x2, x2, __fmul_rn(y2, y2)));
      if ((double)mag <= 1e-3) continue;
      const float d = dist2_ref(__fsub_rn(x2, x1), __fsub_rn(y2, y1), __fsub_rn(z2, z1));
      const float t = fminf(d, temp[k]);
      temp[k] = t;
      if (t > best) { best = t; ibest = i; }
    }
    const unsigned val = best < 0.0f ? 0u : __float_as_uint(best) + 1u;
    const WarpBest w = warp_argmax(val, (brev_t << 23) | (unsigned)ibest);
    if (lane == 0) s_red[buf][warp] = make_uint2(w.val, w.tie);
    __syncthreads();
    uint2 r = make_uint2(0u, 0xffffffffu);
    if (lane < nwarps) r = s_red[buf][lane];
    const WarpBest g = warp_argmax(r.x, r.y);
    if (g.val == 0u) old = 0;
    else {
      const unsigned t = log2bs ? (__brev(g.tie >> 23) >> (32 - log2bs)) : 0u;
      old = (int)((g.tie & 0x7fffffu) * (unsigned)bs + t);
    }
    if (tid == 0) idx[j] = old;
    buf ^= 1;
  }
}

template <int PPT, bool XYZ_REGS>
int launch_onchip(int b, int n, int m, int bs, int log2bs, const float *xyz, int *idx,
                  cudaStream_t stream) {
  const size_t smem = (size_t)n * 3 * sizeof(float);
  auto kern = fps_onchip_kernel<PPT, XYZ_REGS>;
  if (smem + 2048 > 48 * 1024) {  // static __shared__ counts against the 48 KiB default too
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("fps: cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
      return PDR_ERR_CUDA;
    }
  }
  kern<<<b, bs < 32 ? 32 : bs, smem, stream>>>(n, m, bs, log2bs, xyz, idx);
  return check_launch("fps_onchip_kernel");
}

}  // namespace
}  // namespace pdr

extern "C" int pdr_fps_max_onchip_points(void) { return pdr::kFpsMaxOnchip; }

extern "C" int pdr_furthest_point_sampling(int b, int n, int m, const float *xyz, float *temp, int *idx,
                                           void *stream_) {
  using namespace pdr;
  cudaStream_t stream = (cudaStream_t)stream_;
  PDR_REQUIRE(b >= 0 && n >= 1 && m >= 0, "fps: bad sizes b=%d n=%d m=%d", b, n, m);
  PDR_REQUIRE(n < (1 << 30), "fps: n too large");
  if (b == 0 || m == 0) return PDR_OK;
  PDR_REQUIRE(xyz && idx, "fps: null pointer");
  const int bs = fps_ref_block(n);
  int log2bs = 0;
  while ((1 << log2bs) < bs) ++log2bs;
  if (n <= kFpsMaxOnchip) {
    const int ppt = ceil_div(n, bs);
    if (ppt <= 1) return launch_onchip<1, true>(b, n, m, bs, log2bs, xyz, idx, stream);
    if (ppt <= 2) return launch_onchip<2, true>(b, n, m, bs, log2bs, xyz, idx, stream);
    if (ppt <= 4) return launch_onchip<4, true>(b, n, m, bs, log2bs, xyz, idx, stream);
    if (ppt <= 8) return launch_onchip<8, true>(b, n, m, bs, log2bs, xyz, idx, stream);
    if (ppt <= 16) return launch_onchip<16, false>(b, n, m, bs, log2bs, xyz, idx, stream);
    return launch_onchip<32, false>(b, n, m, bs, log2bs, xyz, idx, stream);
  }
  if (!temp) {
    set_error("fps: n=%d > %d needs the (b,n) fp32 temp scratch", n, kFpsMaxOnchip);
    return PDR_ERR_WORKSPACE;
  }
  fps_global_kernel<<<b, bs, 0, stream>>>(n, m, log2bs, xyz, temp, idx);
  return check_launch("fps_global_kernel");
}
