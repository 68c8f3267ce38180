// Nearest-neighbour family for sm_100a: three_nn, knn_points, Chamfer/F1, NmDistance.
//
// Reference: three_nn_kernel (pointnet2_ops_lib/pointnet2_ops/_ext-src/src/interpolate_gpu.cu:9-59),
// NmDistanceKernel (pointnet2/models/pvd/metrics/ChamferDistancePytorch/chamfer3D/chamfer3D.cu:12-134),
// pytorch3d.ops.knn.knn_points (un-vendored; call sites pointnet2/chamfer_loss_new.py:149-150,
// pointnet2_ops/pointnet2_utils.py:365,496-497), calc_cd/fscore (chamfer_loss_new.py:219-245).
//
// Common mapping: one thread owns Q query points in registers; the target cloud streams through
// shared memory in float4 tiles that every lane reads as a broadcast (one LDS.128 per target point per
// warp, amortised over 32*Q pair evaluations); the grid is (query tiles, [direction], b).
// These kernels are FP32-ALU bound (7 instr / pair), not HBM bound: 16 B/point in, 4 B/point out.
#include <math_constants.h>

#include "common.cuh"

namespace pdr {
namespace {

constexpr int kTile = 1024;  // target points per shared-memory tile (16 KiB as float4)

__device__ __forceinline__ void stage_tile(float4 *s, const float *__restrict__ pts, int k0, int cnt) {
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    const float *p = pts + (size_t)(k0 + i) * 3;
    s[i] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0.f);
  }
}

// ---------------------------------------------------------------------------------------------
// three_nn
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
three_nn_kernel(int n, int m, const float *__restrict__ unknown_all, const float *__restrict__ known_all,
                float *__restrict__ dist2_all, int *__restrict__ idx_all) {
  __shared__ float4 tile[kTile];
  const int bi = blockIdx.y;
  const float *known = known_all + (size_t)bi * m * 3;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const bool on = j < n;
  float ux = 0, uy = 0, uz = 0;
  if (on) {
    const float *u = unknown_all + ((size_t)bi * n + j) * 3;
    ux = __ldg(u); uy = __ldg(u + 1); uz = __ldg(u + 2);
  }
  // interpolate_gpu.cu:27 holds the bests in double initialised to 1e40; for finite inputs an fp32
  // +inf start is indistinguishable (every finite d compares below both).
  float best1 = CUDART_INF_F, best2 = CUDART_INF_F, best3 = CUDART_INF_F;
  int besti1 = 0, besti2 = 0, besti3 = 0;
  for (int k0 = 0; k0 < m; k0 += kTile) {
    const int cnt = min(kTile, m - k0);
    __syncthreads();
    stage_tile(tile, known, k0, cnt);
    __syncthreads();
    if (on) {
#pragma unroll 4
      for (int k = 0; k < cnt; ++k) {
        const float4 p = tile[k];
        const float d = dist2_ref(__fsub_rn(ux, p.x), __fsub_rn(uy, p.y), __fsub_rn(uz, p.z));
        if (d < best3) {
          if (d < best1) {
            best3 = best2; besti3 = besti2; best2 = best1; besti2 = besti1; best1 = d; besti1 = k0 + k;
          } else if (d < best2) {
            best3 = best2; besti3 = besti2; best2 = d; besti2 = k0 + k;
          } else {
            best3 = d; besti3 = k0 + k;
          }
        }
      }
    }
  }
  if (on) {
    const size_t o = ((size_t)bi * n + j) * 3;
    dist2_all[o] = best1; dist2_all[o + 1] = best2; dist2_all[o + 2] = best3;
    idx_all[o] = besti1; idx_all[o + 1] = besti2; idx_all[o + 2] = besti3;
  }
}

// ---------------------------------------------------------------------------------------------
// knn_points: exact brute force, sorted ascending, ties keep the lower index first.
// ---------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(K >= 32 ? 128 : 256)
knn_kernel(int p1, int p2, int kout, const float *__restrict__ x_all, const float *__restrict__ y_all,
           float *__restrict__ dists_all, int64_t *__restrict__ idx_all) {
  __shared__ float4 tile[kTile];
  const int bi = blockIdx.y;
  const float *y = y_all + (size_t)bi * p2 * 3;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool on = i < p1;
  float qx = 0, qy = 0, qz = 0;
  if (on) {
    const float *q = x_all + ((size_t)bi * p1 + i) * 3;
    qx = __ldg(q); qy = __ldg(q + 1); qz = __ldg(q + 2);
  }
  float bd[K];
  int bk[K];
#pragma unroll
  for (int t = 0; t < K; ++t) { bd[t] = CUDART_INF_F; bk[t] = 0; }
  for (int k0 = 0; k0 < p2; k0 += kTile) {
    const int cnt = min(kTile, p2 - k0);
    __syncthreads();
    stage_tile(tile, y, k0, cnt);
    __syncthreads();
    if (on) {
      for (int k = 0; k < cnt; ++k) {
        const float4 p = tile[k];
        const float d = dist2_xyz(__fsub_rn(qx, p.x), __fsub_rn(qy, p.y), __fsub_rn(qz, p.z));
        if (d < bd[K - 1]) {
          bd[K - 1] = d; bk[K - 1] = k0 + k;
#pragma unroll
          for (int t = K - 1; t > 0; --t) {
            if (bd[t] < bd[t - 1]) {  // strict: an equal, later point stays behind the earlier one
              const float td = bd[t]; bd[t] = bd[t - 1]; bd[t - 1] = td;
              const int tk = bk[t]; bk[t] = bk[t - 1]; bk[t - 1] = tk;
            }
          }
        }
      }
    }
  }
  if (on) {
    const int have = min(kout, p2);
    float *dd = dists_all + ((size_t)bi * p1 + i) * kout;
    int64_t *di = idx_all + ((size_t)bi * p1 + i) * kout;
#pragma unroll
    for (int t = 0; t < K; ++t)
      if (t < kout) { dd[t] = t < have ? bd[t] : 0.f; di[t] = t < have ? (int64_t)bk[t] : 0; }
  }
}

// ---------------------------------------------------------------------------------------------
// Chamfer + F1, both directions in one launch (blockIdx.y = direction).
// direction 0: queries = xyz2 (gt), targets = xyz1 (output)  -> dist1 (b,m)
// direction 1: queries = xyz1,     targets = xyz2          -> dist2 (b,n)
// Each CTA writes its partial (sum d, sum sqrt d, #(d < thr)) in a fixed slot; a second tiny kernel
// adds the slots in index order, so results are deterministic (no float atomics).
// ---------------------------------------------------------------------------------------------
constexpr int kChQ = 4;          // queries per thread (8 measured equal: r03ch2)
constexpr int kChThreads = 128;  // => 512 queries per CTA

__global__ void __launch_bounds__(kChThreads)
chamfer_min_kernel(int n, int m, const float *__restrict__ xyz1_all, const float *__restrict__ xyz2_all,
                   float thr, float *__restrict__ dist1_all, float *__restrict__ dist2_all,
                   float *__restrict__ partials, int nblk_max) {
  __shared__ float4 tile[kTile];
  __shared__ float s_part[kChThreads / 32][3];
  const int bi = blockIdx.z, dir = blockIdx.y;
  const int nq = dir == 0 ? m : n, nt = dir == 0 ? n : m;
  float *part = partials + (((size_t)bi * 2 + dir) * nblk_max + blockIdx.x) * 3;
  const int q0 = blockIdx.x * (kChThreads * kChQ);
  if (q0 >= nq) return;  // uniform per CTA; slots beyond the direction's own block count are not read
  const float *qs = (dir == 0 ? xyz2_all : xyz1_all) + (size_t)bi * nq * 3;
  const float *ts = (dir == 0 ? xyz1_all : xyz2_all) + (size_t)bi * nt * 3;
  float *dout = dir == 0 ? dist1_all : dist2_all;

  float qx[kChQ], qy[kChQ], qz[kChQ], best[kChQ];
#pragma unroll
  for (int u = 0; u < kChQ; ++u) {
    const int q = min(q0 + u * kChThreads + threadIdx.x, nq - 1);
    qx[u] = __ldg(qs + (size_t)q * 3); qy[u] = __ldg(qs + (size_t)q * 3 + 1); qz[u] = __ldg(qs + (size_t)q * 3 + 2);
    best[u] = CUDART_INF_F;
  }
  for (int k0 = 0; k0 < nt; k0 += kTile) {
    const int cnt = min(kTile, nt - k0);
    __syncthreads();
    stage_tile(tile, ts, k0, cnt);
    __syncthreads();
    // two queries per instruction on packed fp32 pairs (add / mul / fma.rn.f32x2, as in emd.cu): (p - q)^2 == (q - p)^2 exactly and
    // every operation is the .rn form of dist2_xyz's, so the distances are bit-identical; 4.75 issue slots per pair instead of 7
    unsigned long long nqx[kChQ / 2], nqy[kChQ / 2], nqz[kChQ / 2];
#pragma unroll
    for (int h = 0; h < kChQ / 2; ++h) {
      asm("mov.b64 %0, {%1, %2};" : "=l"(nqx[h]) : "f"(-qx[2 * h]), "f"(-qx[2 * h + 1]));
      asm("mov.b64 %0, {%1, %2};" : "=l"(nqy[h]) : "f"(-qy[2 * h]), "f"(-qy[2 * h + 1]));
      asm("mov.b64 %0, {%1, %2};" : "=l"(nqz[h]) : "f"(-qz[2 * h]), "f"(-qz[2 * h + 1]));
    }
#pragma unroll 4
    for (int k = 0; k < cnt; ++k) {
      const float4 p = tile[k];
      unsigned long long px, py, pz;
      asm("mov.b64 %0, {%1, %1};" : "=l"(px) : "f"(p.x));
      asm("mov.b64 %0, {%1, %1};" : "=l"(py) : "f"(p.y));
      asm("mov.b64 %0, {%1, %1};" : "=l"(pz) : "f"(p.z));
#pragma unroll
      for (int h = 0; h < kChQ / 2; ++h) {
        unsigned long long dx, dy, dz, d;
        asm("add.rn.f32x2 %0, %1, %2;" : "=l"(dx) : "l"(px), "l"(nqx[h]));
        asm("add.rn.f32x2 %0, %1, %2;" : "=l"(dy) : "l"(py), "l"(nqy[h]));
        asm("add.rn.f32x2 %0, %1, %2;" : "=l"(dz) : "l"(pz), "l"(nqz[h]));
        asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(d) : "l"(dx));
        asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(d) : "l"(dy));
        asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(d) : "l"(dz));
        float d0, d1;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
        best[2 * h] = fminf(best[2 * h], d0);
        best[2 * h + 1] = fminf(best[2 * h + 1], d1);
      }
    }
  }
  float s_d = 0.f, s_sqrt = 0.f, s_cnt = 0.f;
#pragma unroll
  for (int u = 0; u < kChQ; ++u) {
    const int q = q0 + u * kChThreads + threadIdx.x;
    if (q < nq) {
      if (dout) dout[(size_t)bi * nq + q] = best[u];
      s_d += best[u];
      s_sqrt += sqrtf(best[u]);
      s_cnt += best[u] < thr ? 1.f : 0.f;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s_d += __shfl_xor_sync(0xffffffffu, s_d, o);
    s_sqrt += __shfl_xor_sync(0xffffffffu, s_sqrt, o);
    s_cnt += __shfl_xor_sync(0xffffffffu, s_cnt, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s_part[warp][0] = s_d; s_part[warp][1] = s_sqrt; s_part[warp][2] = s_cnt; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float a = 0.f;
    for (int w = 0; w < kChThreads / 32; ++w) a += s_part[w][threadIdx.x];
    part[threadIdx.x] = a;
  }
}

// cd_p = (mean sqrt d1 + mean sqrt d2)/2 ; cd_t = mean d1 + mean d2 ; f1 = 2 p1 p2/(p1+p2), NaN -> 0
// (chamfer_loss_new.py:219-245).  One warp per cloud.
__global__ void chamfer_finalize_kernel(int b, int n, int m, const float *__restrict__ partials, int nblk_max,
                                        float *__restrict__ cd_p, float *__restrict__ cd_t,
                                        float *__restrict__ f1) {
  const int bi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (bi >= b) return;
  const int lane = threadIdx.x & 31;
  float acc[2][3];
  for (int dir = 0; dir < 2; ++dir) {
    const int nq = dir == 0 ? m : n;
    const int nblk = (nq + kChThreads * kChQ - 1) / (kChThreads * kChQ);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int k = lane; k < nblk; k += 32) {
      const float *p = partials + (((size_t)bi * 2 + dir) * nblk_max + k) * 3;
      a0 += p[0]; a1 += p[1]; a2 += p[2];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, o);
      a1 += __shfl_xor_sync(0xffffffffu, a1, o);
      a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    acc[dir][0] = a0 / (float)nq; acc[dir][1] = a1 / (float)nq; acc[dir][2] = a2 / (float)nq;
  }
  if (lane == 0) {
    cd_p[bi] = (acc[0][1] + acc[1][1]) / 2.f;
    cd_t[bi] = acc[0][0] + acc[1][0];
    const float p1 = acc[0][2], p2 = acc[1][2];
    const float f = 2.f * p1 * p2 / (p1 + p2);
    f1[bi] = (f != f) ? 0.f : f;
  }
}

// ---------------------------------------------------------------------------------------------
// NmDistance (chamfer3D semantics: value + argmin, first minimum wins), both directions per launch.
// ---------------------------------------------------------------------------------------------
constexpr int kNmQ = 2;
constexpr int kNmThreads = 128;

__global__ void __launch_bounds__(kNmThreads)
nm_distance_kernel(int n, int m, const float *__restrict__ xyz1_all, const float *__restrict__ xyz2_all,
                   float *__restrict__ dist1_all, int *__restrict__ idx1_all,
                   float *__restrict__ dist2_all, int *__restrict__ idx2_all) {
  __shared__ float4 tile[kTile];
  const int bi = blockIdx.z, dir = blockIdx.y;
  const int nq = dir == 0 ? n : m, nt = dir == 0 ? m : n;
  const int q0 = blockIdx.x * (kNmThreads * kNmQ);
  if (q0 >= nq) return;
  const float *qs = (dir == 0 ? xyz1_all : xyz2_all) + (size_t)bi * nq * 3;
  const float *ts = (dir == 0 ? xyz2_all : xyz1_all) + (size_t)bi * nt * 3;
  float *dout = (dir == 0 ? dist1_all : dist2_all) + (size_t)bi * nq;
  int *iout = (dir == 0 ? idx1_all : idx2_all) + (size_t)bi * nq;
  float qx[kNmQ], qy[kNmQ], qz[kNmQ], best[kNmQ];
  int besti[kNmQ];
#pragma unroll
  for (int u = 0; u < kNmQ; ++u) {
    const int q = min(q0 + u * kNmThreads + threadIdx.x, nq - 1);
    qx[u] = __ldg(qs + (size_t)q * 3); qy[u] = __ldg(qs + (size_t)q * 3 + 1); qz[u] = __ldg(qs + (size_t)q * 3 + 2);
    best[u] = CUDART_INF_F; besti[u] = 0;
  }
  for (int k0 = 0; k0 < nt; k0 += kTile) {
    const int cnt = min(kTile, nt - k0);
    __syncthreads();
    stage_tile(tile, ts, k0, cnt);
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < cnt; ++k) {
      const float4 p = tile[k];
#pragma unroll
      for (int u = 0; u < kNmQ; ++u) {
        const float d = dist2_ref(__fsub_rn(p.x, qx[u]), __fsub_rn(p.y, qy[u]), __fsub_rn(p.z, qz[u]));
        if (d < best[u]) { best[u] = d; besti[u] = k0 + k; }
      }
    }
  }
#pragma unroll
  for (int u = 0; u < kNmQ; ++u) {
    const int q = q0 + u * kNmThreads + threadIdx.x;
    if (q < nq) { dout[q] = best[u]; iout[q] = besti[u]; }
  }
}

template <int K>
int launch_knn(int b, int p1, int p2, int kout, const float *x, const float *y, float *dists, int64_t *idx,
               cudaStream_t stream) {
  const int threads = K >= 32 ? 128 : 256;
  dim3 grid(ceil_div(p1, threads), b);
  knn_kernel<K><<<grid, threads, 0, stream>>>(p1, p2, kout, x, y, dists, idx);
  return check_launch("knn_kernel");
}

}  // namespace
}  // namespace pdr

using namespace pdr;

extern "C" int pdr_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2,
                            int *idx, void *stream) {
  PDR_REQUIRE(b >= 0 && n >= 0 && m >= 0 && b <= 65535, "three_nn: bad sizes b=%d n=%d m=%d", b, n, m);
  if (b == 0 || n == 0) return PDR_OK;
  PDR_REQUIRE(unknown && (known || m == 0) && dist2 && idx, "three_nn: null pointer");
  dim3 grid(ceil_div(n, 256), b);
  three_nn_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, m, unknown, known, dist2, idx);
  return check_launch("three_nn_kernel");
}

extern "C" int pdr_knn_points(int b, int p1, int p2, int K, const float *x, const float *y, float *dists,
                              int64_t *idx, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PDR_REQUIRE(b >= 0 && p1 >= 0 && p2 >= 0 && b <= 65535, "knn_points: bad sizes b=%d p1=%d p2=%d", b, p1, p2);
  PDR_REQUIRE(K >= 1 && K <= 64, "knn_points: K=%d outside [1,64]", K);
  if (b == 0 || p1 == 0) return PDR_OK;
  PDR_REQUIRE(x && (y || p2 == 0) && dists && idx, "knn_points: null pointer");
  if (K <= 1) return launch_knn<1>(b, p1, p2, K, x, y, dists, idx, stream);
  if (K <= 2) return launch_knn<2>(b, p1, p2, K, x, y, dists, idx, stream);
  if (K <= 4) return launch_knn<4>(b, p1, p2, K, x, y, dists, idx, stream);
  if (K <= 8) return launch_knn<8>(b, p1, p2, K, x, y, dists, idx, stream);
  if (K <= 16) return launch_knn<16>(b, p1, p2, K, x, y, dists, idx, stream);
  if (K <= 32) return launch_knn<32>(b, p1, p2, K, x, y, dists, idx, stream);
  return launch_knn<64>(b, p1, p2, K, x, y, dists, idx, stream);
}

static int chamfer_nblk_max(int n, int m) {
  const int per = kChThreads * kChQ;
  return ceil_div(n > m ? n : m, per);
}

extern "C" size_t pdr_chamfer_f1_workspace_bytes(int b, int n, int m) {
  if (b <= 0 || n <= 0 || m <= 0) return 0;
  return (size_t)b * 2 * chamfer_nblk_max(n, m) * 3 * sizeof(float);
}

extern "C" int pdr_chamfer_f1(int b, int n, int m, const float *xyz1, const float *xyz2, float f1_threshold,
                              float *cd_p, float *cd_t, float *f1, float *dist1, float *dist2,
                              void *workspace, size_t workspace_bytes, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PDR_REQUIRE(b >= 0 && n >= 1 && m >= 1 && b <= 65535, "chamfer_f1: bad sizes b=%d n=%d m=%d", b, n, m);
  if (b == 0) return PDR_OK;
  PDR_REQUIRE(xyz1 && xyz2 && cd_p && cd_t && f1, "chamfer_f1: null pointer");
  const size_t need = pdr_chamfer_f1_workspace_bytes(b, n, m);
  if (!workspace || workspace_bytes < need) {
    set_error("chamfer_f1: workspace %zu B < %zu B", workspace_bytes, need);
    return PDR_ERR_WORKSPACE;
  }
  const int nblk_max = chamfer_nblk_max(n, m);
  dim3 grid(nblk_max, 2, b);
  chamfer_min_kernel<<<grid, kChThreads, 0, stream>>>(n, m, xyz1, xyz2, f1_threshold, dist1, dist2,
                                                      (float *)workspace, nblk_max);
  int rc = check_launch("chamfer_min_kernel");
  if (rc) return rc;
  chamfer_finalize_kernel<<<ceil_div(b, 4), 128, 0, stream>>>(b, n, m, (const float *)workspace, nblk_max, cd_p,
                                                               cd_t, f1);
  return check_launch("chamfer_finalize_kernel");
}

namespace pdr {
namespace {
// Backward of NmDistance (chamfer3D.cu:155-174, called twice by chamfer_cuda_backward :176-196): for the points A of
// one cloud, grad_A[j] = 2 g_a[j] (A_j - B_{idx_a[j]})  -  sum_{k : idx_b[k] == j} 2 g_b[k] (B_k - A_j).
// The reference scatters the second term with atomicAdd (order of the additions undefined); here a thread owns one
// output point and walks the other cloud's index list in order: deterministic, no atomics, O(n m) integer compares.
constexpr int kNmGradThreads = 256, kNmGradTile = 2048;
__global__ void __launch_bounds__(kNmGradThreads)
nm_distance_grad_kernel(int na, int nb, const float *__restrict__ A_all, const float *__restrict__ B_all,
                        const float *__restrict__ ga_all, const int *__restrict__ idxa_all,
                        const float *__restrict__ gb_all, const int *__restrict__ idxb_all, float *__restrict__ grad_all) {
  __shared__ int s_idx[kNmGradTile];
  __shared__ float4 s_b[kNmGradTile];         // (B_k, 2 g_b[k])
  const int cloud = blockIdx.y;
  const float *A = A_all + (size_t)cloud * na * 3, *B = B_all + (size_t)cloud * nb * 3;
  const int j = blockIdx.x * kNmGradThreads + threadIdx.x;
  const bool on = j < na;
  const int jj = on ? j : na - 1;
  const float ax = __ldg(A + jj * 3), ay = __ldg(A + jj * 3 + 1), az = __ldg(A + jj * 3 + 2);
  float gx, gy, gz;
  {
    const int j2 = __ldg(idxa_all + (size_t)cloud * na + jj);
    const float g = __ldg(ga_all + (size_t)cloud * na + jj) * 2.f;
    gx = g * (ax - __ldg(B + j2 * 3)); gy = g * (ay - __ldg(B + j2 * 3 + 1)); gz = g * (az - __ldg(B + j2 * 3 + 2));
  }
  for (int k0 = 0; k0 < nb; k0 += kNmGradTile) {
    const int cnt = min(kNmGradTile, nb - k0);
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += kNmGradThreads) {
      const int k = k0 + i;
      s_idx[i] = __ldg(idxb_all + (size_t)cloud * nb + k);
      s_b[i] = make_float4(__ldg(B + k * 3), __ldg(B + k * 3 + 1), __ldg(B + k * 3 + 2),
                           __ldg(gb_all + (size_t)cloud * nb + k) * 2.f);
    }
    __syncthreads();
    for (int i = 0; i < cnt; ++i) {
      if (s_idx[i] == j) {
        const float4 p = s_b[i];
        gx += -(p.w * (p.x - ax)); gy += -(p.w * (p.y - ay)); gz += -(p.w * (p.z - az));
      }
    }
  }
  if (on) {
    float *o = grad_all + ((size_t)cloud * na + j) * 3;
    o[0] = gx; o[1] = gy; o[2] = gz;
  }
}
}  // namespace
}  // namespace pdr

extern "C" int pdr_nm_distance_grad(int b, int n, int m, const float *xyz1, const float *xyz2, const float *grad_dist1,
                                    const int *idx1, const float *grad_dist2, const int *idx2, float *grad_xyz1,
                                    float *grad_xyz2, void *stream) {
  PDR_REQUIRE(b >= 0 && n >= 1 && m >= 1 && b <= 65535, "nm_distance_grad: bad sizes b=%d n=%d m=%d", b, n, m);
  if (b == 0) return PDR_OK;
  PDR_REQUIRE(xyz1 && xyz2 && grad_dist1 && idx1 && grad_dist2 && idx2 && grad_xyz1 && grad_xyz2, "nm_distance_grad: null pointer");
  nm_distance_grad_kernel<<<dim3(ceil_div(n, kNmGradThreads), b), kNmGradThreads, 0, (cudaStream_t)stream>>>(
      n, m, xyz1, xyz2, grad_dist1, idx1, grad_dist2, idx2, grad_xyz1);
  int rc = check_launch("nm_distance_grad_kernel<1>");
  if (rc) return rc;
  nm_distance_grad_kernel<<<dim3(ceil_div(m, kNmGradThreads), b), kNmGradThreads, 0, (cudaStream_t)stream>>>(
      m, n, xyz2, xyz1, grad_dist2, idx2, grad_dist1, idx1, grad_xyz2);
  return check_launch("nm_distance_grad_kernel<2>");
}

extern "C" int pdr_nm_distance(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist1,
                               int *idx1, float *dist2, int *idx2, void *stream) {
  PDR_REQUIRE(b >= 0 && n >= 1 && m >= 1 && b <= 65535, "nm_distance: bad sizes b=%d n=%d m=%d", b, n, m);
  if (b == 0) return PDR_OK;
  PDR_REQUIRE(xyz1 && xyz2 && dist1 && idx1 && dist2 && idx2, "nm_distance: null pointer");
  dim3 grid(ceil_div(n > m ? n : m, kNmThreads * kNmQ), 2, b);
  nm_distance_kernel<<<grid, kNmThreads, 0, (cudaStream_t)stream>>>(n, m, xyz1, xyz2, dist1, idx1, dist2, idx2);
  return check_launch("nm_distance_kernel");
}
