// Earth Mover's Distance (auction-style approximate matching) for sm_100a.
//
// Replaces approxmatch / matchcost / matchcostgrad{1,2} (reference: PytorchEMD/cuda/emd_kernel.cu:29-161,
// 204-246, 290-358; Python side pointnet2/emd.py:7-21).  Algorithm unchanged: 10 temperature levels
// (-4^7 ... -4^-1, 0), three O(n*m) passes per level.
//
// Mapping differences:
//   reference                                       here
//   ---------                                       ----
//   <<<32,512>>> on the legacy default stream:      one thread-block CLUSTER per cloud (1/2/4/8 CTAs
//   32 CTAs total, clouds serialised per CTA        picked so the grid covers >= 2 waves of 148 SMs);
//                                                   rows are split across the cluster, the four
//                                                   remain/ratio vectors are exchanged through L2 and
//                                                   ordered by barrier.cluster (release/acquire)
//   tile re-staged for every 512-row group          each thread owns up to 8 rows in registers, a tile
//                                                   of the other cloud is staged once per pass
//   match (b*m*n fp32) read-modify-written 10x,     pdr_emd_cost: `match` never exists -- the cost
//   then re-read by matchcost (~88 B / pair)        sum(d^2 * w) is accumulated in the third pass
//                                                   (12(n+m) B / cloud of HBM traffic);
//                                                   pdr_emd_approxmatch: first level stores, later
//                                                   levels accumulate (no memset, one read fewer)
//   __expf with the denormal fix-up sequence        ex2.approx.ftz on a pre-scaled level (the values
//   (FSETP + 2 predicated FMUL per pair)            that differ are < 2^-126 and vanish against the
//                                                   1e-9 regulariser); -DPDR_EMD_EXACT_EXPF restores it
// Per-row accumulation order (sequential over the other cloud's index) is the reference's.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace pdr {
namespace {

constexpr int kEmdThreads = 256;
constexpr int kEmdTile = 1024;
constexpr int kEmdMaxCluster = 8;
constexpr int kEmdSlots = 8;  // per-cloud partial-cost slots in the workspace

__device__ __forceinline__ float emd_exp(float level_scaled, float d) {
#ifdef PDR_EMD_EXACT_EXPF
  return __expf(level_scaled * d);  // level_scaled == level
#else
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(level_scaled * d));  // level_scaled == level*log2(e)
  return r;
#endif
}

// Packed fp32 pairs (add/mul/fma.rn.f32x2, sm_100): one instruction does the work of two rows.  The kernel is bound by FP32
// ISSUE, not by the MUFU pipe (8 FP32 instructions against one ex2 per pair evaluation; profiles/r01 ncu: issue 68 %, xu
// 41.6 %), so the rows a thread owns are processed two at a time.  Every operation is the .rn form of the scalar one it
// replaces (p - x is p + (-x)): results are bit-identical to the scalar code.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// squared distances of one staged point to two rows (held negated), in dist2_ref's rounding sequence, and e = 2^(lvl d)
__device__ __forceinline__ void pair_d_e(const float4 &p, f32x2 nx, f32x2 ny, f32x2 nz, f32x2 lvl2, f32x2 &d, f32x2 &e) {
  const f32x2 dx = add2(pk2(p.x, p.x), nx), dy = add2(pk2(p.y, p.y), ny), dz = add2(pk2(p.z, p.z), nz);
  d = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
  float a0, a1;
  upk2(mul2(lvl2, d), a0, a1);
#ifdef PDR_EMD_EXACT_EXPF
  e = pk2(__expf(a0), __expf(a1));
#else
  float e0, e1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a1));
  e = pk2(e0, e1);
#endif
}

__device__ __forceinline__ void sync_all(int cs) {
  if (cs > 1) cg::this_cluster().sync();
  else __syncthreads();
}

// Stage tile [p0, p0+cnt) of a cloud with its per-point weight from the L2-resident workspace.
__device__ __forceinline__ void stage(float4 *s, const float *__restrict__ pts, const float *w, int p0,
                                      int cnt) {
  for (int i = threadIdx.x; i < cnt; i += kEmdThreads) {
    const float *p = pts + (size_t)(p0 + i) * 3;
    s[i] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldcg(w + p0 + i));
  }
}

template <int R, bool WRITE_MATCH, bool ACC_COST>
__global__ void __launch_bounds__(kEmdThreads)
emd_kernel(int n, int m, int cs, const float *__restrict__ xyz1_all, const float *__restrict__ xyz2_all,
           float *__restrict__ match_all, float *temp_all, float *__restrict__ cost_out) {
  __shared__ float4 tile[kEmdTile];
  __shared__ float tile_w1[kEmdTile];     // fused sweep: remainR of the staged points (the weight of the next level's pass 1)
  __shared__ float s_cost[kEmdThreads / 32];
  const int cloud = blockIdx.x / cs, rank = blockIdx.x % cs, tid = threadIdx.x;
  const float *xyz1 = xyz1_all + (size_t)cloud * n * 3;
  const float *xyz2 = xyz2_all + (size_t)cloud * m * 3;
  float *match = WRITE_MATCH ? match_all + (size_t)cloud * n * m : nullptr;
  float *remainL = temp_all + (size_t)cloud * (2 * (size_t)(n + m) + kEmdSlots);
  float *remainR = remainL + n, *ratioL = remainR + m, *ratioR = ratioL + n, *slots = ratioR + m;

  const int rp1 = (n + cs - 1) / cs, rp2 = (m + cs - 1) / cs;
  const int kb = min(n, rank * rp1), ke = min(n, kb + rp1);
  const int lb = min(m, rank * rp2), le = min(m, lb + rp2);

  float multiL, multiR;  // emd_kernel.cu:31-38 (integer division)
  if (n >= m) { multiL = 1.f; multiR = (float)(n / m); } else { multiL = (float)(m / n); multiR = 1.f; }
  for (int k = kb + tid; k < ke; k += kEmdThreads) __stcg(remainL + k, multiL);
  for (int l = lb + tid; l < le; l += kEmdThreads) __stcg(remainR + l, multiR);
  sync_all(cs);

  float cost_acc = 0.f;
  for (int j = 7; j >= -2; --j) {
    float level = -powf(4.0f, (float)j);
    if (j == -2) level = 0.f;
#ifndef PDR_EMD_EXACT_EXPF
    level *= 1.4426950408889634f;
#endif
    // ---- pass 1: ratioL[k] = remainL[k] / (1e-9 + sum_l e(k,l) * remainR[l]) ---------------------
    // (first level only: for the later levels it rides in the previous level's pass 3 -- one sweep over the same pairs, one
    //  distance evaluation for two exponentials; same operations in the same order per accumulator, bit-identical)
    if (j == 7)
    for (int g = kb; g < ke; g += kEmdThreads * R) {
      float x[R], y[R], z[R], acc[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int k = min(g + r * kEmdThreads + tid, n - 1);
        x[r] = __ldg(xyz1 + k * 3); y[r] = __ldg(xyz1 + k * 3 + 1); z[r] = __ldg(xyz1 + k * 3 + 2);
        acc[r] = 1e-9f;
      }
      for (int l0 = 0; l0 < m; l0 += kEmdTile) {
        const int cnt = min(kEmdTile, m - l0);
        __syncthreads();
        stage(tile, xyz2, remainR, l0, cnt);
        __syncthreads();
        if constexpr (R >= 2) {
          f32x2 nx[R / 2], ny[R / 2], nz[R / 2], ac[R / 2];
#pragma unroll
          for (int q = 0; q < R / 2; ++q) {
            nx[q] = pk2(-x[2 * q], -x[2 * q + 1]); ny[q] = pk2(-y[2 * q], -y[2 * q + 1]); nz[q] = pk2(-z[2 * q], -z[2 * q + 1]);
            ac[q] = pk2(acc[2 * q], acc[2 * q + 1]);
          }
          const f32x2 lvl2 = pk2(level, level);
#pragma unroll 2
          for (int l = 0; l < cnt; ++l) {
            const float4 p = tile[l];
            const f32x2 pw = pk2(p.w, p.w);
#pragma unroll
            for (int q = 0; q < R / 2; ++q) {
              f32x2 d, e;
              pair_d_e(p, nx[q], ny[q], nz[q], lvl2, d, e);
              ac[q] = fma2(e, pw, ac[q]);
            }
          }
#pragma unroll
          for (int q = 0; q < R / 2; ++q) upk2(ac[q], acc[2 * q], acc[2 * q + 1]);
        } else {
#pragma unroll 2
          for (int l = 0; l < cnt; ++l) {
            const float4 p = tile[l];
#pragma unroll
            for (int r = 0; r < R; ++r) {
              const float d = dist2_ref(__fsub_rn(p.x, x[r]), __fsub_rn(p.y, y[r]), __fsub_rn(p.z, z[r]));
              acc[r] = __fmaf_rn(emd_exp(level, d), p.w, acc[r]);
            }
          }
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int k = g + r * kEmdThreads + tid;
        if (k < ke) __stcg(ratioL + k, __ldcg(remainL + k) / acc[r]);
      }
    }
    if (j == 7) sync_all(cs);
    // ---- pass 2: columns.  sumr = remainR[l] * sum_k e * ratioL[k] ---------------------------------
    for (int g = lb; g < le; g += kEmdThreads * R) {
      float x[R], y[R], z[R], acc[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int l = min(g + r * kEmdThreads + tid, m - 1);
        x[r] = __ldg(xyz2 + l * 3); y[r] = __ldg(xyz2 + l * 3 + 1); z[r] = __ldg(xyz2 + l * 3 + 2);
        acc[r] = 0.f;
      }
      for (int k0 = 0; k0 < n; k0 += kEmdTile) {
        const int cnt = min(kEmdTile, n - k0);
        __syncthreads();
        stage(tile, xyz1, ratioL, k0, cnt);
        __syncthreads();
        if constexpr (R >= 2) {
          // ((x - p)^2 == (p - x)^2 exactly: the same packed form as pass 1)
          f32x2 nx[R / 2], ny[R / 2], nz[R / 2], ac[R / 2];
#pragma unroll
          for (int q = 0; q < R / 2; ++q) {
            nx[q] = pk2(-x[2 * q], -x[2 * q + 1]); ny[q] = pk2(-y[2 * q], -y[2 * q + 1]); nz[q] = pk2(-z[2 * q], -z[2 * q + 1]);
            ac[q] = pk2(acc[2 * q], acc[2 * q + 1]);
          }
          const f32x2 lvl2 = pk2(level, level);
#pragma unroll 2
          for (int k = 0; k < cnt; ++k) {
            const float4 p = tile[k];
            const f32x2 pw = pk2(p.w, p.w);
#pragma unroll
            for (int q = 0; q < R / 2; ++q) {
              f32x2 d, e;
              pair_d_e(p, nx[q], ny[q], nz[q], lvl2, d, e);
              ac[q] = fma2(e, pw, ac[q]);
            }
          }
#pragma unroll
          for (int q = 0; q < R / 2; ++q) upk2(ac[q], acc[2 * q], acc[2 * q + 1]);
        } else {
#pragma unroll 2
          for (int k = 0; k < cnt; ++k) {
            const float4 p = tile[k];
#pragma unroll
            for (int r = 0; r < R; ++r) {
              const float d = dist2_ref(__fsub_rn(x[r], p.x), __fsub_rn(y[r], p.y), __fsub_rn(z[r], p.z));
              acc[r] = __fmaf_rn(emd_exp(level, d), p.w, acc[r]);
            }
          }
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int l = g + r * kEmdThreads + tid;
        if (l < le) {
          const float rr = __ldcg(remainR + l);
          const float sumr = acc[r] * rr;
          const float consumption = fminf(rr / (sumr + 1e-9f), 1.0f);
          __stcg(ratioR + l, consumption * rr);
          __stcg(remainR + l, fmaxf(0.0f, rr - sumr));
        }
      }
    }
    sync_all(cs);
    // ---- pass 3: w = e * ratioL[k] * ratioR[l]; match += w; remainL[k] -= sum_l w; cost += d^2 w --
    //      + pass 1 of the NEXT level in the same sweep: acc1[k] = 1e-9 + sum_l e'(k,l) * remainR[l] (remainR as pass 2 just
    //      left it), ratioL[k] = remainL_new[k] / acc1[k]
    const bool fuse_next = j > -2;
    float level1 = (j - 1 == -2) ? 0.f : -powf(4.0f, (float)(j - 1));
#ifndef PDR_EMD_EXACT_EXPF
    level1 *= 1.4426950408889634f;
#endif
    for (int g = kb; g < ke; g += kEmdThreads * R) {
      float x[R], y[R], z[R], acc[R], rl[R], acc1[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int kr = g + r * kEmdThreads + tid;
        const int k = min(kr, n - 1);
        x[r] = __ldg(xyz1 + k * 3); y[r] = __ldg(xyz1 + k * 3 + 1); z[r] = __ldg(xyz1 + k * 3 + 2);
        rl[r] = kr < ke ? __ldcg(ratioL + k) : 0.f;  // rows this CTA does not own contribute w = 0
        acc[r] = 0.f;
        acc1[r] = 1e-9f;
      }
      for (int l0 = 0; l0 < m; l0 += kEmdTile) {
        const int cnt = min(kEmdTile, m - l0);
        __syncthreads();
        stage(tile, xyz2, ratioR, l0, cnt);
        if (fuse_next)
          for (int i = tid; i < cnt; i += kEmdThreads) tile_w1[i] = __ldcg(remainR + l0 + i);
        __syncthreads();
        if constexpr (R >= 2) {
          f32x2 nx[R / 2], ny[R / 2], nz[R / 2], ac[R / 2], rl2[R / 2], a1[R / 2];
#pragma unroll
          for (int q = 0; q < R / 2; ++q) {
            nx[q] = pk2(-x[2 * q], -x[2 * q + 1]); ny[q] = pk2(-y[2 * q], -y[2 * q + 1]); nz[q] = pk2(-z[2 * q], -z[2 * q + 1]);
            ac[q] = pk2(acc[2 * q], acc[2 * q + 1]); rl2[q] = pk2(rl[2 * q], rl[2 * q + 1]);
            a1[q] = pk2(acc1[2 * q], acc1[2 * q + 1]);
          }
          const f32x2 lvl2 = pk2(level, level), lvl1 = pk2(level1, level1);
#pragma unroll 2
          for (int l = 0; l < cnt; ++l) {
            const float4 p = tile[l];
            const f32x2 pw = pk2(p.w, p.w);
            const float w1s = fuse_next ? tile_w1[l] : 0.f;
            const f32x2 pw1 = pk2(w1s, w1s);
#pragma unroll
            for (int q = 0; q < R / 2; ++q) {
              f32x2 d, e;
              pair_d_e(p, nx[q], ny[q], nz[q], lvl2, d, e);
              const f32x2 w = mul2(mul2(e, rl2[q]), pw);
              ac[q] = add2(ac[q], w);
              if (fuse_next) {
                float b0, b1, e0, e1;
                upk2(mul2(lvl1, d), b0, b1);
#ifdef PDR_EMD_EXACT_EXPF
                e0 = __expf(b0); e1 = __expf(b1);
#else
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(b0));
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(b1));
#endif
                a1[q] = fma2(pk2(e0, e1), pw1, a1[q]);
              }
              if (ACC_COST || WRITE_MATCH) {
                float d0, d1, w0, w1;
                upk2(d, d0, d1); upk2(w, w0, w1);
                if (ACC_COST) { cost_acc = __fmaf_rn(d0, w0, cost_acc); cost_acc = __fmaf_rn(d1, w1, cost_acc); }   // row order kept
                if (WRITE_MATCH) {
                  const int k = g + 2 * q * kEmdThreads + tid;
                  if (k < ke) { float *mp = match + (size_t)(l0 + l) * n + k; *mp = (j == 7) ? w0 : *mp + w0; }
                  if (k + kEmdThreads < ke) {
                    float *mp = match + (size_t)(l0 + l) * n + k + kEmdThreads;
                    *mp = (j == 7) ? w1 : *mp + w1;
                  }
                }
              }
            }
          }
#pragma unroll
          for (int q = 0; q < R / 2; ++q) { upk2(ac[q], acc[2 * q], acc[2 * q + 1]); upk2(a1[q], acc1[2 * q], acc1[2 * q + 1]); }
        } else {
#pragma unroll 2
          for (int l = 0; l < cnt; ++l) {
            const float4 p = tile[l];
            const float w1s = fuse_next ? tile_w1[l] : 0.f;
#pragma unroll
            for (int r = 0; r < R; ++r) {
              const float d = dist2_ref(__fsub_rn(p.x, x[r]), __fsub_rn(p.y, y[r]), __fsub_rn(p.z, z[r]));
              const float w = emd_exp(level, d) * rl[r] * p.w;
              acc[r] += w;
              if (fuse_next) acc1[r] = __fmaf_rn(emd_exp(level1, d), w1s, acc1[r]);
              if (ACC_COST) cost_acc = __fmaf_rn(d, w, cost_acc);
              if (WRITE_MATCH) {
                const int k = g + r * kEmdThreads + tid;
                if (k < ke) {
                  float *mp = match + (size_t)(l0 + l) * n + k;
                  *mp = (j == 7) ? w : *mp + w;
                }
              }
            }
          }
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int k = g + r * kEmdThreads + tid;
        if (k < ke) {
          const float left = fmaxf(0.0f, __ldcg(remainL + k) - acc[r]);
          __stcg(remainL + k, left);
          if (fuse_next) __stcg(ratioL + k, left / acc1[r]);
        }
      }
    }
    sync_all(cs);
  }
  if (ACC_COST) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cost_acc += __shfl_xor_sync(0xffffffffu, cost_acc, o);
    if ((tid & 31) == 0) s_cost[tid >> 5] = cost_acc;
    __syncthreads();
    if (tid == 0) {
      float a = 0.f;
      for (int w = 0; w < kEmdThreads / 32; ++w) a += s_cost[w];
      __stcg(slots + rank, a);
    }
    sync_all(cs);
    if (rank == 0 && tid == 0) {
      float a = 0.f;
      for (int r = 0; r < cs; ++r) a += __ldcg(slots + r);
      cost_out[cloud] = a;
    }
  }
}

// matchcost forward: cost[b] = sum_{k,l} d^2(k,l) * match[l][k]   (emd_kernel.cu:204-246)
__global__ void __launch_bounds__(256)
matchcost_kernel(int n, int m, const float *__restrict__ xyz1_all, const float *__restrict__ xyz2_all,
                 const float *__restrict__ match_all, float *__restrict__ partial, int nblk) {
  __shared__ float4 tile[kEmdTile];
  __shared__ float s_red[8];
  const int cloud = blockIdx.y;
  const float *xyz1 = xyz1_all + (size_t)cloud * n * 3;
  const float *xyz2 = xyz2_all + (size_t)cloud * m * 3;
  const float *match = match_all + (size_t)cloud * n * m;
  const int k = blockIdx.x * 256 + threadIdx.x;
  const bool on = k < n;
  const int kk = on ? k : n - 1;
  const float x1 = __ldg(xyz1 + kk * 3), y1 = __ldg(xyz1 + kk * 3 + 1), z1 = __ldg(xyz1 + kk * 3 + 2);
  float sub = 0.f;
  for (int l0 = 0; l0 < m; l0 += kEmdTile) {
    const int cnt = min(kEmdTile, m - l0);
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += 256) {
      const float *p = xyz2 + (size_t)(l0 + i) * 3;
      tile[i] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0.f);
    }
    __syncthreads();
    if (on) {
#pragma unroll 4
      for (int l = 0; l < cnt; ++l) {
        const float4 p = tile[l];
        const float d = dist2_ref(__fsub_rn(p.x, x1), __fsub_rn(p.y, y1), __fsub_rn(p.z, z1));
        sub = __fmaf_rn(d, __ldg(match + (size_t)(l0 + l) * n + k), sub);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sub += __shfl_xor_sync(0xffffffffu, sub, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = sub;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int w = 0; w < 8; ++w) a += s_red[w];
    partial[(size_t)cloud * nblk + blockIdx.x] = a;
  }
}

__global__ void sum_partials_kernel(int b, int nblk, const float *__restrict__ partial, float *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b) return;
  float a = 0.f;
  for (int k = 0; k < nblk; ++k) a += partial[(size_t)i * nblk + k];
  out[i] = a;
}

// matchcost backward (emd_kernel.cu:290-358):
//   grad1[k] = grad_cost * 2 * sum_l match[l][k] * (xyz1[k] - xyz2[l])
//   grad2[l] = grad_cost * 2 * sum_k match[l][k] * (xyz2[l] - xyz1[k])
template <bool FOR_XYZ1>
__global__ void __launch_bounds__(256)
matchcost_grad_kernel(int n, int m, const float *__restrict__ grad_cost, const float *__restrict__ xyz1_all,
                      const float *__restrict__ xyz2_all, const float *__restrict__ match_all,
                      float *__restrict__ grad_all) {
  __shared__ float4 tile[kEmdTile];
  const int cloud = blockIdx.y;
  const int na = FOR_XYZ1 ? n : m, nb = FOR_XYZ1 ? m : n;
  const float *A = (FOR_XYZ1 ? xyz1_all : xyz2_all) + (size_t)cloud * na * 3;
  const float *B = (FOR_XYZ1 ? xyz2_all : xyz1_all) + (size_t)cloud * nb * 3;
  const float *match = match_all + (size_t)cloud * n * m;
  const int a = blockIdx.x * 256 + threadIdx.x;
  const bool on = a < na;
  const int aa = on ? a : na - 1;
  const float ax = __ldg(A + aa * 3), ay = __ldg(A + aa * 3 + 1), az = __ldg(A + aa * 3 + 2);
  float gx = 0.f, gy = 0.f, gz = 0.f;
  for (int b0 = 0; b0 < nb; b0 += kEmdTile) {
    const int cnt = min(kEmdTile, nb - b0);
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += 256) {
      const float *p = B + (size_t)(b0 + i) * 3;
      tile[i] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0.f);
    }
    __syncthreads();
    if (on) {
      for (int t = 0; t < cnt; ++t) {
        const float4 p = tile[t];
        const float w = FOR_XYZ1 ? __ldg(match + (size_t)(b0 + t) * n + a)    // match[l][k], k = a
                                 : __ldg(match + (size_t)a * n + (b0 + t));    // match[l][k], l = a
        gx += (ax - p.x) * w; gy += (ay - p.y) * w; gz += (az - p.z) * w;
      }
    }
  }
  if (on) {
    const float g = 2.f * __ldg(grad_cost + cloud);
    float *o = grad_all + ((size_t)cloud * na + a) * 3;
    o[0] = gx * g; o[1] = gy * g; o[2] = gz * g;
  }
}

int pick_cluster(int b, int n, int m) {
  const int big = n > m ? n : m;
  static const int forced = getenv("PDR_EMD_CLUSTER") ? atoi(getenv("PDR_EMD_CLUSTER")) : 0;     // A/B: 1, 2, 4 or 8
  if (forced == 1 || forced == 2 || forced == 4 || forced == 8) {
    int cs = forced;
    while (cs > 1 && big / cs < 128) cs /= 2;
    return cs;
  }
  int cs = 1;
  while (cs < kEmdMaxCluster && (long long)b * cs < 2 * kNumSMs && big / (cs * 2) >= 128) cs *= 2;
  return cs;
}

template <bool WRITE_MATCH, bool ACC_COST>
int launch_emd(int b, int n, int m, const float *xyz1, const float *xyz2, float *match, float *temp,
               float *cost, cudaStream_t stream) {
  const int cs = pick_cluster(b, n, m);
  const int big = n > m ? n : m;
  const int rp = ceil_div(big, cs);
  const int R = ceil_div(rp, kEmdThreads);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)b * cs);
  cfg.blockDim = dim3(kEmdThreads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e;
#define PDR_EMD_LAUNCH(RR) \
  e = cudaLaunchKernelEx(&cfg, emd_kernel<RR, WRITE_MATCH, ACC_COST>, n, m, cs, xyz1, xyz2, match, temp, cost)
  if (R <= 1) PDR_EMD_LAUNCH(1);
  else if (R <= 2) PDR_EMD_LAUNCH(2);
  else if (R <= 4) PDR_EMD_LAUNCH(4);
  else PDR_EMD_LAUNCH(8);
#undef PDR_EMD_LAUNCH
  if (e != cudaSuccess) {
    set_error("emd_kernel launch (cluster %d): %s", cs, cudaGetErrorString(e));
    return PDR_ERR_CUDA;
  }
  return check_launch("emd_kernel");
}

int emd_check(const char *op, int b, int n, int m, const void *temp, size_t temp_bytes) {
  PDR_REQUIRE(b >= 0 && n >= 1 && m >= 1, "%s: bad sizes b=%d n=%d m=%d", op, b, n, m);
  PDR_REQUIRE((long long)b * kEmdMaxCluster < (1ll << 31), "%s: b too large", op);
  const size_t need = pdr_emd_workspace_bytes(b, n, m);
  if (b > 0 && (!temp || temp_bytes < need)) {
    set_error("%s: workspace %zu B < %zu B", op, temp_bytes, need);
    return PDR_ERR_WORKSPACE;
  }
  return PDR_OK;
}

}  // namespace
}  // namespace pdr

using namespace pdr;

extern "C" size_t pdr_emd_workspace_bytes(int b, int n, int m) {
  if (b <= 0 || n <= 0 || m <= 0) return 0;
  const size_t per_cloud = 2 * ((size_t)n + m) + kEmdSlots;
  const size_t mc = (size_t)ceil_div(n, 256);  // matchcost partials share the same scratch
  return (size_t)b * (per_cloud > mc ? per_cloud : mc) * sizeof(float);
}

extern "C" int pdr_emd_approxmatch(int b, int n, int m, const float *xyz1, const float *xyz2, float *match,
                                   void *temp, size_t temp_bytes, void *stream) {
  int rc = emd_check("emd_approxmatch", b, n, m, temp, temp_bytes);
  if (rc) return rc;
  if (b == 0) return PDR_OK;
  PDR_REQUIRE(xyz1 && xyz2 && match, "emd_approxmatch: null pointer");
  return launch_emd<true, false>(b, n, m, xyz1, xyz2, match, (float *)temp, nullptr, (cudaStream_t)stream);
}

extern "C" int pdr_emd_cost(int b, int n, int m, const float *xyz1, const float *xyz2, float *cost, void *temp,
                            size_t temp_bytes, void *stream) {
  int rc = emd_check("emd_cost", b, n, m, temp, temp_bytes);
  if (rc) return rc;
  if (b == 0) return PDR_OK;
  PDR_REQUIRE(xyz1 && xyz2 && cost, "emd_cost: null pointer");
  return launch_emd<false, true>(b, n, m, xyz1, xyz2, nullptr, (float *)temp, cost, (cudaStream_t)stream);
}

extern "C" int pdr_emd_matchcost(int b, int n, int m, const float *xyz1, const float *xyz2, const float *match,
                                 float *cost, void *temp, size_t temp_bytes, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = emd_check("emd_matchcost", b, n, m, temp, temp_bytes);
  if (rc) return rc;
  PDR_REQUIRE(b <= 65535, "emd_matchcost: b > 65535");
  if (b == 0) return PDR_OK;
  PDR_REQUIRE(xyz1 && xyz2 && match && cost, "emd_matchcost: null pointer");
  const int nblk = ceil_div(n, 256);
  float *partial = (float *)temp;  // (b, nblk) per-CTA partial sums, added in index order below
  matchcost_kernel<<<dim3(nblk, b), 256, 0, stream>>>(n, m, xyz1, xyz2, match, partial, nblk);
  rc = check_launch("matchcost_kernel");
  if (rc) return rc;
  sum_partials_kernel<<<ceil_div(b, 128), 128, 0, stream>>>(b, nblk, partial, cost);
  return check_launch("sum_partials_kernel");
}

extern "C" int pdr_emd_matchcost_backward(int b, int n, int m, const float *grad_cost, const float *xyz1,
                                          const float *xyz2, const float *match, float *grad1, float *grad2,
                                          void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  PDR_REQUIRE(b >= 0 && n >= 1 && m >= 1 && b <= 65535, "emd_matchcost_backward: bad sizes");
  if (b == 0) return PDR_OK;
  PDR_REQUIRE(grad_cost && xyz1 && xyz2 && match && grad1 && grad2, "emd_matchcost_backward: null pointer");
  matchcost_grad_kernel<true><<<dim3(ceil_div(n, 256), b), 256, 0, stream>>>(n, m, grad_cost, xyz1, xyz2, match, grad1);
  int rc = check_launch("matchcost_grad_kernel<1>");
  if (rc) return rc;
  matchcost_grad_kernel<false><<<dim3(ceil_div(m, 256), b), 256, 0, stream>>>(n, m, grad_cost, xyz1, xyz2, match, grad2);
  return check_launch("matchcost_grad_kernel<2>");
}
