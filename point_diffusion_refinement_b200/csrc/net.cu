// Fused denoiser primitives for sm_100a (channels-last activations).
//
// The reference evaluates one warm eps_theta step as ~1000 ATen/cuDNN launches: every 1x1 Conv2d, every
// GroupNorm, ReLU, cat, expand, softmax materialises a (B, C, npoint, nsample) tensor
// (pointnet2_ops/pointnet2_modules.py:69-174, attention.py:70-96).  Here the step is a short chain of
//   gather -> [GEMM with GroupNorm/ReLU/embedding/residual folded into the A-operand load and the
//              statistics of the NEXT GroupNorm folded into the epilogue] -> attention pooling
// so that every activation is written once (raw conv output) and read once.
//
// This file: the fp32 SIMT GEMM (validation mode and small-problem path), GroupNorm finalisation,
// attention pooling, grouping/gather kernels.  The tcgen05 (TF32) GEMM lives in gemm_tc.cu.
#include <string.h>

#include "common.cuh"

namespace pdr {

int launch_gemm_tf32(const PdrGemmArgs &a, cudaStream_t stream);  // gemm_tc.cu

namespace {

// ------------------------------------------------------------------------------------------------
// SIMT GEMM: C[128 x (16*TN)] per CTA, 256 threads, 8 x TN outputs per thread, K stepped by 16.
// ------------------------------------------------------------------------------------------------
constexpr int kTileM = 128;
constexpr int kTileK = 16;
constexpr int kGemmThreads = 256;

__device__ __forceinline__ float pro_apply(int mode, float x, float sc, float sh) {
  if (mode == PDR_PRO_GN_RELU) return fmaxf(fmaf(x, sc, sh), 0.f);
  if (mode == PDR_PRO_RELU_GN) return fmaf(fmaxf(x, 0.f), sc, sh);
  return x;
}

template <int TN>
__global__ void __launch_bounds__(kGemmThreads)
gemm_simt_kernel(const PdrGemmArgs a) {
  constexpr int kTileN = 16 * TN;
  __shared__ __align__(16) float As[kTileK][kTileM + 4];
  __shared__ __align__(16) float Bs[kTileK][kTileN + 4];
  __shared__ float s_red[kGemmThreads / 32][kTileN][4];

  const int tid = threadIdx.x;
  const int tiles_per_sample = (a.rows_per_sample + kTileM - 1) / kTileM;
  const int tile = blockIdx.y;
  const int b = tile / tiles_per_sample;
  const int r0 = (tile % tiles_per_sample) * kTileM;            // first row of the tile inside the sample
  const size_t row_base = (size_t)b * a.rows_per_sample + r0;   // global row of tile row 0
  const int rows_valid = min(kTileM, a.rows_per_sample - r0);
  const int n0 = blockIdx.x * kTileN;

  const int ty = tid >> 4, tx = tid & 15;
  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // global -> register staging (2 float4 of A, kTileN*16/4/256 float4 of W per thread)
  float4 ra[2];
  float4 rw[(kTileN * kTileK / 4 + kGemmThreads - 1) / kGemmThreads];
  constexpr int kWLoads = (kTileN * kTileK / 4 + kGemmThreads - 1) / kGemmThreads;

  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int f = tid + i * kGemmThreads;
      const int row = f >> 2, kq = (f & 3) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const int k = k0 + kq;
      if (row < rows_valid && k < a.K) {
        const size_t grow = row_base + row;
        v = *reinterpret_cast<const float4 *>(a.A + grow * a.lda + k);
        if (a.pro_mode != PDR_PRO_NONE) {
          const float4 s = *reinterpret_cast<const float4 *>(a.sc + (size_t)b * a.ld_scsh + k);
          const float4 h = *reinterpret_cast<const float4 *>(a.sh + (size_t)b * a.ld_scsh + k);
          v.x = pro_apply(a.pro_mode, v.x, s.x, h.x); v.y = pro_apply(a.pro_mode, v.y, s.y, h.y);
          v.z = pro_apply(a.pro_mode, v.z, s.z, h.z); v.w = pro_apply(a.pro_mode, v.w, s.w, h.w);
        }
        if (a.add) {
          const float4 e = *reinterpret_cast<const float4 *>(a.add + (size_t)b * a.ld_add + k);
          v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
        }
        if (a.R) {
          const float4 r = *reinterpret_cast<const float4 *>(a.R + grow * a.ldr + k);
          v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
        }
      }
      ra[i] = v;
    }
#pragma unroll
    for (int i = 0; i < kWLoads; ++i) {
      const int f = tid + i * kGemmThreads;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (f < kTileN * kTileK / 4) {
        const int n = f >> 2, kq = (f & 3) * 4;
        if (n0 + n < a.N && k0 + kq < a.K) v = *reinterpret_cast<const float4 *>(a.W + (size_t)(n0 + n) * a.ldw + k0 + kq);
      }
      rw[i] = v;
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int f = tid + i * kGemmThreads;
      const int row = f >> 2, kq = (f & 3) * 4;
      As[kq + 0][row] = ra[i].x; As[kq + 1][row] = ra[i].y; As[kq + 2][row] = ra[i].z; As[kq + 3][row] = ra[i].w;
    }
#pragma unroll
    for (int i = 0; i < kWLoads; ++i) {
      const int f = tid + i * kGemmThreads;
      if (f < kTileN * kTileK / 4) {
        const int n = f >> 2, kq = (f & 3) * 4;
        Bs[kq + 0][n] = rw[i].x; Bs[kq + 1][n] = rw[i].y; Bs[kq + 2][n] = rw[i].z; Bs[kq + 3][n] = rw[i].w;
      }
    }
  };

  load_tiles(0);
  for (int k0 = 0; k0 < a.K; k0 += kTileK) {
    __syncthreads();
    store_tiles();
    __syncthreads();
    if (k0 + kTileK < a.K) load_tiles(k0 + kTileK);
#pragma unroll
    for (int kk = 0; kk < kTileK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4 *>(&As[kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4 *>(&As[kk][ty * 8 + 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[TN];
#pragma unroll
      for (int j = 0; j < TN; ++j) bv[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }

  // ---- epilogue: bias, broadcast row-add, store, per-column statistics ------------------------------
  float st[TN][4];
#pragma unroll
  for (int j = 0; j < TN; ++j) st[j][0] = st[j][1] = st[j][2] = st[j][3] = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = ty * 8 + i;
    if (row >= rows_valid) continue;
    const size_t grow = row_base + row;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n < a.N) {
        float y = acc[i][j];
        if (a.bias) y += __ldg(a.bias + n);
        if (a.rowadd) y += __ldg(a.rowadd + (grow / a.rowadd_div) * a.ld_rowadd + n);
        a.C[grow * a.ldc + n] = y;
        const float r = fmaxf(y, 0.f);
        st[j][0] += y; st[j][1] = fmaf(y, y, st[j][1]); st[j][2] += r; st[j][3] = fmaf(r, r, st[j][3]);
      } else if (n < a.ldc_zero_to) {
        a.C[grow * a.ldc + n] = 0.f;
      }
    }
  }
  if (a.stats) {
    // threads with equal tx hold the same columns: lanes l and l^16 inside a warp, then 8 warps
#pragma unroll
    for (int j = 0; j < TN; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) st[j][q] += __shfl_xor_sync(0xffffffffu, st[j][q], 16);
    const int warp = tid >> 5, lane = tid & 31;
    if (lane < 16) {
#pragma unroll
      for (int j = 0; j < TN; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) s_red[warp][lane * TN + j][q] = st[j][q];
    }
    __syncthreads();
    for (int f = tid; f < kTileN * 4; f += kGemmThreads) {
      const int col = f >> 2, q = f & 3;
      if (n0 + col < a.N) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kGemmThreads / 32; ++w) s += s_red[w][col][q];
        a.stats[((size_t)tile * a.N + n0 + col) * 4 + q] = s;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// GroupNorm finalisation.  grid = (batch, group splits); each CTA owns a contiguous range of groups.
// The per-tile partial sums written by the GEMM epilogues are added in a FIXED order (warp w takes tiles
// w, w+8, ...; the 8 warp partials are then added in warp order), in double, so the result is
// deterministic; variance is E[x^2] - mean^2 (biased, like nn.GroupNorm).
// ------------------------------------------------------------------------------------------------
// Round-to-nearest (ties away) to TF32, applied by the kernels that PRODUCE the tables the tensor-core GEMMs read raw (gathered
// A operand, raw K tail, prologue-free dense A): cp.async copies them straight into the MMA stage and the tensor core
// TRUNCATES fp32 operands to 10 mantissa bits, so unrounded producers would put a toward-zero bias on every such operand.
__device__ __forceinline__ float rtf32(float x, int on) {
  return on ? __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u) : x;
}

constexpr int kGnThreads = 256;
constexpr int kGnSplit = 4;

// Up to two independent GroupNorms per launch (blockIdx.z): the engine finalises the normalisations that become ready
// together -- (first MLP layer, attention query|key) and (second MLP layer, attention scores) of a stage -- in one launch.
struct GnBatch { PdrGnArgs g[2]; };

__global__ void __launch_bounds__(kGnThreads)
gn_finalize_kernel(const GnBatch batch) {
  pdl_launch_dependents();      // (PDL, common.cuh) the next kernel may ramp up; nothing below touches global memory before
  pdl_wait();                   // every earlier kernel has completed
  extern __shared__ double s_tot[];                 // [channels in range][3] = weighted sum, sum of squares, count
  __shared__ double s_red[kGnThreads][2];
  const PdrGnArgs &a = batch.g[blockIdx.z];
  if ((int)blockIdx.x >= a.batch) return;
  const int b = blockIdx.x;
  const int cpg = a.gn_channels / a.groups;
  const int gpc = (a.groups + gridDim.y - 1) / gridDim.y;           // groups per CTA
  const int g_lo = min(a.groups, (int)blockIdx.y * gpc), g_hi = min(a.groups, g_lo + gpc);
  const int c_lo = g_lo * cpg;
  // the last split also writes the pass-through channels [gn_channels, channels)
  const int c_hi = (blockIdx.y == gridDim.y - 1) ? a.channels : g_hi * cpg;
  const int c_gn_hi = g_hi * cpg;                                   // channels that need statistics: [c_lo, c_gn_hi)

  // accumulate phase: thread = (tile slice, channel), channel fastest, so that all 256 threads read partials even
  // when this CTA owns only a handful of channels; slices are folded in a fixed order (deterministic)
  int src_off = 0;
  for (int s = 0; s < a.nsrc; ++s) {
    const PdrGnSource src = a.src[s];
    const int v_lo = max(c_lo, src_off), v_hi = min(c_gn_hi, src_off + src.ncols);   // virtual channels of this source
    for (int v0 = v_lo; v0 < v_hi; v0 += kGnThreads) {
      const int ncs = min(kGnThreads, v_hi - v0);
      int cw = 1;
      while (cw < ncs) cw <<= 1;                      // channels per slice row, power of two <= 256
      const int nsl = kGnThreads / cw;                // tile slices
      const int cl = threadIdx.x & (cw - 1), slice = threadIdx.x / cw;
      const int v = v0 + cl;
      double sum = 0.0, sq = 0.0;
      if (cl < ncs) {
        const float *p = src.stats + ((size_t)b * src.tiles_per_sample * src.ld_stats + src.col0 + (v - src_off)) * 4 +
                         (src.use_relu ? 2 : 0);
        // 8 loads in flight per thread (the loop is latency-bound: one L2 round trip per tile partial otherwise);
        // the partials are still folded in the same fixed order, so the result is unchanged bit for bit
        const size_t tstride = (size_t)src.ld_stats * 4;
        int t = slice;
        for (; t + 7 * nsl < src.tiles_per_sample; t += 8 * nsl) {
          float2 q[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) q[u] = *reinterpret_cast<const float2 *>(p + (size_t)(t + u * nsl) * tstride);
#pragma unroll
          for (int u = 0; u < 8; ++u) { sum += (double)q[u].x; sq += (double)q[u].y; }
        }
        for (; t < src.tiles_per_sample; t += nsl) {
          const float2 q = *reinterpret_cast<const float2 *>(p + (size_t)t * tstride);
          sum += (double)q.x;
          sq += (double)q.y;
        }
      }
      s_red[threadIdx.x][0] = sum;
      s_red[threadIdx.x][1] = sq;
      __syncthreads();
      if (threadIdx.x < ncs) {
        double ts = 0.0, tq = 0.0;
        for (int w = 0; w < nsl; ++w) { ts += s_red[w * cw + threadIdx.x][0]; tq += s_red[w * cw + threadIdx.x][1]; }
        s_tot[(v0 + threadIdx.x - c_lo) * 3 + 0] = (double)src.mult * ts;
        s_tot[(v0 + threadIdx.x - c_lo) * 3 + 1] = (double)src.mult * tq;
        s_tot[(v0 + threadIdx.x - c_lo) * 3 + 2] = (double)src.mult * (double)src.rows;
      }
      __syncthreads();
    }
    src_off += src.ncols;
  }
  for (int c = c_lo + threadIdx.x; c < c_hi; c += kGnThreads) {
    float sc = 1.f, sh = 0.f;  // MyGroupNorm passes the trailing C % G channels through (attention.py:17-23)
    if (c < a.gn_channels) {
      const int g = c / cpg;
      double sum = 0.0, sq = 0.0, n = 0.0;
      for (int cc = g * cpg; cc < (g + 1) * cpg; ++cc) {
        sum += s_tot[(cc - c_lo) * 3 + 0]; sq += s_tot[(cc - c_lo) * 3 + 1]; n += s_tot[(cc - c_lo) * 3 + 2];
      }
      const double mean = sum / n;
      double var = sq / n - mean * mean;
      if (var < 0.0) var = 0.0;
      const double rstd = 1.0 / sqrt(var + (double)a.eps);
      const double gsc = (double)__ldg(a.gamma + c) * rstd;
      sc = (float)gsc;
      sh = (float)((double)__ldg(a.beta + c) - mean * gsc);
    }
    int s = 0, off = 0;
    while (s + 1 < a.nsrc && c >= off + a.src[s].ncols) { off += a.src[s].ncols; ++s; }
    const int o = a.src[s].out_col0 + (c - off);
    a.sc[(size_t)b * a.ld_out + o] = sc;
    a.sh[(size_t)b * a.ld_out + o] = sh;
  }
}

// out[row, c] = pro(x)(+add)(+R) for c < C.  `out` may be a column slice of a wider matrix, so nothing
// outside [0, C) is touched (pad columns stay zero from allocation).
__global__ void __launch_bounds__(256)
affine_rows_kernel(int rows_per_sample, int C, const float *__restrict__ x, int ldx, int mode,
                   const float *__restrict__ sc, const float *__restrict__ sh, int ld_scsh,
                   const float *__restrict__ add, int ld_add, const float *__restrict__ R, int ldr,
                   float *__restrict__ out, int ldo, long long total, int round_tf32) {
  pdl_launch_dependents();      // (PDL, common.cuh) the next kernel may ramp up; nothing below touches global memory before
  pdl_wait();                   // every earlier kernel has completed
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const long long row = i / C;
  const int b = (int)(row / rows_per_sample);
  float v = x[row * ldx + c];
  if (mode != PDR_PRO_NONE) v = pro_apply(mode, v, __ldg(sc + (size_t)b * ld_scsh + c), __ldg(sh + (size_t)b * ld_scsh + c));
  if (add) v += __ldg(add + (size_t)b * ld_add + c);
  if (R) v += R[row * ldr + c];
  out[row * ldo + c] = rtf32(v, round_tf32);
}

// Attention pooling: one thread per (b, p, c); the K scores/values of a (p, c) are strided by ld, consecutive
// threads take consecutive channels (coalesced).  KT > 0: the K scores live in registers, S and V are read
// exactly once.
template <int KT>
__global__ void __launch_bounds__(256)
attention_pool_kernel(int P, int K, int C, const float *__restrict__ S, int lds, const float *__restrict__ V,
                      int ldv, const float *__restrict__ sc, const float *__restrict__ sh, int ld_scsh,
                      const int *__restrict__ counts, float *__restrict__ out, int ldo, long long total, int round_tf32) {
  pdl_launch_dependents();      // (PDL, common.cuh) the next kernel may ramp up; nothing below touches global memory before
  pdl_wait();                   // every earlier kernel has completed
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const long long bp = i / C;           // b*P + p
  const int b = (int)(bp / P);
  int cnt = K;
  if (counts) { cnt = __ldg(counts + bp); cnt = cnt < 1 ? 1 : cnt; }
  const float gs = __ldg(sc + (size_t)b * ld_scsh + c), gh = __ldg(sh + (size_t)b * ld_scsh + c);
  const float *s = S + (size_t)bp * K * lds + c;
  const float *v = V + (size_t)bp * K * ldv + c;
  float mx = -3.0e38f, den = 0.f, num = 0.f;
  if (KT > 0) {
    float sv[KT > 0 ? KT : 1];
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      sv[k] = k < cnt ? s[(size_t)k * lds] : -1e9f;   // scores*mask + (-1e9)*(1-mask), attention.py:88
      mx = fmaxf(mx, sv[k]);
    }
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      const float e = expf(sv[k] - mx);
      den += e;
      num = fmaf(e, fmaxf(fmaf(v[(size_t)k * ldv], gs, gh), 0.f), num);
    }
  } else {
    for (int k = 0; k < K; ++k) mx = fmaxf(mx, k < cnt ? s[(size_t)k * lds] : -1e9f);
    for (int k = 0; k < K; ++k) {
      const float e = expf((k < cnt ? s[(size_t)k * lds] : -1e9f) - mx);
      den += e;
      num = fmaf(e, fmaxf(fmaf(v[(size_t)k * ldv], gs, gh), 0.f), num);
    }
  }
  out[bp * ldo + c] = rtf32(num / den, round_tf32);
}

// Ball-query grouping.  A warp assembles 32 consecutive output rows: lane u first fetches the neighbour index,
// centre and validity flag of row u, then the warp walks the 32 * ldo output floats as ONE flat, contiguous span
// (lane = consecutive floats, every store instruction writes 128 contiguous bytes whatever the row width) and
// pulls the per-row values from lane u by shuffle.  All loads of an unrolled group are independent, so a lane has
// several gathers in flight.
constexpr int kGbRows = 32;    // rows per warp
__global__ void __launch_bounds__(256)
group_ball_kernel(int n, int P, int K, int C, const float *__restrict__ feat, int ldf,
                  const float *__restrict__ xyz, const float *__restrict__ centres, const int *__restrict__ idx,
                  const int *__restrict__ counts, int fill_missing, float *__restrict__ out, int ldo, int rows) {
  const int lane = threadIdx.x;
  const int row0 = (blockIdx.x * 8 + threadIdx.y) * kGbRows;          // (b*P + p)*K + k
  if (row0 >= rows) return;
  const int nrows = min(kGbRows, rows - row0);
  const int rr = min(row0 + lane, rows - 1);
  const int bp_l = rr / K;
  const int fb_l = (bp_l / P) * n + __ldg(idx + rr);                  // row of the gathered point in feat / xyz
  const int miss_l = (fill_missing && counts && __ldg(counts + bp_l) == 0) ? 1 : 0;
  float *obase = out + (size_t)row0 * ldo;
  const int total = nrows * ldo;
  int u = 0, col = lane;
  while (col >= ldo) { col -= ldo; ++u; }
#pragma unroll 4
  for (int e = lane; e < kGbRows * ldo; e += 32) {
    const int uu = min(u, kGbRows - 1);
    const int fb = __shfl_sync(0xffffffffu, fb_l, uu);
    const int bp = __shfl_sync(0xffffffffu, bp_l, uu);
    const int miss = __shfl_sync(0xffffffffu, miss_l, uu);
    float v = 0.f;
    if (e < total) {
      if (col < C) {
        if (!miss) v = __ldg(feat + (size_t)fb * ldf + col);
      } else {
        const int q = col - C;
        if (q < 9) {
          const int d = q - 3 * (q / 3);
          const float cen = __ldg(centres + (size_t)bp * 3 + d);
          const float ab = miss ? cen : __ldg(xyz + (size_t)fb * 3 + d);
          v = q < 3 ? ab - cen : (q < 6 ? ab : cen);
        }
      }
      obase[e] = v;
    }
    col += 32;
    while (col >= ldo) { col -= ldo; ++u; }
  }
}

// kNN grouping rows: [feat | d2 | w | nn_abs | nn_rel | x | 0-pad], same block shape.
__global__ void __launch_bounds__(256)
group_knn_kernel(int n, int P, int K, int C, const float *__restrict__ feat, int ldf, const float *__restrict__ y,
                 const float *__restrict__ x, const int64_t *__restrict__ idx, const float *__restrict__ dists,
                 float *__restrict__ out, int ldo, int rows) {
  const int row = blockIdx.x * 8 + threadIdx.y;
  if (row >= rows) return;
  const int bp = row / K;
  const int b = bp / P;
  const int src = (int)__ldg(idx + row);
  const float *frow = feat + ((size_t)b * n + src) * ldf;
  float *orow = out + (size_t)row * ldo;
  for (int c = threadIdx.x; c < C; c += 32) orow[c] = __ldg(frow + c);
  const int q = threadIdx.x;
  if (q < ldo - C) {
    float v = 0.f;
    if (q == 0) {
      v = __ldg(dists + row);
    } else if (q == 1) {
      // weight = (1/(d+1e-8)) / sum_k (1/(d_k+1e-8)), summed in neighbour order like torch.sum(dim=2)
      float norm = 0.f;
      for (int k = 0; k < K; ++k) norm += 1.0f / (__ldg(dists + (size_t)bp * K + k) + 1e-8f);
      v = (1.0f / (__ldg(dists + row) + 1e-8f)) / norm;
    } else if (q < 11) {
      const int t = q - 2, d = t % 3;
      const float xv = __ldg(x + (size_t)bp * 3 + d);
      const float yv = __ldg(y + ((size_t)b * n + src) * 3 + d);
      v = t < 3 ? yv : (t < 6 ? yv - xv : xv);
    }
    orow[C + q] = v;
  }
}

__global__ void __launch_bounds__(256)
gather_rows_kernel2(int n, int P, int C, const float *__restrict__ src, int lds, const int *__restrict__ idx,
                    float *__restrict__ out, int ldo, long long total, int round_tf32) {
  pdl_launch_dependents();      // (PDL, common.cuh) the next kernel may ramp up; nothing below touches global memory before
  pdl_wait();                   // every earlier kernel has completed
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const long long bp = i / C;
  const int b = (int)(bp / P);
  const int j = idx ? __ldg(idx + bp) : (int)(bp % P);
  out[bp * ldo + c] = rtf32(__ldg(src + ((size_t)b * n + j) * lds + c), round_tf32);
}

// Geometric channels + table row of every grouped row in one pass, thread per row (the form the gathered-A GEMM
// consumes): geo[row] = [rel(3) | abs(3) | centre(3) | 0 0 0], src_row[row] = b*n + idx or -1 (fill rule).  Same
// arithmetic as group_ball_kernel (rel = abs - centre, one rounding).
__global__ void __launch_bounds__(256)
group_geo_ball_kernel(int n, int P, int K, const float *__restrict__ xyz, const float *__restrict__ centres,
                      const int *__restrict__ idx, const int *__restrict__ counts, int fill_missing,
                      float *__restrict__ geo, int *__restrict__ src_row, int rows, int rt) {
  pdl_launch_dependents();      // (PDL, common.cuh) the next kernel may ramp up; nothing below touches global memory before
  pdl_wait();                   // every earlier kernel has completed
  const int row = blockIdx.x * blockDim.x + threadIdx.x;             // (b*P + p)*K + k
  if (row >= rows) return;
  const int bp = row / K;
  const int fb = (bp / P) * n + __ldg(idx + row);
  const bool miss = fill_missing && counts && __ldg(counts + bp) == 0;
  const float cx = __ldg(centres + (size_t)bp * 3), cy = __ldg(centres + (size_t)bp * 3 + 1),
              cz = __ldg(centres + (size_t)bp * 3 + 2);
  float ax = cx, ay = cy, az = cz;
  if (!miss) { ax = __ldg(xyz + (size_t)fb * 3); ay = __ldg(xyz + (size_t)fb * 3 + 1); az = __ldg(xyz + (size_t)fb * 3 + 2); }
  float4 *o = reinterpret_cast<float4 *>(geo + (size_t)row * 12);
  o[0] = make_float4(rtf32(ax - cx, rt), rtf32(ay - cy, rt), rtf32(az - cz, rt), rtf32(ax, rt));
  o[1] = make_float4(rtf32(ay, rt), rtf32(az, rt), rtf32(cx, rt), rtf32(cy, rt));
  o[2] = make_float4(rtf32(cz, rt), 0.f, 0.f, 0.f);
  src_row[row] = miss ? -1 : fb;
}

// kNN flavour: geo[row] = [d2 | w | nn_abs(3) | nn_rel(3) | x(3) | 0], w as in group_knn_kernel.
__global__ void __launch_bounds__(256)
group_geo_knn_kernel(int n, int P, int K, const float *__restrict__ y, const float *__restrict__ x,
                     const long long *__restrict__ idx, const float *__restrict__ dists, float *__restrict__ geo,
                     int *__restrict__ src_row, int rows, int rt) {
  pdl_launch_dependents();      // (PDL, common.cuh) the next kernel may ramp up; nothing below touches global memory before
  pdl_wait();                   // every earlier kernel has completed
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  const int bp = row / K;
  const int fb = (bp / P) * n + (int)__ldg(idx + row);
  float norm = 0.f;
  for (int k = 0; k < K; ++k) norm += 1.0f / (__ldg(dists + (size_t)bp * K + k) + 1e-8f);
  const float d = __ldg(dists + row);
  const float w = (1.0f / (d + 1e-8f)) / norm;
  const float xx = __ldg(x + (size_t)bp * 3), xy = __ldg(x + (size_t)bp * 3 + 1), xz = __ldg(x + (size_t)bp * 3 + 2);
  const float yx = __ldg(y + (size_t)fb * 3), yy = __ldg(y + (size_t)fb * 3 + 1), yz = __ldg(y + (size_t)fb * 3 + 2);
  float4 *o = reinterpret_cast<float4 *>(geo + (size_t)row * 12);
  o[0] = make_float4(rtf32(d, rt), rtf32(w, rt), rtf32(yx, rt), rtf32(yy, rt));
  o[1] = make_float4(rtf32(yz, rt), rtf32(yx - xx, rt), rtf32(yy - xy, rt), rtf32(yz - xz, rt));
  o[2] = make_float4(rtf32(xx, rt), rtf32(xy, rt), rtf32(xz, rt), 0.f);
  src_row[row] = fb;
}

__global__ void __launch_bounds__(256)
group_src_rows_kernel(int n, int P, int K, const int *__restrict__ idx32, const long long *__restrict__ idx64,
                      const int *__restrict__ counts, int fill_missing, int *__restrict__ src_row, long long rows) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (b*P + p)*K + k
  if (row >= rows) return;
  const long long bp = row / K;
  const int b = (int)(bp / P);
  const int j = idx64 ? (int)__ldg(idx64 + row) : __ldg(idx32 + row);
  const bool miss = fill_missing && counts && __ldg(counts + bp) == 0;
  src_row[row] = miss ? -1 : b * n + j;
}

inline unsigned blocks_for(long long total) { return (unsigned)((total + 255) / 256); }

}  // namespace
}  // namespace pdr

using namespace pdr;

extern "C" int pdr_gemm_tile_rows(void) { return kTileM; }

extern "C" int pdr_gemm_fused(const PdrGemmArgs *args, void *stream_) {
  PDR_REQUIRE(args, "gemm_fused: null args");
  const PdrGemmArgs &a = *args;
  cudaStream_t stream = (cudaStream_t)stream_;
  PDR_REQUIRE(a.A && a.W && (a.C || a.pool_K > 0), "gemm_fused: null pointer");
  PDR_REQUIRE(a.K > 0 && a.N > 0 && a.batch > 0 && a.rows_per_sample > 0, "gemm_fused: bad sizes");
  PDR_REQUIRE(a.K % 4 == 0 && a.lda % 4 == 0 && a.ldw % 4 == 0 && (a.a_rows || a.tail_rows || a.lda >= a.K) && a.ldw >= a.K,
              "gemm_fused: K/lda/ldw must be multiples of 4 (K=%d lda=%d ldw=%d)", a.K, a.lda, a.ldw);
  PDR_REQUIRE(a.ldc >= a.N && a.ldc_zero_to <= a.ldc, "gemm_fused: ldc=%d < N=%d", a.ldc, a.N);
  PDR_REQUIRE(a.pro_mode == PDR_PRO_NONE || (a.sc && a.sh && a.ld_scsh % 4 == 0 && ((uintptr_t)a.sc % 16) == 0 &&
                                             ((uintptr_t)a.sh % 16) == 0),
              "gemm_fused: prologue needs 16-byte aligned sc/sh with ld %% 4 == 0");
  PDR_REQUIRE(!a.add || (a.ld_add % 4 == 0 && ((uintptr_t)a.add % 16) == 0), "gemm_fused: add alignment");
  PDR_REQUIRE(!a.R || a.ldr % 4 == 0, "gemm_fused: ldr must be a multiple of 4");
  PDR_REQUIRE(!a.rowadd || a.rowadd_div > 0, "gemm_fused: rowadd_div");
  PDR_REQUIRE(((uintptr_t)a.A % 16) == 0 && ((uintptr_t)a.W % 16) == 0 && (!a.R || ((uintptr_t)a.R % 16) == 0),
              "gemm_fused: A/W/R must be 16-byte aligned");
  // the tensor-core epilogue stores float4s: it needs 16-byte aligned output rows (and broadcast rows); anything
  // else takes the SIMT kernel
  const bool tc_aligned = a.ldc % 4 == 0 && ((uintptr_t)a.C % 16) == 0 &&
                          (!a.rowadd || (a.ld_rowadd % 4 == 0 && ((uintptr_t)a.rowadd % 16) == 0));
  if (a.tail_rows) {
    PDR_REQUIRE(a.T && a.T2 && a.k_pro > 0 && a.k_pro < a.K && a.k_pro % 32 == 0 && a.t_split > 0 && a.t_split % 4 == 0 &&
                    a.t_split < a.K - a.k_pro && a.ldt % 4 == 0 && a.ldt2 % 4 == 0 && a.ldt >= a.t_split &&
                    a.ldt2 >= a.K - a.k_pro - a.t_split && ((uintptr_t)a.T % 16) == 0 && ((uintptr_t)a.T2 % 16) == 0 &&
                    a.lda >= a.k_pro,
                "gemm_fused: raw K tail needs T/T2, k_pro %% 32 == 0, t_split/ldt/ldt2 multiples of 4");
    PDR_REQUIRE(a.pro_mode != PDR_PRO_NONE && !a.R && !a.a_rows, "gemm_fused: raw K tail needs a prologue and excludes R / a_rows");
    if (!(a.use_tf32 && tc_aligned)) {
      set_error("gemm_fused: the raw K tail is implemented on the tensor-core path only");
      return PDR_ERR_UNSUPPORTED;
    }
  }
  if (a.pool_K > 0) {
    PDR_REQUIRE((a.pool_K == 8 || a.pool_K == 16 || a.pool_K == 32) && a.rows_per_sample % a.pool_K == 0 && a.pool_V &&
                    a.pool_sc && a.pool_sh && a.pool_out && !a.stats && !a.rowadd,
                "gemm_fused: pooling epilogue needs K in {8,16,32} dividing rows_per_sample, V/sc/sh/out, no stats/rowadd");
    if (!a.use_tf32) {
      set_error("gemm_fused: the pooling epilogue is implemented on the tensor-core path only");
      return PDR_ERR_UNSUPPORTED;
    }
    return launch_gemm_tf32(a, stream);
  }
  if (a.a_rows) {
    PDR_REQUIRE(a.A2 && a.k_split > 0 && a.k_split < a.K && a.k_split % 4 == 0 && a.lda2 % 4 == 0 &&
                    a.lda >= a.k_split && a.lda2 >= a.K - a.k_split && ((uintptr_t)a.A2 % 16) == 0,
                "gemm_fused: gathered A needs A2, k_split/lda2 multiples of 4, lda >= k_split, lda2 >= K - k_split");
    PDR_REQUIRE(a.pro_mode == PDR_PRO_NONE && !a.add && !a.R, "gemm_fused: gathered A excludes prologue / add / R");
    if (!(a.use_tf32 && tc_aligned)) {
      set_error("gemm_fused: gathered A is implemented on the tensor-core path only (use_tf32, aligned output)");
      return PDR_ERR_UNSUPPORTED;
    }
  }
  if (a.use_tf32 && tc_aligned) return launch_gemm_tf32(a, stream);
  const int tiles_per_sample = ceil_div(a.rows_per_sample, kTileM);
  const long long tiles = (long long)a.batch * tiles_per_sample;
  PDR_REQUIRE(tiles <= 65535ll * 32768, "gemm_fused: too many tiles");
  if (a.N <= 32) {
    dim3 grid(ceil_div(a.N, 32), (unsigned)tiles);
    gemm_simt_kernel<2><<<grid, kGemmThreads, 0, stream>>>(a);
  } else {
    dim3 grid(ceil_div(a.N, 64), (unsigned)tiles);
    gemm_simt_kernel<4><<<grid, kGemmThreads, 0, stream>>>(a);
  }
  return check_launch("gemm_simt_kernel");
}

static int gn_check(const PdrGnArgs &a) {
  PDR_REQUIRE(a.nsrc >= 1 && a.nsrc <= PDR_GN_MAX_SOURCES && a.batch > 0 && a.channels > 0, "gn_finalize: bad sizes");
  PDR_REQUIRE(a.gn_channels <= a.channels && a.groups > 0 && a.gn_channels % a.groups == 0,
              "gn_finalize: channels=%d gn=%d groups=%d ld=%d", a.channels, a.gn_channels, a.groups, a.ld_out);
  int tot = 0;
  for (int s = 0; s < a.nsrc; ++s) tot += a.src[s].ncols;
  PDR_REQUIRE(tot == a.channels, "gn_finalize: sources cover %d of %d channels", tot, a.channels);
  PDR_REQUIRE(a.gamma && a.beta && a.sc && a.sh, "gn_finalize: null pointer");
  PDR_REQUIRE((size_t)a.channels * 3 * sizeof(double) <= 48 * 1024, "gn_finalize: too many channels");
  return PDR_OK;
}

extern "C" int pdr_gn_finalize_batch(const PdrGnArgs *args, int count, void *stream) {
  PDR_REQUIRE(args && count >= 1 && count <= 2, "gn_finalize: count=%d not in {1, 2}", count);
  GnBatch batch;
  memset(&batch, 0, sizeof(batch));
  int nb = 0, channels = 0, split = kGnSplit;
  for (int i = 0; i < count; ++i) {
    const int rc = gn_check(args[i]);
    if (rc) return rc;
    batch.g[i] = args[i];
    nb = args[i].batch > nb ? args[i].batch : nb;
    channels = args[i].channels > channels ? args[i].channels : channels;
    if (args[i].groups < kGnSplit) split = 1;
  }
  const size_t smem = (size_t)channels * 3 * sizeof(double);      // upper bound for any split
  const cudaError_t e = launch_pdl(gn_finalize_kernel, dim3(nb, split, count), dim3(kGnThreads), smem, (cudaStream_t)stream, batch);
  if (e != cudaSuccess) { set_error("gn_finalize_kernel: %s", cudaGetErrorString(e)); return PDR_ERR_CUDA; }
  return check_launch("gn_finalize_kernel");
}

extern "C" int pdr_gn_finalize(const PdrGnArgs *args, void *stream) {
  PDR_REQUIRE(args, "gn_finalize: null args");
  return pdr_gn_finalize_batch(args, 1, stream);
}

extern "C" int pdr_affine_rows(int batch, int rows_per_sample, int C, const float *x, int ldx, int pro_mode,
                               const float *sc, const float *sh, int ld_scsh, const float *add, int ld_add,
                               const float *R, int ldr, float *out, int ldo, int round_tf32, void *stream) {
  PDR_REQUIRE(batch > 0 && rows_per_sample > 0 && C > 0 && ldo >= C && x && out, "affine_rows: bad arguments");
  const long long total = (long long)batch * rows_per_sample * C;
  launch_pdl(affine_rows_kernel, dim3(blocks_for(total)), dim3(256), 0, (cudaStream_t)stream, rows_per_sample, C, x, ldx, pro_mode, sc,
             sh, ld_scsh, add, ld_add, R, ldr, out, ldo, total, round_tf32);
  return check_launch("affine_rows_kernel");
}

extern "C" int pdr_attention_pool(int batch, int P, int K, int C, const float *S, int lds, const float *V, int ldv,
                                  const float *sc, const float *sh, int ld_scsh, const int *counts, float *out,
                                  int ldo, int round_tf32, void *stream) {
  PDR_REQUIRE(batch > 0 && P > 0 && K > 0 && C > 0 && S && V && sc && sh && out, "attention_pool: bad arguments");
  const long long total = (long long)batch * P * C;
  if (K == 32)
    launch_pdl(attention_pool_kernel<32>, dim3(blocks_for(total)), dim3(256), 0, (cudaStream_t)stream, P, K, C, S, lds, V, ldv, sc, sh,
               ld_scsh, counts, out, ldo, total, round_tf32);
  else if (K == 8)
    launch_pdl(attention_pool_kernel<8>, dim3(blocks_for(total)), dim3(256), 0, (cudaStream_t)stream, P, K, C, S, lds, V, ldv, sc, sh,
               ld_scsh, counts, out, ldo, total, round_tf32);
  else
    launch_pdl(attention_pool_kernel<0>, dim3(blocks_for(total)), dim3(256), 0, (cudaStream_t)stream, P, K, C, S, lds, V, ldv, sc, sh,
               ld_scsh, counts, out, ldo, total, round_tf32);
  return check_launch("attention_pool_kernel");
}

extern "C" int pdr_group_ball(int batch, int n, int P, int K, int C, const float *feat, int ldf, const float *xyz,
                              const float *centres, const int *idx, const int *counts, int fill_missing,
                              float *out, int ldo, void *stream) {
  PDR_REQUIRE(batch > 0 && n > 0 && P > 0 && K > 0 && C >= 0 && ldo >= C + 9, "group_ball: bad sizes");
  PDR_REQUIRE((feat || C == 0) && xyz && centres && idx && out, "group_ball: null pointer");
  const long long rows = (long long)batch * P * K;
  PDR_REQUIRE(rows < (1ll << 31) && ldo - C <= 32, "group_ball: too many rows or pad too wide");
  group_ball_kernel<<<(unsigned)((rows + 8 * kGbRows - 1) / (8 * kGbRows)), dim3(32, 8), 0, (cudaStream_t)stream>>>(
      n, P, K, C, feat, ldf, xyz, centres, idx, counts, fill_missing, out, ldo, (int)rows);
  return check_launch("group_ball_kernel");
}

extern "C" int pdr_group_knn(int batch, int n, int P, int K, int C, const float *feat, int ldf, const float *y,
                             const float *x, const int64_t *idx, const float *dists, float *out, int ldo,
                             void *stream) {
  PDR_REQUIRE(batch > 0 && n > 0 && P > 0 && K > 0 && C >= 0 && ldo >= C + 11, "group_knn: bad sizes");
  PDR_REQUIRE((feat || C == 0) && y && x && idx && dists && out, "group_knn: null pointer");
  const long long rows = (long long)batch * P * K;
  PDR_REQUIRE(rows < (1ll << 31) && ldo - C <= 32, "group_knn: too many rows or pad too wide");
  group_knn_kernel<<<(unsigned)((rows + 7) / 8), dim3(32, 8), 0, (cudaStream_t)stream>>>(n, P, K, C, feat, ldf, y, x, idx,
                                                                                          dists, out, ldo, (int)rows);
  return check_launch("group_knn_kernel");
}

extern "C" int pdr_group_geo_ball(int batch, int n, int P, int K, const float *xyz, const float *centres, const int *idx,
                                  const int *counts, int fill_missing, float *geo, int *src_row, int round_tf32,
                                  void *stream) {
  PDR_REQUIRE(batch > 0 && n > 0 && P > 0 && K > 0 && xyz && centres && idx && geo && src_row, "group_geo_ball: bad arguments");
  const long long rows = (long long)batch * P * K;
  PDR_REQUIRE(rows < (1ll << 31) && (long long)batch * n < (1ll << 31) && ((uintptr_t)geo % 16) == 0,
              "group_geo_ball: too many rows or unaligned output");
  launch_pdl(group_geo_ball_kernel, dim3(blocks_for(rows)), dim3(256), 0, (cudaStream_t)stream, n, P, K, xyz, centres, idx, counts,
             fill_missing, geo, src_row, (int)rows, round_tf32);
  return check_launch("group_geo_ball_kernel");
}

extern "C" int pdr_group_geo_knn(int batch, int n, int P, int K, const float *y, const float *x, const int64_t *idx,
                                 const float *dists, float *geo, int *src_row, int round_tf32, void *stream) {
  PDR_REQUIRE(batch > 0 && n > 0 && P > 0 && K > 0 && y && x && idx && dists && geo && src_row, "group_geo_knn: bad arguments");
  const long long rows = (long long)batch * P * K;
  PDR_REQUIRE(rows < (1ll << 31) && (long long)batch * n < (1ll << 31) && ((uintptr_t)geo % 16) == 0,
              "group_geo_knn: too many rows or unaligned output");
  launch_pdl(group_geo_knn_kernel, dim3(blocks_for(rows)), dim3(256), 0, (cudaStream_t)stream, n, P, K, y, x, (const long long *)idx,
             dists, geo, src_row, (int)rows, round_tf32);
  return check_launch("group_geo_knn_kernel");
}

extern "C" int pdr_group_src_rows(int batch, int n, int P, int K, const void *idx, int idx_is_int64, const int *counts,
                                  int fill_missing, int *src_row, void *stream) {
  PDR_REQUIRE(batch > 0 && n > 0 && P > 0 && K > 0 && idx && src_row, "group_src_rows: bad arguments");
  PDR_REQUIRE((long long)batch * n < (1ll << 31), "group_src_rows: table too large");
  const long long rows = (long long)batch * P * K;
  group_src_rows_kernel<<<blocks_for(rows), 256, 0, (cudaStream_t)stream>>>(
      n, P, K, idx_is_int64 ? nullptr : (const int *)idx, idx_is_int64 ? (const long long *)idx : nullptr, counts,
      fill_missing, src_row, rows);
  return check_launch("group_src_rows_kernel");
}

extern "C" int pdr_gather_rows(int batch, int n, int P, int C, const float *src, int lds, const int *idx, float *out,
                               int ldo, int round_tf32, void *stream) {
  PDR_REQUIRE(batch > 0 && n > 0 && P > 0 && C > 0 && src && out, "gather_rows: bad arguments");
  const long long total = (long long)batch * P * C;
  launch_pdl(gather_rows_kernel2, dim3(blocks_for(total)), dim3(256), 0, (cudaStream_t)stream, n, P, C, src, lds, idx, out, ldo, total,
             round_tf32);
  return check_launch("gather_rows_kernel");
}
