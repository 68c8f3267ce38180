/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference's hot-path kernels.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library.  The product package (point_diffusion_refinement_b200/) never does; it fails loudly
 * when its CUDA library is missing.
 *
 * The reference (ZhaoyangLyu/Point_Diffusion_Refinement) has NO CPU path for any of these ops
 * (every binding ends in AT_ASSERT(false, "CPU not supported"), e.g.
 * pointnet2_ops_lib/pointnet2_ops/_ext-src/src/sampling.cpp:83), so this file restates the CUDA
 * kernels statement by statement, including the FMA contraction nvcc 12.9 chooses for them
 * (checked with cuobjdump -sass on the reference sources compiled for sm_100a):
 *     mag = fma(z,z, fma(x,x, y*y))            d = fma(dz,dz, fma(dx,dx, dy*dy))
 * Pinning: see oracle/README.md (chamfer vs float64 golden, EMD 2x2 known answer, and bit-exact
 * comparison with the compiled reference kernels in oracle/_ref on the GPU box).
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC (see oracle/Makefile).  -ffp-contract=off
 * matters: every fused multiply-add below is explicit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TOTAL_THREADS 512

/* cuda_utils.h:13-19 opt_n_threads */
static int opt_n_threads(int work_size) {
  int pow_2 = (int)(log((double)work_size) / log(2.0));
  int v = 1 << pow_2;
  if (v > TOTAL_THREADS) v = TOTAL_THREADS;
  if (v < 1) v = 1;
  return v;
}
int oracle_opt_n_threads(int work_size) { return opt_n_threads(work_size); }

/* squared distance with the reference's SASS contraction: dy*dy rounded, then dx and dz fused */
static inline float dist2_ref(float dx, float dy, float dz) {
  return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/* ------------------------------------------------------------------------------------------------
 * furthest_point_sampling: sampling_gpu.cu:69-173, launch rule :175-229, temp init sampling.cpp:74-76.
 * Emulates the block of `bs` threads literally: strided per-thread scan, then the shared-memory
 * tree (__update, sampling_gpu.cu:59-65) level by level, so tie-breaking is the reference's.
 * ---------------------------------------------------------------------------------------------- */
void oracle_fps(int b, int n, int m, const float *xyz_all, int *idx_all) {
  if (m <= 0) return;
  const int bs = opt_n_threads(n);
#pragma omp parallel for schedule(dynamic, 1)
  for (int bi = 0; bi < b; ++bi) {
    const float *dataset = xyz_all + (size_t)bi * n * 3;
    int *idxs = idx_all + (size_t)bi * m;
    float *temp = (float *)malloc(sizeof(float) * (size_t)n);
    float *dists = (float *)malloc(sizeof(float) * (size_t)bs);
    int *dists_i = (int *)malloc(sizeof(int) * (size_t)bs);
    for (int k = 0; k < n; ++k) temp[k] = 1e10f;
    int old = 0;
    idxs[0] = old;
    for (int j = 1; j < m; ++j) {
      const float x1 = dataset[old * 3 + 0], y1 = dataset[old * 3 + 1], z1 = dataset[old * 3 + 2];
      for (int tid = 0; tid < bs; ++tid) {
        int besti = 0;
        float best = -1.0f;
        for (int k = tid; k < n; k += bs) {
          const float x2 = dataset[k * 3 + 0], y2 = dataset[k * 3 + 1], z2 = dataset[k * 3 + 2];
          const float mag = fmaf(z2, z2, fmaf(x2, x2, y2 * y2));
          if ((double)mag <= 1e-3) continue; /* :100-101, float promoted to double */
          const float d = dist2_ref(x2 - x1, y2 - y1, z2 - z1);
          const float d2 = fminf(d, temp[k]);
          temp[k] = d2;
          besti = d2 > best ? k : besti;
          best = d2 > best ? d2 : best;
        }
        dists[tid] = best;
        dists_i[tid] = besti;
      }
      for (int s = bs >> 1; s >= 1; s >>= 1) {
        for (int tid = 0; tid < s; ++tid) {
          const float v1 = dists[tid], v2 = dists[tid + s];
          const int i1 = dists_i[tid], i2 = dists_i[tid + s];
          dists[tid] = fmaxf(v1, v2);
          dists_i[tid] = v2 > v1 ? i2 : i1;
        }
      }
      old = dists_i[0];
      idxs[j] = old;
    }
    free(temp);
    free(dists);
    free(dists_i);
  }
}

/* gather_points: sampling_gpu.cu:8-20.  points (b,c,n), idx (b,m) -> out (b,c,m) */
void oracle_gather_points(int b, int c, int n, int m, const float *points, const int *idx,
                          float *out) {
#pragma omp parallel for collapse(2)
  for (int i = 0; i < b; ++i)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < m; ++j) {
        int a = idx[(size_t)i * m + j];
        out[((size_t)i * c + l) * m + j] = points[((size_t)i * c + l) * n + a];
      }
}

/* ball_query: ball_query_gpu.cu:9-47; idx/counts zero-init ball_query.cpp:21-27 (done here). */
void oracle_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz_all,
                       const float *xyz_all, int *idx_all, int *counts_all) {
  const float radius2 = radius * radius;
#pragma omp parallel for schedule(static)
  for (int bi = 0; bi < b; ++bi) {
    const float *xyz = xyz_all + (size_t)bi * n * 3;
    const float *new_xyz = new_xyz_all + (size_t)bi * m * 3;
    int *idx = idx_all + (size_t)bi * m * nsample;
    int *counts = counts_all + (size_t)bi * m;
    memset(idx, 0, sizeof(int) * (size_t)m * nsample);
    memset(counts, 0, sizeof(int) * (size_t)m);
    for (int j = 0; j < m; ++j) {
      const float nx = new_xyz[j * 3 + 0], ny = new_xyz[j * 3 + 1], nz = new_xyz[j * 3 + 2];
      for (int k = 0, cnt = 0; k < n && cnt < nsample; ++k) {
        const float x = xyz[k * 3 + 0], y = xyz[k * 3 + 1], z = xyz[k * 3 + 2];
        const float d2 = dist2_ref(nx - x, ny - y, nz - z);
        if (d2 < radius2) {
          if (cnt == 0)
            for (int l = 0; l < nsample; ++l) idx[j * nsample + l] = k;
          idx[j * nsample + cnt] = k;
          ++cnt;
          counts[j] = cnt;
        }
      }
    }
  }
}

/* group_points: group_points_gpu.cu:8-28.  points (b,c,n), idx (b,np,ns) -> out (b,c,np,ns) */
void oracle_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                         const int *idx, float *out) {
#pragma omp parallel for collapse(2)
  for (int bi = 0; bi < b; ++bi)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < npoints; ++j)
        for (int k = 0; k < nsample; ++k) {
          int ii = idx[((size_t)bi * npoints + j) * nsample + k];
          out[(((size_t)bi * c + l) * npoints + j) * nsample + k] = points[((size_t)bi * c + l) * n + ii];
        }
}

/* three_nn: interpolate_gpu.cu:9-59 (double-held bests, strict '<' cascade, returns d^2) */
void oracle_three_nn(int b, int n, int m, const float *unknown_all, const float *known_all,
                     float *dist2_all, int *idx_all) {
#pragma omp parallel for schedule(static)
  for (int bi = 0; bi < b; ++bi) {
    const float *unknown = unknown_all + (size_t)bi * n * 3;
    const float *known = known_all + (size_t)bi * m * 3;
    float *dist2 = dist2_all + (size_t)bi * n * 3;
    int *idx = idx_all + (size_t)bi * n * 3;
    for (int j = 0; j < n; ++j) {
      const float ux = unknown[j * 3 + 0], uy = unknown[j * 3 + 1], uz = unknown[j * 3 + 2];
      double best1 = 1e40, best2 = 1e40, best3 = 1e40;
      int besti1 = 0, besti2 = 0, besti3 = 0;
      for (int k = 0; k < m; ++k) {
        const float x = known[k * 3 + 0], y = known[k * 3 + 1], z = known[k * 3 + 2];
        const float d = dist2_ref(ux - x, uy - y, uz - z);
        if (d < best1) {
          best3 = best2; besti3 = besti2;
          best2 = best1; besti2 = besti1;
          best1 = d; besti1 = k;
        } else if (d < best2) {
          best3 = best2; besti3 = besti2;
          best2 = d; besti2 = k;
        } else if (d < best3) {
          best3 = d; besti3 = k;
        }
      }
      dist2[j * 3 + 0] = (float)best1; dist2[j * 3 + 1] = (float)best2; dist2[j * 3 + 2] = (float)best3;
      idx[j * 3 + 0] = besti1; idx[j * 3 + 1] = besti2; idx[j * 3 + 2] = besti3;
    }
  }
}

/* three_interpolate: interpolate_gpu.cu:72-101; SASS order fma(p3,w3, fma(p1,w1, p2*w2)). */
void oracle_three_interpolate(int b, int c, int m, int n, const float *points_all,
                              const int *idx_all, const float *weight_all, float *out_all) {
#pragma omp parallel for collapse(2)
  for (int bi = 0; bi < b; ++bi)
    for (int l = 0; l < c; ++l) {
      const float *points = points_all + ((size_t)bi * c + l) * m;
      const int *idx = idx_all + (size_t)bi * n * 3;
      const float *weight = weight_all + (size_t)bi * n * 3;
      float *out = out_all + ((size_t)bi * c + l) * n;
      for (int j = 0; j < n; ++j) {
        const float w1 = weight[j * 3 + 0], w2 = weight[j * 3 + 1], w3 = weight[j * 3 + 2];
        const int i1 = idx[j * 3 + 0], i2 = idx[j * 3 + 1], i3 = idx[j * 3 + 2];
        out[j] = fmaf(points[i3], w3, fmaf(points[i1], w1, points[i2] * w2));
      }
    }
}

/* ------------------------------------------------------------------------------------------------
 * knn_points (pytorch3d.ops.knn, NOT vendored in the reference; version unpinned by setup_env.sh:5).
 * Published semantics restated: exact brute-force K nearest on squared L2 in fp32, ascending, ties
 * -> lower index (a later equal distance never displaces an earlier one); idx int64; when K > P2
 * the tail is padded with idx 0 / dist 0 as pytorch3d does for short clouds.
 * Call sites: pointnet2/chamfer_loss_new.py:149-150, pointnet2_ops/pointnet2_utils.py:365,496-497.
 * d = ((dx*dx) + dy*dy) + dz*dz accumulated per dimension -> fma(dz,dz, fma(dy,dy, dx*dx)).
 * ---------------------------------------------------------------------------------------------- */
void oracle_knn(int b, int p1, int p2, int K, const float *x_all, const float *y_all, float *dists_all,
                int64_t *idx_all) {
#pragma omp parallel for schedule(static)
  for (int bi = 0; bi < b; ++bi) {
    const float *x = x_all + (size_t)bi * p1 * 3;
    const float *y = y_all + (size_t)bi * p2 * 3;
    float *bd = (float *)malloc(sizeof(float) * (size_t)K);
    int64_t *bk = (int64_t *)malloc(sizeof(int64_t) * (size_t)K);
    for (int i = 0; i < p1; ++i) {
      int cnt = 0;
      const float qx = x[i * 3 + 0], qy = x[i * 3 + 1], qz = x[i * 3 + 2];
      for (int k = 0; k < p2; ++k) {
        const float dx = qx - y[k * 3 + 0], dy = qy - y[k * 3 + 1], dz = qz - y[k * 3 + 2];
        const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        if (cnt < K) {
          int p = cnt++;
          while (p > 0 && bd[p - 1] > d) { bd[p] = bd[p - 1]; bk[p] = bk[p - 1]; --p; }
          bd[p] = d; bk[p] = k;
        } else if (d < bd[K - 1]) {
          int p = K - 1;
          while (p > 0 && bd[p - 1] > d) { bd[p] = bd[p - 1]; bk[p] = bk[p - 1]; --p; }
          bd[p] = d; bk[p] = k;
        }
      }
      for (int t = 0; t < K; ++t) {
        dists_all[((size_t)bi * p1 + i) * K + t] = t < cnt ? bd[t] : 0.0f;
        idx_all[((size_t)bi * p1 + i) * K + t] = t < cnt ? bk[t] : 0;
      }
    }
    free(bd);
    free(bk);
  }
}

/* ------------------------------------------------------------------------------------------------
 * NmDistance (vendored chamfer3D, the only in-tree Chamfer kernel):
 * pointnet2/models/pvd/metrics/ChamferDistancePytorch/chamfer3D/chamfer3D.cu:12-134.
 * One direction: for each of n points in xyz, nearest of m points in xyz2 (squared), first minimum
 * wins (strict '<').  d = x2*x2 + y2*y2 + z2*z2 with x2 = buf - x1 -> fma(z,z, fma(x,x, y*y))
 * [same contraction nvcc picks for every 3-term sum of squares here; checked in oracle/_ref SASS].
 * The 512-point tiling of the kernel changes nothing observable: tiles are visited in index order
 * and a later tile only replaces on strict '>' (:127), i.e. global first minimum.
 * ---------------------------------------------------------------------------------------------- */
void oracle_nm_distance(int b, int n, const float *xyz, int m, const float *xyz2, float *result,
                        int *result_i) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < b; ++i)
    for (int j = 0; j < n; ++j) {
      const float x1 = xyz[((size_t)i * n + j) * 3 + 0], y1 = xyz[((size_t)i * n + j) * 3 + 1],
                  z1 = xyz[((size_t)i * n + j) * 3 + 2];
      float best = 0;
      int best_i = 0;
      for (int k = 0; k < m; ++k) {
        const float x2 = xyz2[((size_t)i * m + k) * 3 + 0] - x1, y2 = xyz2[((size_t)i * m + k) * 3 + 1] - y1,
                    z2 = xyz2[((size_t)i * m + k) * 3 + 2] - z1;
        const float d = dist2_ref(x2, y2, z2);
        if (k == 0 || d < best) { best = d; best_i = k; }
      }
      result[(size_t)i * n + j] = best;
      result_i[(size_t)i * n + j] = best_i;
    }
}

/* ------------------------------------------------------------------------------------------------
 * approxmatch: PytorchEMD/cuda/emd_kernel.cu:29-161 (launch <<<32,512>>> :191).
 * Each virtual thread owns rows k (resp. columns l) and accumulates sequentially over the other
 * axis in index order -- the same order the kernel's inner smem loop uses -- so the only deviation
 * from the GPU is __expf (MUFU.EX2 approximation) vs expf here.
 * match is (b, m, n) laid out match[i*n*m + l*n + k].
 * ---------------------------------------------------------------------------------------------- */
static inline float pair_d2(const float *p1, const float *p2) { /* (x2-x1)^2+(y2-y1)^2+(z2-z1)^2 */
  return dist2_ref(p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]);
}

void oracle_approxmatch(int b, int n, int m, const float *xyz1_all, const float *xyz2_all,
                        float *match_all) {
  float multiL, multiR;
  if (n >= m) { multiL = 1; multiR = (float)(n / m); } /* integer division, :33-38 */
  else { multiL = (float)(m / n); multiR = 1; }
#pragma omp parallel for schedule(dynamic, 1)
  for (int i = 0; i < b; ++i) {
    const float *xyz1 = xyz1_all + (size_t)i * n * 3;
    const float *xyz2 = xyz2_all + (size_t)i * m * 3;
    float *match = match_all + (size_t)i * n * m;
    float *remainL = (float *)malloc(sizeof(float) * (size_t)(n + m) * 2);
    float *remainR = remainL + n, *ratioL = remainL + n + m, *ratioR = remainL + n + m + n;
    memset(match, 0, sizeof(float) * (size_t)n * m);
    for (int j = 0; j < n; ++j) remainL[j] = multiL;
    for (int j = 0; j < m; ++j) remainR[j] = multiR;
    for (int j = 7; j >= -2; --j) {
      float level = -powf(4.0f, (float)j);
      if (j == -2) level = 0;
      for (int k = 0; k < n; ++k) {
        float suml = 1e-9f;
        for (int l = 0; l < m; ++l) {
          const float d = level * pair_d2(xyz1 + k * 3, xyz2 + l * 3);
          suml += expf(d) * remainR[l];
        }
        ratioL[k] = remainL[k] / suml;
      }
      for (int l = 0; l < m; ++l) {
        float sumr = 0;
        for (int k = 0; k < n; ++k)
          sumr += expf(level * pair_d2(xyz1 + k * 3, xyz2 + l * 3)) * ratioL[k];
        sumr *= remainR[l];
        const float consumption = fminf(remainR[l] / (sumr + 1e-9f), 1.0f);
        ratioR[l] = consumption * remainR[l];
        remainR[l] = fmaxf(0.0f, remainR[l] - sumr);
      }
      for (int k = 0; k < n; ++k) {
        float suml = 0;
        const float rl = ratioL[k];
        for (int l = 0; l < m; ++l) {
          const float w = expf(level * pair_d2(xyz1 + k * 3, xyz2 + l * 3)) * rl * ratioR[l];
          match[(size_t)l * n + k] += w;
          suml += w;
        }
        remainL[k] = fmaxf(0.0f, remainL[k] - suml);
      }
    }
    free(remainL);
  }
}

/* matchcost: emd_kernel.cu:204-246, 512 virtual threads then the pairwise tree at :236-242. */
void oracle_matchcost(int b, int n, int m, const float *xyz1_all, const float *xyz2_all,
                      const float *match_all, float *out) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int i = 0; i < b; ++i) {
    const float *xyz1 = xyz1_all + (size_t)i * n * 3;
    const float *xyz2 = xyz2_all + (size_t)i * m * 3;
    const float *match = match_all + (size_t)i * n * m;
    float allsum[512];
    for (int t = 0; t < 512; ++t) {
      float subsum = 0;
      for (int k = t; k < n; k += 512)
        for (int l = 0; l < m; ++l)
          subsum += pair_d2(xyz1 + k * 3, xyz2 + l * 3) * match[(size_t)l * n + k];
      allsum[t] = subsum;
    }
    for (int j = 1; j < 512; j <<= 1)
      for (int t = 0; t < 512; ++t)
        if ((t & j) == 0 && t + j < 512 && (t & (j - 1)) == 0) allsum[t] += allsum[t + j];
    out[i] = allsum[0];
  }
}
