/*
 * pdr_b200.h -- C ABI of libpdr_b200.so, the B200 (sm_100a) replacement for the native kernels on the
 * hot path of ZhaoyangLyu/Point_Diffusion_Refinement:
 *
 *   pointnet2_ops._ext   pointnet2_ops_lib/pointnet2_ops/_ext-src/src/bindings.cpp:6-19
 *   emd_cuda             PytorchEMD/cuda/emd.cpp:24-28
 *   pytorch3d.ops.knn    (third-party, un-vendored) call sites pointnet2/chamfer_loss_new.py:149-150,
 *                        pointnet2_ops_lib/pointnet2_ops/pointnet2_utils.py:365,496-497
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes, no torch types; every pointer is a DEVICE pointer on the current
 *     device unless stated otherwise; tensors are dense, row-major, fp32 / int32 (int64 where the
 *     reference API says so);
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); nothing allocates,
 *     nothing synchronises, so every call is CUDA-graph capturable and re-entrant;
 *   - return value: 0 on success, PDR_ERR_* (< 0) otherwise.  Unlike the reference
 *     (cuda_utils.h:30-39 prints and exit(-1)s) no call aborts the process;
 *     pdr_last_error_string() describes the last failure on the calling thread.
 *   - outputs are fully written by the kernels (no reliance on zero-initialised buffers; the
 *     reference relies on torch::zeros at ball_query.cpp:21-27).
 */
#ifndef PDR_B200_H_
#define PDR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDR_OK 0
#define PDR_ERR_INVALID_ARGUMENT (-1)
#define PDR_ERR_CUDA (-2)
#define PDR_ERR_WORKSPACE (-3)
#define PDR_ERR_UNSUPPORTED (-4)

/* library identity ------------------------------------------------------------------------------ */
int pdr_version(void);                      /* e.g. 100 = 0.1.0 */
const char *pdr_last_error_string(void);    /* thread-local, never NULL */
int pdr_built_for_sm(void);                 /* 100 */
/* Diagnostic (no reference counterpart): HBM ceilings the store-dominated kernels are held against, measured with
 * incompressible data.  mode 0: write-only, 16-byte stores; mode 1: the same plus one 512-byte segment in three read
 * from src (1 read : 3 writes); mode 2: write-only through 4 KiB TMA bulk stores from shared memory.  n_floats a
 * multiple of 1024, dst (and src) 16-byte aligned with n_floats elements.  bench.py / scripts/bw_probe.py time it. */
int pdr_probe_hbm(int mode, float *dst, const float *src, size_t n_floats, void *stream);

/* ---- pointnet2_ops._ext forward ops ----------------------------------------------------------- */

/* furthest_point_sampling.  Replaces furthest_point_sampling_kernel_wrapper(b,n,m,dataset,temp,idxs)
 * (sampling.cpp:11-13, kernel sampling_gpu.cu:69-173).  xyz (b,n,3) -> idx (b,m) int32.
 * Bit-exact with the reference, including the |p|^2 <= 1e-3 skip rule and its tie-breaking.
 * `temp` (b,n) fp32 scratch is only needed when n > pdr_fps_max_onchip_points(); pass NULL below
 * that.  It need not be initialised (the reference wants it pre-filled with 1e10). */
int pdr_fps_max_onchip_points(void);
int pdr_furthest_point_sampling(int b, int n, int m, const float *xyz, float *temp, int *idx,
                                void *stream);

/* gather_points: points (b,c,n), idx (b,m) -> out (b,c,m).  sampling.cpp:4-6, sampling_gpu.cu:8-30. */
int pdr_gather_points(int b, int c, int n, int m, const float *points, const int *idx, float *out,
                      void *stream);
/* gather_points_grad: grad_out (b,c,m), idx (b,m) -> grad_points (b,c,n) (zeroed inside). */
int pdr_gather_points_grad(int b, int c, int n, int m, const float *grad_out, const int *idx,
                           float *grad_points, void *stream);

/* ball_query.  Replaces query_ball_point_kernel_wrapper(b,n,m,radius,nsample,new_xyz,xyz,idx,counts)
 * (ball_query.cpp:6-8, kernel ball_query_gpu.cu:9-47).  new_xyz (b,m,3), xyz (b,n,3) ->
 * idx (b,m,nsample) int32, counts (b,m) int32.  Bit-exact.  Rows without any neighbour are all 0. */
int pdr_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                   const float *xyz, int *idx, int *counts, void *stream);

/* group_points: points (b,c,n), idx (b,npoints,nsample) -> out (b,c,npoints,nsample).
 * group_points.cpp:4-6, group_points_gpu.cu:8-40. */
int pdr_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                     const int *idx, float *out, void *stream);
int pdr_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                          const int *idx, float *grad_points, void *stream);

/* three_nn: unknown (b,n,3), known (b,m,3) -> dist2 (b,n,3) SQUARED distances, idx (b,n,3).
 * interpolate.cpp:4-6, interpolate_gpu.cu:9-68.  Bit-exact indices and distances. */
int pdr_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2,
                 int *idx, void *stream);
/* three_interpolate: points (b,c,m), idx (b,n,3), weight (b,n,3) -> out (b,c,n).
 * interpolate.cpp:7-9, interpolate_gpu.cu:72-111. */
int pdr_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx,
                          const float *weight, float *out, void *stream);
int pdr_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx,
                               const float *weight, float *grad_points, void *stream);

/* ---- pytorch3d.ops.knn surface ---------------------------------------------------------------- */

/* knn_points: p1 (b,P1,3), p2 (b,P2,3) -> dists (b,P1,K) squared L2 ascending, idx (b,P1,K) int64;
 * ties keep the lower index first; K <= 64.  If K > P2 the tail is (0, 0). */
int pdr_knn_points(int b, int p1, int p2, int K, const float *x, const float *y, float *dists,
                   int64_t *idx, void *stream);

/* ---- Chamfer / F1 (pointnet2/chamfer_loss_new.py:219-256) -------------------------------------- */

/* Fused Chamfer_F1.forward(xyz1=output (b,n,3), xyz2=gt (b,m,3)) -> cd_p, cd_t, f1, each (b).
 * dist1 (b,m) [gt -> output] and dist2 (b,n) [output -> gt] are optional outputs (NULL to skip).
 * workspace: pdr_chamfer_f1_workspace_bytes(b,n,m) bytes of device scratch. */
size_t pdr_chamfer_f1_workspace_bytes(int b, int n, int m);
int pdr_chamfer_f1(int b, int n, int m, const float *xyz1, const float *xyz2, float f1_threshold,
                   float *cd_p, float *cd_t, float *f1, float *dist1, float *dist2, void *workspace,
                   size_t workspace_bytes, void *stream);

/* NmDistance of the vendored chamfer3D (ChamferDistancePytorch/chamfer3D/chamfer3D.cu:12-147):
 * dist1/idx1 (b,n): nearest of xyz2 for each xyz1 point; dist2/idx2 (b,m) the converse.
 * First minimum wins; bit-exact with that kernel. */
int pdr_nm_distance(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist1,
                    int *idx1, float *dist2, int *idx2, void *stream);

/* Backward of NmDistance (chamfer_cuda_backward, chamfer3D.cu:155-196): grad_xyz1 (b,n,3), grad_xyz2 (b,m,3) from the
 * gradients of dist1 (b,n) / dist2 (b,m) and the saved argmin indices.  Fully written (no memset needed); the scatter
 * term is accumulated in index order instead of by atomicAdd, so the result is deterministic. */
int pdr_nm_distance_grad(int b, int n, int m, const float *xyz1, const float *xyz2, const float *grad_dist1,
                         const int *idx1, const float *grad_dist2, const int *idx2, float *grad_xyz1,
                         float *grad_xyz2, void *stream);

/* ---- emd_cuda (PytorchEMD/cuda/emd_kernel.cu) -------------------------------------------------- */

/* temp: pdr_emd_workspace_bytes(b,n,m) bytes of device scratch (the reference allocates
 * (b, 2(n+m)) floats, emd_kernel.cu:186). */
size_t pdr_emd_workspace_bytes(int b, int n, int m);
/* approxmatch_forward: xyz1 (b,n,3), xyz2 (b,m,3) -> match (b,m,n).  emd_kernel.cu:29-196. */
int pdr_emd_approxmatch(int b, int n, int m, const float *xyz1, const float *xyz2, float *match,
                        void *temp, size_t temp_bytes, void *stream);
/* matchcost_forward -> cost (b) = sum d^2 * match (NOT yet divided by max(n,m)). :204-282. */
int pdr_emd_matchcost(int b, int n, int m, const float *xyz1, const float *xyz2, const float *match,
                      float *cost, void *temp, size_t temp_bytes, void *stream);
/* fused approxmatch + matchcost without materialising `match` (b*m*n floats never touch HBM). */
int pdr_emd_cost(int b, int n, int m, const float *xyz1, const float *xyz2, float *cost, void *temp,
                 size_t temp_bytes, void *stream);
/* matchcost_backward: grad_cost (b) -> grad1 (b,n,3), grad2 (b,m,3).  emd_kernel.cu:290-401. */
int pdr_emd_matchcost_backward(int b, int n, int m, const float *grad_cost, const float *xyz1,
                               const float *xyz2, const float *match, float *grad1, float *grad2,
                               void *stream);

/* ---- DDPM reverse-step elementwise math (pointnet2/util.py:242-249) ---------------------------- */

/* x <- (x - c_eps*eps) * inv_sqrt_alpha + sigma * z, over `count` floats.  z is either `noise`
 * (injected, parity mode) or, when noise == NULL, N(0,1) from Philox4x32-10(seed, offset) generated
 * in the kernel (no host RNG, no H2D per step; reference: util.py:118-123 draws on the CPU). */
int pdr_ddpm_update(size_t count, float *x, const float *eps, float c_eps, float inv_sqrt_alpha,
                    float sigma, const float *noise, uint64_t seed, uint64_t offset, void *stream);
/* General form used by FastDPM (pointnet2/util_fastdpmv2.py:364-373):
 * x <- x*scale_x + eps*scale_eps + sigma*z. */
int pdr_affine_noise_update(size_t count, float *x, const float *eps, float scale_x, float scale_eps,
                            float sigma, const float *noise, uint64_t seed, uint64_t offset,
                            void *stream);
/* fill with N(0,1) (x_T), same generator. */
int pdr_normal_fill(size_t count, float *x, uint64_t seed, uint64_t offset, void *stream);

/* point_upsample (pointnet2/models/point_upsample_module.py:4-27), the output stage of the refinement
 * network.  coarse (b,n,3); displacement (b,n,3*factor) when include_centre != 0, else (b,n,3*(factor+1));
 * columns 0:3 move the coarse point (mid = coarse + d*out_scale), the remaining reps = factor-1 / factor
 * triples are grid offsets: up[j] = mid + (d_j*grid_scale)*out_scale with grid_scale = 1/sqrt(factor).
 * refined (b, n*reps [+ n], 3): point-major copies, then (include_centre) the n mid points;
 * intermediate (b,n,3) = mid, may be NULL.  Separate roundings (no FMA), bit-identical to the reference. */
int pdr_point_upsample(int b, int n, int factor, int include_centre, const float *coarse,
                       const float *displacement, float grid_scale, float out_scale, float *refined,
                       float *intermediate, void *stream);

/* ================================================================================================
 * Fused denoiser primitives (channels-LAST activations: a tensor (B, P, K, C) is the row-major matrix
 * [B*P*K, ld] with ld >= C, ld % 4 == 0 and pad columns holding zeros).
 * They replace, for the warm eps_theta step, the per-layer cuDNN/ATen launches of
 *   Mlp_plus_t_emb / build_shared_mlp / MyGroupNorm   pointnet2_ops/pointnet2_modules.py:23-174
 *   AttentionModule                                    pointnet2_ops/attention.py:35-96
 *   QueryAndGroup / group_knn                          pointnet2_ops/pointnet2_utils.py:307-438,487-514
 * ============================================================================================== */

/* prologue applied to A while it is loaded (x = A[row, c], b = sample of the row):
 *   PDR_PRO_NONE     x
 *   PDR_PRO_GN_RELU  relu(x*sc[b,c] + sh[b,c])      (conv -> GroupNorm -> ReLU stacks)
 *   PDR_PRO_RELU_GN  relu(x)*sc[b,c] + sh[b,c]      (AttentionModule.weight_conv: ReLU -> GroupNorm -> conv)
 * then  + add[b,c] (per-sample embedding, may be NULL)  + R[row,c] (residual, may be NULL). */
#define PDR_PRO_NONE 0
#define PDR_PRO_GN_RELU 1
#define PDR_PRO_RELU_GN 2

typedef struct PdrGemmArgs {
  /* C[M, N] = pro(A)[M, K] * W[N, K]^T + bias[N] + rowadd[(row / rowadd_div), N] */
  const float *A; int lda; int K;        /* K and lda multiples of 4 */
  const float *W; int ldw;               /* (N, ldw) row-major = Conv2d weight (Cout, Cin) zero-padded */
  const float *bias;                     /* (N) or NULL */
  float *C; int ldc; int N;              /* columns [N, ldc_zero_to) are written as zeros */
  int ldc_zero_to;
  int batch; int rows_per_sample;        /* M = batch * rows_per_sample; tiles never straddle samples */
  int pro_mode;
  const float *sc; const float *sh; int ld_scsh;   /* (batch, ld_scsh), ld_scsh % 4 == 0 */
  const float *add; int ld_add;          /* (batch, ld_add) or NULL */
  const float *R; int ldr;               /* (M, ldr) or NULL */
  const float *rowadd; int ld_rowadd; int rowadd_div;   /* or NULL */
  /* per-tile column statistics for the GroupNorm that follows: (batch*tiles_per_sample, N, 4) =
   * sum y, sum y^2, sum relu(y), sum relu(y)^2 over the tile's valid rows; NULL = not needed */
  float *stats;
  int use_tf32;                          /* 0: fp32 SIMT FMA; 1: tensor cores (TF32 inputs, fp32 accumulate) */
  /* hint: statistics nobody will read.  bit 0: the consumer does not need (sum y, sum y^2); bit 1: it does not need
   * the relu pair.  Slots a kernel skips hold zeros; a kernel may ignore the hint and compute everything. */
  int stats_skip;
  /* gathered A operand (tensor-core path, no prologue / add / R): when a_rows != NULL, row r of A is
   *   [ A[a_rows[r], 0:k_split] | A2[r, 0:K-k_split] ]        (a_rows[r] < 0: the first part is zeros)
   * i.e. the grouped tensor of QueryAndGroup / group_knn is assembled while the GEMM loads its operand instead of
   * being written to and re-read from HBM: A is the (points, lda) feature table, a_rows the flat neighbour row of
   * every grouped row (pdr_group_src_rows), A2 the (M, lda2) geometric channels (pdr_group_ball / _knn with C = 0).
   * k_split, lda2 multiples of 4; A2 16-byte aligned. */
  const int *a_rows; const float *A2; int lda2; int k_split;
  /* pooling epilogue (tensor-core path): when pool_K > 0 this GEMM computes the attention SCORES and, instead of
   * storing them (C may be NULL), pools straight away -- pdr_attention_pool fused into the epilogue:
   *   pool_out[point, n] = sum_k softmax_k(score[point*K + k, n] masked to k < max(count[point], 1))
   *                              * relu(pool_V[point*K + k, n] * pool_sc[b, n] + pool_sh[b, n])
   * pool_K in {8, 16, 32} and divides rows_per_sample; stats and rowadd must be NULL. */
  /* raw gathered K tail (tensor-core path, with a prologue, R == NULL): when tail_rows != NULL only the first k_pro
   * columns of the operand are A with the prologue applied; columns [k_pro, K) are, untransformed,
   *   [ T[tail_rows[r], 0:t_split] | T2[r, 0:K-k_pro-t_split] ]      (tail_rows[r] < 0: zeros)
   * i.e. a second, gathered operand contracted in the same accumulator -- used to fold the residual convolution of
   * Mlp_plus_t_emb into the values GEMM: V = Wv.act(y) + (Wv.Wres).X0.  k_pro multiple of 32; t_split, ldt, ldt2
   * multiples of 4; T, T2 16-byte aligned. */
  const int *tail_rows; const float *T; int ldt; const float *T2; int ldt2; int t_split; int k_pro;
  int pool_K; const float *pool_V; int pool_ldv; const float *pool_sc; const float *pool_sh; int pool_ld_scsh;
  const int *pool_counts; float *pool_out; int pool_ldo;
  /* tensor-core path: upper bound on the CTAs of the persistent grid (0 = one per SM).  A caller that runs another
   * kernel next to this GEMM on a second stream (the engine's geometry chain, PDR_GEOM_OVERLAP) leaves it some SMs:
   * CTAs of this kernel take a whole SM each and never share it. */
  int max_ctas;
  /* 1: W (and bias) are not written by any kernel still in flight on the stream (inference weights).  The tensor-core kernel
   * is launched as a programmatic dependent of the kernel in front of it and may then stage a resident W while that kernel
   * drains; with 0 it touches nothing before the previous kernel has completed. */
  int w_static;
  /* gathered A: number of rows of the table A (every a_rows[r] < table_rows).  > 0 lets the tensor-core kernel fetch whole
   * 32-column chunks of the table part with TMA tile::gather4; 0 = unknown (cp.async pieces). */
  int table_rows;
  /* stats_skip per 32-column block of the output: bits 2 j, 2 j + 1 = the stats_skip bits of columns [32 j, 32 j + 32), j < 32
   * (OR-ed with stats_skip; a merged GEMM whose column ranges feed different normalisations computes each pair only where
   * some consumer reads it). */
  unsigned long long stats_skip_blocks;
} PdrGemmArgs;
int pdr_gemm_tile_rows(void);            /* rows per tile (tiles_per_sample = ceil(rows_per_sample / this)) */
int pdr_gemm_fused(const PdrGemmArgs *args, void *stream);

/* GroupNorm statistics -> per-sample per-channel affine (sc, sh), for a channel concatenation of up to
 * PDR_GN_MAX_SOURCES sources (e.g. [query conv | key conv] in AttentionModule, the key conv possibly produced by
 * several GEMM calls).  Channels >= gn_channels pass through
 * (sc = 1, sh = 0; MyGroupNorm, attention.py:6-23); pad columns of sc/sh are never written (keep them 0). */
#define PDR_GN_MAX_SOURCES 4
typedef struct PdrGnSource {
  const float *stats; int tiles_per_sample; int ld_stats;  /* (batch*tiles, ld_stats, 4) from pdr_gemm_fused */
  int col0; int ncols;                                     /* columns of that GEMM output used here */
  int out_col0;                                            /* where these channels start in sc/sh (each source
                                                              may be padded to a multiple of 4 columns) */
  int use_relu;                                            /* take the statistics of relu(y) instead of y */
  int rows; float mult;                                    /* rows the stats cover per sample; multiplicity of
                                                              each row in the normalised tensor (K for the query
                                                              of AttentionModule, which is expanded over K) */
} PdrGnSource;
typedef struct PdrGnArgs {
  PdrGnSource src[PDR_GN_MAX_SOURCES]; int nsrc;
  int batch; int channels; int gn_channels; int groups;
  const float *gamma; const float *beta; float eps;        /* (gn_channels) */
  float *sc; float *sh; int ld_out;                        /* (batch, ld_out) */
} PdrGnArgs;
int pdr_gn_finalize(const PdrGnArgs *args, void *stream);
/* `count` (1 or 2) independent finalisations in ONE launch: args[0 .. count). */
int pdr_gn_finalize_batch(const PdrGnArgs *args, int count, void *stream);

/* `round_tf32` (pdr_affine_rows, pdr_attention_pool, pdr_gather_rows, pdr_group_geo_*): round the values written to the
 * nearest TF32 number.  Set by callers whose output is read RAW by a tensor-core GEMM (gathered A operand, raw K tail,
 * prologue-free A): the tensor core truncates fp32 operands, so unrounded tables would carry a toward-zero bias. */
/* out[row, c] = pro(x[row, c]) (+add +R) materialised; same prologue semantics as the GEMM. */
int pdr_affine_rows(int batch, int rows_per_sample, int C, const float *x, int ldx, int pro_mode,
                    const float *sc, const float *sh, int ld_scsh, const float *add, int ld_add, const float *R,
                    int ldr, float *out, int ldo, int round_tf32, void *stream);

/* Soft-attention pooling over the K neighbours (attention.py:85-96):
 * out[b,p,c] = sum_k softmax_k(S[b,p,k,c] masked to k < max(count[b,p],1)) * relu(V[b,p,k,c]*sc[b,c]+sh[b,c]).
 * counts == NULL means 'all'.  S, V: (B*P*K, ld); out: (B*P, ldo) written at column offset 0 of `out`. */
int pdr_attention_pool(int batch, int P, int K, int C, const float *S, int lds, const float *V, int ldv,
                       const float *sc, const float *sh, int ld_scsh, const int *counts, float *out, int ldo,
                       int round_tf32, void *stream);

/* Ball-query grouping into channels-last rows [feat(C) | rel(3) | abs(3) | centre(3) | 0-pad] with the
 * subset=False fill rule (pointnet2_utils.py:376-410): counts == 0 -> feat = 0, abs = centre, rel = 0.
 * feat (B, n, ldf), xyz (B, n, 3), centres (B, P, 3), idx (B, P, K) int32, counts (B, P) or NULL. */
int pdr_group_ball(int batch, int n, int P, int K, int C, const float *feat, int ldf, const float *xyz,
                   const float *centres, const int *idx, const int *counts, int fill_missing, float *out,
                   int ldo, void *stream);
/* kNN grouping rows [feat(C) | d2 | w | nn_abs(3) | nn_rel(3) | x(3) | 0-pad], w = normalised 1/(d2+1e-8)
 * (pointnet2_utils.py:487-514).  idx (B, P, K) int64 and dists (B, P, K) from pdr_knn_points. */
int pdr_group_knn(int batch, int n, int P, int K, int C, const float *feat, int ldf, const float *y,
                  const float *x, const int64_t *idx, const float *dists, float *out, int ldo, void *stream);
/* The two inputs of the gathered-A GEMM in one pass (thread per grouped row): geo (B*P*K, 12) =
 * [rel(3) | abs(3) | centre(3) | 0 0 0] (ball) or [d2 | w | nn_abs(3) | nn_rel(3) | x(3) | 0] (kNN), bit-identical to
 * the geometric channels pdr_group_ball / pdr_group_knn write, and src_row as pdr_group_src_rows. */
int pdr_group_geo_ball(int batch, int n, int P, int K, const float *xyz, const float *centres, const int *idx,
                       const int *counts, int fill_missing, float *geo, int *src_row, int round_tf32, void *stream);
int pdr_group_geo_knn(int batch, int n, int P, int K, const float *y, const float *x, const int64_t *idx,
                      const float *dists, float *geo, int *src_row, int round_tf32, void *stream);
/* Flat feature-table row of every grouped row: src_row[(b*P+p)*K+k] = b*n + idx[b,p,k], or -1 where the subset=False
 * fill rule zeroes the features (fill_missing != 0 and counts[b,p] == 0).  idx is int32 (ball query) or, with
 * idx_is_int64 != 0, int64 (pdr_knn_points).  Feeds PdrGemmArgs.a_rows. */
int pdr_group_src_rows(int batch, int n, int P, int K, const void *idx, int idx_is_int64, const int *counts,
                       int fill_missing, int *src_row, void *stream);
/* out[b, j, 0:C] = src[b, idx[b,j], 0:C] (rows); idx == NULL copies row j.  Used for FPS centre features and
 * for placing a feature block into a column slice of a wider buffer (free concatenation). */
int pdr_gather_rows(int batch, int n, int P, int C, const float *src, int lds, const int *idx, float *out,
                    int ldo, int round_tf32, void *stream);

/* ================================================================================================
 * Fused stage: one grouped stage of the network -- the chain of 1x1 convolutions over the same grouped rows that
 * Mlp_plus_t_emb + AttentionModule apply (pointnet2_ops/pointnet2_modules.py:57-65,129-174, attention.py:70-96) --
 * evaluated in sweeps that keep every intermediate on chip (csrc/stage_chain.cu).  A sweep is a step program:
 * per 128-row tile, step s issues its MMAs (accumulators in TMEM), then its epilogue operations run over blocks of 32
 * accumulator columns:
 *   XFORM  y = acc + bias (+ rowadd[row / group_k]);  t = pro(y; sc[b], sh[b]) + emb[b], rounded to TF32, written back
 *          to TMEM as the A operand of a later MMA (never stored)
 *   STATS  per-tile column statistics of y for the GroupNorm that follows, same layout as PdrGemmArgs.stats:
 *          stats[(tile * stats_n + stat_col0 + c) * 4 + {sum, sum^2, relu-sum, relu-sum^2}]
 *   POOL   out[point, c] = sum_k softmax_k(y[point*group_k + k, c] masked to k < max(counts[point], 1))
 *                               * relu((acc_v + v_bias) * v_sc[b] + v_sh[b])         (pdr_attention_pool)
 * All column numbers are relative to the 256 TMEM columns of the tile's group; widths are padded to multiples of 32
 * (pad columns of an XFORM come out as zeros because the per-column arrays are zero there).  rows_per_sample % 128 == 0,
 * group_k in {8, 16, 32}.
 * ============================================================================================== */
#define PDR_CHAIN_MAX_STEPS 4
#define PDR_CHAIN_MAX_MMA 4
#define PDR_CHAIN_MAX_EPI 3
#define PDR_CHAIN_XFORM 1
#define PDR_CHAIN_STATS 2
#define PDR_CHAIN_POOL 3

typedef struct PdrChainMma {
  int d_col; int n;                  /* accumulator columns [d_col, d_col + n), n a multiple of 16, <= 256 */
  int a_tmem; int a_col; int k;      /* A = TMEM columns [a_col, a_col + k) (a_tmem != 0) or columns [a_col, a_col + k) of the
                                        gathered X0 tile in shared memory; k a multiple of 8 */
  int w_off; int w_rows;             /* B: matrix at byte w_off of the weight image, w_rows rows (see w_image) */
  int w_row0; int w_k0;              /* first row (multiple of 8) and first column (multiple of 8) used */
  int accumulate;                    /* 0: the first K step overwrites the accumulator */
} PdrChainMma;

typedef struct PdrChainEpi {
  int kind; int d_col; int ncols;                          /* ncols valid columns from d_col; every per-column array below is
                                                              read in whole 32-column blocks: readable, and zero, up to the
                                                              next multiple of 32 */
  const float *bias;                                       /* (32-padded ncols) or NULL */
  const float *rowadd; int ld_rowadd;                      /* (points, ld_rowadd) or NULL */
  int pro_mode; const float *sc; const float *sh; int ld_scsh;   /* XFORM: PDR_PRO_*, (batch, ld_scsh) */
  const float *emb; int ld_emb;                            /* XFORM: (batch, ld_emb) or NULL */
  int a_col;                                               /* XFORM: destination TMEM columns */
  int stat_col0; int stat_skip;                            /* STATS: column offset in `stats`; bit 0 / 1 as stats_skip */
  int v_col; const float *v_bias; const float *v_sc; const float *v_sh; int v_ld_scsh;   /* POOL: the values */
} PdrChainEpi;

typedef struct PdrChainStep {
  int n_mma; int n_epi;
  int release_x0;                    /* the X0 tile is not read after this step's MMAs (exactly one step sets it) */
  PdrChainMma mma[PDR_CHAIN_MAX_MMA];
  PdrChainEpi epi[PDR_CHAIN_MAX_EPI];
} PdrChainStep;

typedef struct PdrChainArgs {
  /* gathered operand X0[r] = [ table[src_rows[r], 0:k_split] | geo[r, 0:k0-k_split] ]  (src_rows[r] < 0: zeros), as
   * PdrGemmArgs.a_rows: pdr_group_geo_ball / pdr_group_geo_knn produce src_rows and geo */
  const float *table; int ld_table; int k_split;
  const int *src_rows; const float *geo; int ld_geo; int k0;
  /* every weight matrix of the stage, TF32-rounded, in the shared-memory image the tensor core reads: per matrix
   * (rows R, a multiple of 32; K padded to a multiple of 32) chunk kc = columns [32 kc, 32 kc + 32) is R rows of 128
   * bytes, the 16-byte piece p of row r stored at piece p ^ (r & 7) (K-major SWIZZLE_128B); chunks follow each other.
   * w_bytes a multiple of 1024. */
  const void *w_image; int w_bytes;
  int batch; int rows_per_sample; int group_k;
  float *stats; int stats_n;         /* STATS target: (batch * rows_per_sample / 128, stats_n, 4) */
  unsigned int stats_relu_mask[4];   /* bit c: column c of `stats` carries the relu pair (slots 2, 3), else the plain pair
                                        (slots 0, 1); the other pair is written as zeros */
  const int *counts; float *out; int ld_out;      /* POOL: counts (points) or NULL, out (points, ld_out) */
  int max_ctas;                      /* upper bound on the CTAs of the persistent grid (0 = one per SM), as PdrGemmArgs.max_ctas */
  int round_out;                     /* POOL writes `out` rounded to TF32 (nearest), like pdr_attention_pool's round_tf32 */
  int n_steps; PdrChainStep steps[PDR_CHAIN_MAX_STEPS];
} PdrChainArgs;
int pdr_stage_chain_tile_rows(void);
int pdr_stage_chain(const PdrChainArgs *args, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PDR_B200_H_ */
