// TEST INFRASTRUCTURE ONLY (never imported by the product package).
//
// extern "C" forwarding shim around the REFERENCE's own CUDA launchers, which are compiled
// unmodified from /root/reference by oracle/build_ref.sh.  It lets tests/ and bench.py call the
// reference kernels on raw device pointers through ctypes, so our sm_100a kernels can be compared
// with the reference's bit for bit on the B200 box.
//
// Forward declarations below restate the signatures found at:
//   pointnet2_ops_lib/pointnet2_ops/_ext-src/src/sampling.cpp:4-13
//   pointnet2_ops_lib/pointnet2_ops/_ext-src/src/ball_query.cpp:6-8
//   pointnet2_ops_lib/pointnet2_ops/_ext-src/src/group_points.cpp:4-6
//   pointnet2_ops_lib/pointnet2_ops/_ext-src/src/interpolate.cpp:4-12
//   PytorchEMD/cuda/emd.cpp:8-16
//   pointnet2/models/pvd/metrics/ChamferDistancePytorch/chamfer3D/chamfer3D.cu:136
#include <torch/torch.h>
#include <vector>

void gather_points_kernel_wrapper(int b, int c, int n, int npoints, const float *points,
                                  const int *idx, float *out);
void furthest_point_sampling_kernel_wrapper(int b, int n, int m, const float *dataset, float *temp,
                                            int *idxs);
void query_ball_point_kernel_wrapper(int b, int n, int m, float radius, int nsample,
                                     const float *new_xyz, const float *xyz, int *idx, int *counts);
void group_points_kernel_wrapper(int b, int c, int n, int npoints, int nsample, const float *points,
                                 const int *idx, float *out);
void three_nn_kernel_wrapper(int b, int n, int m, const float *unknown, const float *known,
                             float *dist2, int *idx);
void three_interpolate_kernel_wrapper(int b, int c, int m, int n, const float *points,
                                      const int *idx, const float *weight, float *out);
at::Tensor ApproxMatchForward(const at::Tensor xyz1, const at::Tensor xyz2);
at::Tensor MatchCostForward(const at::Tensor xyz1, const at::Tensor xyz2, const at::Tensor match);
int chamfer_cuda_forward(at::Tensor xyz1, at::Tensor xyz2, at::Tensor dist1, at::Tensor dist2,
                         at::Tensor idx1, at::Tensor idx2);

static at::Tensor wrap_f(const float *p, std::vector<int64_t> sizes) {
  return torch::from_blob(const_cast<float *>(p), sizes,
                          torch::TensorOptions().dtype(torch::kFloat32).device(torch::kCUDA));
}
static at::Tensor wrap_i(const int *p, std::vector<int64_t> sizes) {
  return torch::from_blob(const_cast<int *>(p), sizes,
                          torch::TensorOptions().dtype(torch::kInt32).device(torch::kCUDA));
}

extern "C" {

// temp must be pre-filled with 1e10 by the caller (sampling.cpp:74-76 does torch::full).
void ref_fps(int b, int n, int m, const float *xyz, float *temp, int *idx) {
  furthest_point_sampling_kernel_wrapper(b, n, m, xyz, temp, idx);
}
void ref_gather_points(int b, int c, int n, int m, const float *points, const int *idx, float *out) {
  gather_points_kernel_wrapper(b, c, n, m, points, idx, out);
}
// idx and counts must be zero-filled by the caller (ball_query.cpp:21-27).
void ref_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                    const float *xyz, int *idx, int *counts) {
  query_ball_point_kernel_wrapper(b, n, m, radius, nsample, new_xyz, xyz, idx, counts);
}
void ref_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                      const int *idx, float *out) {
  group_points_kernel_wrapper(b, c, n, npoints, nsample, points, idx, out);
}
void ref_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2,
                  int *idx) {
  three_nn_kernel_wrapper(b, n, m, unknown, known, dist2, idx);
}
void ref_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx,
                           const float *weight, float *out) {
  three_interpolate_kernel_wrapper(b, c, m, n, points, idx, weight, out);
}
// match: (b, m, n) output, cost: (b) output.  The reference launches on the legacy default stream.
void ref_emd(int b, int n, int m, const float *xyz1, const float *xyz2, float *match, float *cost) {
  at::Tensor t1 = wrap_f(xyz1, {b, n, 3}), t2 = wrap_f(xyz2, {b, m, 3});
  at::Tensor mt = ApproxMatchForward(t1, t2);
  at::Tensor c = MatchCostForward(t1, t2, mt);
  if (match) wrap_f(match, {b, m, n}).copy_(mt);
  wrap_f(cost, {b}).copy_(c);
}
// chamfer3D NmDistance both directions (dist1/idx1: (b,n); dist2/idx2: (b,m)).
int ref_chamfer3d(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist1,
                  float *dist2, int *idx1, int *idx2) {
  return chamfer_cuda_forward(wrap_f(xyz1, {b, n, 3}), wrap_f(xyz2, {b, m, 3}), wrap_f(dist1, {b, n}),
                              wrap_f(dist2, {b, m}), wrap_i(idx1, {b, n}), wrap_i(idx2, {b, m}));
}
}
