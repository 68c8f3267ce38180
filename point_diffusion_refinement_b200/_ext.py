"""Drop-in for ``pointnet2_ops._ext`` (reference: pointnet2_ops_lib/pointnet2_ops/_ext-src/src/
bindings.cpp:6-19): the same nine names, positional signatures, dtypes and return containers, backed by
the hand-written sm_100a kernels of libpdr_b200.so through its C ABI.

Install with ``point_diffusion_refinement_b200.dropin.install()`` *before* ``pointnet2_ops`` is imported
(the reference imports it at pointnet2_utils.py:9-10).  CPU tensors are rejected exactly like the
reference (``AT_ASSERT(false, "CPU not supported")``, sampling.cpp:34).
"""
import contextlib
import ctypes

import torch

from . import _lib
from ._lib import call, check_cuda_f32, check_cuda_i32, dptr, stream_ptr


@contextlib.contextmanager
def _on_device_of(t):
    if t.device.index is not None and t.device.index != torch.cuda.current_device():
        with torch.cuda.device(t.device):
            yield
    else:
        yield


def furthest_point_sampling(points, nsamples):
    """(B,N,3) f32 -> (B,nsamples) int32.  sampling.cpp:66-87."""
    check_cuda_f32(points, "points")
    b, n, d = points.shape
    if d != 3:
        raise RuntimeError("points must be (B, N, 3)")
    nsamples = int(nsamples)
    out = torch.empty((b, nsamples), dtype=torch.int32, device=points.device)
    temp = None
    if n > _lib.lib().pdr_fps_max_onchip_points():
        temp = torch.empty((b, n), dtype=torch.float32, device=points.device)
    with _on_device_of(points):
        call("pdr_furthest_point_sampling", b, n, nsamples, dptr(points), dptr(temp), dptr(out),
             stream_ptr(points))
    return out


def gather_points(points, idx):
    """(B,C,N) f32, (B,M) int32 -> (B,C,M).  sampling.cpp:15-38."""
    check_cuda_f32(points, "points")
    check_cuda_i32(idx, "idx")
    b, c, n = points.shape
    m = idx.shape[1]
    out = torch.empty((b, c, m), dtype=torch.float32, device=points.device)
    with _on_device_of(points):
        call("pdr_gather_points", b, c, n, m, dptr(points), dptr(idx), dptr(out), stream_ptr(points))
    return out


def gather_points_grad(grad_out, idx, n):
    check_cuda_f32(grad_out, "grad_out")
    check_cuda_i32(idx, "idx")
    b, c, m = grad_out.shape
    out = torch.empty((b, c, int(n)), dtype=torch.float32, device=grad_out.device)
    with _on_device_of(grad_out):
        call("pdr_gather_points_grad", b, c, int(n), m, dptr(grad_out), dptr(idx), dptr(out),
             stream_ptr(grad_out))
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    """new_xyz (B,M,3), xyz (B,N,3) -> (idx (B,M,nsample) int32, counts (B,M) int32).
    Argument order as in ball_query.cpp:10 (centres first)."""
    check_cuda_f32(new_xyz, "new_xyz")
    check_cuda_f32(xyz, "xyz")
    b, n, _ = xyz.shape
    m = new_xyz.shape[1]
    nsample = int(nsample)
    idx = torch.empty((b, m, nsample), dtype=torch.int32, device=xyz.device)
    counts = torch.empty((b, m), dtype=torch.int32, device=xyz.device)
    with _on_device_of(xyz):
        call("pdr_ball_query", b, n, m, ctypes.c_float(radius), nsample, dptr(new_xyz), dptr(xyz), dptr(idx),
             dptr(counts), stream_ptr(xyz))
    return idx, counts


def group_points(points, idx):
    """(B,C,N) f32, (B,npoints,nsample) int32 -> (B,C,npoints,nsample).  group_points.cpp:12-36."""
    check_cuda_f32(points, "points")
    check_cuda_i32(idx, "idx")
    b, c, n = points.shape
    _, npoints, nsample = idx.shape
    out = torch.empty((b, c, npoints, nsample), dtype=torch.float32, device=points.device)
    with _on_device_of(points):
        call("pdr_group_points", b, c, n, npoints, nsample, dptr(points), dptr(idx), dptr(out),
             stream_ptr(points))
    return out


def group_points_grad(grad_out, idx, n):
    check_cuda_f32(grad_out, "grad_out")
    check_cuda_i32(idx, "idx")
    b, c, npoints, nsample = grad_out.shape
    out = torch.empty((b, c, int(n)), dtype=torch.float32, device=grad_out.device)
    with _on_device_of(grad_out):
        call("pdr_group_points_grad", b, c, int(n), npoints, nsample, dptr(grad_out), dptr(idx), dptr(out),
             stream_ptr(grad_out))
    return out


def three_nn(unknowns, knows):
    """unknown (B,n,3), known (B,m,3) -> [dist2 (B,n,3) f32 (SQUARED), idx (B,n,3) int32].
    interpolate.cpp:14-40; the sqrt is applied by the Python caller (pointnet2_utils.py:153)."""
    check_cuda_f32(unknowns, "unknowns")
    check_cuda_f32(knows, "knows")
    b, n, _ = unknowns.shape
    m = knows.shape[1]
    dist2 = torch.empty((b, n, 3), dtype=torch.float32, device=unknowns.device)
    idx = torch.empty((b, n, 3), dtype=torch.int32, device=unknowns.device)
    with _on_device_of(unknowns):
        call("pdr_three_nn", b, n, m, dptr(unknowns), dptr(knows), dptr(dist2), dptr(idx), stream_ptr(unknowns))
    return [dist2, idx]


def three_interpolate(points, idx, weight):
    """points (B,C,m), idx (B,n,3) int32, weight (B,n,3) -> (B,C,n).  interpolate.cpp:42-70."""
    check_cuda_f32(points, "points")
    check_cuda_i32(idx, "idx")
    check_cuda_f32(weight, "weight")
    b, c, m = points.shape
    n = idx.shape[1]
    out = torch.empty((b, c, n), dtype=torch.float32, device=points.device)
    with _on_device_of(points):
        call("pdr_three_interpolate", b, c, m, n, dptr(points), dptr(idx), dptr(weight), dptr(out),
             stream_ptr(points))
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    check_cuda_f32(grad_out, "grad_out")
    check_cuda_i32(idx, "idx")
    check_cuda_f32(weight, "weight")
    b, c, n = grad_out.shape
    out = torch.empty((b, c, int(m)), dtype=torch.float32, device=grad_out.device)
    with _on_device_of(grad_out):
        call("pdr_three_interpolate_grad", b, c, n, int(m), dptr(grad_out), dptr(idx), dptr(weight), dptr(out),
             stream_ptr(grad_out))
    return out
