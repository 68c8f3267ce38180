"""TEST INFRASTRUCTURE ONLY -- ctypes front end of oracle/_ref/libpdr_ref_cuda.so: the REFERENCE's own
CUDA kernels (sampling_gpu.cu, ball_query_gpu.cu, group_points_gpu.cu, interpolate_gpu.cu,
emd_kernel.cu, chamfer3D.cu) compiled unmodified for sm_100a by oracle/build_ref.sh.

Used on the GPU box as the bit-exactness comparator ("reference pointnet2_ops CUDA recompiled for
sm_100a") and for reference-vs-ours kernel timings.  The .so is built in the build container (where
/root/reference exists) and travels with the snapshot; `available()` is False if it is missing.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libpdr_ref_cuda.so")
_lib = None


def available():
    return os.path.exists(_PATH) and torch.cuda.is_available()


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_PATH)  # torch must already be imported (libc10/libtorch resolve through it)
    return _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def furthest_point_sampling(xyz, m):
    b, n, _ = xyz.shape
    idx = torch.zeros(b, m, dtype=torch.int32, device=xyz.device)
    temp = torch.full((b, n), 1e10, dtype=torch.float32, device=xyz.device)
    lib().ref_fps(b, n, int(m), _p(xyz), _p(temp), _p(idx))
    return idx


def gather_points(points, idx):
    b, c, n = points.shape
    m = idx.shape[1]
    out = torch.zeros(b, c, m, device=points.device)
    lib().ref_gather_points(b, c, n, m, _p(points), _p(idx), _p(out))
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    b, n, _ = xyz.shape
    m = new_xyz.shape[1]
    idx = torch.zeros(b, m, nsample, dtype=torch.int32, device=xyz.device)
    counts = torch.zeros(b, m, dtype=torch.int32, device=xyz.device)
    lib().ref_ball_query(b, n, m, ctypes.c_float(radius), int(nsample), _p(new_xyz), _p(xyz), _p(idx), _p(counts))
    return idx, counts


def group_points(points, idx):
    b, c, n = points.shape
    _, npoints, nsample = idx.shape
    out = torch.zeros(b, c, npoints, nsample, device=points.device)
    lib().ref_group_points(b, c, n, npoints, nsample, _p(points), _p(idx), _p(out))
    return out


def three_nn(unknown, known):
    b, n, _ = unknown.shape
    m = known.shape[1]
    d = torch.zeros(b, n, 3, device=unknown.device)
    i = torch.zeros(b, n, 3, dtype=torch.int32, device=unknown.device)
    lib().ref_three_nn(b, n, m, _p(unknown), _p(known), _p(d), _p(i))
    return d, i


def three_interpolate(points, idx, weight):
    b, c, m = points.shape
    n = idx.shape[1]
    out = torch.zeros(b, c, n, device=points.device)
    lib().ref_three_interpolate(b, c, m, n, _p(points), _p(idx), _p(weight), _p(out))
    return out


def emd(xyz1, xyz2, want_match=False):
    """-> cost (b) [not divided by max(n,m)], match (b,m,n) or None.  Launches on the legacy default
    stream like the reference (emd_kernel.cu:191,277)."""
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    cost = torch.zeros(b, device=xyz1.device)
    match = torch.zeros(b, m, n, device=xyz1.device) if want_match else None
    torch.cuda.synchronize()
    lib().ref_emd(b, n, m, _p(xyz1), _p(xyz2), _p(match) if want_match else None, _p(cost))
    torch.cuda.synchronize()
    return cost, match


def chamfer3d(xyz1, xyz2):
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    d1 = torch.zeros(b, n, device=xyz1.device); d2 = torch.zeros(b, m, device=xyz1.device)
    i1 = torch.zeros(b, n, dtype=torch.int32, device=xyz1.device); i2 = torch.zeros(b, m, dtype=torch.int32, device=xyz1.device)
    torch.cuda.synchronize()
    lib().ref_chamfer3d(b, n, m, _p(xyz1), _p(xyz2), _p(d1), _p(d2), _p(i1), _p(i2))
    torch.cuda.synchronize()
    return d1, d2, i1, i2
