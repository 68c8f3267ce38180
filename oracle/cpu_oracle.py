"""TEST INFRASTRUCTURE ONLY -- ctypes front end of oracle/libpdr_oracle.so (the CPU restatement).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module.  The product package never does.

Functions take and return CPU ``torch`` tensors and mirror the reference operator surface
(``pointnet2_ops._ext``: pointnet2_ops_lib/pointnet2_ops/_ext-src/src/bindings.cpp:6-19;
``emd_cuda``: PytorchEMD/cuda/emd.cpp:24-28; ``pytorch3d.ops.knn.knn_points``).
"""
import ctypes
import os
import subprocess
from collections import namedtuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpdr_oracle.so")
_lib = None


def build(force=False):
    """Compile oracle/pdr_oracle.c with gcc (see oracle/Makefile)."""
    src = os.path.join(_HERE, "pdr_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(src) > os.path.getmtime(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "libpdr_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _f(t):
    assert t.dtype == torch.float32 and t.is_contiguous() and t.device.type == "cpu"
    return ctypes.c_void_p(t.data_ptr())


def _i(t):
    assert t.dtype in (torch.int32, torch.int64) and t.is_contiguous() and t.device.type == "cpu"
    return ctypes.c_void_p(t.data_ptr())


def set_threads(n):
    os.environ["OMP_NUM_THREADS"] = str(n)


def opt_n_threads(n):
    return lib().oracle_opt_n_threads(int(n))


def furthest_point_sampling(xyz, m):
    xyz = xyz.contiguous().float()
    b, n, _ = xyz.shape
    idx = torch.zeros(b, m, dtype=torch.int32)
    lib().oracle_fps(b, n, int(m), _f(xyz), _i(idx))
    return idx


def gather_points(points, idx):
    points = points.contiguous().float()
    idx = idx.contiguous().int()
    b, c, n = points.shape
    m = idx.shape[1]
    out = torch.zeros(b, c, m, dtype=torch.float32)
    lib().oracle_gather_points(b, c, n, m, _f(points), _i(idx), _f(out))
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    new_xyz = new_xyz.contiguous().float()
    xyz = xyz.contiguous().float()
    b, n, _ = xyz.shape
    m = new_xyz.shape[1]
    idx = torch.zeros(b, m, nsample, dtype=torch.int32)
    counts = torch.zeros(b, m, dtype=torch.int32)
    lib().oracle_ball_query(b, n, m, ctypes.c_float(radius), int(nsample), _f(new_xyz), _f(xyz),
                            _i(idx), _i(counts))
    return idx, counts


def group_points(points, idx):
    points = points.contiguous().float()
    idx = idx.contiguous().int()
    b, c, n = points.shape
    _, npoints, nsample = idx.shape
    out = torch.zeros(b, c, npoints, nsample, dtype=torch.float32)
    lib().oracle_group_points(b, c, n, npoints, nsample, _f(points), _i(idx), _f(out))
    return out


def three_nn(unknown, known):
    unknown = unknown.contiguous().float()
    known = known.contiguous().float()
    b, n, _ = unknown.shape
    m = known.shape[1]
    dist2 = torch.zeros(b, n, 3, dtype=torch.float32)
    idx = torch.zeros(b, n, 3, dtype=torch.int32)
    lib().oracle_three_nn(b, n, m, _f(unknown), _f(known), _f(dist2), _i(idx))
    return dist2, idx


def three_interpolate(points, idx, weight):
    points = points.contiguous().float()
    idx = idx.contiguous().int()
    weight = weight.contiguous().float()
    b, c, m = points.shape
    n = idx.shape[1]
    out = torch.zeros(b, c, n, dtype=torch.float32)
    lib().oracle_three_interpolate(b, c, m, n, _f(points), _i(idx), _f(weight), _f(out))
    return out


_KNN = namedtuple("KNN", "dists idx knn")


def knn_gather(x, idx):
    """x (N,P2,C), idx (N,P1,K) int64 -> (N,P1,K,C)."""
    N, P1, K = idx.shape
    C = x.shape[2]
    return x.gather(1, idx.reshape(N, P1 * K, 1).expand(-1, -1, C)).reshape(N, P1, K, C)


def knn_points(p1, p2, lengths1=None, lengths2=None, K=1, return_nn=False):
    assert lengths1 is None and lengths2 is None
    p1 = p1.contiguous().float()
    p2 = p2.contiguous().float()
    b, n1, _ = p1.shape
    n2 = p2.shape[1]
    dists = torch.zeros(b, n1, K, dtype=torch.float32)
    idx = torch.zeros(b, n1, K, dtype=torch.int64)
    lib().oracle_knn(b, n1, n2, int(K), _f(p1), _f(p2), _f(dists), _i(idx))
    nn = knn_gather(p2, idx) if return_nn else None
    return _KNN(dists, idx, nn)


def nm_distance(xyz1, xyz2):
    """chamfer3D forward: returns dist1 (b,n), dist2 (b,m), idx1, idx2 (int32)."""
    xyz1 = xyz1.contiguous().float()
    xyz2 = xyz2.contiguous().float()
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    d1 = torch.zeros(b, n); i1 = torch.zeros(b, n, dtype=torch.int32)
    d2 = torch.zeros(b, m); i2 = torch.zeros(b, m, dtype=torch.int32)
    lib().oracle_nm_distance(b, n, _f(xyz1), m, _f(xyz2), _f(d1), _i(i1))
    lib().oracle_nm_distance(b, m, _f(xyz2), n, _f(xyz1), _f(d2), _i(i2))
    return d1, d2, i1, i2


def approxmatch_forward(xyz1, xyz2):
    xyz1 = xyz1.contiguous().float()
    xyz2 = xyz2.contiguous().float()
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    match = torch.zeros(b, m, n, dtype=torch.float32)
    lib().oracle_approxmatch(b, n, m, _f(xyz1), _f(xyz2), _f(match))
    return match


def matchcost_forward(xyz1, xyz2, match):
    xyz1 = xyz1.contiguous().float()
    xyz2 = xyz2.contiguous().float()
    match = match.contiguous().float()
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    cost = torch.zeros(b, dtype=torch.float32)
    lib().oracle_matchcost(b, n, m, _f(xyz1), _f(xyz2), _f(match), _f(cost))
    return cost


def emd_distance(xyz1, xyz2):
    """pointnet2/emd.py:7-21: cost / max(n, m)."""
    match = approxmatch_forward(xyz1, xyz2)
    cost = matchcost_forward(xyz1, xyz2, match)
    return cost / max(xyz1.shape[1], xyz2.shape[1])


def chamfer_f1(xyz1, xyz2, f1_threshold=1e-4):
    """pointnet2/chamfer_loss_new.py:219-256 (calc_cd(output=xyz1, gt=xyz2) -> chamfer(gt, output))."""
    d1 = knn_points(xyz2, xyz1, K=1).dists[..., 0]
    d2 = knn_points(xyz1, xyz2, K=1).dists[..., 0]
    cd_p = (torch.sqrt(d1).mean(1) + torch.sqrt(d2).mean(1)) / 2
    cd_t = d1.mean(1) + d2.mean(1)
    p1 = (d1 < f1_threshold).float().mean(1)
    p2 = (d2 < f1_threshold).float().mean(1)
    f1 = 2 * p1 * p2 / (p1 + p2)
    f1[torch.isnan(f1)] = 0
    return cd_p, cd_t, f1


def point_upsample(coarse, displacement, factor, include_centre, out_scale):
    """pointnet2/models/point_upsample_module.py:4-27 restated with explicit fp32 roundings (numpy float32
    arithmetic rounds every product and sum separately, as the reference's chain of torch kernels does)."""
    import numpy as np
    c = coarse.numpy().astype(np.float32)
    d = displacement.numpy().astype(np.float32)
    B, N, _ = c.shape
    g = np.float32(1 / np.sqrt(factor))            # :8 -- float64 scalar rounded to fp32 by the tensor op
    s = np.float32(out_scale)
    mid = c + d[:, :, 0:3] * s                     # :10-11
    reps = factor - 1 if include_centre else factor
    grid = (d[:, :, 3:] * g).reshape(B, N, reps, 3)    # :9, :15-18
    up = (mid[:, :, None, :] + grid * s).reshape(B, -1, 3)   # :20-22
    refined = np.concatenate([up, mid], axis=1) if include_centre else up   # :23-26
    return torch.from_numpy(np.ascontiguousarray(refined)), torch.from_numpy(np.ascontiguousarray(mid))


def mirror_and_concat(partial, axis=2, num_points=(2048, 3072)):
    """pointnet2/data_utils/mirror_partial.py:5-37 on the oracle's FPS / gather."""
    B, N, _ = partial.shape
    mirrored = partial.clone()
    mirrored[:, :, axis] = -mirrored[:, :, axis]                                   # :5-9
    ones = torch.ones(B, N, 1)
    concat = torch.cat([torch.cat([partial, ones], 2), torch.cat([mirrored, -ones], 2)], 1)   # :28-31
    out = [concat]
    for n in num_points:                                                           # :34-36, :11-20
        idx = furthest_point_sampling(concat[:, :, 0:3].contiguous(), n)
        out.append(gather_points(concat.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous())
    return tuple(out)
