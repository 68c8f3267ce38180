"""Chamfer distance / F1 behind the reference's ``chamfer_loss_new`` surface (reference:
pointnet2/chamfer_loss_new.py).  ``Chamfer_F1`` / ``calc_cd`` run ONE fused sm_100a launch pair for both
directions and the cd_p / cd_t / F1 reductions (pdr_chamfer_f1) instead of two pytorch3d ``knn_points``
calls plus ~10 reduction kernels.  ``chamfer_distance`` keeps the pytorch3d-style signature for the
homogeneous-batch case the reference actually uses (``batch_reduction=None, point_reduction=None``).
"""
import torch
import torch.nn as nn

from ._ext import _on_device_of
from ._lib import call, check_cuda_f32, dptr, lib, stream_ptr
from .knn import knn_points


def _fused(xyz1, xyz2, f1_threshold, want_dists=False):
    """xyz1 = output (B,n,3), xyz2 = gt (B,m,3) -> cd_p, cd_t, f1 [, dist1 (B,m), dist2 (B,n)]."""
    xyz1 = xyz1.contiguous().float()
    xyz2 = xyz2.contiguous().float()
    check_cuda_f32(xyz1, "xyz1")
    check_cuda_f32(xyz2, "xyz2")
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    dev = xyz1.device
    out = torch.empty((3, b), dtype=torch.float32, device=dev)
    nbytes = lib().pdr_chamfer_f1_workspace_bytes(b, n, m)
    ws = torch.empty((max(nbytes, 4) + 3) // 4, dtype=torch.float32, device=dev)
    d1 = torch.empty((b, m), dtype=torch.float32, device=dev) if want_dists else None
    d2 = torch.empty((b, n), dtype=torch.float32, device=dev) if want_dists else None
    import ctypes
    with _on_device_of(xyz1):
        call("pdr_chamfer_f1", b, n, m, dptr(xyz1), dptr(xyz2), ctypes.c_float(f1_threshold), dptr(out[0]),
             dptr(out[1]), dptr(out[2]), dptr(d1), dptr(d2), dptr(ws), nbytes, stream_ptr(xyz1))
    if want_dists:
        return out[0], out[1], out[2], d1, d2
    return out[0], out[1], out[2]


def chamfer_distance(x, y, x_lengths=None, y_lengths=None, x_normals=None, y_normals=None, weights=None,
                     batch_reduction="mean", point_reduction="mean"):
    """cham_x (x -> nearest y), cham_y (y -> nearest x), None.  chamfer_loss_new.py:67-217 restricted to
    equal-length clouds without normals/weights (all the reference's call sites)."""
    if any(a is not None for a in (x_lengths, y_lengths, x_normals, y_normals, weights)):
        raise NotImplementedError("chamfer_distance: lengths/normals/weights are not supported")
    if batch_reduction is not None and batch_reduction not in ("mean", "sum"):
        raise ValueError('batch_reduction must be one of ["mean", "sum"] or None')
    if point_reduction is not None and point_reduction not in ("mean", "sum"):
        raise ValueError('point_reduction must be one of ["mean", "sum"]')
    if point_reduction is None and batch_reduction is not None:
        raise ValueError("batch_reduction must be set to None if point_reduction is already None")
    cham_x = knn_points(x, y, K=1).dists[..., 0]
    cham_y = knn_points(y, x, K=1).dists[..., 0]
    if point_reduction is not None:
        cham_x, cham_y = cham_x.sum(1), cham_y.sum(1)
        if point_reduction == "mean":
            cham_x, cham_y = cham_x / x.shape[1], cham_y / y.shape[1]
    if batch_reduction is not None:
        cham_x, cham_y = cham_x.sum(), cham_y.sum()
        if batch_reduction == "mean":
            cham_x, cham_y = cham_x / x.shape[0], cham_y / x.shape[0]
    return cham_x, cham_y, None


def fscore(dist1, dist2, threshold=0.0001):
    """F-score from squared nearest-neighbour distances.  chamfer_loss_new.py:219-232."""
    p1 = torch.mean((dist1 < threshold).float(), dim=1)
    p2 = torch.mean((dist2 < threshold).float(), dim=1)
    f = 2 * p1 * p2 / (p1 + p2)
    f[torch.isnan(f)] = 0
    return f, p1, p2


def calc_cd(output, gt, calc_f1=False, f1_threshold=0.0001):
    """cd_p = (mean sqrt d1 + mean sqrt d2)/2, cd_t = mean d1 + mean d2 [, f1].  :234-245."""
    cd_p, cd_t, f1 = _fused(output, gt, f1_threshold)
    return (cd_p, cd_t, f1) if calc_f1 else (cd_p, cd_t)


class Chamfer_F1(nn.Module):
    def __init__(self, f1_threshold=0.0001):
        super().__init__()
        self.f1_threshold = f1_threshold

    def forward(self, xyz1, xyz2):
        return calc_cd(xyz1, xyz2, calc_f1=True, f1_threshold=self.f1_threshold)
