"""Drop-in for the part of ``pytorch3d.ops.knn`` the reference uses (pytorch3d is un-vendored and
unpinned, setup_env.sh:5): ``knn_points`` / ``knn_gather`` with the same namedtuple result.
Call sites: pointnet2/chamfer_loss_new.py:149-150, pointnet2_ops/pointnet2_utils.py:365,496-497.

Semantics held: exact brute force on squared L2 in fp32, ascending, ties keep the lower index,
``idx`` int64, ``dists`` squared.  ``lengths1/2`` (heterogeneous batches) are not used by the reference
call sites and are rejected.
"""
from collections import namedtuple

import torch

from ._ext import _on_device_of
from ._lib import call, check_cuda_f32, dptr, stream_ptr

_KNN = namedtuple("KNN", "dists idx knn")


class _KnnFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p1, p2, K):
        b, n1, _ = p1.shape
        n2 = p2.shape[1]
        dists = torch.empty((b, n1, K), dtype=torch.float32, device=p1.device)
        idx = torch.empty((b, n1, K), dtype=torch.int64, device=p1.device)
        with _on_device_of(p1):
            call("pdr_knn_points", b, n1, n2, K, dptr(p1), dptr(p2), dptr(dists), dptr(idx), stream_ptr(p1))
        ctx.save_for_backward(p1, p2, idx)
        ctx.mark_non_differentiable(idx)
        return dists, idx

    @staticmethod
    def backward(ctx, grad_dists, grad_idx):
        p1, p2, idx = ctx.saved_tensors
        nn = knn_gather(p2, idx)                      # (b,n1,K,3)
        diff = p1.unsqueeze(2) - nn                   # d = |p1 - nn|^2
        g = 2.0 * grad_dists.unsqueeze(-1) * diff     # (b,n1,K,3)
        grad_p1 = g.sum(2)
        grad_p2 = torch.zeros_like(p2)
        b, n1, K = idx.shape
        grad_p2.scatter_add_(1, idx.reshape(b, n1 * K, 1).expand(-1, -1, 3), (-g).reshape(b, n1 * K, 3))
        return grad_p1, grad_p2, None


def knn_points(p1, p2, lengths1=None, lengths2=None, K=1, version=-1, return_nn=False, return_sorted=True):
    # The reference's chamfer_distance always passes lengths (chamfer_loss_new.py:149-150), full ones for padded tensors
    # (_handle_pointcloud_input, :52-57): full lengths are the homogeneous case; anything shorter is not implemented.
    for lengths, p in ((lengths1, p1), (lengths2, p2)):
        if lengths is not None and not bool((torch.as_tensor(lengths) == p.shape[1]).all()):
            raise NotImplementedError("knn_points: heterogeneous batches (lengths1/lengths2 < P) are not supported")
    if p1.shape[0] != p2.shape[0] or p1.shape[2] != 3 or p2.shape[2] != 3:
        raise ValueError("knn_points expects (N,P1,3) and (N,P2,3)")
    p1c = p1.contiguous().float()
    p2c = p2.contiguous().float()
    check_cuda_f32(p1c, "p1")
    check_cuda_f32(p2c, "p2")
    dists, idx = _KnnFunction.apply(p1c, p2c, int(K))
    nn = knn_gather(p2c, idx) if return_nn else None
    return _KNN(dists=dists, idx=idx, knn=nn)


def knn_gather(x, idx, lengths=None):
    """x (N,P2,C), idx (N,P1,K) int64 -> (N,P1,K,C)."""
    N, P1, K = idx.shape
    C = x.shape[2]
    return x.gather(1, idx.reshape(N, P1 * K, 1).expand(-1, -1, C)).reshape(N, P1, K, C)


class Pointclouds:  # only used for an isinstance() check at chamfer_loss_new.py:40
    pass
