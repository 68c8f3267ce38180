"""Earth mover's distance behind the reference's ``pointnet2/emd.py`` surface:
``EarthMoverDistanceFunction``, ``earth_mover_distance``, ``EMD_distance``.

Forward without ``return_match`` and without autograd runs the fused cost kernel (no (B,m,n) match
tensor); when gradients or the match are requested the reference's two-step path is used."""
import torch
import torch.nn as nn

from . import emd_cuda


class EarthMoverDistanceFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2, return_match=False):
        # (asked of ctx, not of the tensors: inside forward grad mode is off, and .contiguous() of a non-contiguous
        #  input -- transpose=True -- is a copy with requires_grad False)
        need_match = return_match or ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        xyz1 = xyz1.contiguous()
        xyz2 = xyz2.contiguous()
        assert xyz1.is_cuda and xyz2.is_cuda, "Only support cuda currently."
        scale = max(xyz1.shape[1], xyz2.shape[1])
        if not need_match:
            return emd_cuda.emd_cost_forward(xyz1, xyz2) / scale
        match = emd_cuda.approxmatch_forward(xyz1, xyz2)
        cost = emd_cuda.matchcost_forward(xyz1, xyz2, match) / scale
        ctx.save_for_backward(xyz1, xyz2, match)
        return (cost, match) if return_match else cost

    @staticmethod
    def backward(ctx, grad_cost, *unused):
        xyz1, xyz2, match = ctx.saved_tensors
        g1, g2 = emd_cuda.matchcost_backward(grad_cost.contiguous(), xyz1, xyz2, match)
        return g1, g2, None


def _prepare(xyz1, xyz2, transpose):
    if xyz1.dim() == 2:
        xyz1 = xyz1.unsqueeze(0)
    if xyz2.dim() == 2:
        xyz2 = xyz2.unsqueeze(0)
    if transpose:
        xyz1, xyz2 = xyz1.transpose(1, 2), xyz2.transpose(1, 2)
    return xyz1, xyz2


def earth_mover_distance(xyz1, xyz2, transpose=False, return_match=False):
    """xyz1 (b,n,3), xyz2 (b,m,3) -> cost (b) [, match (b,m,n)].  emd.py:33-56."""
    xyz1, xyz2 = _prepare(xyz1, xyz2, transpose)
    return EarthMoverDistanceFunction.apply(xyz1, xyz2, bool(return_match))


class EMD_distance(nn.Module):
    def forward(self, xyz1, xyz2, transpose=False, return_match=False):
        return earth_mover_distance(xyz1, xyz2, transpose=transpose, return_match=return_match)
