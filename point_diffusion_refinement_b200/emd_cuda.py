"""Drop-in for the ``emd_cuda`` extension (reference: PytorchEMD/cuda/emd.cpp:24-28): the same three
names and signatures, plus ``emd_cost_forward`` -- the fused approxmatch+matchcost path that never
materialises the (B, m, n) match tensor."""
import torch

from ._ext import _on_device_of
from ._lib import call, check_cuda_f32, dptr, lib, stream_ptr


def _check(xyz1, xyz2):
    check_cuda_f32(xyz1, "xyz1")
    check_cuda_f32(xyz2, "xyz2")
    if xyz1.dim() != 3 or xyz2.dim() != 3 or xyz1.shape[2] != 3 or xyz2.shape[2] != 3:
        raise RuntimeError("xyz1/xyz2 must be (B,n,3)/(B,m,3)")
    if xyz1.shape[0] != xyz2.shape[0]:
        raise RuntimeError("batch sizes differ")
    return xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]


def _workspace(b, n, m, device):
    nbytes = lib().pdr_emd_workspace_bytes(b, n, m)
    return torch.empty((max(nbytes, 4) + 3) // 4, dtype=torch.float32, device=device), nbytes


def approxmatch_forward(xyz1, xyz2):
    """xyz1 (B,n,3), xyz2 (B,m,3) -> match (B,m,n).  emd_kernel.cu:174-196."""
    b, n, m = _check(xyz1, xyz2)
    match = torch.empty((b, m, n), dtype=torch.float32, device=xyz1.device)
    temp, nbytes = _workspace(b, n, m, xyz1.device)
    with _on_device_of(xyz1):
        call("pdr_emd_approxmatch", b, n, m, dptr(xyz1), dptr(xyz2), dptr(match), dptr(temp), nbytes,
             stream_ptr(xyz1))
    return match


def matchcost_forward(xyz1, xyz2, match):
    """-> cost (B) = sum d^2 * match.  emd_kernel.cu:260-282."""
    b, n, m = _check(xyz1, xyz2)
    check_cuda_f32(match, "match")
    cost = torch.empty((b,), dtype=torch.float32, device=xyz1.device)
    temp, nbytes = _workspace(b, n, m, xyz1.device)
    with _on_device_of(xyz1):
        call("pdr_emd_matchcost", b, n, m, dptr(xyz1), dptr(xyz2), dptr(match), dptr(cost), dptr(temp), nbytes,
             stream_ptr(xyz1))
    return cost


def matchcost_backward(grad_cost, xyz1, xyz2, match):
    """-> [grad1 (B,n,3), grad2 (B,m,3)].  emd_kernel.cu:376-401."""
    b, n, m = _check(xyz1, xyz2)
    check_cuda_f32(grad_cost, "grad_cost")
    check_cuda_f32(match, "match")
    g1 = torch.empty_like(xyz1)
    g2 = torch.empty_like(xyz2)
    with _on_device_of(xyz1):
        call("pdr_emd_matchcost_backward", b, n, m, dptr(grad_cost), dptr(xyz1), dptr(xyz2), dptr(match),
             dptr(g1), dptr(g2), stream_ptr(xyz1))
    return [g1, g2]


def emd_cost_forward(xyz1, xyz2):
    """Fused approxmatch + matchcost: cost (B), match never written to HBM."""
    b, n, m = _check(xyz1, xyz2)
    cost = torch.empty((b,), dtype=torch.float32, device=xyz1.device)
    temp, nbytes = _workspace(b, n, m, xyz1.device)
    with _on_device_of(xyz1):
        call("pdr_emd_cost", b, n, m, dptr(xyz1), dptr(xyz2), dptr(cost), dptr(temp), nbytes, stream_ptr(xyz1))
    return cost
