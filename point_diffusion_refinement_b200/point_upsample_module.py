"""Output stage of the refinement network behind the reference's ``point_upsample`` surface (reference:
pointnet2/models/point_upsample_module.py:4-27; caller completion_eval.py:159-168).

One sm_100a launch (``pdr_point_upsample``) replaces the reference's slice / scale / view / broadcast-add /
reshape / cat chain; every product and sum is rounded separately, exactly like that chain of elementwise
torch kernels, so the result is bit-identical.
"""
import ctypes

import numpy as np
import torch

from ._ext import _on_device_of
from ._lib import call, check_cuda_f32, dptr, stream_ptr


def point_upsample(coarse, displacement, point_upsample_factor, include_displacement_center_to_final_output,
                   output_scale_factor_value):
    """coarse (B,N,3), displacement (B,N,3*factor) or (B,N,3*(factor+1)) -> (refined_X, intermediate_refined_X).

    refined_X is (B, N*factor, 3): per coarse point its ``factor`` (or ``factor-1``) displaced copies, followed --
    when ``include_displacement_center_to_final_output`` -- by the N intermediate points themselves."""
    coarse = coarse.contiguous()
    displacement = displacement.contiguous()
    check_cuda_f32(coarse, "coarse")
    check_cuda_f32(displacement, "displacement")
    B, N, _ = coarse.shape
    factor = int(point_upsample_factor)
    centre = bool(include_displacement_center_to_final_output)
    reps = factor - 1 if centre else factor
    if displacement.shape != (B, N, 3 * (reps + 1)):
        raise RuntimeError("displacement must be (B, N, %d), got %s" % (3 * (reps + 1), tuple(displacement.shape)))
    # the reference multiplies an fp32 tensor by the float64 python scalar 1/np.sqrt(factor): torch rounds the
    # scalar to fp32 first (point_upsample_module.py:8-9)
    grid_scale = float(np.float32(1 / np.sqrt(factor)))
    refined = torch.empty((B, N * (reps + (1 if centre else 0)), 3), dtype=torch.float32, device=coarse.device)
    mid = torch.empty_like(coarse)
    with _on_device_of(coarse):
        call("pdr_point_upsample", B, N, factor, int(centre), dptr(coarse), dptr(displacement),
             ctypes.c_float(grid_scale), ctypes.c_float(float(np.float32(output_scale_factor_value))),
             dptr(refined), dptr(mid), stream_ptr(coarse))
    return refined, mid
