"""Mirror-partial preprocessing behind the reference's ``data_utils.mirror_partial`` surface (reference:
pointnet2/data_utils/mirror_partial.py:5-37; offline caller mvp_dataloader/generate_mirrored_partial.py:44).

The partial cloud is mirrored across one axis, tagged with a +1 / -1 flag channel, concatenated (2N points)
and down-sampled by furthest point sampling to the 2048- / 3072-point conditions the networks consume.
FPS and the 4-channel gather run on the sm_100a kernels (bit-exact index parity with the reference FPS).
"""
import torch

from . import pointnet2_utils


def mirror(partial, axis=1):
    """partial (B,N,3) -> copy with coordinate `axis` negated.  mirror_partial.py:5-9."""
    partial_mirror = partial.clone()
    partial_mirror[:, :, axis] = -partial_mirror[:, :, axis]
    return partial_mirror


def down_sample_points(xyz, npoints):
    """xyz (B,N,4) = coordinates + flag -> (B,npoints,4), FPS on the coordinates.  mirror_partial.py:11-20."""
    xyz_flipped = xyz.transpose(1, 2).contiguous()
    ori_xyz = xyz[:, :, 0:3].contiguous()
    idx = pointnet2_utils.furthest_point_sample(ori_xyz, npoints)
    new_xyz = pointnet2_utils.gather_operation(xyz_flipped, idx)
    return new_xyz.transpose(1, 2).contiguous()


def mirror_and_concat(partial, axis=2, num_points=(2048, 3072)):
    """partial (B,N,3) -> (concat (B,2N,4), down-sampled (B,n,4) for n in num_points).  mirror_partial.py:22-37.
    The reference moves the concatenation to the current CUDA device (`.cuda()`); so does this."""
    B, N, _ = partial.size()
    partial_mirror = mirror(partial, axis=axis)
    ones = torch.ones(B, N, 1, device=partial.device, dtype=partial.dtype)
    concat = torch.cat([torch.cat([partial, ones], dim=2), torch.cat([partial_mirror, -ones], dim=2)], dim=1)
    if not concat.is_cuda:
        concat = concat.cuda()
    down_sampled = [concat]
    for n in num_points:
        down_sampled.append(down_sample_points(concat, n))
    return tuple(down_sampled)
