"""Host-side mirror of ``pointnet2_ops.pointnet2_utils`` (reference:
pointnet2_ops_lib/pointnet2_ops/pointnet2_utils.py) on top of the sm_100a kernels.

Same public names, argument meaning and return types: ``furthest_point_sample, gather_operation,
three_nn, three_interpolate, grouping_operation, ball_query, QueryAndGroup, GroupAll, group_knn,
count_to_mask, average_feature``.  Ops are looked up on the ``_ext`` / ``knn`` modules at call time.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.autograd import Function

from . import _ext, knn


def count_to_mask(count, K):
    """count (B,npoint) -> bool mask (B,npoint,K), mask[..., k] = k < count.  (pointnet2_utils.py:36-44)"""
    return torch.arange(K, device=count.device, dtype=count.dtype).view(1, 1, K) < count.unsqueeze(-1)


def average_feature(feature, count, K):
    """Masked mean over the neighbour axis.  feature (B,C,npoint,K) -> (B,C,npoint).  (:46-60)"""
    if isinstance(count, str) and count == "all":
        return feature.mean(dim=-1)
    count = torch.clamp(count, min=1)
    mask = count_to_mask(count, K).unsqueeze(1)
    return (feature * mask).sum(dim=-1) / count.unsqueeze(1)


class FurthestPointSampling(Function):
    @staticmethod
    def forward(ctx, xyz, npoint):
        out = _ext.furthest_point_sampling(xyz, npoint)
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        return None, None


furthest_point_sample = FurthestPointSampling.apply


class GatherOperation(Function):
    @staticmethod
    def forward(ctx, features, idx):
        ctx.save_for_backward(idx)
        ctx.n = features.size(2)
        return _ext.gather_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        return _ext.gather_points_grad(grad_out.contiguous(), idx, ctx.n), None


gather_operation = GatherOperation.apply


class ThreeNN(Function):
    @staticmethod
    def forward(ctx, unknown, known):
        dist2, idx = _ext.three_nn(unknown, known)
        dist = torch.sqrt(dist2)  # pointnet2_utils.py:153
        ctx.mark_non_differentiable(dist, idx)
        return dist, idx

    @staticmethod
    def backward(ctx, grad_dist, grad_idx):
        return None, None


three_nn = ThreeNN.apply


class ThreeInterpolate(Function):
    @staticmethod
    def forward(ctx, features, idx, weight):
        ctx.save_for_backward(idx, weight)
        ctx.m = features.size(2)
        return _ext.three_interpolate(features, idx, weight)

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight = ctx.saved_tensors
        g = _ext.three_interpolate_grad(grad_out.contiguous(), idx, weight, ctx.m)
        return g, None, None


three_interpolate = ThreeInterpolate.apply


class GroupingOperation(Function):
    @staticmethod
    def forward(ctx, features, idx):
        ctx.save_for_backward(idx)
        ctx.n = features.size(2)
        return _ext.group_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        return _ext.group_points_grad(grad_out.contiguous(), idx, ctx.n), None


grouping_operation = GroupingOperation.apply


class BallQuery(Function):
    @staticmethod
    def forward(ctx, radius, nsample, xyz, new_xyz):
        # note the argument swap: Python passes (xyz, new_xyz), the extension wants centres first
        # (pointnet2_utils.py:273-297 vs ball_query.cpp:10)
        idx, counts = _ext.ball_query(new_xyz, xyz, radius, nsample)
        ctx.mark_non_differentiable(idx, counts)
        return idx, counts

    @staticmethod
    def backward(ctx, *grads):
        return None, None, None, None


ball_query = BallQuery.apply


class QueryAndGroup(nn.Module):
    """Ball-query (or kNN) grouping that builds ``[features, rel_xyz, abs_xyz?, centre_xyz?]``.

    Mirrors pointnet2_utils.py:307-438.  No parameters, so nothing enters the state_dict.
    """

    def __init__(self, radius, nsample, use_xyz=True, include_abs_coordinate=False,
                 include_center_coordinate=False, neighbor_def="radius"):
        super().__init__()
        assert neighbor_def in ("radius", "nn")
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz
        self.include_abs_coordinate = include_abs_coordinate
        self.include_center_coordinate = include_center_coordinate
        self.neighbor_def = neighbor_def
        self.neighbor_stats = None
        self.neighbor_num_quantile = None
        self.quantile = torch.linspace(0, 1, 11)

    def neighbours(self, xyz, new_xyz):
        """-> idx (B,npoint,K) int32, counts ((B,npoint) int32 or the string 'all')."""
        if self.neighbor_def == "radius":
            return ball_query(self.radius, self.nsample, xyz, new_xyz)
        k = min(self.nsample, xyz.shape[1])
        idx = knn.knn_points(new_xyz, xyz, K=k).idx.int()
        return idx, "all"

    def forward(self, xyz, new_xyz, features=None, subset=True, record_neighbor_stats=False,
                return_counts=False):
        idx, counts = self.neighbours(xyz, new_xyz)
        radius_mode = self.neighbor_def == "radius"
        centre = new_xyz.transpose(1, 2).unsqueeze(-1)                      # (B,3,npoint,1)
        abs_xyz = grouping_operation(xyz.transpose(1, 2).contiguous(), idx)  # (B,3,npoint,K)
        fill_missing = (not subset) and radius_mode
        if fill_missing:
            # a centre without any neighbour is its own (zero-feature) neighbour  (:376-386)
            have = (counts > 0).float().unsqueeze(1).unsqueeze(-1).detach()
            abs_xyz = have * abs_xyz + (1 - have) * centre
        parts = [abs_xyz - centre]
        if self.include_abs_coordinate:
            parts.append(abs_xyz)
        if self.include_center_coordinate:
            parts.append(centre.expand(-1, -1, -1, abs_xyz.shape[3]))
        grouped_xyz = torch.cat(parts, dim=1) if len(parts) > 1 else parts[0]

        if features is not None:
            grouped = grouping_operation(features, idx)
            if fill_missing:
                grouped = have * grouped
            new_features = torch.cat([grouped, grouped_xyz], dim=1) if self.use_xyz else grouped
        else:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            new_features = grouped_xyz

        if record_neighbor_stats and radius_mode:
            with torch.no_grad():
                c = counts.float()
                self.neighbor_stats = torch.stack([c.min(), c.mean(), c.max()])
                self.neighbor_num_quantile = torch.quantile(c, self.quantile.to(c.device)).long()
        if return_counts:
            return new_features, counts
        return new_features


class GroupAll(nn.Module):
    """Single group holding every point (pointnet2_utils.py:441-484)."""

    def __init__(self, use_xyz=True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz, new_xyz, features=None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is None:
            return grouped_xyz
        grouped = features.unsqueeze(2)
        return torch.cat([grouped, grouped_xyz], dim=1) if self.use_xyz else grouped


def group_knn(x, y, features_at_y, K, transpose=False):
    """K nearest points of y for every point of x, with the 11 geometric channels appended:
    ``[feat(C), d2(1), w(1), nn_abs(3), nn_rel(3), x(3)]``, w = normalised 1/(d2+1e-8) with d2 the
    SQUARED distance (pointnet2_utils.py:487-514).  transpose=True: features (B,C,N2) in,
    (B,C+11,N1,K) out."""
    feats = features_at_y.transpose(1, 2).contiguous() if transpose else features_at_y
    dist, idx, nn_abs = knn.knn_points(x, y, K=K, return_nn=True)
    nn_feats = knn.knn_gather(feats, idx)                  # (B,N1,K,C)
    x_rep = x.unsqueeze(2).expand(-1, -1, K, -1)           # (B,N1,K,3)
    dist = dist.unsqueeze(3)
    recip = 1.0 / (dist + 1e-8)
    weight = recip / recip.sum(dim=2, keepdim=True)
    out = torch.cat([nn_feats, dist, weight, nn_abs, nn_abs - x_rep, x_rep], dim=3)
    if transpose:
        out = out.permute(0, 3, 1, 2)
    return out
