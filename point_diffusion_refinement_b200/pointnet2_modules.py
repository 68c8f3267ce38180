"""Host-side mirror of ``pointnet2_ops.pointnet2_modules`` (reference:
pointnet2_ops_lib/pointnet2_ops/pointnet2_modules.py): ``Mlp_plus_t_emb``, ``build_shared_mlp``,
``PointnetSAModule(MSG)``, ``PointnetFPModule``, ``PointnetKnnFPModule``, ``FeatureMapModule``,
``pooling_features``.  Constructor arguments, forward signatures and -- because checkpoints are loaded
with ``load_state_dict`` (generate_samples.py:178-179) -- every parameter name match the reference.
"""
import copy
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import pointnet2_utils
from .attention import AttentionModule, GlobalAttentionModule, MyGroupNorm


def swish(x):
    return x * torch.sigmoid(x)


class Swish(nn.Module):
    def forward(self, x):
        return swish(x)


def _act(name):
    assert name in ("relu", "swish")
    return nn.ReLU(True) if name == "relu" else Swish()


def build_shared_mlp(mlp_spec: List[int], bn: bool = True, bn_first: bool = False, bias: bool = False,
                     activation: str = "relu"):
    """1x1-conv stack.  post-norm (bn_first=False): [conv, GN(32), act]*; pre-norm: [GN, act, conv]*.
    pointnet2_modules.py:42-67."""
    layers = []
    for cin, cout in zip(mlp_spec[:-1], mlp_spec[1:]):
        conv = nn.Conv2d(cin, cout, kernel_size=1, bias=bias)
        if bn_first:
            if bn:
                layers.append(MyGroupNorm(min(32, cin), cin))
            layers += [_act(activation), conv]
        else:
            layers.append(conv)
            if bn:
                layers.append(MyGroupNorm(32, cout))
            layers.append(_act(activation))
    return nn.Sequential(*layers)


def _extra_xyz_channels(use_xyz, include_abs, include_center):
    return (3 + (3 if include_abs else 0) + (3 if include_center else 0)) if use_xyz else 0


class Mlp_plus_t_emb(nn.Module):
    """Shared MLP with per-sample additive embeddings: +fc(t_emb) after layer 1, +fc_condition after
    layer 2, +fc_second_condition at the end, optional 1x1-conv residual.  pointnet2_modules.py:69-174."""

    def __init__(self, mlp_spec, bn, t_dim=128, include_t=True, bn_first=False, bias=False, first_conv=False,
                 first_conv_in_channel=0, res_connect=False, include_condition=False, condition_dim=128,
                 include_second_condition=False, second_condition_dim=128, activation="relu"):
        super().__init__()
        assert len(mlp_spec) >= 3
        if include_second_condition:
            assert len(mlp_spec) >= 4
        self.include_t = include_t
        if include_t:
            self.fc = nn.Linear(t_dim, mlp_spec[1])
        self.include_condition = include_condition
        if include_condition:
            self.fc_condition = nn.Linear(condition_dim, mlp_spec[2])
        self.include_second_condition = include_second_condition
        if include_second_condition:
            self.fc_second_condition = nn.Linear(second_condition_dim, mlp_spec[-1])
        self.first_conv_bool = first_conv
        if first_conv:
            self.first_conv = nn.Conv2d(first_conv_in_channel, mlp_spec[0], kernel_size=1, bias=bias)
        self.res_connect_bool = res_connect
        if res_connect:
            self.res_connect = (None if mlp_spec[0] == mlp_spec[-1]
                                else nn.Conv2d(mlp_spec[0], mlp_spec[-1], kernel_size=1, bias=bias))
        kw = dict(bn_first=bn_first, bias=bias, activation=activation)
        self.first_mlp = build_shared_mlp(mlp_spec[0:2], bn, **kw)
        self.second_mlp = build_shared_mlp(mlp_spec[1:3], bn, **kw)
        self.rest_mlp = build_shared_mlp(mlp_spec[2:], bn, **kw) if len(mlp_spec) > 3 else None

    @staticmethod
    def _per_sample(fc, emb, what):
        if emb is None:
            raise Exception("Should pass %s to the forward function" % what)
        return fc(emb).unsqueeze(2).unsqueeze(3)

    def forward(self, feature, t_emb=None, condition_emb=None, second_condition_emb=None):
        if self.first_conv_bool:
            feature = self.first_conv(feature)
        h = self.first_mlp(feature)
        if self.include_t:
            h = h + self._per_sample(self.fc, t_emb, "t_emb")
        elif t_emb is not None:
            raise Exception("This module does not include t but t_emb is given")
        h = self.second_mlp(h)
        if self.include_condition:
            h = h + self._per_sample(self.fc_condition, condition_emb, "condition_emb")
        elif condition_emb is not None:
            raise Exception("This module does not include condition but condition_emb is given")
        if self.rest_mlp is not None:
            h = self.rest_mlp(h)
        if self.include_second_condition:
            h = h + self._per_sample(self.fc_second_condition, second_condition_emb, "second_condition_emb")
        elif second_condition_emb is not None:
            raise Exception("This module does not include condition but condition_emb is given")
        if self.res_connect_bool:
            h = h + (feature if self.res_connect is None else self.res_connect(feature))
        return h


def pooling_features(feature, count=None, pooling="max"):
    """(B,C,npoint,K) -> (B,C,npoint) by max / masked avg / half-and-half.  pointnet2_modules.py:177-206."""
    assert pooling in ("max", "avg", "avg_max", "max_avg")
    K = feature.size(3)
    if pooling == "max":
        return feature.max(dim=-1)[0]
    if pooling == "avg":
        return pointnet2_utils.average_feature(feature, count, K)
    half = feature.shape[1] // 2
    return torch.cat([feature[:, :half].max(dim=-1)[0],
                      pointnet2_utils.average_feature(feature[:, half:], count, K)], dim=1)


class PointnetSAModuleMSG(nn.Module):
    """Set abstraction: FPS -> gather centres -> (ball query -> group -> MLP -> attention/pool)* -> cat.
    pointnet2_modules.py:209-388 (base forward :220-280)."""

    def __init__(self, npoint, radii, nsamples, mlps, bn=True, use_xyz=True, t_dim=128, include_t=False,
                 include_abs_coordinate=False, include_center_coordinate=False, bn_first=False, bias=False,
                 first_conv=False, first_conv_in_channel=0, res_connect=False, include_condition=False,
                 condition_dim=128, include_second_condition=False, second_condition_dim=128,
                 neighbor_def="radius", activation="relu", attention_setting=None,
                 global_attention_setting=None):
        super().__init__()
        assert len(radii) == len(nsamples) == len(mlps)
        self.npoint = npoint
        self.include_t, self.t_dim = include_t, t_dim
        self.include_condition, self.condition_dim = include_condition, condition_dim
        self.include_second_condition, self.second_condition_dim = include_second_condition, second_condition_dim
        self.use_attention_module = bool(attention_setting and attention_setting["use_attention_module"])
        self.use_global_attention_module = bool(
            global_attention_setting and global_attention_setting["use_global_attention_module"])
        self.groupers = nn.ModuleList()
        self.mlps = nn.ModuleList()
        self.attention_modules = nn.ModuleList() if self.use_attention_module else None
        self.global_attention_modules = nn.ModuleList() if self.use_global_attention_module else None
        extra = _extra_xyz_channels(use_xyz, include_abs_coordinate, include_center_coordinate)
        for radius, nsample, mlp_spec in zip(radii, nsamples, mlps):
            self.groupers.append(
                pointnet2_utils.QueryAndGroup(radius, nsample, use_xyz=use_xyz,
                                              include_abs_coordinate=include_abs_coordinate,
                                              include_center_coordinate=include_center_coordinate,
                                              neighbor_def=neighbor_def)
                if npoint is not None else pointnet2_utils.GroupAll(use_xyz))
            query_dim = first_conv_in_channel if first_conv else mlp_spec[0]
            in_conv = first_conv_in_channel + extra if first_conv else first_conv_in_channel
            if not first_conv:
                mlp_spec[0] += extra  # in place, like the reference (callers rely on it)
            self.mlps.append(Mlp_plus_t_emb(
                mlp_spec, bn, t_dim=t_dim, include_t=include_t, bn_first=bn_first, bias=bias,
                first_conv=first_conv, first_conv_in_channel=in_conv, res_connect=res_connect,
                include_condition=include_condition, condition_dim=condition_dim,
                include_second_condition=include_second_condition, second_condition_dim=second_condition_dim,
                activation=activation))
            if self.use_attention_module:
                key_dim = in_conv if first_conv else mlp_spec[0]
                self.attention_modules.append(AttentionModule(
                    query_dim, key_dim, query_dim, key_dim, mlp_spec[-1],
                    attention_bn=attention_setting["attention_bn"],
                    transform_grouped_feat_out=attention_setting["transform_grouped_feat_out"],
                    last_activation=attention_setting["last_activation"]))
            if self.use_global_attention_module:
                self.global_attention_modules.append(GlobalAttentionModule(
                    mlp_spec[-1], additional_dim=3, attention_bn=global_attention_setting["attention_bn"],
                    last_activation=global_attention_setting["last_activation"]))

    def forward(self, xyz, features, t_emb=None, condition_emb=None, second_condition_emb=None, subset=True,
                record_neighbor_stats=False, pooling="max"):
        assert self.npoint is not None
        fps_idx = pointnet2_utils.furthest_point_sample(xyz, self.npoint)
        new_xyz = pointnet2_utils.gather_operation(xyz.transpose(1, 2).contiguous(), fps_idx)
        new_xyz = new_xyz.transpose(1, 2).contiguous()
        if self.use_attention_module:
            centre_feat = pointnet2_utils.gather_operation(features, fps_idx)
        t_emb = t_emb if self.include_t else None
        condition_emb = condition_emb if self.include_condition else None
        second_condition_emb = second_condition_emb if self.include_second_condition else None
        outs = []
        for i, grouper in enumerate(self.groupers):
            grouped, count = grouper(xyz, new_xyz, features, subset=subset,
                                     record_neighbor_stats=record_neighbor_stats, return_counts=True)
            out = self.mlps[i](grouped, t_emb=t_emb, condition_emb=condition_emb,
                               second_condition_emb=second_condition_emb)
            if self.use_attention_module:
                new_features = self.attention_modules[i](centre_feat, grouped, out, count)
            else:
                new_features = pooling_features(out, count=count, pooling=pooling)
            if self.use_global_attention_module:
                new_features = self.global_attention_modules[i](
                    torch.cat([new_features, new_xyz.transpose(1, 2)], dim=1))
            outs.append(new_features)
        return new_xyz, torch.cat(outs, dim=1)


class PointnetSAModule(PointnetSAModuleMSG):
    """Single-scale set abstraction (pointnet2_modules.py:391-442)."""

    def __init__(self, mlp, npoint=None, radius=None, nsample=None, **kw):
        super().__init__(npoint=npoint, radii=[radius], nsamples=[nsample], mlps=[mlp], **kw)


class PointnetFPModule(nn.Module):
    """Feature propagation by inverse-distance three-NN interpolation (pointnet2_modules.py:445-576)."""

    def __init__(self, mlp, bn=True, t_dim=128, include_t=False, bn_first=False, bias=False, first_conv=False,
                 first_conv_in_channel=0, res_connect=False, include_condition=False, condition_dim=128,
                 include_second_condition=False, second_condition_dim=128, include_grouper=False, radius=0,
                 nsample=32, use_xyz=True, include_abs_coordinate=True, include_center_coordinate=False,
                 neighbor_def="radius", activation="relu"):
        super().__init__()
        self.include_t, self.t_dim = include_t, t_dim
        self.include_condition, self.condition_dim = include_condition, condition_dim
        self.include_second_condition, self.second_condition_dim = include_second_condition, second_condition_dim
        self.include_grouper = include_grouper
        if include_grouper:
            extra = _extra_xyz_channels(use_xyz, include_abs_coordinate, include_center_coordinate)
            if first_conv:
                first_conv_in_channel += extra
            else:
                mlp[0] += extra
            self.grouper = pointnet2_utils.QueryAndGroup(
                radius, nsample, use_xyz=use_xyz, include_abs_coordinate=include_abs_coordinate,
                include_center_coordinate=include_center_coordinate, neighbor_def=neighbor_def)
        self.mlp = Mlp_plus_t_emb(
            mlp, bn, t_dim=t_dim, include_t=include_t, bn_first=bn_first, bias=bias, first_conv=first_conv,
            first_conv_in_channel=first_conv_in_channel, res_connect=res_connect,
            include_condition=include_condition, condition_dim=condition_dim,
            include_second_condition=include_second_condition, second_condition_dim=second_condition_dim,
            activation=activation)

    def forward(self, unknown, known, unknow_feats, known_feats, t_emb=None, condition_emb=None,
                second_condition_emb=None, record_neighbor_stats=False, pooling="max"):
        if known is not None:
            dist, idx = pointnet2_utils.three_nn(unknown, known)
            recip = 1.0 / (dist + 1e-8)
            weight = recip / recip.sum(dim=2, keepdim=True)
            interpolated = pointnet2_utils.three_interpolate(known_feats, idx, weight)
        else:
            interpolated = known_feats.expand(*(list(known_feats.size()[0:2]) + [unknown.size(1)]))
        new_features = interpolated if unknow_feats is None else torch.cat([interpolated, unknow_feats], dim=1)
        if self.include_grouper:
            new_features, count = self.grouper(unknown, unknown, new_features, subset=True,
                                               record_neighbor_stats=record_neighbor_stats, return_counts=True)
        else:
            new_features = new_features.unsqueeze(-1)
        new_features = self.mlp(new_features,
                                t_emb=t_emb if self.include_t else None,
                                condition_emb=condition_emb if self.include_condition else None,
                                second_condition_emb=second_condition_emb if self.include_second_condition else None)
        if self.include_grouper:
            return pooling_features(new_features, count=count, pooling=pooling)
        return new_features.squeeze(-1)


class FeatureMapModule(nn.Module):
    """Feature transfer: pool features living at ``xyz`` onto arbitrary query points ``new_xyz`` through
    ball-query grouping + MLP + attention (pointnet2_modules.py:579-649)."""

    def __init__(self, mlp, radius, K, use_xyz=True, include_abs_coordinate=True, include_center_coordinate=False,
                 bn=True, bn_first=True, bias=True, res_connect=True, first_conv=False, first_conv_in_channel=0,
                 neighbor_def="radius", activation="relu", attention_setting=None, query_feature_dim=None):
        super().__init__()
        self.use_attention_module = bool(attention_setting and attention_setting["use_attention_module"])
        extra = _extra_xyz_channels(use_xyz, include_abs_coordinate, include_center_coordinate)
        if first_conv:
            first_conv_in_channel += extra
        else:
            mlp[0] += extra
        self.mlp = Mlp_plus_t_emb(mlp, bn, include_t=False, bn_first=bn_first, bias=bias, first_conv=first_conv,
                                  first_conv_in_channel=first_conv_in_channel, res_connect=res_connect,
                                  include_condition=False, activation=activation)
        self.mapper = pointnet2_utils.QueryAndGroup(
            radius, K, use_xyz=use_xyz, include_abs_coordinate=include_abs_coordinate,
            include_center_coordinate=include_center_coordinate, neighbor_def=neighbor_def)
        if self.use_attention_module:
            key_dim = first_conv_in_channel if first_conv else mlp[0]
            self.attention_module = AttentionModule(
                query_feature_dim, key_dim, query_feature_dim, key_dim, mlp[-1],
                attention_bn=attention_setting["attention_bn"],
                transform_grouped_feat_out=attention_setting["transform_grouped_feat_out"],
                last_activation=attention_setting["last_activation"])

    def forward(self, xyz, features, new_xyz, subset=False, record_neighbor_stats=True, pooling="max",
                features_at_new_xyz=None):
        grouped, count = self.mapper(xyz, new_xyz, features, subset=subset,
                                     record_neighbor_stats=record_neighbor_stats, return_counts=True)
        out = self.mlp(grouped)
        if self.use_attention_module:
            return self.attention_module(features_at_new_xyz, grouped, out, count)
        return pooling_features(out, count=count, pooling=pooling)


class PointnetKnnFPModule(nn.Module):
    """Feature propagation through K nearest neighbours: group_knn -> mlp1 -> attention pool over K ->
    cat skip + xyz -> mlp2 (pointnet2_modules.py:652-839)."""

    def __init__(self, mlp1, mlp2, K, bn=True, t_dim=128, include_t=False, bn_first=False, bias=False,
                 first_conv=False, first_conv_in_channel1=0, first_conv_in_channel2=0, res_connect=False,
                 include_condition=False, condition_dim=128, include_second_condition=False,
                 second_condition_dim=128, include_grouper=False, radius=0, nsample=32, use_xyz=True,
                 include_abs_coordinate=True, include_center_coordinate=False, neighbor_def="radius",
                 activation="relu", attention_setting=None, global_attention_setting=None):
        super().__init__()
        self.include_t, self.t_dim = include_t, t_dim
        self.include_condition, self.condition_dim = include_condition, condition_dim
        self.include_second_condition, self.second_condition_dim = include_second_condition, second_condition_dim
        self.K = K
        if first_conv:
            first_conv_in_channel1 += 11
        else:
            mlp1[0] += 11
        # mlp1 carries the SECOND condition in its `condition` slot and never sees t (reference :690-695)
        self.mlp1 = Mlp_plus_t_emb(mlp1, bn, t_dim=t_dim, include_t=False, bn_first=bn_first, bias=bias,
                                   first_conv=first_conv, first_conv_in_channel=first_conv_in_channel1,
                                   res_connect=res_connect, include_condition=include_second_condition,
                                   condition_dim=second_condition_dim, activation=activation)
        self.use_attention_module = bool(attention_setting and attention_setting["use_attention_module"])
        if self.use_attention_module:
            query_dim = (first_conv_in_channel2 if first_conv else mlp2[0]) - mlp1[-1]  # dim of the skip feature
            key_dim = first_conv_in_channel1 if first_conv else mlp1[0]
            self.attention_module = AttentionModule(
                query_dim, key_dim, query_dim, key_dim, mlp1[-1],
                attention_bn=attention_setting["attention_bn"],
                transform_grouped_feat_out=attention_setting["transform_grouped_feat_out"],
                last_activation=attention_setting["last_activation"])
        self.include_grouper = include_grouper
        if include_grouper:
            extra = _extra_xyz_channels(use_xyz, include_abs_coordinate, include_center_coordinate)
            self.grouper = pointnet2_utils.QueryAndGroup(
                radius, nsample, use_xyz=use_xyz, include_abs_coordinate=include_abs_coordinate,
                include_center_coordinate=include_center_coordinate, neighbor_def=neighbor_def)
        else:
            extra = 3
        if first_conv:
            first_conv_in_channel2 += extra
        else:
            mlp2[0] += extra
        self.mlp2 = Mlp_plus_t_emb(mlp2, bn, t_dim=t_dim, include_t=include_t, bn_first=bn_first, bias=bias,
                                   first_conv=first_conv, first_conv_in_channel=first_conv_in_channel2,
                                   res_connect=res_connect, include_condition=include_condition,
                                   condition_dim=condition_dim, activation=activation)
        self.use_global_attention_module = bool(
            global_attention_setting and global_attention_setting["use_global_attention_module"])
        if self.use_global_attention_module:
            self.global_attention_module = GlobalAttentionModule(
                mlp2[-1], additional_dim=3, attention_bn=global_attention_setting["attention_bn"],
                last_activation=global_attention_setting["last_activation"])

    def forward(self, unknown, known, unknow_feats, known_feats, t_emb=None, condition_emb=None,
                second_condition_emb=None, record_neighbor_stats=False, pooling="max"):
        if self.use_attention_module or self.use_global_attention_module:
            assert known is not None and unknown is not None
            if self.use_global_attention_module:
                assert not self.include_grouper
        if known is not None:
            grouped = pointnet2_utils.group_knn(unknown, known, known_feats, self.K, transpose=True)
            c2 = second_condition_emb if self.include_second_condition else None
            grouped_out = self.mlp1(grouped, t_emb=None, condition_emb=c2)
            if self.use_attention_module:
                interpolated = self.attention_module(unknow_feats, grouped, grouped_out, count="all")
            else:
                interpolated = pooling_features(grouped_out, count="all", pooling=pooling)
        else:
            interpolated = known_feats.expand(*(list(known_feats.size()[0:2]) + [unknown.size(1)]))
        new_features = interpolated if unknow_feats is None else torch.cat([interpolated, unknow_feats], dim=1)
        if self.include_grouper:
            new_features, count = self.grouper(unknown, unknown, new_features, subset=True,
                                               record_neighbor_stats=record_neighbor_stats, return_counts=True)
        else:
            new_features = torch.cat([new_features, unknown.transpose(1, 2)], dim=1).unsqueeze(-1)
        new_features = self.mlp2(new_features,
                                 t_emb=t_emb if self.include_t else None,
                                 condition_emb=condition_emb if self.include_condition else None)
        if self.include_grouper:
            return pooling_features(new_features, count=count, pooling=pooling)
        new_features = new_features.squeeze(-1)
        if self.use_global_attention_module:
            new_features = self.global_attention_module(
                torch.cat([new_features, unknown.transpose(1, 2)], dim=1))
        return new_features
