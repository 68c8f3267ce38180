"""Host-side mirror of ``pointnet2_ops.attention`` (reference: pointnet2_ops_lib/pointnet2_ops/
attention.py): ``MyGroupNorm`` and ``AttentionModule`` with the reference's parameter names
(``feat_conv``, ``grouped_feat_conv``, ``weight_conv.{1,4}.group_norm``, ``weight_conv.{2,5}``,
``feat_out_conv.{0,1}``) so reference checkpoints load key for key.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class MyGroupNorm(nn.Module):
    """GroupNorm over the first ``C - C % G`` channels; trailing channels pass through unchanged
    (e.g. 79 -> 64 normalised + 15 untouched).  attention.py:6-23."""

    def __init__(self, num_groups, num_channels):
        super().__init__()
        self.num_groups = num_groups
        self.num_channels = num_channels - num_channels % num_groups
        self.group_norm = nn.GroupNorm(self.num_groups, self.num_channels)

    def forward(self, x):
        c = self.num_channels
        if x.shape[1] == c:
            return self.group_norm(x)
        return torch.cat([self.group_norm(x[:, :c]), x[:, c:]], dim=1)


def masked_softmax_pool(scores, values, count):
    """softmax over the neighbour axis K with slots >= clamp(count,1) masked out, then the weighted
    sum of ``values``.  scores/values (B,C,N,K), count (B,N) or 'all' -> (B,C,N).  attention.py:85-96."""
    if not (isinstance(count, str) and count == "all"):
        K = scores.shape[-1]
        count = torch.clamp(count, min=1)
        mask = (torch.arange(K, device=count.device, dtype=count.dtype).view(1, 1, K)
                < count.unsqueeze(-1)).unsqueeze(1).float()
        scores = scores * mask + (-1e9) * (1 - mask)
    weight = F.softmax(scores, dim=-1)
    return (values * weight).sum(dim=-1)


class AttentionModule(nn.Module):
    """Neighbour soft-attention pooling: query = centre feature, key = grouped input feature, value =
    MLP output.  attention.py:35-96."""

    def __init__(self, C_in1, C_in2, C1, C2, C_out, attention_bn=True, transform_grouped_feat_out=True,
                 last_activation=True):
        super().__init__()
        C1, C2 = max(C1, 32), max(C2, 32)
        inter_C = min(C1 + C2, C_out)
        self.feat_conv = nn.Conv2d(C_in1, C1, kernel_size=1)
        self.grouped_feat_conv = nn.Conv2d(C_in2, C2, kernel_size=1)
        layers = [nn.ReLU(inplace=True)]
        if attention_bn:
            layers.append(MyGroupNorm(min(32, C1 + C2), C1 + C2))
        layers += [nn.Conv2d(C1 + C2, inter_C, kernel_size=1), nn.ReLU(inplace=True)]
        if attention_bn:
            layers.append(MyGroupNorm(min(32, inter_C), inter_C))
        layers.append(nn.Conv2d(inter_C, C_out, kernel_size=1))
        self.weight_conv = nn.Sequential(*layers)
        self.transform_grouped_feat_out = transform_grouped_feat_out
        if transform_grouped_feat_out:
            out = [nn.Conv2d(C_out, C_out, kernel_size=1)]
            if last_activation:
                if attention_bn:
                    out.append(MyGroupNorm(min(32, C_out), C_out))
                out.append(nn.ReLU(inplace=True))
            self.feat_out_conv = nn.Sequential(*out)

    def forward(self, feat, grouped_feat, grouped_feat_out, count):
        K = grouped_feat.shape[-1]
        q = self.feat_conv(feat.unsqueeze(-1)).expand(-1, -1, -1, K)
        k = self.grouped_feat_conv(grouped_feat)
        scores = self.weight_conv(torch.cat([q, k], dim=1))
        if self.transform_grouped_feat_out:
            grouped_feat_out = self.feat_out_conv(grouped_feat_out)
        return masked_softmax_pool(scores, grouped_feat_out, count)


class GlobalAttentionModule(nn.Module):
    """All-pairs attention over the points of one level (attention.py:98-154).  Not enabled by any
    shipped config; kept for API completeness (O(N^2) memory like the reference)."""

    def __init__(self, C, additional_dim=0, attention_bn=True, last_activation=True):
        super().__init__()
        self.key_conv = nn.Conv2d(C + additional_dim, C, kernel_size=1)
        self.query_conv = nn.Conv2d(C + additional_dim, C, kernel_size=1)
        value = [nn.Conv2d(C + additional_dim, C, kernel_size=1)]
        if last_activation:
            if attention_bn:
                value.append(MyGroupNorm(min(32, C), C))
            value.append(nn.ReLU(inplace=True))
        self.value_conv = nn.Sequential(*value)
        layers = [nn.ReLU(inplace=True)]
        if attention_bn:
            layers.append(MyGroupNorm(min(32, 2 * C), 2 * C))
        layers += [nn.Conv2d(2 * C, C, kernel_size=1), nn.ReLU(inplace=True)]
        if attention_bn:
            layers.append(MyGroupNorm(min(32, C), C))
        layers.append(nn.Conv2d(C, C, kernel_size=1))
        self.weight_conv = nn.Sequential(*layers)

    def forward(self, feat):
        N = feat.shape[2]
        f = feat.unsqueeze(-1)
        key = self.key_conv(f).squeeze(-1).unsqueeze(-2).expand(-1, -1, N, -1)
        query = self.query_conv(f).expand(-1, -1, -1, N)
        value = self.value_conv(f).squeeze(-1)
        weight = F.softmax(self.weight_conv(torch.cat([query, key], dim=1)), dim=-1)
        return (value.unsqueeze(-1) * weight).sum(dim=-1)
