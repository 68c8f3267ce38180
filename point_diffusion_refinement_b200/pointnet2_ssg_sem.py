"""``PointNet2SemSegSSG``: the unconditional PointNet++ denoiser and the builder base class of the
conditional network (reference: pointnet2/models/pointnet2_ssg_sem.py).  Accepts the reference's
``hparams`` / ``pointnet_config`` dict verbatim and registers parameters under the reference's names
(``SA_modules``, ``FP_modules``, ``fc_t1``, ``fc_t2``, ``class_emb``, ``fc_lyaer`` [sic])."""
import math

import torch
import torch.nn as nn

from .pointnet2_modules import PointnetFPModule, PointnetKnnFPModule, PointnetSAModule


def swish(x):
    return x * torch.sigmoid(x)


def calc_t_emb(ts, t_emb_dim):
    """Sinusoidal step embedding (B,) -> (B, t_emb_dim): [sin(t*f_i), cos(t*f_i)], f_i = 1e4^(-i/(h-1)).
    pointnet2_ssg_sem.py:14-31."""
    assert t_emb_dim % 2 == 0
    half = t_emb_dim // 2
    freq = torch.exp(torch.arange(half) * -(math.log(10000) / (half - 1))).to(ts.device)
    arg = ts.unsqueeze(1) * freq
    return torch.cat((torch.sin(arg), torch.cos(arg)), 1)


class PointNet2SemSegSSG(nn.Module):
    def __init__(self, hparams):
        super().__init__()
        self.hparams = hparams
        self._build_model()

    @staticmethod
    def _break_up_pc(pc):
        xyz = pc[..., 0:3].contiguous()
        features = pc[..., 3:].transpose(1, 2).contiguous() if pc.size(-1) > 3 else None
        return xyz, features

    # -- builders (pointnet2_ssg_sem.py:47-177) --------------------------------------------------------
    def _condition_plan(self, include_class_condition, class_condition_dim, include_global_feature,
                        global_feature_dim):
        cdim = self.hparams["class_condition_dim"] if class_condition_dim is None else class_condition_dim
        if include_global_feature:
            return dict(include_condition=True, condition_dim=global_feature_dim,
                        include_second_condition=include_class_condition, second_condition_dim=cdim)
        return dict(include_condition=include_class_condition, condition_dim=cdim,
                    include_second_condition=False, second_condition_dim=None)

    def build_SA_model(self, npoint, radius, nsample, feature_dim, mlp_depth, in_fea_dim, include_t,
                       include_class_condition, class_condition_dim=None, include_global_feature=False,
                       global_feature_dim=None, additional_fea_dim=None, neighbor_def="radius",
                       activation="relu", bn=True, attention_setting=None, global_attention_setting=None):
        hp = self.hparams
        if not isinstance(neighbor_def, list):
            neighbor_def = [neighbor_def] * len(radius)
        plan = self._condition_plan(include_class_condition, class_condition_dim, include_global_feature,
                                    global_feature_dim)
        modules = nn.ModuleList()
        for i in range(len(npoint)):
            mlp_spec = [feature_dim[i]] * mlp_depth + [feature_dim[i + 1]]
            if additional_fea_dim is not None:
                mlp_spec[0] += additional_fea_dim[i]
            first_conv = hp["bn_first"] and i == 0
            if i == 0 and not first_conv:
                mlp_spec[0] = in_fea_dim
            use_ga = bool(global_attention_setting and global_attention_setting["use_global_attention_module"]
                          and i in global_attention_setting["global_attention_layer_index"])
            modules.append(PointnetSAModule(
                npoint=npoint[i], radius=radius[i], nsample=nsample[i], mlp=mlp_spec,
                use_xyz=hp["model.use_xyz"], t_dim=4 * hp["t_dim"], include_t=include_t,
                include_abs_coordinate=self.include_abs_coordinate,
                include_center_coordinate=hp.get("include_center_coordinate", False),
                bn_first=hp["bn_first"], first_conv=first_conv, first_conv_in_channel=in_fea_dim,
                res_connect=hp["res_connect"], bias=hp["bias"], neighbor_def=neighbor_def[i],
                activation=activation, bn=bn, attention_setting=attention_setting,
                global_attention_setting=global_attention_setting if use_ga else None, **plan))
        return modules

    def build_FP_model(self, decoder_feature_dim, decoder_mlp_depth, feature_dim, in_fea_dim, include_t,
                       include_class_condition, class_condition_dim=None, include_global_feature=False,
                       global_feature_dim=None, additional_fea_dim=None, use_knn_FP=False, K=3,
                       include_grouper=False, radius=[0], nsample=[32], neighbor_def="radius",
                       activation="relu", bn=True, attention_setting=None, global_attention_setting=None):
        hp = self.hparams
        if not isinstance(neighbor_def, list):
            neighbor_def = [neighbor_def] * len(radius)
        plan = self._condition_plan(include_class_condition, class_condition_dim, include_global_feature,
                                    global_feature_dim)
        common = dict(first_conv=False, bn=bn, t_dim=4 * hp["t_dim"], include_t=include_t, bn_first=hp["bn_first"],
                      res_connect=hp["res_connect"], bias=hp["bias"], include_grouper=include_grouper,
                      use_xyz=hp["model.use_xyz"], include_abs_coordinate=self.include_abs_coordinate,
                      include_center_coordinate=hp.get("include_center_coordinate", False),
                      activation=activation, **plan)
        modules = nn.ModuleList()
        for i in range(len(decoder_feature_dim) - 1):
            skip_dim = in_fea_dim if i == 0 else feature_dim[i]
            extra_in = additional_fea_dim[i] if additional_fea_dim is not None else 0
            if use_knn_FP:
                mlp1 = [decoder_feature_dim[i + 1] + extra_in] + [decoder_feature_dim[i]] * decoder_mlp_depth
                mlp2 = [decoder_feature_dim[i] + skip_dim] + [decoder_feature_dim[i]] * decoder_mlp_depth
                use_ga = bool(global_attention_setting and global_attention_setting["use_global_attention_module"]
                              and i in global_attention_setting["global_attention_layer_index"])
                modules.append(PointnetKnnFPModule(
                    mlp1=mlp1, mlp2=mlp2, K=K, radius=radius[i], nsample=nsample[i], neighbor_def=neighbor_def[i],
                    attention_setting=attention_setting,
                    global_attention_setting=global_attention_setting if use_ga else None, **common))
            else:
                mlp = [decoder_feature_dim[i + 1] + skip_dim + extra_in] + [decoder_feature_dim[i]] * decoder_mlp_depth
                modules.append(PointnetFPModule(mlp=mlp, radius=radius[i], nsample=nsample[i],
                                                neighbor_def=neighbor_def[i], **common))
        return modules

    def _build_head(self, in_dim, activation_module=None, bn=True):
        hp = self.hparams
        act = activation_module if activation_module is not None else nn.ReLU(True)
        if hp["bn_first"]:
            return nn.Sequential(act, nn.Conv1d(in_dim, hp["out_dim"], kernel_size=1))
        layers = [nn.Conv1d(in_dim, 128, kernel_size=1, bias=hp["bias"])]
        if bn:
            layers.append(nn.GroupNorm(32, 128))
        layers += [act, nn.Conv1d(128, hp["out_dim"], kernel_size=1)]
        return nn.Sequential(*layers)

    def _build_model(self):
        hp = self.hparams
        self.record_neighbor_stats = hp["record_neighbor_stats"]
        self.scale_factor = hp["scale_factor"]
        if hp["include_class_condition"]:
            self.class_emb = nn.Embedding(hp["num_class"], hp["class_condition_dim"])
        self.attach_position_to_input_feature = hp["attach_position_to_input_feature"]
        in_fea_dim = hp["in_fea_dim"] + (3 if self.attach_position_to_input_feature else 0)
        self.include_abs_coordinate = hp["include_abs_coordinate"]
        t_dim = hp["t_dim"]
        self.fc_t1 = nn.Linear(t_dim, 4 * t_dim)
        self.fc_t2 = nn.Linear(4 * t_dim, 4 * t_dim)
        self.activation = swish
        arch = hp["architecture"]
        self.SA_modules = self.build_SA_model(arch["npoint"], arch["radius"], arch["nsample"], arch["feature_dim"],
                                              arch["mlp_depth"], in_fea_dim, hp["include_t"],
                                              hp["include_class_condition"])
        dec = arch["decoder_feature_dim"]
        assert dec[-1] == arch["feature_dim"][-1]
        self.use_knn_FP = hp.get("use_knn_FP", False)
        self.K = hp.get("K", 3)
        self.FP_modules = self.build_FP_model(dec, arch["decoder_mlp_depth"], arch["feature_dim"], in_fea_dim,
                                              hp["include_t"], hp["include_class_condition"],
                                              use_knn_FP=self.use_knn_FP, K=self.K)
        self.fc_lyaer = self._build_head(dec[0] + (3 if self.use_knn_FP else 0))

    def embed_t(self, ts):
        if ts is None or not self.hparams["include_t"]:
            return None
        t = self.activation(self.fc_t1(calc_t_emb(ts, self.hparams["t_dim"])))
        return self.activation(self.fc_t2(t))

    def forward(self, pointcloud, ts=None, label=None):
        """(B,N,3+C) -> (B,N,out_dim).  pointnet2_ssg_sem.py:240-312."""
        if self.attach_position_to_input_feature:
            pointcloud = torch.cat([pointcloud, pointcloud[:, :, 0:3] / self.scale_factor], dim=2)
        xyz, features = self._break_up_pc(pointcloud)
        xyz = xyz / self.scale_factor
        t_emb = self.embed_t(ts)
        class_emb = (self.class_emb(label) if (label is not None and self.hparams["include_class_condition"])
                     else None)
        l_xyz, l_features = [xyz], [features]
        for sa in self.SA_modules:
            li_xyz, li_features = sa(l_xyz[-1], l_features[-1], t_emb=t_emb, condition_emb=class_emb,
                                     record_neighbor_stats=self.record_neighbor_stats)
            l_xyz.append(li_xyz)
            l_features.append(li_features)
        for i in range(-1, -(len(self.FP_modules) + 1), -1):
            l_features[i - 1] = self.FP_modules[i](l_xyz[i - 1], l_xyz[i], l_features[i - 1], l_features[i],
                                                   t_emb=t_emb, condition_emb=class_emb,
                                                   record_neighbor_stats=self.record_neighbor_stats)
        out_feature = l_features[0]
        if self.use_knn_FP:
            out_feature = torch.cat([out_feature, xyz.transpose(1, 2)], dim=1)
        return self.fc_lyaer(out_feature).transpose(1, 2)
