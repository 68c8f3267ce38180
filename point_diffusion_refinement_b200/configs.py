"""The model hyper-parameters of the shipped experiments, as plain dicts in the form the reference's
models take (``pointnet_config`` after ``json_reader.restore_string_to_list_in_a_dict``).

Values restate pointnet2/exp_configs/mvp_configs/
config_standard_attention_real_3072_partial_points_rot_90_scale_1.2_translation_0.1.json:2-78 (DDPM) and
config_refine_and_upsample_*_standard_attention_10_trials.json (refinement; they differ from the DDPM
net only by ``include_t: false`` and the upsampling factor).  They live here because /root/reference
does not exist on the GPU box.
"""
import copy

DIFFUSION_CONFIG = {"T": 1000, "beta_0": 0.0001, "beta_T": 0.02}


def _pyramid(points_in=2048, cond_points=3072, scale=1.0):
    """4-level pyramid: 1024/256/64/16 centres, radii .1/.2/.4/.8, 32 neighbours, kNN(8) decoder."""
    return dict(npoint=[1024, 256, 64, 16], radius=[0.1, 0.2, 0.4, 0.8], neighbor_definition="radius",
                nsample=[32, 32, 32, 32], mlp_depth=3, include_grouper=False, decoder_mlp_depth=2,
                use_knn_FP=True, K=8)


def ddpm_pointnet_config():
    attention = dict(use_attention_module=True, attention_bn=True, transform_grouped_feat_out=True,
                     last_activation=True, add_attention_to_FeatureMapper_module=True)
    arch = dict(_pyramid(), feature_dim=[32, 64, 128, 256, 512], decoder_feature_dim=[128, 128, 256, 256, 512])
    cond = dict(_pyramid(), feature_dim=[32, 32, 64, 64, 128], decoder_feature_dim=[32, 32, 64, 64, 128])
    mapper = dict(neighbor_definition="radius", encoder_feature_map_dim=[32, 32, 64, 64], encoder_mlp_depth=2,
                  encoder_radius=[0.1, 0.2, 0.4, 0.8], encoder_nsample=[32, 32, 32, 32],
                  decoder_feature_map_dim=[32, 32, 64, 64, 128], decoder_mlp_depth=2,
                  decoder_radius=[0.1, 0.2, 0.4, 0.8, 1.6], decoder_nsample=[32, 32, 32, 32, 32])
    cfg = {
        "model_name": "shape_completion_mirror_rot_90_scale_1.2_translation_0.1",
        "in_fea_dim": 0, "partial_in_fea_dim": 1, "out_dim": 3, "include_t": True, "t_dim": 128,
        "model.use_xyz": True, "attach_position_to_input_feature": True, "include_abs_coordinate": True,
        "include_center_coordinate": True, "record_neighbor_stats": False, "bn_first": False, "bias": True,
        "res_connect": True, "include_class_condition": True, "num_class": 16, "class_condition_dim": 128,
        "bn": True, "include_local_feature": True, "include_global_feature": True,
        "global_feature_remove_last_activation": False,
        "pnet_global_feature_architecture": [[4, 128, 256], [512, 1024]],
        "attention_setting": attention, "architecture": arch, "condition_net_architecture": cond,
        "feature_mapper_architecture": mapper,
    }
    return copy.deepcopy(cfg)


def refine_pointnet_config(point_upsample_factor=1):
    cfg = ddpm_pointnet_config()
    cfg["include_t"] = False
    if point_upsample_factor > 1:
        cfg["point_upsample_factor"] = point_upsample_factor
        cfg["include_displacement_center_to_final_output"] = False
        cfg["intermediate_refined_X_loss_weight"] = 0
    return cfg


def tiny_pointnet_config():
    """Same topology, 4x fewer points per level and halved radii scale unchanged -- for CPU-sized tests."""
    cfg = ddpm_pointnet_config()
    for key in ("architecture", "condition_net_architecture"):
        cfg[key]["npoint"] = [128, 64, 32, 16]
        cfg[key]["radius"] = [0.3, 0.5, 0.8, 1.2]
    m = cfg["feature_mapper_architecture"]
    m["encoder_radius"] = [0.3, 0.5, 0.8, 1.2]
    m["decoder_radius"] = [0.3, 0.5, 0.8, 1.2, 1.6]
    return cfg
