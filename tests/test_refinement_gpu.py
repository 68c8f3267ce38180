"""GPU: SURVEY 8f rows -- refinement network on the compiled program, point_upsample, mirror-partial
preprocessing and the evaluate() loop, against the reference-generated fixtures and the CPU oracle.

Tolerances: point_upsample / mirror_and_concat bit-exact (separately rounded fp32 elementwise ops; FPS indices);
refiner displacement rtol = atol = 1e-4 in fp32 (accumulation order of the 1x1 convolutions)."""
import os

import pytest
import torch

from tests import common as C

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _net(cfg, seed):
    from point_diffusion_refinement_b200.pointnet2_with_pcld_condition import PointNet2CloudCondition
    return C.fill_parameters_(PointNet2CloudCondition(cfg).eval(), seed=seed).to(DEV)


def test_point_upsample_bit_exact(golden_dir, oracle):
    from point_diffusion_refinement_b200.point_upsample_module import point_upsample
    g = torch.load(golden_dir + "/refinement_io.pt")
    for case in g["upsample"]:
        refined, mid = point_upsample(case["coarse"].to(DEV), case["disp"].to(DEV), case["factor"], case["centre"], case["scale"])
        assert torch.equal(refined.cpu(), case["refined"]) and torch.equal(mid.cpu(), case["mid"]), case["factor"]
    # the shipped configuration: 2048 -> 16384 points (factor 8, no centre), scale 0.001
    gen = torch.Generator().manual_seed(3)
    coarse = torch.rand(4, 2048, 3, generator=gen) * 2 - 1
    disp = torch.randn(4, 2048, 27, generator=gen)
    refined, mid = point_upsample(coarse.to(DEV), disp.to(DEV), 8, False, 0.001)
    o_ref, o_mid = oracle.point_upsample(coarse, disp, 8, False, 0.001)
    assert refined.shape == (4, 16384, 3) and torch.equal(refined.cpu(), o_ref) and torch.equal(mid.cpu(), o_mid)
    with pytest.raises(RuntimeError):
        point_upsample(coarse.to(DEV), disp.to(DEV), 8, True, 0.001)       # wrong displacement width
    with pytest.raises(RuntimeError):
        point_upsample(coarse, disp, 8, False, 0.001)                      # CPU tensors are rejected


def test_mirror_and_concat_bit_exact(golden_dir, oracle):
    from point_diffusion_refinement_b200.mirror_partial import mirror_and_concat
    g = torch.load(golden_dir + "/refinement_io.pt")
    out = mirror_and_concat(g["partial"].to(DEV), axis=2, num_points=[256, 384])
    assert all(torch.equal(a.cpu(), b) for a, b in zip(out, g["mirror_axis2_256_384"]))
    out = mirror_and_concat(g["partial"], axis=1, num_points=[128])        # CPU input is moved like the reference does
    assert all(torch.equal(a.cpu(), b) for a, b in zip(out, g["mirror_axis1_128"]))
    # full size: 2048-point partial -> 4096 -> 2048 / 3072 (BASELINE configs[0] shape: FPS over 4096 points)
    partial = torch.rand(3, 2048, 3, generator=torch.Generator().manual_seed(5)) * 2 - 1
    ours = mirror_and_concat(partial.to(DEV), axis=2, num_points=[2048, 3072])
    ref = oracle.mirror_and_concat(partial, axis=2, num_points=[2048, 3072])
    assert [tuple(t.shape) for t in ours] == [(3, 4096, 4), (3, 2048, 4), (3, 3072, 4)]
    assert all(torch.equal(a.cpu(), b) for a, b in zip(ours, ref))


@pytest.mark.parametrize("use_graph", [False, True])
def test_refiner_on_compiled_program_matches_modules_and_fixture(golden_dir, use_graph):
    """Refinement network = cold call with ts=None (completion_eval.py:159-163): x-branch on the compiled program."""
    from point_diffusion_refinement_b200 import configs
    gold = torch.load(golden_dir + "/refiner_tiny.pt")
    cfg = configs.tiny_pointnet_config()
    cfg.update(include_t=False, point_upsample_factor=2, include_displacement_center_to_final_output=False)
    net = _net(cfg, gold["param_seed"])
    x, cond, ts, label = [t.to(DEV) for t in C.denoiser_inputs(2, 256, 384, seed=gold["input_seed"])]
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            mod = net(x, cond, ts=None, label=label)
            net.enable_fused(True, use_tf32=False, use_graph=use_graph, fuse_cold=True)
            fused = net(x, cond, ts=None, label=label)
            # a second batch with another condition cloud through the same compiled program
            x2, cond2, _, label2 = [t.to(DEV) for t in C.denoiser_inputs(2, 256, 384, seed=9)]
            fused2 = net(x2, cond2, ts=None, label=label2)
            again = net(x, cond, ts=None, label=label)
            net.enable_fused(False)
            mod2 = net(x2, cond2, ts=None, label=label2)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    assert fused.shape == (2, 256, 9) and torch.equal(fused, again)
    torch.testing.assert_close(mod.cpu(), gold["disp"], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(fused.cpu(), gold["disp"], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(fused2, mod2, rtol=1e-4, atol=1e-4)


def test_evaluate_loop_refine_and_completion(tmp_path, oracle):
    """evaluate() (completion_eval.py:65-330): refine_completion with x2 upsampling and completion from a
    precomputed x_T; metrics re-derived with the CPU oracle from the clouds it saved."""
    from point_diffusion_refinement_b200 import completion_eval, configs, results_io, util
    cfg = configs.tiny_pointnet_config()
    cfg.update(include_t=False, point_upsample_factor=2, include_displacement_center_to_final_output=True)
    refiner = _net(cfg, 2)
    gen = torch.Generator().manual_seed(21)
    loader = []
    for b in range(2):
        _, cond, _, label = C.denoiser_inputs(2, 256, 384, seed=30 + b)
        complete = torch.rand(2, 512, 3, generator=gen) * 2 - 1
        coarse = complete[:, :256] + 0.01 * torch.randn(2, 256, 3, generator=gen)
        loader.append({"label": label, "partial": cond, "complete": complete, "generated": coarse,
                       "XT": torch.randn(2, 256, 3, generator=gen)})
    dh = util.calc_diffusion_hyperparams(**configs.DIFFUSION_CONFIG)
    save_dir = str(tmp_path / "refine")
    cd, emd, meta, cd_all, emd_all = completion_eval.evaluate(
        refiner, loader, dh, dataset="mvp_dataset", scale=1, task="refine_completion", refine_output_scale_factor=0.001,
        point_upsample_factor=2, include_displacement_center_to_final_output=True, save_generated_samples=True,
        save_dir=save_dir, use_tf32=False)
    assert meta.tolist() == torch.cat([d["label"] for d in loader]).tolist() and cd_all.shape == (4,)
    saved = torch.from_numpy(results_io.load_generated(os.path.join(save_dir, "mvp_generated_data_512pts.h5")))
    assert saved.shape == (4, 512, 3)
    gt = torch.cat([d["complete"] for d in loader]) / 2
    _, o_cd, _ = oracle.chamfer_f1(saved, gt)
    torch.testing.assert_close(cd_all.cpu(), o_cd, rtol=1e-5, atol=1e-8)
    torch.testing.assert_close(emd_all.cpu(), oracle.emd_distance(saved, gt), rtol=2e-3, atol=1e-6)
    assert abs(cd - o_cd.mean().item()) < 1e-6 and abs(emd - emd_all.mean().item()) < 1e-6
    # the refined cloud is coarse + 0.001 * (small displacement): it stays close to the coarse input
    coarse_all = torch.cat([d["generated"] for d in loader]) / 2
    assert (saved[:, 256:] - coarse_all).abs().max() < 0.05
    # completion from a precomputed x_T (3 reverse steps), no EMD
    denoiser = _net(configs.tiny_pointnet_config(), 1)
    for d in loader:
        d["complete"] = d["complete"][:, :256].contiguous()
    out = completion_eval.evaluate(denoiser, loader, dh, dataset="mvp_dataset", task="completion", use_a_precomputed_XT=True,
                                   T_step=3, compute_emd=False, return_all_metrics=True, seed=4, use_tf32=False)
    assert out[3]["emd_distance"].abs().max() == 0 and torch.isfinite(out[3]["cd_distance"]).all()
    assert set(out[3]) == {"cd_distance", "emd_distance", "cd_p", "f1"}


@pytest.mark.parametrize("tag", ["tiny", "full"])
def test_cold_step_on_compiled_condition_program(golden_dir, tag):
    """fuse_cold: the condition branch (SA_modules_condition / FP_modules_condition) as a compiled program feeding
    the x-branch program -- cold and warm eps against the reference-Python fixture, fp32."""
    from point_diffusion_refinement_b200 import configs
    gold = torch.load(golden_dir + "/denoiser_%s.pt" % tag)
    cfg = configs.tiny_pointnet_config() if tag == "tiny" else configs.ddpm_pointnet_config()
    net = _net(cfg, gold["param_seed"])
    x, cond, ts, label = [t.to(DEV) for t in C.denoiser_inputs(gold["B"], gold["N"], gold["M"], seed=gold["input_seed"])]
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            net.enable_fused(True, use_tf32=False, use_graph=True, fuse_cold=True)
            cold = net(x, cond, ts=ts, label=label, use_retained_condition_feature=True)
            assert net.l_uvw is not None and len(net.encoder_cond_features) == 5
            fused_state = [f.clone() for f in net.encoder_cond_features + net.decoder_cond_features]
            x2 = x + 0.05 * gold["eps_cold"].to(DEV)
            warm = net(x2, cond, ts=ts - 1, label=label, use_retained_condition_feature=True)
            net.reset_cond_features()
            net.enable_fused(False)
            net(x, cond, ts=ts, label=label, use_retained_condition_feature=True)
            mod_state = net.encoder_cond_features + net.decoder_cond_features
            for a, b in zip(fused_state, mod_state):       # retained condition features, layer by layer
                torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-4)
            net.reset_cond_features()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    torch.testing.assert_close(cold.cpu(), gold["eps_cold"], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(warm.cpu(), gold["eps_warm"], rtol=1e-4, atol=1e-4)
