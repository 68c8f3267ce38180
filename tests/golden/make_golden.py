"""Generate the committed golden fixtures by running the REFERENCE's own Python (imported from
/root/reference, never copied) on CPU, with its native modules (`pointnet2_ops._ext`,
`pytorch3d.ops.knn`, `emd_cuda`) bound to the CPU oracle (oracle/cpu_oracle.py).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Outputs (small, committed): tests/golden/*.pt, tests/golden/state_dict_keys.json
(`--only-refinement-io` regenerates refinement_io.pt alone, `--only-fastdpm` fastdpm_loops.pt alone.)
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("PDR_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, ROOT)

from oracle import cpu_oracle as O  # noqa: E402
from tests import common as C  # noqa: E402


def bind_reference_to_oracle():
    """sys.modules shims so that the unmodified reference Python runs on the CPU oracle."""
    import point_diffusion_refinement_b200.dropin as dropin
    ext = types.ModuleType("oracle_ext")
    for name in ("furthest_point_sampling", "gather_points", "ball_query", "group_points", "three_nn",
                 "three_interpolate"):
        setattr(ext, name, getattr(O, name))
    ext.three_nn = lambda u, k: list(O.three_nn(u, k))
    knn = types.ModuleType("oracle_knn")
    knn.knn_points, knn.knn_gather = O.knn_points, O.knn_gather
    knn.Pointclouds = type("Pointclouds", (), {})
    emd = types.ModuleType("oracle_emd")
    emd.approxmatch_forward, emd.matchcost_forward = O.approxmatch_forward, O.matchcost_forward
    dropin.install(ext=ext, knn_module=knn, emd_module=emd)
    for p in (os.path.join(REF, "pointnet2"), REF, os.path.join(REF, "pointnet2_ops_lib")):
        if p not in sys.path:
            sys.path.insert(0, p)


def refinement_io_fixtures():
    """8f rows 1-2: point_upsample (models/point_upsample_module.py:4-27) and mirror_and_concat
    (data_utils/mirror_partial.py:22-37), run from the reference's own files."""
    from pointnet2.models.point_upsample_module import point_upsample as ref_upsample
    from pointnet2.data_utils.mirror_partial import mirror_and_concat as ref_mirror
    g = torch.Generator().manual_seed(11)
    cases = []
    for factor, centre, scale in ((8, False, 0.001), (8, True, 0.001), (2, False, 0.01), (1, False, 1.0), (3, True, 0.5)):
        coarse = torch.rand(3, 50, 3, generator=g) * 2 - 1
        reps = factor - 1 if centre else factor
        disp = torch.randn(3, 50, 3 * (reps + 1), generator=g)
        refined, mid = ref_upsample(coarse, disp, factor, centre, scale)
        cases.append({"factor": factor, "centre": centre, "scale": scale, "coarse": coarse, "disp": disp,
                      "refined": refined.clone(), "mid": mid.clone()})
    partial = torch.rand(2, 300, 3, generator=g) * 2 - 1
    real_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self      # mirror_partial.py:32 calls .cuda()
    try:
        mirrored = [t.clone() for t in ref_mirror(partial, axis=2, num_points=[256, 384])]
        mirrored_ax1 = [t.clone() for t in ref_mirror(partial, axis=1, num_points=[128])]
    finally:
        torch.Tensor.cuda = real_cuda
    torch.save({"upsample": cases, "partial": partial, "mirror_axis2_256_384": mirrored, "mirror_axis1_128": mirrored_ax1},
               os.path.join(HERE, "refinement_io.pt"))
    print("refinement_io.pt written")


def stub_net(x, ts=None, label=None):
    """Deterministic stand-in for eps_theta used by the FastDPM loop fixture: nonlinear in x and in the
    (continuous) step, no parameters, roughly x / sqrt(1 - alpha_bar) so that the chain stays O(1).  tests/test_host_model.py and tests/test_model_gpu.py hold the same
    three lines; the reference's own loop checker uses `lambda x, ts, label: x` (util_fastdpmv2.py:480)."""
    t = ts.to(x.dtype).view(-1, 1, 1) / 1000.0
    return 0.9 * x / torch.sqrt(1.001 - torch.exp(-(0.1 * t + 10.0 * t * t))) + 0.1 * torch.tanh(x + t)


def fastdpm_fixtures():
    """a14: the reference's VAR_sampling / STEP_sampling update loops (util_fastdpmv2.py:307-452) run from the
    reference's own file with a stub eps_theta and an injected noise bank (its std_normal replaced by a
    replay of pre-drawn tensors, its .cuda() calls made no-ops)."""
    import pointnet2.util as ref_util
    import pointnet2.util_fastdpmv2 as ref_fast
    from point_diffusion_refinement_b200 import configs
    dh = ref_util.calc_diffusion_hyperparams(**configs.DIFFUSION_CONFIG)
    dh64 = dict(dh, Beta=dh["Beta"].double())       # float64 Stirling: see schedules.pt "taus_f64"
    size = (3, 40, 3)
    cases = []
    real_cuda, real_normal = torch.Tensor.cuda, ref_fast.std_normal
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        for method, schedule, length, kappa in (("var", "quadratic", 12, 0.5), ("var", "linear", 8, 0.0),
                                                ("var", "quadratic", 50, 1.0), ("step", "quadratic", 12, 0.5),
                                                ("step", "linear", 9, 0.2), ("step", "quadratic", 50, 0.0)):
            g = torch.Generator().manual_seed(1000 + len(cases))
            bank = [torch.randn(size, generator=g) for _ in range(length + 1)]   # [x_T, z of step 0, 1, ...]
            drawn = []

            def replay(sz):
                assert tuple(sz) == size
                drawn.append(len(drawn))
                return bank[len(drawn) - 1].clone()
            ref_fast.std_normal = replay
            seen = []

            def net(x, ts=None, label=None):
                seen.append(x.clone())
                return stub_net(x, ts=ts, label=label)
            if method == "var":
                eta = ref_fast.get_VAR_noise(length, configs.DIFFUSION_CONFIG, schedule)
                taus = ref_fast._precompute_VAR_steps(dh64, eta)
                x = ref_fast.VAR_sampling(net, size, dh, eta, kappa, taus, label=None, verbose=False)
            else:
                steps = ref_fast.get_STEP_step(length, configs.DIFFUSION_CONFIG, schedule)
                x = ref_fast.STEP_sampling(net, size, dh, steps, kappa, label=None, verbose=False)
            assert len(drawn) == length + 1     # the reference draws z every step, also when sigma == 0
            cases.append({"method": method, "schedule": schedule, "length": length, "kappa": kappa,
                          "bank": torch.stack(bank), "x_in": torch.stack(seen), "x0": x.clone()})
            print("fastdpm", method, schedule, length, kappa, "x0 abs-mean %.4f" % x.abs().mean().item())
    finally:
        torch.Tensor.cuda, ref_fast.std_normal = real_cuda, real_normal
    torch.save({"size": size, "cases": cases}, os.path.join(HERE, "fastdpm_loops.pt"))
    print("fastdpm_loops.pt written")


def main():
    bind_reference_to_oracle()
    if "--only-refinement-io" in sys.argv:
        refinement_io_fixtures()
        return
    if "--only-fastdpm" in sys.argv:
        fastdpm_fixtures()
        return
    from pointnet2.models.pointnet2_with_pcld_condition import PointNet2CloudCondition as RefNet
    from pointnet2.models.pointnet2_ssg_sem import PointNet2SemSegSSG as RefSSG
    import pointnet2.util as ref_util  # reference util.py (schedule)
    import pointnet2.util_fastdpmv2 as ref_fast
    from point_diffusion_refinement_b200 import configs

    out = {}
    # ---- 1. state_dict contract of the shipped DDPM config -------------------------------------------
    torch.manual_seed(0)
    net = RefNet(configs.ddpm_pointnet_config()).eval()
    keys = {k: list(v.shape) for k, v in net.state_dict().items()}
    refine = RefNet(configs.refine_pointnet_config(8)).eval()
    keys_refine = {k: list(v.shape) for k, v in refine.state_dict().items()}
    with open(os.path.join(HERE, "state_dict_keys.json"), "w") as f:
        json.dump({"ddpm": keys, "refine_x8": keys_refine}, f, indent=0, sort_keys=True)
    print("ddpm params: %.3f M in %d tensors" % (sum(np.prod(s) for s in keys.values()) / 1e6, len(keys)))

    # ---- 2. one eps_theta call, tiny pyramid (CPU-test sized), cold and warm --------------------------
    for tag, cfg, B, N, M in (("tiny", configs.tiny_pointnet_config(), 2, 256, 384),
                              ("full", configs.ddpm_pointnet_config(), 2, 2048, 3072)):
        net = RefNet(cfg).eval()
        C.fill_parameters_(net, seed=1)
        x, cond, ts, label = C.denoiser_inputs(B, N, M, seed=3)
        with torch.no_grad():
            cold = net(x, cond, ts=ts, label=label, use_retained_condition_feature=True)
            x2 = x + 0.05 * cold
            warm = net(x2, cond, ts=ts - 1, label=label, use_retained_condition_feature=True)
            net.reset_cond_features()
            plain = net(x, cond, ts=ts, label=label, use_retained_condition_feature=False)
        assert torch.equal(plain, cold)
        torch.save({"cfg": tag, "B": B, "N": N, "M": M, "param_seed": 1, "input_seed": 3,
                    "eps_cold": cold.clone(), "eps_warm": warm.clone()},
                   os.path.join(HERE, "denoiser_%s.pt" % tag))
        print(tag, "eps cold/warm abs-mean", cold.abs().mean().item(), warm.abs().mean().item())

    # refinement net (include_t False, upsample x2 head)
    cfg = configs.tiny_pointnet_config()
    cfg["include_t"] = False
    cfg["point_upsample_factor"] = 2
    cfg["include_displacement_center_to_final_output"] = False
    net = RefNet(cfg).eval()
    C.fill_parameters_(net, seed=2)
    x, cond, ts, label = C.denoiser_inputs(2, 256, 384, seed=4)
    with torch.no_grad():
        disp = net(x, cond, ts=None, label=label)
    torch.save({"disp": disp.clone(), "param_seed": 2, "input_seed": 4}, os.path.join(HERE, "refiner_tiny.pt"))

    # ---- 3. unconditional PointNet2SemSegSSG with three_nn / three_interpolate decoder ----------------
    ssg_cfg = C.ssg_config()
    net = RefSSG(ssg_cfg).eval()
    C.fill_parameters_(net, seed=5)
    g = torch.Generator().manual_seed(6)
    pc = torch.randn(2, 256, 3, generator=g)
    with torch.no_grad():
        y = net(pc, ts=torch.tensor([10.0, 500.0]), label=torch.tensor([1, 7]))
    torch.save({"out": y.clone(), "keys": {k: list(v.shape) for k, v in net.state_dict().items()}},
               os.path.join(HERE, "ssg_tiny.pt"))

    # ---- 4. schedules -------------------------------------------------------------------------------
    dh = ref_util.calc_diffusion_hyperparams(**configs.DIFFUSION_CONFIG)
    sched = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in dh.items()}
    fast = {}
    real_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self  # reference calls .cuda() on schedule tensors
    try:
        for length, schedule in ((50, "quadratic"), (20, "linear")):
            eta = ref_fast.get_VAR_noise(length, configs.DIFFUSION_CONFIG, schedule)
            taus = ref_fast._precompute_VAR_steps(dh, eta)          # as it runs under this numpy (fp32 Stirling)
            dh64 = dict(dh, Beta=dh["Beta"].double())               # same fp32 endpoints, float64 Stirling
            taus64 = ref_fast._precompute_VAR_steps(dh64, eta)
            steps = ref_fast.get_STEP_step(length, configs.DIFFUSION_CONFIG, schedule)
            fast["%d_%s" % (length, schedule)] = {"eta": torch.from_numpy(np.asarray(eta)),
                                                  "taus_as_run_here": torch.tensor(taus, dtype=torch.float64),
                                                  "taus_f64": torch.tensor(taus64, dtype=torch.float64),
                                                  "steps": steps}
    finally:
        torch.Tensor.cuda = real_cuda
    torch.save({"ddpm": sched, "fast": fast}, os.path.join(HERE, "schedules.pt"))

    # ---- 5. reference-owned known answers ------------------------------------------------------------
    # (a) PytorchEMD/test_emd_loss.py:6-19: 2x2 clouds, optimal assignment cost 0.71 -> /max(n,m) = 0.355
    # (b) ChamferDistancePytorch/unit_test.py:14-35: chamfer vs float64 brute force, rand(4,100,3) x rand(4,200,3)
    g = torch.Generator().manual_seed(7)
    p1, p2 = torch.rand(4, 100, 3, generator=g), torch.rand(4, 200, 3, generator=g)
    P = ((p1.double()[:, :, None, :] - p2.double()[:, None, :, :]) ** 2).sum(-1)
    torch.save({"p1": p1, "p2": p2, "dist1": P.min(2)[0].float(), "dist2": P.min(1)[0].float(),
                "idx1": P.min(2)[1].int(), "idx2": P.min(1)[1].int()}, os.path.join(HERE, "chamfer_f64.pt"))
    refinement_io_fixtures()
    fastdpm_fixtures()
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
