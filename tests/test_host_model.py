"""CPU: host-side logic against fixtures produced by the reference's own Python (tests/golden/
make_golden.py).  The package's native entry points are bound to the CPU oracle for these tests only."""
import json

import numpy as np
import pytest
import torch

from tests import common as C


@pytest.fixture(scope="module")
def keys(golden_dir):
    return json.load(open(golden_dir + "/state_dict_keys.json"))


def test_state_dict_contract(keys):
    """Reference checkpoints load key for key (generate_samples.py:178-179)."""
    from point_diffusion_refinement_b200 import configs
    from point_diffusion_refinement_b200.pointnet2_with_pcld_condition import PointNet2CloudCondition
    net = PointNet2CloudCondition(configs.ddpm_pointnet_config())
    assert {k: list(v.shape) for k, v in net.state_dict().items()} == keys["ddpm"]
    assert sum(p.numel() for p in net.parameters()) == 9758959 or abs(sum(p.numel() for p in net.parameters()) - 9.759e6) < 2e3
    net = PointNet2CloudCondition(configs.refine_pointnet_config(8))
    assert {k: list(v.shape) for k, v in net.state_dict().items()} == keys["refine_x8"]
    assert net.fc_lyaer[-1].out_channels == 27       # 3 * (8 + 1) displacement channels


@pytest.mark.parametrize("tag", ["tiny", "full"])
def test_denoiser_matches_reference_python(oracle, golden_dir, tag):
    from point_diffusion_refinement_b200 import configs
    from point_diffusion_refinement_b200.pointnet2_with_pcld_condition import PointNet2CloudCondition
    gold = torch.load(golden_dir + "/denoiser_%s.pt" % tag)
    cfg = configs.tiny_pointnet_config() if tag == "tiny" else configs.ddpm_pointnet_config()
    net = C.fill_parameters_(PointNet2CloudCondition(cfg).eval(), seed=gold["param_seed"])
    x, cond, ts, label = C.denoiser_inputs(gold["B"], gold["N"], gold["M"], seed=gold["input_seed"])
    with C.package_bound_to_oracle(), torch.no_grad():
        cold = net(x, cond, ts=ts, label=label, use_retained_condition_feature=True)
        assert net.l_uvw is not None and len(net.encoder_cond_features) == 5
        warm = net(x + 0.05 * cold, cond, ts=ts - 1, label=label, use_retained_condition_feature=True)
        net.reset_cond_features()
        assert net.l_uvw is None
    # identical ops + identical module arithmetic on the same CPU -> identical bits
    assert torch.equal(cold, gold["eps_cold"]) and torch.equal(warm, gold["eps_warm"])


def test_refiner_and_ssg_match_reference_python(oracle, golden_dir):
    from point_diffusion_refinement_b200 import configs
    from point_diffusion_refinement_b200.pointnet2_with_pcld_condition import PointNet2CloudCondition
    from point_diffusion_refinement_b200.pointnet2_ssg_sem import PointNet2SemSegSSG
    gold = torch.load(golden_dir + "/refiner_tiny.pt")
    cfg = configs.tiny_pointnet_config()
    cfg.update(include_t=False, point_upsample_factor=2, include_displacement_center_to_final_output=False)
    net = C.fill_parameters_(PointNet2CloudCondition(cfg).eval(), seed=gold["param_seed"])
    x, cond, ts, label = C.denoiser_inputs(2, 256, 384, seed=gold["input_seed"])
    with C.package_bound_to_oracle(), torch.no_grad():
        disp = net(x, cond, ts=None, label=label)
    assert disp.shape == (2, 256, 9) and torch.equal(disp, gold["disp"])

    gold = torch.load(golden_dir + "/ssg_tiny.pt")
    net = PointNet2SemSegSSG(C.ssg_config()).eval()
    assert {k: list(v.shape) for k, v in net.state_dict().items()} == gold["keys"]
    C.fill_parameters_(net, seed=5)
    pc = torch.randn(2, 256, 3, generator=torch.Generator().manual_seed(6))
    with C.package_bound_to_oracle(), torch.no_grad():
        y = net(pc, ts=torch.tensor([10.0, 500.0]), label=torch.tensor([1, 7]))
    assert torch.equal(y, gold["out"])          # exercises three_nn + three_interpolate (PointnetFPModule)


def test_schedules_match_reference(golden_dir):
    from point_diffusion_refinement_b200 import configs, util, util_fastdpmv2 as fast
    gold = torch.load(golden_dir + "/schedules.pt")
    dh = util.calc_diffusion_hyperparams(**configs.DIFFUSION_CONFIG)
    for k in ("Beta", "Alpha", "Alpha_bar", "Sigma"):
        assert torch.equal(dh[k], gold["ddpm"][k]), k
    for key, g in gold["fast"].items():
        length, schedule = key.split("_")
        eta = fast.get_VAR_noise(int(length), configs.DIFFUSION_CONFIG, schedule)
        np.testing.assert_allclose(eta, g["eta"].numpy(), rtol=1e-12)
        taus = fast._precompute_VAR_steps(dh, eta)
        # the reference evaluated with a float64 Beta (see util_fastdpmv2._precompute_VAR_steps for why)
        np.testing.assert_allclose(taus, g["taus_f64"].numpy(), rtol=0, atol=1e-4)
        # ... and within the fp32 noise of the reference exactly as it runs under this container's numpy
        np.testing.assert_allclose(taus, g["taus_as_run_here"].numpy(), atol=1.0)
        assert fast.get_STEP_step(int(length), configs.DIFFUSION_CONFIG, schedule) == g["steps"]
        assert all(a > b for a, b in zip(taus[:-1], taus[1:])) and abs(taus[-1]) < 0.1


def test_fastdpm_loops_match_reference_python(golden_dir, monkeypatch):
    """a14: VAR_sampling / STEP_sampling (reference util_fastdpmv2.py:307-452) step by step against the fixture the
    reference's own loops produced with the same stub eps_theta and noise bank.  Host logic only: the fused device
    update x*a + c*eps + sigma*z is replaced by the same expression in torch (the GPU twin of this test,
    tests/test_model_gpu.py, runs the real kernel).  Tolerance: 2e-6 relative + 1e-6 absolute per step (fp32 ulp level: the reference
    updates x in two rounded steps, x *= a; x += c*eps + sigma*z)."""
    from point_diffusion_refinement_b200 import util
    gold = torch.load(golden_dir + "/fastdpm_loops.pt")

    def affine_update(self, x, eps, scale_x, scale_eps, sigma, noise=None):
        x.mul_(scale_x).add_(eps, alpha=scale_eps)
        if sigma != 0.0:
            x.add_(noise.to(x.dtype), alpha=sigma)
        return x
    monkeypatch.setattr(util.DeviceNoise, "affine_update", affine_update)
    size = tuple(gold["size"])
    assert len(gold["cases"]) == 6
    for case in gold["cases"]:
        seen = []
        x0 = C.run_fastdpm_case(case, size, torch.device("cpu"), seen)
        assert len(seen) == case["length"] == case["x_in"].shape[0]
        torch.testing.assert_close(torch.stack(seen), case["x_in"], rtol=2e-6, atol=1e-6)
        torch.testing.assert_close(x0, case["x0"], rtol=2e-6, atol=1e-6)


def test_ddim_coefficients_reduce_to_ddpm_limits():
    from point_diffusion_refinement_b200.util_fastdpmv2 import _ddim_coefficients
    sx, c, s = _ddim_coefficients(0.5, None, 0.7, last=True)       # last step: x0 = (x - sqrt(1-a) eps)/sqrt(a)
    assert s == 0.0 and abs(sx - 2 ** 0.5) < 1e-6 and abs(c + (0.5 ** 0.5) * 2 ** 0.5) < 1e-6
    sx, c, s = _ddim_coefficients(0.5, 0.8, 0.0, last=False)       # kappa = 0: deterministic DDIM
    assert s == 0.0 and abs(c - ((0.2 ** 0.5) - (0.5 ** 0.5) * (1.6 ** 0.5))) < 1e-6


def test_json_reader_and_configs_roundtrip():
    from point_diffusion_refinement_b200 import configs, json_reader
    cfg = configs.ddpm_pointnet_config()
    s = json_reader.replace_list_with_string_in_a_dict(json.loads(json.dumps(cfg)))
    assert isinstance(s["architecture"]["npoint"], str)
    assert json_reader.restore_string_to_list_in_a_dict(s) == cfg


def test_query_and_group_fill_rule(oracle):
    """subset=False: a centre with no neighbour becomes its own zero-feature neighbour (pointnet2_utils.py:376-410)."""
    from point_diffusion_refinement_b200.pointnet2_utils import QueryAndGroup
    xyz = torch.tensor([[[0.0, 0, 0], [0.05, 0, 0], [5.0, 5, 5]]])
    centres = torch.tensor([[[0.0, 0, 0], [9.0, 9, 9]]])
    feats = torch.arange(6, dtype=torch.float32).reshape(1, 2, 3) + 1
    q = QueryAndGroup(0.1, 4, use_xyz=True, include_abs_coordinate=True, include_center_coordinate=True)
    with C.package_bound_to_oracle():
        out, cnt = q(xyz, centres, feats, subset=False, return_counts=True)
    assert out.shape == (1, 2 + 9, 2, 4) and cnt.tolist() == [[2, 0]]
    assert out[0, :2, 1].abs().sum() == 0                      # features zeroed
    assert out[0, 2:5, 1].abs().sum() == 0                     # relative xyz = 0
    assert torch.equal(out[0, 5:8, 1, 0], centres[0, 1])       # abs xyz = the centre itself
    assert out[0, 0, 0].tolist() == [1.0, 2.0, 1.0, 1.0]       # hits 0,1 then padded with the first hit


def test_result_formats_round_trip(tmp_path):
    import os
    """SURVEY 8f row 3: file names of completion_eval.py:283-315 and the pickled dict of
    generate_samples.py:247-252 / generate_samples_distributed.py:84-93."""
    import numpy as np
    from point_diffusion_refinement_b200 import results_io as R
    assert R.generated_file_name("mvp_dataset", 2048) == "mvp_generated_data_2048pts.h5"
    assert R.generated_file_name("shapenet_chunk", 16384, 100) == "shapenet_generated_data_16384pts_T100.h5"
    data = np.random.RandomState(0).rand(5, 64, 3).astype(np.float32)
    path = str(tmp_path / R.generated_file_name("mvp40", 64))
    assert R.save_generated(path, data) == path and os.path.exists(path)
    assert np.array_equal(R.load_generated(path), data)
    # the container itself: HDF5 signature, version-0 superblock with 8-byte offsets, end-of-file address = file size,
    # dataset 'data' found through the root group's B-tree / heap / symbol node, raw data 8-byte aligned at the tail
    raw = open(path, "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and raw[8] == 0 and raw[13] == 8 and raw[14] == 8
    assert int.from_bytes(raw[40:48], "little") == len(raw) and (len(raw) - data.nbytes) % 8 == 0
    assert raw[len(raw) - data.nbytes:] == data.tobytes()
    for shape in ((1, 2048, 3), (3, 16384, 3), (2, 5)):
        arr = np.random.RandomState(1).randn(*shape).astype(np.float32)
        p2 = str(tmp_path / ("t%d.h5" % len(shape)))
        R.write_hdf5_dataset(p2, arr)
        back = R.read_hdf5_dataset(p2)
        assert back.dtype == np.float32 and np.array_equal(back, arr)
    with __import__("pytest").raises(KeyError):
        R.read_hdf5_dataset(path, "other")
    parts = [R.eval_result_dict(np.arange(3) + 3 * r, np.full(3, 0.1 * (r + 1), np.float32), np.full(3, 0.2, np.float32),
                                np.full(3, 0.5, np.float32), 545999) for r in range(2)]
    assert set(parts[0]) == {"meta", "cd_distance", "emd_distance", "f1", "avg_cd", "avg_emd", "iter"}
    files = [R.save_eval_result(str(tmp_path / ("eval_result_rank_%d.pkl" % r)), p) for r, p in enumerate(parts)]
    merged = R.gather_eval_results([R.load_eval_result(f) for f in files])
    assert merged["meta"].tolist() == list(range(6)) and merged["iter"] == 545999
    assert abs(merged["avg_cd"] - 0.15) < 1e-6 and merged["cd_distance"].shape == (6,)


def test_compiled_programs_construct_without_a_gpu(cuda_lib):
    """Host logic of fused.py: both static programs (x-branch, condition branch) are laid out from the module tree
    alone -- every channel-width / alignment assertion of the builder runs here; nothing is launched."""
    import collections
    from point_diffusion_refinement_b200 import configs
    from point_diffusion_refinement_b200.fused import FusedDenoiser
    from point_diffusion_refinement_b200.pointnet2_with_pcld_condition import PointNet2CloudCondition
    for cfg, n_gemm in ((configs.tiny_pointnet_config(), 134), (dict(configs.tiny_pointnet_config(), include_t=False), 133)):
        net = PointNet2CloudCondition(cfg).eval()
        eng = FusedDenoiser(net, 2, 256, use_tf32=False, use_graph=False)
        eng.build(384)
        main = collections.Counter(n for n, _ in eng.meta)
        cond = collections.Counter(n for n, _ in eng.cond_meta)
        assert main["pdr_gemm_fused"] == n_gemm and main["pdr_furthest_point_sampling"] == 4
        assert main["pdr_ball_query"] == 9 and main["pdr_attention_pool"] == 17        # 13 queries in the reference
        assert cond["pdr_gemm_fused"] == 68 and cond["pdr_attention_pool"] == 8 and cond["pdr_group_knn"] == 4
        assert eng.condition_shapes(384) == ([384, 128, 64, 32, 16], [4, 32, 64, 64, 128], [32, 32, 64, 64, 128])
        assert [v.C for v in eng._enc_cl] == [4, 32, 64, 64, 128] and eng.eps_out.shape == (2, 256, 3)


def test_geometry_overlap_program_layout(cuda_lib, monkeypatch):
    """PDR_GEOM_OVERLAP (default on): same multiset of calls; FPS chain, centre gathers, 8 ball queries and the kNN calls are
    tagged for the side stream; the level-0 mapper query stays first on the main stream; the main stream joins right
    before the first set-abstraction block; only the GEMMs of the first mapper block carry the CTA cap."""
    import collections
    from point_diffusion_refinement_b200 import configs, fused
    from point_diffusion_refinement_b200.pointnet2_with_pcld_condition import PointNet2CloudCondition
    layouts = {}
    monkeypatch.setattr(fused, "_STAGE_CHAIN", False)         # layout of the per-layer engine
    for on in (False, True):
        monkeypatch.setattr(fused, "_GEOM_OVERLAP", on)
        eng = fused.FusedDenoiser(PointNet2CloudCondition(configs.tiny_pointnet_config()).eval(), 2, 256, use_tf32=True,
                                  use_graph=False)
        eng.build(384)
        layouts[on] = eng
    off, on = layouts[False], layouts[True]
    assert collections.Counter(n for n, _ in off.meta) == collections.Counter(n for n, _ in on.meta)
    assert not off.side_ops and off._join_at is None
    names = [n for n, _ in on.meta]
    side = collections.Counter(names[k] for k in on.side_ops)
    assert side == {"pdr_furthest_point_sampling": 4, "pdr_gather_rows": 4, "pdr_ball_query": 8, "pdr_knn_points": 4}
    first_side = min(on.side_ops)
    assert names[first_side - 1] == "pdr_ball_query" and max(on.side_ops) < on._join_at
    assert names[on._join_at] == "pdr_group_geo_ball" and names[on._join_at - 1] in ("pdr_attention_pool", "pdr_stage_chain")
    caps = [g.max_ctas for g in on.keep if isinstance(g, fused.GemmArgs)]
    assert sum(1 for c in caps if c) == 7 and set(caps) == {0, fused._GEOM_OVERLAP_CTAS}
    assert all(g.max_ctas == 0 for g in off.keep if isinstance(g, fused.GemmArgs))


def test_gemm_argument_flags_of_the_compiled_program(cuda_lib, monkeypatch):
    """Launch hints the engine sets on every PdrGemmArgs: weights are static (the kernel may stage them before its
    programmatic-dependency wait); the TMA gather of the table chunks is an opt-in (PDR_GEMM_TMA_GATHER); the statistics pairs are
    requested per 32-column block -- a merged first GEMM computes the plain pair under the columns a GroupNorm->ReLU reads and the
    relu pair under the ReLU->GroupNorm ones, never both under the same block unless two consumers overlap there."""
    from point_diffusion_refinement_b200 import configs, fused
    from point_diffusion_refinement_b200.pointnet2_with_pcld_condition import PointNet2CloudCondition
    monkeypatch.setattr(fused, "_STAGE_CHAIN", False)
    engines = {}
    for tma in (False, True):
        monkeypatch.setattr(fused, "_TMA_GATHER", tma)
        eng = fused.FusedDenoiser(PointNet2CloudCondition(configs.tiny_pointnet_config()).eval(), 2, 256, use_tf32=True,
                                  use_graph=False)
        eng.build(384)
        engines[tma] = [g for g in eng.keep if isinstance(g, fused.GemmArgs)]
    off, on = engines[False], engines[True]
    assert len(off) == len(on) and all(g.w_static == 1 for g in off)
    assert all(g.table_rows == 0 for g in off)
    assert all((g.table_rows > 0) == bool(g.a_rows) for g in on)
    with_stats = [g for g in off if g.stats]
    assert with_stats and all(g.stats_skip == 0 for g in with_stats)
    mixed = 0
    for g in with_stats:
        bits = [(g.stats_skip_blocks >> (2 * b)) & 3 for b in range((g.N + 31) // 32)]
        assert any(b != 3 for b in bits), (g.N, bits)                  # a statistics GEMM has at least one consumer
        mixed += len(set(bits)) > 1
    assert mixed >= 4                                                  # the merged first GEMMs of the grouped stages


def test_builtin_hdf5_writer_is_readable_by_h5py(tmp_path):
    """the built-in HDF5 writer against the real library -- runs wherever h5py is installed (not in this image)."""
    h5py = pytest.importorskip("h5py")
    from point_diffusion_refinement_b200 import results_io
    data = np.random.default_rng(0).standard_normal((5, 64, 3)).astype(np.float32)
    path = str(tmp_path / "mvp_generated_data_64pts.h5")
    results_io.write_hdf5_dataset(path, data, "data")
    with h5py.File(path, "r") as hf:
        assert list(hf.keys()) == ["data"] and hf["data"].dtype == np.float32
        assert np.array_equal(np.array(hf["data"]), data)
