"""Shared helpers for the test-suite and tests/golden/make_golden.py (test infrastructure only)."""
import contextlib
import math
import zlib

import torch


def fill_parameters_(module, seed=0):
    """Deterministic, NAME-keyed parameter values: independent of construction order, so the reference
    model (at fixture time) and our model (at test time) get identical weights without committing 39 MB."""
    sd = module.state_dict()
    for name in sorted(sd.keys()):
        t = sd[name]
        if not t.dtype.is_floating_point:
            continue
        g = torch.Generator().manual_seed((zlib.crc32(name.encode()) + 7919 * seed) & 0x7FFFFFFF)
        if name.endswith("group_norm.weight") or (name.endswith(".weight") and t.dim() == 1):
            v = 1.0 + 0.1 * torch.randn(t.shape, generator=g)
        elif name.endswith(".bias"):
            v = 0.1 * torch.randn(t.shape, generator=g)
        elif t.dim() >= 2:
            fan_in = t[0].numel()
            v = torch.randn(t.shape, generator=g) / math.sqrt(max(fan_in, 1))
        else:
            v = 0.1 * torch.randn(t.shape, generator=g)
        t.copy_(v)
    return module


def denoiser_inputs(B, N, M, seed=0):
    """x_t (B,N,3) ~ N(0,1); cond (B,M,4): xyz ~ U[-1,1]^3 + mirror flag (+1 first half, -1 second half);
    ts, label.  """
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, N, 3, generator=g)
    uvw = torch.rand(B, M, 3, generator=g) * 2 - 1
    flag = torch.ones(B, M, 1)
    flag[:, M // 2:] = -1
    cond = torch.cat([uvw, flag], dim=2)
    ts = torch.full((B,), 500.0)
    ts[0] = 37.0
    label = torch.arange(B) % 16
    return x, cond, ts, label


def stub_eps(x, ts=None, label=None, **_):
    """The parameter-free eps_theta stand-in of tests/golden/make_golden.py::stub_net (same three lines)."""
    t = ts.to(x.dtype).view(-1, 1, 1) / 1000.0
    return 0.9 * x / torch.sqrt(1.001 - torch.exp(-(0.1 * t + 10.0 * t * t))) + 0.1 * torch.tanh(x + t)


def run_fastdpm_case(case, size, device, record):
    """Replay one case of tests/golden/fastdpm_loops.pt through the package's VAR_sampling / STEP_sampling with the
    fixture's noise bank; `record` receives every x the loop hands to eps_theta."""
    from point_diffusion_refinement_b200 import configs, util, util_fastdpmv2 as fast
    dh = util.calc_diffusion_hyperparams(**configs.DIFFUSION_CONFIG)
    bank = case["bank"]

    def net(x, ts=None, label=None):
        record.append(x.detach().cpu().clone())
        return stub_eps(x, ts=ts)
    kw = dict(label=None, verbose=False, condition=None, noise=lambda i, s: bank[i + 1], device=device)
    if case["method"] == "var":
        eta = fast.get_VAR_noise(case["length"], configs.DIFFUSION_CONFIG, case["schedule"])
        taus = fast._precompute_VAR_steps(dh, eta)
        return fast.VAR_sampling(net, size, dh, eta, case["kappa"], taus, **kw)
    steps = fast.get_STEP_step(case["length"], configs.DIFFUSION_CONFIG, case["schedule"])
    return fast.STEP_sampling(net, size, dh, steps, case["kappa"], **kw)


def ssg_config():
    """Small unconditional PointNet2SemSegSSG with the three_nn / three_interpolate decoder
    (reference default, pointnet2_ssg_sem.py:213)."""
    return {
        "in_fea_dim": 0, "out_dim": 3, "include_t": True, "t_dim": 32, "model.use_xyz": True,
        "attach_position_to_input_feature": True, "include_abs_coordinate": True, "record_neighbor_stats": False,
        "bn_first": False, "bias": True, "res_connect": True, "include_class_condition": True, "num_class": 16,
        "class_condition_dim": 32, "scale_factor": 1,
        # one level only: the reference's base class indexes its default radius=[0] per level
        # (pointnet2_ssg_sem.py:172), so deeper unconditional pyramids cannot be built there either
        "architecture": {"npoint": [64], "radius": [0.5], "nsample": [16],
                         "feature_dim": [32, 64], "mlp_depth": 3, "decoder_feature_dim": [32, 64],
                         "decoder_mlp_depth": 2},
    }


@contextlib.contextmanager
def package_bound_to_oracle():
    """TESTS ONLY: run the package's host-side modules on CPU by binding its native entry points to the
    CPU oracle.  The product never does this -- without libpdr_b200.so + a GPU its ops raise."""
    from oracle import cpu_oracle as O
    from point_diffusion_refinement_b200 import _ext, knn
    saved = {}
    patches = {
        _ext: dict(furthest_point_sampling=O.furthest_point_sampling, gather_points=O.gather_points,
                   ball_query=O.ball_query, group_points=O.group_points,
                   three_nn=lambda u, k: list(O.three_nn(u, k)), three_interpolate=O.three_interpolate),
        knn: dict(knn_points=O.knn_points, knn_gather=O.knn_gather),
    }
    for mod, table in patches.items():
        for name, fn in table.items():
            saved[(mod, name)] = getattr(mod, name)
            setattr(mod, name, fn)
    try:
        yield
    finally:
        for (mod, name), fn in saved.items():
            setattr(mod, name, fn)


@contextlib.contextmanager
def package_bound_to_reference_cuda():
    """TESTS / bench.py's `gpu_reference` baseline leg ONLY: the package's per-layer module path with its native entry
    points bound to the REFERENCE's own CUDA kernels recompiled for sm_100a (oracle/_ref/libpdr_ref_cuda.so) -- FPS, gather,
    ball query, grouping, three_nn / three_interpolate -- i.e. the reference's design (pointnet2_ops kernels + cuDNN 1x1
    convolutions + ATen GroupNorm, ~1000 launches per step) on this box.  pytorch3d's kNN is not vendored by the reference,
    so kNN stays on the package's kernel."""
    from oracle import ref_cuda as R
    from point_diffusion_refinement_b200 import _ext
    table = dict(furthest_point_sampling=R.furthest_point_sampling, gather_points=R.gather_points,
                 ball_query=R.ball_query, group_points=R.group_points,
                 three_nn=lambda u, k: list(R.three_nn(u, k)), three_interpolate=R.three_interpolate)
    saved = {name: getattr(_ext, name) for name in table}
    for name, fn in table.items():
        setattr(_ext, name, fn)
    try:
        yield
    finally:
        for name, fn in saved.items():
            setattr(_ext, name, fn)
