"""GPU: the LITERAL drop-in.  The reference's own ``pointnet2/completion_eval.py`` (and, through it, its own
``pointnet2_ops`` modules, ``util.sampling``, ``Chamfer_F1``, ``EMD_distance``) is imported UNMODIFIED from the staging
directory ``oracle/_ref/pyref`` (git-ignored test infrastructure written by oracle/build_ref.sh in the build container; the
GPU box has no /root/reference) and executed on libpdr_b200.so through ``dropin.install()``.

  * task='refine_completion' is deterministic: the reference's evaluate() and the package's evaluate() must return the same
    CD / EMD / F1 for the same weights and inputs (same kernels underneath, reference module tree vs ours);
  * task='completion' runs the reference's sampling loop (its CPU-generator noise) for a short schedule end to end.
Libraries the reference's data / plotting modules import but the hot path never touches (matplotlib, h5py, transforms3d,
open3d ...) are stubbed with empty modules."""
import contextlib
import os
import sys
import types

import pytest
import torch

from tests import common as C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PYREF = os.path.join(ROOT, "oracle", "_ref", "pyref")


@contextlib.contextmanager
def reference_drivers():
    """sys.path / sys.modules set up so that `import completion_eval` resolves to the reference's file."""
    from point_diffusion_refinement_b200 import dropin
    saved_path, saved_mods = list(sys.path), set(sys.modules)
    stubs = {}
    for name in ("matplotlib", "matplotlib.pyplot", "h5py", "transforms3d", "open3d", "tensorboardX", "termcolor"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.use = lambda *a, **k: None
            stubs[name] = m
    if "matplotlib" in stubs:
        stubs["matplotlib"].pyplot = stubs["matplotlib.pyplot"]
    sys.modules.update(stubs)
    dropin.install()
    sys.path[:0] = [os.path.join(PYREF, "pointnet2"), os.path.join(PYREF, "pointnet2_ops_lib"), PYREF]
    try:
        with dropin.reference_torch_version("1.7.1"):
            import completion_eval as ref_eval
            yield ref_eval
    finally:
        sys.path[:] = saved_path
        for name in set(sys.modules) - saved_mods:         # only what came from the staging directory or is a stub
            mod = sys.modules[name]
            if name in stubs or (getattr(mod, "__file__", None) or "").startswith(PYREF):
                del sys.modules[name]
        dropin.uninstall()


def _loader(n_batches, B, N, M, with_generated):
    g = torch.Generator().manual_seed(21)
    out = []
    for _ in range(n_batches):
        partial = torch.cat([torch.rand(B, M, 3, generator=g) * 2 - 1, torch.ones(B, M, 1)], dim=2)
        complete = torch.rand(B, N, 3, generator=g) * 2 - 1
        d = {"label": torch.randint(0, 16, (B,), generator=g), "partial": partial, "complete": complete}
        if with_generated:
            d["generated"] = complete + 0.05 * torch.randn(B, N, 3, generator=g)
        out.append(d)
    return out


def test_reference_staging_is_importable_without_a_gpu_call():
    """(also runs on the GPU box only: it needs the staged reference files)"""
    if not os.path.isdir(PYREF):
        pytest.skip("oracle/_ref/pyref not staged (run oracle/build_ref.sh where /root/reference exists)")
    with reference_drivers() as ref_eval:
        assert ref_eval.__file__.startswith(PYREF) and callable(ref_eval.evaluate)
        import pointnet2_ops.pointnet2_utils as ref_utils
        assert ref_utils.__file__.startswith(PYREF)
        from point_diffusion_refinement_b200 import _ext
        assert ref_utils._ext is _ext                                  # the reference module is bound to our kernels


@pytest.mark.gpu
def test_reference_completion_eval_runs_on_our_kernels():
    if not os.path.isdir(PYREF):
        pytest.skip("oracle/_ref/pyref not staged (run oracle/build_ref.sh where /root/reference exists)")
    from point_diffusion_refinement_b200 import completion_eval as our_eval, configs, util
    from point_diffusion_refinement_b200.pointnet2_with_pcld_condition import PointNet2CloudCondition as OurNet
    cfg = configs.tiny_pointnet_config()
    cfg.update(include_t=False, point_upsample_factor=1)
    loader = _loader(2, 2, 256, 384, with_generated=True)
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with reference_drivers() as ref_eval:
            from models.pointnet2_with_pcld_condition import PointNet2CloudCondition as RefNet
            assert RefNet.__module__ != OurNet.__module__
            ref_net = C.fill_parameters_(RefNet(dict(cfg)).eval(), seed=2).cuda()
            with torch.no_grad():
                r_cd, r_emd, r_meta, r_cds, r_emds = ref_eval.evaluate(
                    ref_net, loader, None, parallel=False, dataset="mvp_dataset", scale=1, task="refine_completion",
                    refine_output_scale_factor=0.001, compute_emd=True, compute_cd=True)
            # the reference's own sampling loop + metrics, short schedule, end to end
            cfg_t = configs.tiny_pointnet_config()
            ddpm = C.fill_parameters_(RefNet(dict(cfg_t)).eval(), seed=1).cuda()
            import util as ref_util
            dh = ref_util.calc_diffusion_hyperparams(T=6, beta_0=1e-4, beta_T=0.02)
            dh = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in dh.items()}
            torch.manual_seed(0)
            with torch.no_grad():
                s_cd, s_emd, s_meta, _, _ = ref_eval.evaluate(ddpm, _loader(1, 2, 256, 384, False), dh, parallel=False,
                                                              dataset="mvp_dataset", scale=1, task="completion",
                                                              print_every_n_steps=100)
        our_net = C.fill_parameters_(OurNet(dict(cfg)).eval(), seed=2).cuda()
        with torch.no_grad():
            o_cd, o_emd, o_meta, o_cds, o_emds = our_eval.evaluate(
                our_net, loader, None, parallel=False, dataset="mvp_dataset", scale=1, task="refine_completion",
                refine_output_scale_factor=0.001, compute_emd=True, compute_cd=True, use_fused=False)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    assert list(r_meta) == list(o_meta) and len(r_meta) == 4
    torch.testing.assert_close(r_cds, o_cds, rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(r_emds, o_emds, rtol=2e-3, atol=1e-6)
    assert abs(r_cd - o_cd) <= 1e-4 * abs(o_cd) + 1e-7 and abs(r_emd - o_emd) <= 2e-3 * abs(o_emd) + 1e-6
    assert s_cd == s_cd and s_emd == s_emd and len(s_meta) == 2          # finite: the reference chain ran to x_0
