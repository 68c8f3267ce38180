"""GPU: the fused stage kernel (csrc/stage_chain.cu, pdr_stage_chain).

1. Every sweep of synthetic stages (2- and 3-layer MLPs, embeddings, 32- and 8-neighbour groups, masked counts, rows
   gathered with missing neighbours) against the numpy emulator of chain.py -- which tests/test_chain_host.py holds
   against a direct float64 evaluation of the stage on the CPU.  Tolerance: TF32 operands (10-bit mantissa), fp32
   accumulation: statistics 2e-3 relative to their scale, pooled rows 5e-3 absolute on O(1) values, mean error 10x lower.
2. The compiled denoiser with the stages on the fused kernel against the same program on the per-layer GEMMs."""
import ctypes

import numpy as np
import pytest
import torch

from point_diffusion_refinement_b200 import chain as CH
from tests import common as C
from tests.test_chain_host import _gn_affine, _make_stage

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _run_sweep(args):
    from point_diffusion_refinement_b200._lib import call
    call("pdr_stage_chain", ctypes.c_void_p(ctypes.addressof(args)),
         ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()


@pytest.mark.parametrize("case", [
    # max_ctas caps the persistent grid so that a CTA walks many tiles: both tile groups, odd tile counts, ring wrap-around
    dict(B=2, P=64, K=32, C=4, widths=[32, 32], ck=32, ci=32, co=32, emb=False, max_ctas=2),
    dict(B=3, P=40, K=32, C=35, widths=[32, 32], ck=41, ci=32, co=32, emb=False, max_ctas=7),
    dict(B=2, P=256, K=8, C=35, widths=[32, 32, 64], ck=44, ci=64, co=64, emb=True, max_ctas=3),
    dict(B=5, P=36, K=32, C=4, widths=[32, 32, 32], ck=32, ci=32, co=32, emb=True, max_ctas=0),
])
def test_every_sweep_against_the_emulator(case):
    B, P, K, Cc = case["B"], case["P"], case["K"], case["C"]
    spec, host, gn = _make_stage(11, B, P, K, Cc, case["widths"], case["ck"], case["ci"], case["co"], case["emb"])
    g = np.random.default_rng(5)
    M = B * P * K
    Cp = CH.r4(Cc)
    n_table = 1000
    table = np.zeros((n_table, Cp + 4), dtype=np.float32)                 # ld > Cp: rows are slices of a wider table
    table[:, :Cc] = g.standard_normal((n_table, Cc))
    src = g.integers(0, n_table, M).astype(np.int32)
    src[g.random(M) < 0.05] = -1                                           # the subset=False fill rule: zero rows
    geo = np.zeros((M, 12), dtype=np.float32)
    geo[:, :9] = g.standard_normal((M, 9))
    t32 = lambda a: CH.tf32_round(torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))).numpy()
    table, geo = t32(table), t32(geo)                                      # exactly representable: the first GEMM is exact
    X0 = np.zeros((M, spec.k0))
    X0[:, :Cp] = np.where(src[:, None] >= 0, table[np.maximum(src, 0), :Cp], 0)
    X0[:, Cp:] = geo
    host["X0"] = X0
    dev = torch.device(DEV)
    up = lambda a, dt=torch.float32: torch.from_numpy(np.ascontiguousarray(a)).to(dt).to(dev)
    d_table, d_src, d_geo = up(table), up(src, torch.int32), up(geo)
    d_rowadd, d_counts = up(host["rowadd"]), up(host["counts"], torch.int32)
    d_out = torch.full((B * P, CH.r4(case["co"]) + 4), -7.0, device=dev)
    plan = CH.StagePlan(spec, dev)
    assert plan.fits()
    L, rps = spec.L, P * K
    tiles_ps = rps // CH.TILE_ROWS
    d_emb = {l: (up(host["emb"][l]) if host["emb"][l] is not None else None) for l in host["emb"]}
    rt = dict(table=(d_table.data_ptr(), d_table.shape[1]), src_rows=d_src.data_ptr(), geo=(d_geo.data_ptr(), 12), batch=B,
              rows_per_sample=rps, group_k=K, gn={}, emb={l: ((t.data_ptr(), t.shape[1]) if t is not None else None)
                                                          for l, t in d_emb.items()},
              rowadd=(d_rowadd.data_ptr(), d_rowadd.shape[1]), counts=d_counts.data_ptr(), out=(d_out.data_ptr(), d_out.shape[1]),
              max_ctas=case["max_ctas"])
    host["gn"] = {}
    keep = []
    for d in range(1, L + 3):
        n = plan.sweep_stats_n(d)
        d_stats = torch.full((B * tiles_ps, max(n, 1), 4), -3.0, device=dev)
        rt["stats"] = d_stats.data_ptr()
        args, steps = plan.build_sweep(d, rt)
        want = CH.emulate_sweep(plan, d, steps, host)
        _run_sweep(args)
        if d <= L + 1:
            got = d_stats.cpu().double().numpy()
            scale = np.abs(want).max(axis=(0, 1), keepdims=True) + 1e-6
            err = np.abs(got - want) / scale
            assert err.max() < 2e-3 and err.mean() < 2e-4, (d, err.max(), err.mean())
            for name, c0, nc, relu in plan.sweep_stat_columns(d):
                st = want.reshape(B, tiles_ps, -1, 4).sum(1)[:, c0:c0 + nc]     # the emulator's statistics drive both sides
                gam, bet, groups = gn[name]
                sc, sh = _gn_affine(st[:, :, 2 if relu else 0], st[:, :, 3 if relu else 1], rps, gam, bet, groups)
                host["gn"][name] = (sc, sh)
                ld = CH.p32(nc) + 4              # read in whole 32-column blocks: zero padded
                d_sc, d_sh = torch.zeros(B, ld, device=dev), torch.zeros(B, ld, device=dev)
                d_sc[:, :nc], d_sh[:, :nc] = up(sc), up(sh)
                keep += [d_sc, d_sh]
                rt["gn"][name] = (d_sc.data_ptr(), d_sh.data_ptr(), ld)
        else:
            got = d_out.cpu().double().numpy()
            assert np.all(got[:, case["co"]:] == -7.0)                      # nothing outside the stage's columns
            err = np.abs(got[:, :case["co"]] - want)
            assert err.max() < 5e-3 and err.mean() < 5e-4, (err.max(), err.mean())


def test_denoiser_on_fused_stages_matches_the_per_layer_engine(monkeypatch):
    from point_diffusion_refinement_b200 import configs, fused
    from point_diffusion_refinement_b200.pointnet2_with_pcld_condition import PointNet2CloudCondition
    x, cond, ts, label = [t.to(DEV) for t in C.denoiser_inputs(2, 2048, 3072, seed=3)]
    res = {}
    x2 = None
    for on in (False, True):
        monkeypatch.setattr(fused, "_STAGE_CHAIN", on)
        net = C.fill_parameters_(PointNet2CloudCondition(configs.ddpm_pointnet_config()).eval(), seed=1).to(DEV)
        net.enable_fused(True, use_tf32=True, use_graph=True, fuse_cold=True)
        with torch.no_grad():
            cold = net(x, cond, ts=ts, label=label, use_retained_condition_feature=True)
            if x2 is None:
                x2 = x + 0.05 * cold         # the same warm input for both engines (neighbour lists are functions of it)
            warm = net(x2, cond, ts=ts - 1, label=label, use_retained_condition_feature=True)
            again = net(x2, cond, ts=ts - 1, label=label, use_retained_condition_feature=True)
        assert torch.equal(warm, again)
        eng = net._fused_engine
        names = [n for n, _ in eng.meta]
        assert ("pdr_stage_chain" in names) == on
        res[on] = (cold, warm, eng._Fl[0].t.clone(), eng._Fl[1].t.clone(), eng._Gl[0].t.clone())
    for i, what in enumerate(("eps cold", "eps warm", "level-0 features (enc_map0)", "level-1 features (sa0)",
                              "decoder level 0 (dec_map0)")):
        a, b = res[False][i], res[True][i]
        err = (a - b).abs()
        print("%s: fused stages vs per-layer |diff| mean %.2e max %.2e (|ref| mean %.2e)" % (what, err.mean(), err.max(), a.abs().mean()))
    # The first fused stage sees bit-identical inputs: its output differs from the per-layer GEMMs by fp32 association only
    # (statistics summed in another order, the conv bias folded into the GroupNorm shift) -- except where that last-bit
    # difference flips the TF32 rounding of an activation (one TF32 ulp = 5e-4 relative on that element): mean 2e-7, max 6e-4
    # measured.  Further down the two TF32 engines drift apart like either drifts from fp32: judged with the TF32
    # distribution bars of test_model_gpu.py.
    from tests.test_model_gpu import TF32_MAX, TF32_MEDIAN, TF32_P999, _tf32_error_profile
    first = (res[False][2] - res[True][2]).abs()
    assert first.mean() < 2e-6 and first.max() < 3e-3, (first.mean(), first.max())
    for i in (0, 1):
        med, p999, mx = _tf32_error_profile(res[True][i], res[False][i])
        assert med <= TF32_MEDIAN and p999 <= TF32_P999 and mx <= TF32_MAX, (i, med, p999, mx)
