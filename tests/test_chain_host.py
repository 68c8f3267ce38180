"""CPU: host side of the fused stage kernel (point_diffusion_refinement_b200/chain.py) -- the weight image, the TMEM
column plan and the step programs of every sweep, executed by the numpy emulator (which follows csrc/stage_chain.cu
step for step) and held against a direct float64 evaluation of the same grouped stage (Mlp_plus_t_emb over grouped rows
+ AttentionModule pooling; reference pointnet2_modules.py:129-174, attention.py:70-96).  No GPU, no kernel launch."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch

from point_diffusion_refinement_b200 import chain as CH

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gn_affine(stat_sum, stat_sq, rows, gamma, beta, groups, eps=1e-5):
    """Per-sample GroupNorm affine from per-column sums (what pdr_gn_finalize computes)."""
    C = stat_sum.shape[-1]
    cpg = C // groups
    s = stat_sum.reshape(-1, groups, cpg).sum(-1)
    q = stat_sq.reshape(-1, groups, cpg).sum(-1)
    n = rows * cpg
    mean = s / n
    var = np.maximum(q / n - mean * mean, 0)
    rstd = 1 / np.sqrt(var + eps)
    sc = gamma[None] * np.repeat(rstd, cpg, axis=1)
    sh = beta[None] - np.repeat(mean, cpg, axis=1) * sc
    return sc, sh


def _make_stage(seed, B, P, K, C, widths, ck, ci, co, with_emb):
    g = np.random.default_rng(seed)
    Cp = CH.r4(C)
    k0 = Cp + 12
    rnd = lambda *s: g.standard_normal(s) / np.sqrt(s[-1])
    mlp, cin = [], k0
    for c in widths:
        mlp.append((torch.tensor(rnd(c, cin), dtype=torch.float32), torch.tensor(0.1 * g.standard_normal(c), dtype=torch.float32)))
        cin = c
    f32 = lambda a: torch.tensor(a, dtype=torch.float32)
    key = (f32(rnd(ck, k0)), f32(0.1 * g.standard_normal(ck)))
    w1k = (f32(rnd(ci, ck)), f32(0.1 * g.standard_normal(ci)))
    ws = (f32(rnd(co, ci)), f32(0.1 * g.standard_normal(co)))
    wv = (f32(rnd(co, widths[-1])), f32(rnd(co, k0)), f32(0.1 * g.standard_normal(co)))
    spec = CH.StageSpec(k0, Cp, mlp, key, w1k, ws, wv)
    M = B * P * K
    X0 = np.zeros((M, k0))
    X0[:, :C] = g.standard_normal((M, C))
    X0[:, Cp:Cp + 9] = g.standard_normal((M, 9))
    host = dict(X0=X0, batch=B, rows_per_sample=P * K, group_k=K,
                rowadd=g.standard_normal((B * P, ci)), counts=g.integers(0, K + 1, B * P),
                emb={l + 1: (g.standard_normal((B, c)) if with_emb else None) for l, c in enumerate(widths)})
    gn = {}
    for name, c in [("y%d" % (l + 1), c) for l, c in enumerate(widths)] + [("key", ck), ("s1", ci), ("V", co)]:
        gn[name] = (1 + 0.1 * g.standard_normal(c), 0.1 * g.standard_normal(c), 4 if c % 4 == 0 else 1)
    return spec, host, gn


def _direct(spec, host, gn):
    """The stage evaluated layer by layer in float64 with TF32-rounded weights (as the image holds them)."""
    B, rps, K = host["batch"], host["rows_per_sample"], host["group_k"]
    X0 = host["X0"]
    W = lambda t: CH.tf32_round(t.float()).double().numpy()
    bof = lambda r: np.repeat(np.arange(B), r)
    norm = {}

    def affine(name, y, relu_first, rows):
        gam, bet, groups = gn[name]
        t = np.maximum(y, 0) if relu_first else y
        s = np.stack([t[bof(rows) == b].sum(0) for b in range(B)]); q = np.stack([(t[bof(rows) == b] ** 2).sum(0) for b in range(B)])
        sc, sh = _gn_affine(s, q, rows, gam, bet, groups)
        norm[name] = (sc, sh)
        return sc[bof(rows)], sh[bof(rows)]

    a = X0
    for l, (w, bb) in enumerate(spec.mlp):
        y = a @ W(w).T + bb.double().numpy()
        sc, sh = affine("y%d" % (l + 1), y, False, rps)
        a = np.maximum(y * sc + sh, 0)
        e = host["emb"][l + 1]
        if e is not None:
            a = a + e[bof(rps)]
    V = a @ W(spec.wv[0]).T + X0 @ W(spec.wv[1]).T + spec.wv[2].double().numpy()
    sc, sh = affine("V", V, False, rps)
    vv = np.maximum(V * sc + sh, 0)
    key = X0 @ W(spec.key[0]).T + spec.key[1].double().numpy()
    sc, sh = affine("key", key, True, rps)
    k1 = np.maximum(key, 0) * sc + sh
    s1 = k1 @ W(spec.w1k[0]).T + spec.w1k[1].double().numpy() + np.repeat(host["rowadd"], K, axis=0)
    sc, sh = affine("s1", s1, True, rps)
    S = (np.maximum(s1, 0) * sc + sh) @ W(spec.ws[0]).T + spec.ws[1].double().numpy()
    out = np.zeros((X0.shape[0] // K, spec.co))
    for pt in range(out.shape[0]):
        cnt = max(int(host["counts"][pt]), 1)
        sk = np.where(np.arange(K)[:, None] < cnt, S[pt * K:(pt + 1) * K], -1e9)
        ex = np.exp(sk - sk.max(0))
        out[pt] = (ex * vv[pt * K:(pt + 1) * K]).sum(0) / ex.sum(0)
    return out, norm


@pytest.mark.parametrize("case", [
    dict(B=2, P=8, K=32, C=4, widths=[32, 32], ck=32, ci=32, co=32, emb=False),      # encoder mapper, level 0
    dict(B=2, P=8, K=32, C=35, widths=[32, 32], ck=44, ci=32, co=32, emb=False),     # mapper with a 44-wide key
    dict(B=1, P=32, K=8, C=35, widths=[32, 32, 64], ck=44, ci=64, co=64, emb=True),  # SA-style: 3 layers, embeddings
    dict(B=2, P=16, K=16, C=67, widths=[64, 64], ck=76, ci=64, co=64, emb=True, fits=False),   # weights too large for smem
])
def test_sweeps_reproduce_the_stage(case):
    spec, host, gn = _make_stage(7, case["B"], case["P"], case["K"], case["C"], case["widths"], case["ck"], case["ci"],
                                 case["co"], case["emb"])
    want, norm = _direct(spec, host, gn)
    plan = CH.StagePlan(spec, torch.device("cpu"))
    B, rps = host["batch"], host["rows_per_sample"]
    tiles_per_sample = rps // CH.TILE_ROWS
    # runtime bindings: pointers are never dereferenced on the host (dummy non-null values), the emulator reads `host`
    host["gn"] = {}
    rt = dict(table=(1 << 20, CH.r4(case["C"])), src_rows=1 << 21, geo=(1 << 22, 12), batch=B, rows_per_sample=rps,
              group_k=case["K"], stats=1 << 23, gn={}, emb={l: ((1 << 24, 64) if host["emb"][l] is not None else None)
                                                           for l in host["emb"]},
              rowadd=(1 << 25, CH.r4(case["ci"])), counts=1 << 26, out=(1 << 27, CH.r4(case["co"])))
    L = spec.L
    assert plan.n_sweeps() == L + 2 and plan.fits() == case.get("fits", True)
    if not plan.fits():
        return
    for d in range(1, L + 3):
        for name in list(gn):
            if name in host["gn"]:
                rt["gn"][name] = (1 << 28, 1 << 29, 64)
        args, steps = plan.build_sweep(d, rt)
        assert args.n_steps == len(steps) and sum(st["release"] for st in steps) == 1
        res = CH.emulate_sweep(plan, d, steps, host)
        if d <= L + 1:
            # fold the per-tile partials like pdr_gn_finalize and hand the affine to the next sweeps
            for name, c0, nc, relu in plan.sweep_stat_columns(d):
                st = res.reshape(B, tiles_per_sample, -1, 4).sum(1)[:, c0:c0 + nc]
                gam, bet, groups = gn[name]
                sc, sh = _gn_affine(st[:, :, 2 if relu else 0], st[:, :, 3 if relu else 1], rps, gam, bet, groups)
                np.testing.assert_allclose(sc, norm[name][0], rtol=1e-9, atol=1e-12, err_msg="%s sweep %d" % (name, d))
                np.testing.assert_allclose(sh, norm[name][1], rtol=1e-9, atol=1e-10, err_msg="%s sweep %d" % (name, d))
                host["gn"][name] = (sc, sh)
                # the pair the consumer does not read is written as zeros
                assert np.all(res[:, c0:c0 + nc, 0 if relu else 2] == 0)
        else:
            np.testing.assert_allclose(res, want, rtol=1e-9, atol=1e-11)


def test_weight_image_round_trip_and_swizzle():
    g = torch.Generator().manual_seed(0)
    img = CH.WeightImage()
    w = torch.randn(44, 76, generator=g)
    off0, rows0 = img.add("a", torch.randn(32, 16, generator=g))
    off, rows = img.add("b", w)
    assert off0 == 0 and rows0 == 32 and off == 32 * 32 * 4 and rows == 64 and img.bytes % 1024 == 0
    flat = img.tensor("cpu").numpy()
    back = CH.WeightImage.decode(flat, off, rows, 96)
    assert np.array_equal(back[:44, :76], CH.tf32_round(w).numpy()) and not back[44:].any() and not back[:, 76:].any()
    # element (n, k) sits at chunk k // 32, row n, 16-byte piece ((k % 32) // 4) ^ (n & 7)
    n, k = 13, 41
    pos = off // 4 + (k // 32) * rows * 32 + n * 32 + ((((k % 32) // 4) ^ (n & 7)) * 4) + k % 4
    assert flat[pos] == CH.tf32_round(w)[n, k]


def test_tmem_plan_respects_lifetimes():
    p = CH.TmemPlan()
    a = p.alloc("a", 64, 0, 1)
    b = p.alloc("b", 64, 1, 2)          # alive together with a at time 1
    c = p.alloc("c", 64, 2, 3)          # a is dead: its columns are free again
    assert a == 0 and b == 64 and c == 0
    with pytest.raises(MemoryError):
        p.alloc("d", 256, 2, 2)


def test_chain_struct_layouts_match_the_header(tmp_path):
    """PdrChainArgs is passed by address: the ctypes mirrors must have the header's layout field for field."""
    header = os.path.join(ROOT, "include", "pdr_b200.h")
    text = re.sub(r"/\*.*?\*/", "", open(header).read(), flags=re.S)

    def c_fields(struct):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), text, flags=re.S).group(1)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if decl:
                names.append(re.sub(r"\[.*\]", "", decl.split()[-1].lstrip("*")))
        return names

    src = ['#include <stddef.h>', '#include <stdio.h>', '#include "pdr_b200.h"', 'int main(void) {']
    want = {}
    for cname, mirror in (("PdrChainMma", CH.ChainMma), ("PdrChainEpi", CH.ChainEpi), ("PdrChainStep", CH.ChainStep),
                          ("PdrChainArgs", CH.ChainArgs)):
        fields = c_fields(cname)
        assert fields == [f[0] for f in mirror._fields_], (cname, fields, [f[0] for f in mirror._fields_])
        src.append('  printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        want[cname] = ctypes.sizeof(mirror)
        for f in fields:
            src.append('  printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, f, cname, f))
            want["%s.%s" % (cname, f)] = getattr(mirror, f).offset
    src += ["  return 0;", "}"]
    cfile = tmp_path / "layout.c"
    cfile.write_text("\n".join(src))
    exe = str(tmp_path / "layout")
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, str(cfile)])
    got = dict((k, int(v)) for k, v in (l.split() for l in subprocess.check_output([exe], text=True).splitlines()))
    assert got == want, sorted(k for k in want if got.get(k) != want[k])
    assert want["PdrChainArgs"] <= 4000          # passed to the kernel by value as a __grid_constant__ parameter
