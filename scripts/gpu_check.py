"""First-contact GPU check: every kernel vs the CPU oracle and vs the reference CUDA kernels."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import cpu_oracle as O, ref_cuda as R
from point_diffusion_refinement_b200 import _ext, knn, emd_cuda, chamfer_loss_new
from point_diffusion_refinement_b200._lib import call, dptr, stream_ptr

dev = torch.device("cuda")
print(torch.cuda.get_device_name(0), "ref cuda available:", R.available())
g = torch.Generator().manual_seed(0)


def t_ms(fn, n=5):
    fn(); torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n


for (B, N, M, dist) in ((1, 4096, 1024, "U"), (1, 4096, 1024, "G"), (4, 2048, 1024, "G"), (3, 3072, 1024, "U"),
                        (2, 1000, 300, "G"), (2, 256, 64, "G"), (2, 64, 16, "U"), (2, 20, 7, "G"), (1, 8192, 512, "G")):
    x = (torch.rand(B, N, 3, generator=g) * 2 - 1) if dist == "U" else torch.randn(B, N, 3, generator=g)
    if dist == "U": x[:, 5] = 0.001  # a skipped point (|p|^2 <= 1e-3)
    xc = x.to(dev)
    a = _ext.furthest_point_sampling(xc, M).cpu()
    o = O.furthest_point_sampling(x, M)
    r = R.furthest_point_sampling(xc, M).cpu() if R.available() else o
    print("fps", B, N, M, dist, "oracle-eq", torch.equal(a, o), "ref-eq", torch.equal(a, r), "oracle==ref", torch.equal(o, r))
    new = _ext.gather_points(xc.transpose(1, 2).contiguous(), _ext.furthest_point_sampling(xc, M))
    print("  gather eq", torch.equal(new.cpu(), O.gather_points(x.transpose(1, 2).contiguous(), o)))
    newc = new.transpose(1, 2).contiguous()
    for rad, ns in ((0.1, 32), (0.4, 32), (0.8, 16), (3.0, 5)):
        i1, c1 = _ext.ball_query(newc, xc, rad, ns)
        i2, c2 = O.ball_query(newc.cpu(), x, rad, ns)
        ok_r = True
        if R.available():
            i3, c3 = R.ball_query(newc, xc, rad, ns)
            ok_r = torch.equal(i1, i3) and torch.equal(c1, c3)
        print("  ball", rad, ns, "oracle-eq", torch.equal(i1.cpu(), i2) and torch.equal(c1.cpu(), c2), "ref-eq", ok_r, "mean cnt", c2.float().mean().item())
    feats = torch.randn(B, 7, N, generator=g)
    gi = i1
    print("  group eq", torch.equal(_ext.group_points(feats.to(dev), gi).cpu(), O.group_points(feats, gi.cpu())))
    d2a, ia = _ext.three_nn(xc, newc)
    d2o, io = O.three_nn(x, newc.cpu())
    ok_r = True
    if R.available():
        d2r, ir = R.three_nn(xc, newc); ok_r = torch.equal(ia, ir) and torch.equal(d2a, d2r)
    print("  three_nn oracle-eq", torch.equal(ia.cpu(), io), torch.equal(d2a.cpu(), d2o), "ref-eq", ok_r)
    w = torch.rand(B, N, 3, generator=g); w = (w / w.sum(2, keepdim=True)).to(dev)
    f2 = torch.randn(B, 9, M, generator=g).to(dev)
    ta = _ext.three_interpolate(f2, ia, w)
    to = O.three_interpolate(f2.cpu(), io, w.cpu())
    ok_r = torch.equal(ta, R.three_interpolate(f2, ia, w)) if R.available() else True
    print("  three_interp oracle-eq", torch.equal(ta.cpu(), to), "ref-eq", ok_r)
    for K in (1, 3, 8, 20):
        k1 = knn.knn_points(xc, newc, K=K)
        k2 = O.knn_points(x, newc.cpu(), K=K)
        print("  knn", K, torch.equal(k1.idx.cpu(), k2.idx), torch.equal(k1.dists.cpu(), k2.dists))

# chamfer
gold = torch.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "chamfer_f64.pt"))
p1, p2 = gold["p1"].to(dev), gold["p2"].to(dev)
d1 = torch.empty(4, 100, device=dev); d2 = torch.empty(4, 200, device=dev)
i1 = torch.empty(4, 100, dtype=torch.int32, device=dev); i2 = torch.empty(4, 200, dtype=torch.int32, device=dev)
call("pdr_nm_distance", 4, 100, 200, dptr(p1), dptr(p2), dptr(d1), dptr(i1), dptr(d2), dptr(i2), stream_ptr(p1))
print("nm_distance vs f64: mse", ((d1.cpu() - gold["dist1"]) ** 2).mean().item() + ((d2.cpu() - gold["dist2"]) ** 2).mean().item(),
      "idx eq", torch.equal(i1.cpu(), gold["idx1"]) and torch.equal(i2.cpu(), gold["idx2"]))
if R.available():
    r1, r2, ri1, ri2 = R.chamfer3d(p1, p2)
    print("  vs ref chamfer3D bit-exact", torch.equal(d1, r1), torch.equal(d2, r2), torch.equal(i1, ri1), torch.equal(i2, ri2))
for (B, n, m) in ((3, 500, 700), (2, 2048, 2048)):
    a = torch.rand(B, n, 3, generator=g); b = torch.rand(B, m, 3, generator=g) * 0.9
    cp, ct, f1 = chamfer_loss_new.Chamfer_F1(f1_threshold=1e-3)(a.to(dev), b.to(dev))
    ocp, oct_, of1 = O.chamfer_f1(a, b, 1e-3)
    print("chamfer_f1", B, n, m, (cp.cpu() - ocp).abs().max().item(), (ct.cpu() - oct_).abs().max().item(), (f1.cpu() - of1).abs().max().item())

# EMD
a = torch.tensor([[[1.7, -0.1, 0.1], [0.1, 1.2, 0.3]]]).repeat(3, 1, 1); b = torch.tensor([[[0.3, 1.8, 0.2], [1.2, -0.2, 0.3]]]).repeat(3, 1, 1)
from point_diffusion_refinement_b200.emd import EMD_distance
print("emd KAT (expect ~0.355):", EMD_distance()(a.to(dev), b.to(dev)).cpu())
for (B, n, m) in ((2, 256, 256), (3, 512, 300), (2, 300, 1024), (40, 1024, 1024), (300, 256, 256), (2, 2048, 2048)):
    a = torch.rand(B, n, 3, generator=g); b = torch.rand(B, m, 3, generator=g)
    ac, bc = a.to(dev), b.to(dev)
    fused = emd_cuda.emd_cost_forward(ac, bc)
    match = emd_cuda.approxmatch_forward(ac, bc)
    two = emd_cuda.matchcost_forward(ac, bc, match)
    if B * n * m <= 3 * 512 * 512:
        om = O.approxmatch_forward(a, b); oc = O.matchcost_forward(a, b, om)
        print("emd", B, n, m, "fused vs oracle rel", ((fused.cpu() - oc).abs() / oc).max().item(), "two-step", ((two.cpu() - oc).abs() / oc).max().item(),
              "match maxabs", (match.cpu() - om).abs().max().item())
    if R.available():
        rc, rm = R.emd(ac, bc, want_match=True)
        print("emd", B, n, m, "fused vs ref rel", ((fused - rc).abs() / rc).max().item(), "two-step", ((two - rc).abs() / rc).max().item(), "match maxabs", (match - rm).abs().max().item())

# timings
print("== timings (ms)")
x = torch.randn(32, 2048, 3, generator=g).to(dev)
print("fps 32x2048->1024 ours", t_ms(lambda: _ext.furthest_point_sampling(x, 1024)), "ref", t_ms(lambda: R.furthest_point_sampling(x, 1024)) if R.available() else None)
c = torch.rand(32, 3072, 3, generator=g).to(dev) * 2 - 1
print("ball 2048x3072 r=.1 ours", t_ms(lambda: _ext.ball_query(x, c, 0.1, 32)), "ref", t_ms(lambda: R.ball_query(x, c, 0.1, 32)) if R.available() else None)
a = torch.rand(256, 2048, 3, generator=g).to(dev); b = torch.rand(256, 2048, 3, generator=g).to(dev)
cf = chamfer_loss_new.Chamfer_F1()
print("chamfer_f1 256x2048x2048 ours", t_ms(lambda: cf(a, b)), "ref chamfer3D", t_ms(lambda: R.chamfer3d(a, b), 2) if R.available() else None)
print("emd fused 256x2048x2048 ours", t_ms(lambda: emd_cuda.emd_cost_forward(a, b), 2))
if R.available():
    a32, b32 = a[:32].contiguous(), b[:32].contiguous()
    print("emd 32x2048^2 ours fused", t_ms(lambda: emd_cuda.emd_cost_forward(a32, b32), 2), "ref approxmatch+matchcost", t_ms(lambda: R.emd(a32, b32, False), 1))
print("DONE")
