"""GPU parity tests proper: every kernel, called through the C ABI (via the ctypes shims), against the CPU
oracle on the same seeded inputs -- and, when oracle/_ref/libpdr_ref_cuda.so travelled with the snapshot,
against the reference's own CUDA kernels compiled for sm_100a.

Bars: indices / copies bit-exact; three_interpolate and NmDistance bit-exact (same rounding sequence);
Chamfer/F1 rtol 1e-5; EMD rtol 2e-3 (the ex2.approx path and summation order differ)."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_cuda
    return ref_cuda if ref_cuda.available() else None


def _cloud(B, N, dist, seed, skip=False):
    g = torch.Generator().manual_seed(seed)
    x = (torch.rand(B, N, 3, generator=g) * 2 - 1) if dist == "U" else torch.randn(B, N, 3, generator=g)
    if skip and N > 5:
        x[:, 5] = 0.001          # |p|^2 <= 1e-3: skipped by FPS (sampling_gpu.cu:100-101)
    return x


FPS_CASES = [(1, 4096, 1024, "U"), (1, 4096, 1024, "G"), (4, 2048, 1024, "G"), (3, 3072, 1024, "U"),
             (2, 1000, 300, "G"), (2, 1024, 256, "G"), (2, 256, 64, "G"), (2, 64, 16, "U"), (2, 20, 7, "G"),
             (1, 1, 1, "G"), (1, 8192, 512, "G"), (1, 16384, 64, "U"), (1, 20000, 40, "G")]


@pytest.mark.parametrize("B,N,M,dist", FPS_CASES)
def test_fps_bit_exact(oracle, ref, B, N, M, dist):
    from point_diffusion_refinement_b200 import _ext
    x = _cloud(B, N, dist, seed=N + M, skip=True)
    got = _ext.furthest_point_sampling(x.to(DEV), M)
    assert got.dtype == torch.int32 and got.shape == (B, M)
    assert torch.equal(got.cpu(), oracle.furthest_point_sampling(x, M))
    if ref is not None:
        assert torch.equal(got, ref.furthest_point_sampling(x.to(DEV), M))


def test_fps_ties_and_degenerate(oracle, ref):
    """Exact ties exercise the tie-break contract (bit-reversed residue, then quotient)."""
    from point_diffusion_refinement_b200 import _ext
    g = torch.Generator().manual_seed(0)
    base = torch.randint(-2, 3, (2, 2048, 3), generator=g).float()       # lattice -> masses of equal distances
    dup = torch.randn(1, 600, 3, generator=g).repeat(1, 3, 1)             # every point three times
    zero = torch.zeros(1, 64, 3)                                           # everything skipped
    for x, m in ((base, 200), (dup, 700), (zero, 9), (base[:, :700].contiguous(), 300)):
        got = _ext.furthest_point_sampling(x.to(DEV), m)
        assert torch.equal(got.cpu(), oracle.furthest_point_sampling(x, m))
        if ref is not None:
            assert torch.equal(got, ref.furthest_point_sampling(x.to(DEV), m))


@pytest.mark.parametrize("B,N,M,dist", [(1, 4096, 1024, "U"), (2, 3072, 2048, "U"), (2, 2048, 1024, "G"),
                                        (3, 300, 77, "G"), (2, 64, 64, "U"), (1, 20000, 100, "U")])
@pytest.mark.parametrize("radius,nsample", [(0.1, 32), (0.2, 32), (0.4, 32), (0.8, 16), (3.0, 5), (0.05, 64)])
def test_ball_query_bit_exact(oracle, ref, B, N, M, dist, radius, nsample):
    from point_diffusion_refinement_b200 import _ext
    xyz = _cloud(B, N, dist, seed=N)
    centres = xyz[:, :M].contiguous() if M <= N // 2 else _cloud(B, M, dist, seed=M + 1)
    idx, cnt = _ext.ball_query(centres.to(DEV), xyz.to(DEV), radius, nsample)
    oi, oc = oracle.ball_query(centres, xyz, radius, nsample)
    assert torch.equal(idx.cpu(), oi) and torch.equal(cnt.cpu(), oc)
    if ref is not None:
        ri, rc = ref.ball_query(centres.to(DEV), xyz.to(DEV), radius, nsample)
        assert torch.equal(idx, ri) and torch.equal(cnt, rc)


def test_ball_query_no_neighbours_rows_are_zero():
    from point_diffusion_refinement_b200 import _ext
    xyz = torch.rand(2, 100, 3, device=DEV)
    idx, cnt = _ext.ball_query(xyz[:, :10].contiguous() + 10, xyz, 0.1, 8)
    assert idx.abs().sum() == 0 and cnt.sum() == 0


@pytest.mark.parametrize("B,C,N,NP,NS", [(2, 3, 2048, 1024, 32), (2, 35, 1024, 256, 32), (1, 320, 64, 16, 32),
                                         (3, 7, 50, 13, 5), (1, 1, 1, 1, 1)])
def test_gather_and_group_are_exact_copies(oracle, ref, B, C, N, NP, NS):
    from point_diffusion_refinement_b200 import _ext
    g = torch.Generator().manual_seed(C)
    f = torch.randn(B, C, N, generator=g)
    idx = torch.randint(0, N, (B, NP, NS), generator=g, dtype=torch.int32)
    out = _ext.group_points(f.to(DEV), idx.to(DEV))
    assert torch.equal(out.cpu(), oracle.group_points(f, idx))
    i1 = idx[:, :, 0].contiguous()
    assert torch.equal(_ext.gather_points(f.to(DEV), i1.to(DEV)).cpu(), oracle.gather_points(f, i1))
    if ref is not None:
        assert torch.equal(out, ref.group_points(f.to(DEV), idx.to(DEV)))


@pytest.mark.parametrize("B,n,m", [(2, 2048, 1024), (2, 1024, 256), (3, 100, 33), (1, 50, 2), (1, 5000, 3000)])
def test_three_nn_and_interpolate_bit_exact(oracle, ref, B, n, m):
    from point_diffusion_refinement_b200 import _ext
    u, k = _cloud(B, n, "G", 1), _cloud(B, m, "G", 2)
    d2, idx = _ext.three_nn(u.to(DEV), k.to(DEV))
    od, oi = oracle.three_nn(u, k)
    assert torch.equal(idx.cpu(), oi) and torch.equal(d2.cpu(), od)
    g = torch.Generator().manual_seed(3)
    w = torch.rand(B, n, 3, generator=g)
    w = w / w.sum(2, keepdim=True)
    f = torch.randn(B, 19, m, generator=g)
    out = _ext.three_interpolate(f.to(DEV), idx, w.to(DEV))
    assert torch.equal(out.cpu(), oracle.three_interpolate(f, oi, w))
    if ref is not None:
        rd, ri = ref.three_nn(u.to(DEV), k.to(DEV))
        assert torch.equal(idx, ri) and torch.equal(d2, rd)
        assert torch.equal(out, ref.three_interpolate(f.to(DEV), idx, w.to(DEV)))


@pytest.mark.parametrize("K", [1, 3, 8, 20, 32, 40])
@pytest.mark.parametrize("B,p1,p2", [(2, 2048, 1024), (2, 64, 16), (1, 300, 1500), (1, 10, 5)])
def test_knn_points_exact(oracle, B, p1, p2, K):
    from point_diffusion_refinement_b200 import knn
    x, y = _cloud(B, p1, "G", 4), _cloud(B, p2, "G", 5)
    r = knn.knn_points(x.to(DEV), y.to(DEV), K=K, return_nn=True)
    o = oracle.knn_points(x, y, K=K, return_nn=True)
    assert r.idx.dtype == torch.int64
    assert torch.equal(r.idx.cpu(), o.idx) and torch.equal(r.dists.cpu(), o.dists) and torch.equal(r.knn.cpu(), o.knn)


def test_knn_exact_ties_keep_lower_index():
    from point_diffusion_refinement_b200 import knn
    y = torch.tensor([[[1.0, 0, 0], [0, 1.0, 0], [-1.0, 0, 0], [0, 0, 2.0]]], device=DEV)
    r = knn.knn_points(torch.zeros(1, 1, 3, device=DEV), y, K=3)
    assert r.idx[0, 0].tolist() == [0, 1, 2]


def test_group_knn_layout(oracle):
    """[feat(C), d2, w, nn_abs(3), nn_rel(3), x(3)] -> (B, C+11, N1, K)  (pointnet2_utils.py:487-514)."""
    from point_diffusion_refinement_b200.pointnet2_utils import group_knn
    g = torch.Generator().manual_seed(6)
    x, y, f = torch.randn(2, 50, 3, generator=g), torch.randn(2, 20, 3, generator=g), torch.randn(2, 6, 20, generator=g)
    out = group_knn(x.to(DEV), y.to(DEV), f.to(DEV), 8, transpose=True)
    assert out.shape == (2, 17, 50, 8)
    o = oracle.knn_points(x, y, K=8, return_nn=True)
    torch.testing.assert_close(out[:, 6].cpu(), o.dists, rtol=0, atol=0)
    torch.testing.assert_close(out[:, 7].sum(-1).cpu(), torch.ones(2, 50), rtol=1e-5, atol=1e-6)
    assert torch.equal(out[:, 8:11].cpu(), o.knn.permute(0, 3, 1, 2))


def test_chamfer_unit_test_of_the_reference(golden_dir, ref):
    """ChamferDistancePytorch/unit_test.py:14-35 verbatim criteria."""
    from point_diffusion_refinement_b200._lib import call, dptr, stream_ptr
    g = torch.load(golden_dir + "/chamfer_f64.pt")
    p1, p2 = g["p1"].to(DEV), g["p2"].to(DEV)
    d1 = torch.empty(4, 100, device=DEV); d2 = torch.empty(4, 200, device=DEV)
    i1 = torch.empty(4, 100, dtype=torch.int32, device=DEV); i2 = torch.empty(4, 200, dtype=torch.int32, device=DEV)
    call("pdr_nm_distance", 4, 100, 200, dptr(p1), dptr(p2), dptr(d1), dptr(i1), dptr(d2), dptr(i2), stream_ptr(p1))
    assert ((d1.cpu() - g["dist1"]) ** 2).mean() + ((d2.cpu() - g["dist2"]) ** 2).mean() < 1e-8
    assert torch.equal(i1.cpu(), g["idx1"]) and torch.equal(i2.cpu(), g["idx2"])
    if ref is not None:
        r1, r2, ri1, ri2 = ref.chamfer3d(p1, p2)
        assert torch.equal(d1, r1) and torch.equal(d2, r2) and torch.equal(i1, ri1) and torch.equal(i2, ri2)


@pytest.mark.parametrize("B,n,m", [(3, 500, 700), (2, 2048, 2048), (5, 1, 3), (2, 1500, 100)])
def test_chamfer_f1_fused(oracle, B, n, m):
    from point_diffusion_refinement_b200.chamfer_loss_new import Chamfer_F1, calc_cd, chamfer_distance
    g = torch.Generator().manual_seed(n)
    a, b = torch.rand(B, n, 3, generator=g), torch.rand(B, m, 3, generator=g) * 0.9
    cp, ct, f1 = Chamfer_F1(f1_threshold=1e-3)(a.to(DEV), b.to(DEV))
    ocp, oct_, of1 = oracle.chamfer_f1(a, b, 1e-3)
    torch.testing.assert_close(cp.cpu(), ocp, rtol=1e-5, atol=1e-8)
    torch.testing.assert_close(ct.cpu(), oct_, rtol=1e-5, atol=1e-8)
    torch.testing.assert_close(f1.cpu(), of1, rtol=1e-5, atol=1e-7)
    d1, d2, _ = chamfer_distance(b.to(DEV), a.to(DEV), batch_reduction=None, point_reduction=None)
    torch.testing.assert_close((d1.mean(1) + d2.mean(1)).cpu(), oct_, rtol=1e-5, atol=1e-8)
    assert len(calc_cd(a.to(DEV), b.to(DEV))) == 2


def test_chamfer_large_properties():
    """BASELINE config 4 sizes (16384 x 16384): size-independent properties instead of a CPU oracle."""
    from point_diffusion_refinement_b200.chamfer_loss_new import Chamfer_F1
    g = torch.Generator().manual_seed(0)
    a = torch.rand(4, 16384, 3, generator=g).to(DEV)
    perm = torch.randperm(16384, generator=g).to(DEV)
    cf = Chamfer_F1()
    cp, ct, f1 = cf(a, a[:, perm].contiguous())                 # same set, permuted -> zero distance, F1 = 1
    assert cp.abs().max() == 0 and ct.abs().max() == 0 and torch.all(f1 == 1)
    b = a + 0.01
    cp1, ct1, _ = cf(a, b)
    cp2, ct2, _ = cf(b, a)                                        # symmetric in its arguments
    torch.testing.assert_close(ct1, ct2, rtol=1e-6, atol=0)
    torch.testing.assert_close(cp1, cp2, rtol=1e-6, atol=0)
    assert torch.all(ct1 <= 2 * 3 * 0.01 ** 2 * 1.0001)           # each point has a neighbour at distance <= |shift|


def test_chamfer_16384_against_the_oracle(oracle):
    """BASELINE config 4's second shape (16384 x 16384) against the CPU oracle on one cloud pair (< 1 s of host work)."""
    from point_diffusion_refinement_b200.chamfer_loss_new import Chamfer_F1
    g = torch.Generator().manual_seed(3)
    a, b = torch.rand(1, 16384, 3, generator=g) * 2 - 1, torch.rand(1, 16384, 3, generator=g) * 2 - 1
    cp, ct, f1 = Chamfer_F1(f1_threshold=1e-3)(a.to(DEV), b.to(DEV))
    ocp, oct_, of1 = oracle.chamfer_f1(a, b, 1e-3)
    torch.testing.assert_close(cp.cpu(), ocp, rtol=1e-5, atol=1e-8)
    torch.testing.assert_close(ct.cpu(), oct_, rtol=1e-5, atol=1e-8)
    torch.testing.assert_close(f1.cpu(), of1, rtol=1e-5, atol=1e-7)


def test_emd_gradient_through_a_non_contiguous_input():
    """transpose=True hands forward() a non-contiguous tensor; the match must still be saved for backward."""
    from point_diffusion_refinement_b200.emd import earth_mover_distance
    g = torch.Generator().manual_seed(1)
    a = torch.rand(2, 3, 64, generator=g).to(DEV).requires_grad_(True)
    b = torch.rand(2, 3, 64, generator=g).to(DEV)
    cost = earth_mover_distance(a, b, transpose=True)
    cost.sum().backward()
    assert a.grad is not None and a.grad.shape == a.shape and torch.isfinite(a.grad).all() and a.grad.abs().sum() > 0
    a2 = a.detach().transpose(1, 2).contiguous().requires_grad_(True)
    earth_mover_distance(a2, b.transpose(1, 2).contiguous()).sum().backward()
    torch.testing.assert_close(a.grad.transpose(1, 2), a2.grad, rtol=1e-5, atol=1e-7)


def test_emd_known_answer_and_api():
    from point_diffusion_refinement_b200.emd import EMD_distance, earth_mover_distance
    p1 = torch.tensor([[[1.7, -0.1, 0.1], [0.1, 1.2, 0.3]]], device=DEV).repeat(3, 1, 1)
    p2 = torch.tensor([[[0.3, 1.8, 0.2], [1.2, -0.2, 0.3]]], device=DEV).repeat(3, 1, 1)
    d = EMD_distance()(p1, p2)
    torch.testing.assert_close(d.cpu(), torch.full((3,), 0.355), rtol=2e-3, atol=0)   # test_emd_loss.py KAT / 2
    d2, match = earth_mover_distance(p1, p2, return_match=True)
    assert match.shape == (3, 2, 2)
    torch.testing.assert_close(d2, d, rtol=1e-5, atol=0)
    torch.testing.assert_close(EMD_distance()(p1.transpose(1, 2), p2.transpose(1, 2), transpose=True), d)
    assert EMD_distance()(p1[0], p2[0]).shape == (1,)


@pytest.mark.parametrize("B,n,m", [(2, 256, 256), (3, 512, 300), (2, 300, 1024), (40, 1024, 1024), (300, 128, 128),
                                   (2, 2048, 2048), (1, 7, 5)])
def test_emd_against_oracle_and_reference(oracle, ref, B, n, m):
    from point_diffusion_refinement_b200 import emd_cuda
    g = torch.Generator().manual_seed(n + m)
    a, b = torch.rand(B, n, 3, generator=g), torch.rand(B, m, 3, generator=g)
    ac, bc = a.to(DEV), b.to(DEV)
    fused = emd_cuda.emd_cost_forward(ac, bc)
    match = emd_cuda.approxmatch_forward(ac, bc)
    two = emd_cuda.matchcost_forward(ac, bc, match)
    torch.testing.assert_close(fused, two, rtol=1e-4, atol=1e-6)
    if B * n * m <= 3 * 512 * 512:
        om = oracle.approxmatch_forward(a, b)
        oc = oracle.matchcost_forward(a, b, om)
        torch.testing.assert_close(fused.cpu(), oc, rtol=2e-3, atol=1e-6)
        torch.testing.assert_close(match.cpu(), om, rtol=0, atol=2e-3)
    if ref is not None:
        rc, rm = ref.emd(ac, bc, want_match=True)
        torch.testing.assert_close(fused, rc, rtol=2e-3, atol=1e-6)
        torch.testing.assert_close(match, rm, rtol=0, atol=1e-4)


def test_emd_gradient_matches_autograd_of_matchcost():
    from point_diffusion_refinement_b200 import emd_cuda
    g = torch.Generator().manual_seed(1)
    a, b = torch.rand(2, 64, 3, generator=g).to(DEV), torch.rand(2, 96, 3, generator=g).to(DEV)
    match = emd_cuda.approxmatch_forward(a, b)
    a1, b1 = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    d = ((b1.unsqueeze(2) - a1.unsqueeze(1)) ** 2).sum(-1)           # (B, m, n)
    gc = torch.tensor([0.5, 2.0], device=DEV)
    ((d * match).sum((1, 2)) * gc).sum().backward()
    g1, g2 = emd_cuda.matchcost_backward(gc, a, b, match)
    torch.testing.assert_close(g1, a1.grad, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(g2, b1.grad, rtol=1e-4, atol=1e-6)


def test_backward_kernels_match_autograd():
    from point_diffusion_refinement_b200 import pointnet2_utils as pu
    g = torch.Generator().manual_seed(2)
    f = torch.randn(2, 5, 40, generator=g).to(DEV).requires_grad_(True)
    idx = torch.randint(0, 40, (2, 9, 4), generator=g, dtype=torch.int32).to(DEV)
    out = pu.grouping_operation(f, idx)
    w = torch.randn_like(out)
    (out * w).sum().backward()
    ref = torch.zeros_like(f)
    ref.scatter_add_(2, idx.long().reshape(2, 1, 36).expand(-1, 5, -1), w.reshape(2, 5, 36))
    torch.testing.assert_close(f.grad, ref, rtol=1e-5, atol=1e-6)
    f.grad = None
    out = pu.gather_operation(f, idx[:, :, 0].contiguous())
    out.sum().backward()
    ref = torch.zeros_like(f)
    ref.scatter_add_(2, idx[:, :, 0].long().unsqueeze(1).expand(-1, 5, -1), torch.ones(2, 5, 9, device=DEV))
    torch.testing.assert_close(f.grad, ref)


def test_device_noise_statistics_and_replay():
    from point_diffusion_refinement_b200.util import DeviceNoise
    z = DeviceNoise(seed=7).normal((64, 2048, 3), torch.device(DEV))
    assert abs(z.mean().item()) < 5e-3 and abs(z.std().item() - 1) < 5e-3
    assert abs((z ** 4).mean().item() - 3) < 0.05 and z.abs().max() < 7
    z2 = DeviceNoise(seed=7).normal((64, 2048, 3), torch.device(DEV))
    assert torch.equal(z, z2) and not torch.equal(z, DeviceNoise(seed=8).normal((64, 2048, 3), torch.device(DEV)))
    x = torch.randn(1000, device=DEV); e = torch.randn(1000, device=DEV); n = torch.randn(1000, device=DEV)
    y = DeviceNoise(0).affine_update(x.clone(), e, 1.5, -0.25, 0.1, noise=n)
    torch.testing.assert_close(y, x * 1.5 - 0.25 * e + 0.1 * n, rtol=1e-6, atol=1e-6)


def test_input_validation_raises():
    from point_diffusion_refinement_b200 import _ext
    x = torch.rand(1, 8, 3, device=DEV)
    with pytest.raises(RuntimeError, match="contiguous"):
        _ext.furthest_point_sampling(x.transpose(1, 2).transpose(1, 2)[:, ::2], 2)
    with pytest.raises(RuntimeError, match="float"):
        _ext.furthest_point_sampling(x.double(), 2)
    with pytest.raises(RuntimeError, match="int"):
        _ext.gather_points(x.transpose(1, 2).contiguous(), torch.zeros(1, 2, dtype=torch.int64, device=DEV))
