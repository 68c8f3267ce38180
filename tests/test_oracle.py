"""CPU: pin the oracle (oracle/pdr_oracle.c) against the reference-owned known answers and against
independent float64 restatements.  These run without a GPU."""
import numpy as np
import torch


def _fps_numpy(x, m):
    """Independent float64-free restatement of FPS semantics for tie-free data (no block emulation)."""
    n = x.shape[0]
    x = x.astype(np.float32)
    temp = np.full(n, 1e10, np.float32)
    mag = np.float32(x[:, 2] * x[:, 2]) + (np.float32(x[:, 0] * x[:, 0]) + np.float32(x[:, 1] * x[:, 1]))
    valid = mag.astype(np.float64) > 1e-3
    out = [0]
    for _ in range(1, m):
        d = ((x - x[out[-1]]) ** 2).sum(1).astype(np.float32)
        temp = np.where(valid, np.minimum(temp, d), temp)
        cand = np.where(valid, temp, -1)
        out.append(int(np.argmax(cand)))
    return np.array(out)


def test_chamfer_against_float64_golden(oracle, golden_dir):
    # ChamferDistancePytorch/unit_test.py:14-35: MSE-sum < 1e-8 and identical argmin
    g = torch.load(golden_dir + "/chamfer_f64.pt")
    d1, d2, i1, i2 = oracle.nm_distance(g["p1"], g["p2"])
    assert ((d1 - g["dist1"]) ** 2).mean() + ((d2 - g["dist2"]) ** 2).mean() < 1e-8
    assert torch.equal(i1, g["idx1"]) and torch.equal(i2, g["idx2"])
    k = oracle.knn_points(g["p1"], g["p2"], K=1)
    assert torch.equal(k.idx[..., 0].int(), g["idx1"])
    torch.testing.assert_close(k.dists[..., 0], g["dist1"], rtol=1e-5, atol=1e-7)


def test_emd_known_answer(oracle):
    # PytorchEMD/test_emd_loss.py:6-19: optimal assignment costs 0.30 + 0.41 = 0.71; this fork divides by
    # max(n, m) = 2 (pointnet2/emd.py:14-16) -> 0.355 per cloud
    p1 = torch.tensor([[[1.7, -0.1, 0.1], [0.1, 1.2, 0.3]]]).repeat(3, 1, 1)
    p2 = torch.tensor([[[0.3, 1.8, 0.2], [1.2, -0.2, 0.3]]]).repeat(3, 1, 1)
    torch.testing.assert_close(oracle.emd_distance(p1, p2), torch.full((3,), 0.355), rtol=2e-3, atol=0)


def test_emd_properties(oracle):
    g = torch.Generator().manual_seed(0)
    x = torch.rand(2, 64, 3, generator=g)
    assert oracle.emd_distance(x, x).abs().max() < 1e-4          # identical clouds
    y = torch.rand(2, 96, 3, generator=g)
    m = oracle.approxmatch_forward(x, y)                           # (b, 96, 64)
    # n < m: every xyz1 point ships mass m/n (integer division -> 1); columns of the smaller side sum to <= 1
    assert m.min() >= 0 and m.sum(1).max() <= 1.0 + 1e-4


def test_fps_matches_plain_restatement(oracle):
    g = torch.Generator().manual_seed(1)
    for n, m in ((4096, 1024), (700, 100), (64, 16), (5, 5)):
        x = torch.randn(1, n, 3, generator=g)
        idx = oracle.furthest_point_sampling(x, m)[0].numpy()
        assert idx[0] == 0 and len(set(idx.tolist())) == m
        np.testing.assert_array_equal(idx, _fps_numpy(x[0].numpy(), m))


def test_fps_skip_rule_and_degenerate(oracle):
    x = torch.zeros(1, 16, 3)                                       # every point skipped -> always index 0
    assert oracle.furthest_point_sampling(x, 5).abs().sum() == 0
    g = torch.Generator().manual_seed(2)
    x = torch.randn(1, 128, 3, generator=g)
    x[0, 7] = 0.01                                                  # |p|^2 = 3e-4 <= 1e-3: never selectable
    assert 7 not in oracle.furthest_point_sampling(x, 128)[0, 1:].tolist()


def test_ball_query_semantics(oracle):
    g = torch.Generator().manual_seed(3)
    xyz = torch.rand(2, 300, 3, generator=g)
    centres = torch.rand(2, 40, 3, generator=g)
    idx, cnt = oracle.ball_query(centres, xyz, 0.25, 8)
    d = ((centres[:, :, None] - xyz[:, None]) ** 2).sum(-1)
    for b in range(2):
        for j in range(40):
            hits = torch.nonzero(d[b, j] < 0.25 ** 2 - 1e-6)[:, 0]
            c = int(cnt[b, j])
            assert c == min(len(torch.nonzero(d[b, j] < 0.0625)[:, 0]), 8) or abs(c - min(len(hits), 8)) <= 1
            row = idx[b, j]
            if c == 0:
                assert row.abs().sum() == 0
            else:
                assert torch.all(row[:c][1:] > row[:c][:-1]) and torch.all(row[c:] == row[0])
    # radius so small that nothing matches, and empty-ish input
    idx, cnt = oracle.ball_query(centres, xyz + 10, 0.1, 4)
    assert idx.abs().sum() == 0 and cnt.sum() == 0


def test_three_nn_and_interpolate(oracle):
    g = torch.Generator().manual_seed(4)
    u, k = torch.rand(2, 50, 3, generator=g), torch.rand(2, 20, 3, generator=g)
    d2, idx = oracle.three_nn(u, k)
    full = ((u[:, :, None] - k[:, None]) ** 2).sum(-1)
    top = full.topk(3, dim=2, largest=False)
    assert torch.equal(idx.long(), top.indices)
    torch.testing.assert_close(d2, top.values, rtol=1e-5, atol=1e-7)
    w = torch.rand(2, 50, 3, generator=g)
    f = torch.randn(2, 6, 20, generator=g)
    out = oracle.three_interpolate(f, idx, w)
    ref = sum(f.gather(2, idx[:, :, t].long().unsqueeze(1).expand(-1, 6, -1)) * w[:, :, t].unsqueeze(1) for t in range(3))
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-6)
    # fewer than 3 known points: the reference leaves (1e40 -> inf, index 0) in the unused slots
    d2, idx = oracle.three_nn(u, k[:, :2].contiguous())
    assert torch.isinf(d2[..., 2]).all() and (idx[..., 2] == 0).all()


def test_knn_ties_prefer_lower_index(oracle):
    y = torch.tensor([[[1.0, 0, 0], [0, 1.0, 0], [-1.0, 0, 0], [0, 0, 2.0]]])
    x = torch.zeros(1, 1, 3)
    k = oracle.knn_points(x, y, K=3)
    assert k.idx[0, 0].tolist() == [0, 1, 2]
    k = oracle.knn_points(x, y, K=6)                                # K > P2 -> zero padded tail
    assert k.idx[0, 0].tolist() == [0, 1, 2, 3, 0, 0] and k.dists[0, 0, 4:].abs().sum() == 0


def test_gather_group_are_pure_copies(oracle):
    g = torch.Generator().manual_seed(5)
    f = torch.randn(2, 5, 30, generator=g)
    idx = torch.randint(0, 30, (2, 7, 4), generator=g, dtype=torch.int32)
    out = oracle.group_points(f, idx)
    ref = f.gather(2, idx.long().reshape(2, 1, 28).expand(-1, 5, -1)).reshape(2, 5, 7, 4)
    assert torch.equal(out, ref)
    assert torch.equal(oracle.gather_points(f, idx[:, :, 0].contiguous()), ref[..., 0])


def test_point_upsample_and_mirror_partial_against_reference_fixture(oracle, golden_dir):
    """the oracle's restatements against outputs of the reference's own
    point_upsample_module.py / mirror_partial.py (tests/golden/make_golden.py), bit for bit."""
    g = torch.load(golden_dir + "/refinement_io.pt")
    for case in g["upsample"]:
        refined, mid = oracle.point_upsample(case["coarse"], case["disp"], case["factor"], case["centre"], case["scale"])
        assert torch.equal(refined, case["refined"]) and torch.equal(mid, case["mid"]), case["factor"]
        assert refined.shape == (3, 50 * case["factor"], 3)
    out = oracle.mirror_and_concat(g["partial"], axis=2, num_points=[256, 384])
    assert len(out) == 3 and all(torch.equal(a, b) for a, b in zip(out, g["mirror_axis2_256_384"]))
    out = oracle.mirror_and_concat(g["partial"], axis=1, num_points=[128])
    assert all(torch.equal(a, b) for a, b in zip(out, g["mirror_axis1_128"]))
    assert (out[0][:, 300:, 3] == -1).all() and (out[0][:, :300, 3] == 1).all()
