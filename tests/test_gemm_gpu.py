"""GPU: the fused GEMM primitive (pdr_gemm_fused) -- fp32 SIMT path against a float64 torch restatement, and
the tcgen05 TF32 path against the SIMT path -- over the shapes, prologues and epilogues the denoiser uses."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _run(lib, A, W, bias, B, rps, N, pro, sc, sh, add, R, rowadd, div, use_tf32, want_stats=True, ldc=None, stats_skip=0, stats_skip_blocks=0):
    from point_diffusion_refinement_b200.fused import GemmArgs
    M, K = A.shape
    ldc = ldc or (N + 3) // 4 * 4
    C = torch.full((M, ldc), float("nan"), device=DEV)
    tiles = (rps + lib.pdr_gemm_tile_rows() - 1) // lib.pdr_gemm_tile_rows()
    stats = torch.zeros(B * tiles, N, 4, device=DEV)
    g = GemmArgs()
    g.A, g.lda, g.K = A.data_ptr(), A.stride(0), K
    g.W, g.ldw = W.data_ptr(), W.stride(0)
    g.bias = bias.data_ptr() if bias is not None else None
    g.C, g.ldc, g.N, g.ldc_zero_to = C.data_ptr(), ldc, N, ldc
    g.batch, g.rows_per_sample, g.pro_mode = B, rps, pro
    if pro:
        g.sc, g.sh, g.ld_scsh = sc.data_ptr(), sh.data_ptr(), sc.stride(0)
    if add is not None:
        g.add, g.ld_add = add.data_ptr(), add.stride(0)
    if R is not None:
        g.R, g.ldr = R.data_ptr(), R.stride(0)
    if rowadd is not None:
        g.rowadd, g.ld_rowadd, g.rowadd_div = rowadd.data_ptr(), rowadd.stride(0), div
    g.stats = stats.data_ptr() if want_stats else None
    g.use_tf32 = int(use_tf32)
    g.stats_skip = stats_skip
    g.stats_skip_blocks = stats_skip_blocks
    rc = lib.pdr_gemm_fused(ctypes.c_void_p(ctypes.addressof(g)), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, lib.pdr_last_error_string()
    torch.cuda.synchronize()
    return C, stats.view(B, tiles, N, 4).sum(1)


def _reference(A, W, bias, B, rps, N, pro, sc, sh, add, R, rowadd, div):
    M, K = A.shape
    x = A.double().view(B, rps, K)
    if pro == 1:
        x = torch.relu(x * sc.double()[:, None, :K] + sh.double()[:, None, :K])
    elif pro == 2:
        x = torch.relu(x) * sc.double()[:, None, :K] + sh.double()[:, None, :K]
    if add is not None:
        x = x + add.double()[:, None, :K]
    x = x.reshape(M, K)
    if R is not None:
        x = x + R.double()[:, :K]
    y = x @ W.double()[:, :K].t()
    if bias is not None:
        y = y + bias.double()
    if rowadd is not None:
        y = y + rowadd.double()[:, :N].repeat_interleave(div, dim=0)
    st = torch.stack([y.view(B, rps, N).sum(1), (y ** 2).view(B, rps, N).sum(1), torch.relu(y).view(B, rps, N).sum(1),
                      (torch.relu(y) ** 2).view(B, rps, N).sum(1)], dim=-1)
    return y, st


CASES = [  # B, rows_per_sample, K, N, pro, add, R, rowadd_div
    (2, 4096, 44, 140, 0, False, False, 0),      # SA0 merged first|res|key
    (2, 4096, 32, 32, 1, True, False, 0),        # GN+ReLU prologue + t embedding
    (3, 1024, 64, 64, 1, True, True, 0),         # feat_out_conv with residual
    (2, 2048, 44, 64, 2, False, False, 8),       # weight_conv key part + broadcast query part
    (2, 512, 652, 256, 1, False, False, 0),      # deep K
    (1, 4160, 332, 300, 0, False, False, 0),     # N > 256 (two column tiles), ragged rows (4160 = 32*130)
    (5, 200, 128, 3, 1, False, False, 0),        # head: N = 3, rows not a multiple of 128
    (2, 16, 36, 36, 0, False, False, 0),         # tiny query conv
    (2, 2048, 332, 300, 1, True, False, 0),      # prologue + streamed weights (cp.async into the claimed stage)
    (2, 1024, 512, 512, 1, True, True, 0),       # residual + wide N: planner falls back to 128-column tiles
    (2, 4096, 172, 128, 1, True, False, 0),      # prologue, weights resident (96 KiB)
    (1, 8192, 332, 588, 0, False, False, 0),     # no prologue, 3 column tiles, streamed weights
    (3, 640, 44, 64, 2, False, False, 5),        # broadcast row groups that straddle the 16-row epilogue halves
    (2, 200, 32, 32, 1, True, False, 8),         # narrow tile, ragged rows, row groups of 8    (1, 32, 512, 1248, 0, False, False, 0),      # t-embedding GEMM: one row tile -> spread over 39 column tiles of 32
    (4, 128, 512, 512, 1, True, True, 0),        # deepest level: 4 row tiles -> 32-column tiles, prologue + residual
    (2, 1024, 256, 256, 1, True, False, 0),      # 16 row tiles -> 64-column tiles
    (2, 4096, 256, 256, 2, False, False, 32),    # 64 row tiles -> 128-column tiles, broadcast row groups
]


@pytest.mark.parametrize("case", CASES)
def test_gemm_simt_matches_float64(cuda_lib, case):
    B, rps, K, N, pro, use_add, use_R, div = case
    g = torch.Generator().manual_seed(K * N + rps)
    M = B * rps
    A = torch.randn(M, K, generator=g).to(DEV)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    sc = (1 + 0.2 * torch.randn(B, K, generator=g)).to(DEV)
    sh = (0.2 * torch.randn(B, K, generator=g)).to(DEV)
    add = torch.randn(B, K, generator=g).to(DEV) if use_add else None
    R = torch.randn(M, K, generator=g).to(DEV) if use_R else None
    rowadd = torch.randn(M // div, (N + 3) // 4 * 4, generator=g).to(DEV) if div else None
    y64, st64 = _reference(A, W, bias, B, rps, N, pro, sc, sh, add, R, rowadd, div)
    C, st = _run(cuda_lib, A, W, bias, B, rps, N, pro, sc, sh, add, R, rowadd, div, use_tf32=False)
    torch.testing.assert_close(C[:, :N].double(), y64, rtol=1e-4, atol=1e-4)
    assert C[:, N:].abs().sum() == 0                                  # pad columns are written as zeros
    torch.testing.assert_close(st.double(), st64, rtol=1e-3, atol=1e-2)


@pytest.mark.parametrize("case", CASES)
def test_gemm_tcgen05_matches_simt(cuda_lib, case):
    B, rps, K, N, pro, use_add, use_R, div = case
    g = torch.Generator().manual_seed(K * N + rps + 1)
    M = B * rps
    A = torch.randn(M, K, generator=g).to(DEV)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    sc = (1 + 0.2 * torch.randn(B, K, generator=g)).to(DEV)
    sh = (0.2 * torch.randn(B, K, generator=g)).to(DEV)
    add = torch.randn(B, K, generator=g).to(DEV) if use_add else None
    R = torch.randn(M, K, generator=g).to(DEV) if use_R else None
    rowadd = torch.randn(M // div, (N + 3) // 4 * 4, generator=g).to(DEV) if div else None
    C0, st0 = _run(cuda_lib, A, W, bias, B, rps, N, pro, sc, sh, add, R, rowadd, div, use_tf32=False)
    # engine contract (fused.FusedDenoiser._pack): weights handed to the tensor-core path are pre-rounded to TF32
    # (round-to-nearest), because the hardware truncates raw fp32 operands and a truncated W biases every row of a
    # column the same way
    from point_diffusion_refinement_b200.fused import tf32_round
    Wt = tf32_round(W)
    C1, st1 = _run(cuda_lib, A, Wt, bias, B, rps, N, pro, sc, sh, add, R, rowadd, div, use_tf32=True)
    # TF32: 10-bit mantissa inputs, fp32 accumulate -> ~1e-3 relative to the row scale
    torch.testing.assert_close(C1[:, :N], C0[:, :N], rtol=5e-3, atol=5e-3)
    # statistics: TF32 noise is ~1e-3 of every summand, so compare against the L1 mass of the column
    l1 = C0[:, :N].abs().view(B, rps, N).sum(1)
    l2 = st0[..., 1]
    assert ((st1[..., 0] - st0[..., 0]).abs() <= 3e-3 * l1 + 1e-2).all()
    assert ((st1[..., 2] - st0[..., 2]).abs() <= 3e-3 * l1 + 1e-2).all()
    # ... and exactly (to fp32 summation order) against the values the kernel itself stored
    own = C1[:, :N].double().view(B, rps, N)
    assert ((st1[..., 0].double() - own.sum(1)).abs() <= 1e-5 * l1.double() + 1e-2).all()
    assert ((st1[..., 3].double() - (own.clamp(min=0) ** 2).sum(1)).abs() <= 1e-5 * l2.double() + 1e-2).all()
    assert ((st1[..., 1] - st0[..., 1]).abs() <= 5e-3 * l2 + 1e-2).all()
    assert ((st1[..., 3] - st0[..., 3]).abs() <= 5e-3 * l2 + 1e-2).all()
    C2, _ = _run(cuda_lib, A, Wt, bias, B, rps, N, pro, sc, sh, add, R, rowadd, div, use_tf32=True)
    assert torch.equal(C1[:, :N], C2[:, :N])                          # deterministic


@pytest.mark.parametrize("case", [CASES[0], CASES[1], CASES[3], CASES[5], CASES[12]])
def test_gemm_stats_skip_hint(cuda_lib, case):
    """PdrGemmArgs.stats_skip: the pair a consumer asked for is bit-identical to the full computation, the output
    matrix does not change, whichever pair the epilogue leaves out."""
    B, rps, K, N, pro, use_add, use_R, div = case
    g = torch.Generator().manual_seed(K * N + rps + 2)
    M = B * rps
    A = torch.randn(M, K, generator=g).to(DEV)
    from point_diffusion_refinement_b200.fused import tf32_round
    W = tf32_round((torch.randn(N, K, generator=g) / K ** 0.5).to(DEV))
    bias = torch.randn(N, generator=g).to(DEV)
    sc = (1 + 0.2 * torch.randn(B, K, generator=g)).to(DEV)
    sh = (0.2 * torch.randn(B, K, generator=g)).to(DEV)
    add = torch.randn(B, K, generator=g).to(DEV) if use_add else None
    R = torch.randn(M, K, generator=g).to(DEV) if use_R else None
    rowadd = torch.randn(M // div, (N + 3) // 4 * 4, generator=g).to(DEV) if div else None
    C0, st0 = _run(cuda_lib, A, W, bias, B, rps, N, pro, sc, sh, add, R, rowadd, div, use_tf32=True)
    C1, st1 = _run(cuda_lib, A, W, bias, B, rps, N, pro, sc, sh, add, R, rowadd, div, use_tf32=True, stats_skip=2)
    C2, st2 = _run(cuda_lib, A, W, bias, B, rps, N, pro, sc, sh, add, R, rowadd, div, use_tf32=True, stats_skip=1)
    C3, _ = _run(cuda_lib, A, W, bias, B, rps, N, pro, sc, sh, add, R, rowadd, div, use_tf32=True, stats_skip=3)
    assert torch.equal(C0, C1) and torch.equal(C0, C2) and torch.equal(C0, C3)
    assert torch.equal(st1[..., :2], st0[..., :2]) and torch.equal(st2[..., 2:], st0[..., 2:])
    # per 32-column block (stats_skip_blocks): block 0 keeps the plain pair only, block 1 the relu pair only, the rest everything
    C4, st4 = _run(cuda_lib, A, W, bias, B, rps, N, pro, sc, sh, add, R, rowadd, div, use_tf32=True, stats_skip_blocks=2 | (1 << 2))
    assert torch.equal(C0, C4)
    assert torch.equal(st4[:, :32, :2], st0[:, :32, :2]) and torch.equal(st4[:, 32:64, 2:], st0[:, 32:64, 2:])
    assert torch.equal(st4[:, 64:], st0[:, 64:])


@pytest.mark.parametrize("shape", [(2, 4096, 35, 9, 140, 3000), (3, 1000, 4, 9, 96, 700), (2, 2048, 160, 11, 428, 512),
                                   (1, 8320, 320, 11, 588, 64), (2, 640, 32, 9, 32, 5000), (2, 1002, 64, 9, 64, 900)])
def test_gemm_gathered_operand_equals_materialised(cuda_lib, shape):
    """PdrGemmArgs.a_rows: A assembled by the producers from (feature table, row index, geometric channels) gives
    bit for bit the GEMM over the materialised grouped tensor -- including rows whose index is -1 (zero features),
    ragged last tiles and several column tiles."""
    import ctypes
    from point_diffusion_refinement_b200.fused import GemmArgs, tf32_round
    B, rps, C, n_geo, N, table_rows = shape
    g = torch.Generator().manual_seed(C * N + rps)
    M, Cp = B * rps, (C + 3) // 4 * 4
    K = Cp + 12
    table = torch.zeros(table_rows, Cp + 4)                      # wider than Cp: a column slice of a wider buffer
    table[:, :C] = torch.randn(table_rows, C, generator=g)
    src = torch.randint(0, table_rows, (M,), generator=g, dtype=torch.int32)
    src[torch.rand(M, generator=g) < 0.05] = -1
    geo = torch.zeros(M, 12)
    geo[:, :n_geo] = torch.randn(M, n_geo, generator=g)
    W = tf32_round((torch.randn(N, K, generator=g) / K ** 0.5)).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    X = torch.zeros(M, K)
    ok = src >= 0
    X[ok, :Cp] = table[src[ok].long(), :Cp]
    X[:, Cp:] = geo
    table, src, geo, X = table.to(DEV), src.to(DEV), geo.to(DEV), X.to(DEV)
    C0, st0 = _run(cuda_lib, X, W, bias, B, rps, N, 0, None, None, None, None, None, 0, use_tf32=True)
    ldc = (N + 3) // 4 * 4
    Cg = torch.full((M, ldc), float("nan"), device=DEV)
    tiles = (rps + cuda_lib.pdr_gemm_tile_rows() - 1) // cuda_lib.pdr_gemm_tile_rows()
    stats = torch.zeros(B * tiles, N, 4, device=DEV)
    a = GemmArgs()
    a.A, a.lda, a.K = table.data_ptr(), table.stride(0), K
    a.W, a.ldw, a.bias = W.data_ptr(), W.stride(0), bias.data_ptr()
    a.C, a.ldc, a.N, a.ldc_zero_to = Cg.data_ptr(), ldc, N, ldc
    a.batch, a.rows_per_sample, a.pro_mode = B, rps, 0
    a.stats, a.use_tf32 = stats.data_ptr(), 1
    a.a_rows, a.A2, a.lda2, a.k_split = src.data_ptr(), geo.data_ptr(), 12, Cp
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    # table_rows = 0: the table part comes by cp.async pieces; > 0: whole 32-column chunks by TMA tile::gather4 (index -1 and
    # the rows beyond a ragged tile read as zeros through the out-of-bounds fill)
    for rows_known in (0, table_rows):
        a.table_rows = rows_known
        Cg.fill_(float("nan")); stats.zero_()
        assert cuda_lib.pdr_gemm_fused(ctypes.c_void_p(ctypes.addressof(a)), stream) == 0, cuda_lib.pdr_last_error_string()
        torch.cuda.synchronize()
        assert torch.equal(Cg, C0), rows_known
        assert torch.equal(stats.view(B, tiles, N, 4).sum(1), st0), rows_known
    a.use_tf32 = 0                                               # the fp32 SIMT path has no gathered operand
    assert cuda_lib.pdr_gemm_fused(ctypes.c_void_p(ctypes.addressof(a)), stream) == -4


def test_group_src_rows_and_geometric_channels(cuda_lib):
    """pdr_group_src_rows + pdr_group_ball(C = 0): together they carry exactly what the materialised grouping holds."""
    import ctypes
    from point_diffusion_refinement_b200 import _ext
    from point_diffusion_refinement_b200._lib import call, dptr, stream_ptr
    g = torch.Generator().manual_seed(4)
    B, n, P, K, C = 2, 300, 64, 16, 8
    xyz = (torch.rand(B, n, 3, generator=g) * 2 - 1).to(DEV)
    centres = (torch.rand(B, P, 3, generator=g) * 2 - 1).to(DEV)
    centres[0, 5] = 9.0                                          # a centre without neighbours
    feat = torch.randn(B * n, C, generator=g).to(DEV)
    idx, cnt = _ext.ball_query(centres, xyz, 0.4, K)
    assert cnt[0, 5] == 0
    for fill in (0, 1):
        full = torch.zeros(B * P * K, C + 12, device=DEV)
        call("pdr_group_ball", B, n, P, K, C, dptr(feat), C, dptr(xyz), dptr(centres), dptr(idx), dptr(cnt), fill, dptr(full),
             C + 12, stream_ptr(xyz))
        geo = torch.zeros(B * P * K, 12, device=DEV)
        call("pdr_group_ball", B, n, P, K, 0, None, 0, dptr(xyz), dptr(centres), dptr(idx), dptr(cnt), fill, dptr(geo), 12,
             stream_ptr(xyz))
        src = torch.empty(B * P * K, dtype=torch.int32, device=DEV)
        call("pdr_group_src_rows", B, n, P, K, dptr(idx), 0, dptr(cnt), fill, dptr(src), stream_ptr(xyz))
        assert torch.equal(geo[:, :9], full[:, C:C + 9]) and geo[:, 9:].abs().sum() == 0
        gathered = torch.where((src >= 0)[:, None], feat[src.clamp(min=0).long()], torch.zeros((), device=DEV))
        assert torch.equal(gathered, full[:, :C])
        assert (src.view(B, P, K)[0, 5] == -1).all() == bool(fill)


def test_group_geo_kernels_match_the_grouping_kernels(cuda_lib):
    """pdr_group_geo_ball / pdr_group_geo_knn (one pass, thread per row) against pdr_group_ball / pdr_group_knn with C = 0
    and pdr_group_src_rows: bit-identical."""
    from point_diffusion_refinement_b200 import _ext, knn
    from point_diffusion_refinement_b200._lib import call, dptr, stream_ptr
    g = torch.Generator().manual_seed(8)
    B, n, P, K = 3, 500, 200, 32
    xyz = (torch.rand(B, n, 3, generator=g) * 2 - 1).to(DEV)
    centres = (torch.rand(B, P, 3, generator=g) * 2 - 1).to(DEV)
    centres[1, 7] = 5.0
    idx, cnt = _ext.ball_query(centres, xyz, 0.3, K)
    rows = B * P * K
    for fill in (0, 1):
        ref = torch.zeros(rows, 12, device=DEV); src_ref = torch.empty(rows, dtype=torch.int32, device=DEV)
        call("pdr_group_ball", B, n, P, K, 0, None, 0, dptr(xyz), dptr(centres), dptr(idx), dptr(cnt), fill, dptr(ref), 12, stream_ptr(xyz))
        call("pdr_group_src_rows", B, n, P, K, dptr(idx), 0, dptr(cnt), fill, dptr(src_ref), stream_ptr(xyz))
        geo = torch.full((rows, 12), float("nan"), device=DEV); src = torch.empty(rows, dtype=torch.int32, device=DEV)
        call("pdr_group_geo_ball", B, n, P, K, dptr(xyz), dptr(centres), dptr(idx), dptr(cnt), fill, dptr(geo), dptr(src), 0, stream_ptr(xyz))
        assert torch.equal(geo, ref) and torch.equal(src, src_ref)
        # round_tf32: the same values rounded to the nearest TF32 number (what the tensor core would otherwise truncate)
        call("pdr_group_geo_ball", B, n, P, K, dptr(xyz), dptr(centres), dptr(idx), dptr(cnt), fill, dptr(geo), dptr(src), 1, stream_ptr(xyz))
        assert torch.equal(geo, ((ref.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32))
    Kn = 8
    kn = knn.knn_points(centres, xyz, K=Kn)
    rows = B * P * Kn
    ref = torch.zeros(rows, 12, device=DEV); src_ref = torch.empty(rows, dtype=torch.int32, device=DEV)
    call("pdr_group_knn", B, n, P, Kn, 0, None, 0, dptr(xyz), dptr(centres), dptr(kn.idx), dptr(kn.dists), dptr(ref), 12, stream_ptr(xyz))
    call("pdr_group_src_rows", B, n, P, Kn, dptr(kn.idx), 1, None, 0, dptr(src_ref), stream_ptr(xyz))
    geo = torch.full((rows, 12), float("nan"), device=DEV); src = torch.empty(rows, dtype=torch.int32, device=DEV)
    call("pdr_group_geo_knn", B, n, P, Kn, dptr(xyz), dptr(centres), dptr(kn.idx), dptr(kn.dists), dptr(geo), dptr(src), 0, stream_ptr(xyz))
    assert torch.equal(geo, ref) and torch.equal(src, src_ref)


@pytest.mark.parametrize("shape", [(2, 4096, 32, 32, 32, True), (2, 2048, 64, 128, 8, False), (3, 1024, 128, 300, 8, False),
                                   (2, 4000, 32, 64, 32, True), (2, 640, 256, 256, 16, True)])
def test_gemm_pooling_epilogue_equals_attention_pool(cuda_lib, shape):
    """PdrGemmArgs.pool_*: the score GEMM that pools in its epilogue gives bit for bit what the stored scores +
    pdr_attention_pool give (same tensor-core accumulators, same softmax operation order)."""
    import ctypes
    from point_diffusion_refinement_b200._lib import call, dptr, stream_ptr
    from point_diffusion_refinement_b200.fused import GemmArgs, tf32_round
    B, rps, K, N, PK, with_counts = shape
    g = torch.Generator().manual_seed(K * N + rps + PK)
    M, P = B * rps, rps // PK
    A = torch.randn(M, K, generator=g).to(DEV)
    W = tf32_round((torch.randn(N, K, generator=g) / K ** 0.5)).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    sc = (1 + 0.2 * torch.randn(B, K, generator=g)).to(DEV)
    sh = (0.2 * torch.randn(B, K, generator=g)).to(DEV)
    ldn = (N + 3) // 4 * 4
    V = torch.randn(M, ldn, generator=g).to(DEV)
    vsc = (1 + 0.2 * torch.randn(B, ldn, generator=g)).to(DEV)
    vsh = (0.2 * torch.randn(B, ldn, generator=g)).to(DEV)
    counts = torch.randint(0, PK + 1, (B * P,), generator=g, dtype=torch.int32).to(DEV) if with_counts else None
    S, _ = _run(cuda_lib, A, W, bias, B, rps, N, 2, sc, sh, None, None, None, 0, use_tf32=True, want_stats=False)
    ref = torch.zeros(B * P, ldn + 4, device=DEV)
    call("pdr_attention_pool", B, P, PK, N, dptr(S), S.stride(0), dptr(V), ldn, dptr(vsc), dptr(vsh), ldn, dptr(counts),
         dptr(ref), ldn + 4, 0, stream_ptr(S))
    out = torch.zeros(B * P, ldn + 4, device=DEV)
    a = GemmArgs()
    a.A, a.lda, a.K = A.data_ptr(), A.stride(0), K
    a.W, a.ldw, a.bias = W.data_ptr(), W.stride(0), bias.data_ptr()
    a.C, a.ldc, a.N, a.ldc_zero_to = None, ldn, N, N
    a.batch, a.rows_per_sample, a.pro_mode = B, rps, 2
    a.sc, a.sh, a.ld_scsh = sc.data_ptr(), sh.data_ptr(), sc.stride(0)
    a.use_tf32 = 1
    a.pool_K, a.pool_V, a.pool_ldv = PK, V.data_ptr(), ldn
    a.pool_sc, a.pool_sh, a.pool_ld_scsh = vsc.data_ptr(), vsh.data_ptr(), ldn
    a.pool_counts = counts.data_ptr() if counts is not None else None
    a.pool_out, a.pool_ldo = out.data_ptr(), ldn + 4
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert cuda_lib.pdr_gemm_fused(ctypes.c_void_p(ctypes.addressof(a)), stream) == 0, cuda_lib.pdr_last_error_string()
    torch.cuda.synchronize()
    assert torch.isfinite(ref).all() and torch.equal(out, ref)
    a.use_tf32 = 0
    assert cuda_lib.pdr_gemm_fused(ctypes.c_void_p(ctypes.addressof(a)), stream) == -4


@pytest.mark.parametrize("shape", [(2, 4096, 64, 35, 9, 64, 2000, True), (3, 1000, 32, 4, 9, 32, 300, False),
                                   (2, 2048, 128, 160, 11, 128, 512, True), (1, 8320, 256, 320, 11, 300, 64, True)])
def test_gemm_raw_gathered_tail(cuda_lib, shape):
    """PdrGemmArgs.tail_rows: C = pro(A[:, :k_pro]) W1^T + [T[rows] | T2] W2^T + bias against float64, TF32 tolerance;
    deterministic; the fp32 path refuses."""
    import ctypes
    from point_diffusion_refinement_b200.fused import GemmArgs, tf32_round
    B, rps, k_pro, C, n_geo, N, table_rows, use_add = shape
    g = torch.Generator().manual_seed(k_pro * N + rps)
    M, Cp = B * rps, (C + 3) // 4 * 4
    K = k_pro + Cp + 12
    A = torch.randn(M, k_pro, generator=g)
    table = torch.zeros(table_rows, Cp)
    table[:, :C] = torch.randn(table_rows, C, generator=g)
    src = torch.randint(0, table_rows, (M,), generator=g, dtype=torch.int32)
    src[torch.rand(M, generator=g) < 0.05] = -1
    geo = torch.zeros(M, 12)
    geo[:, :n_geo] = torch.randn(M, n_geo, generator=g)
    W = tf32_round(torch.randn(N, K, generator=g) / K ** 0.5)
    bias = torch.randn(N, generator=g)
    sc = 1 + 0.2 * torch.randn(B, k_pro, generator=g)
    sh = 0.2 * torch.randn(B, k_pro, generator=g)
    add = torch.randn(B, k_pro, generator=g) if use_add else None
    x = torch.relu(A.double().view(B, rps, k_pro) * sc.double()[:, None] + sh.double()[:, None])
    if add is not None:
        x = x + add.double()[:, None]
    tail = torch.zeros(M, Cp + 12, dtype=torch.float64)
    ok = src >= 0
    tail[ok, :Cp] = table[src[ok].long()].double()
    tail[:, Cp:] = geo.double()
    y64 = x.reshape(M, k_pro) @ W.double()[:, :k_pro].t() + tail @ W.double()[:, k_pro:].t() + bias.double()
    A, table, src, geo, W, bias, sc, sh = [t.to(DEV) for t in (A, table, src, geo, W, bias, sc, sh)]
    add = add.to(DEV) if add is not None else None
    ldc = (N + 3) // 4 * 4
    tiles = (rps + cuda_lib.pdr_gemm_tile_rows() - 1) // cuda_lib.pdr_gemm_tile_rows()

    def run():
        Cg = torch.full((M, ldc), float("nan"), device=DEV)
        stats = torch.zeros(B * tiles, N, 4, device=DEV)
        a = GemmArgs()
        a.A, a.lda, a.K = A.data_ptr(), A.stride(0), K
        a.W, a.ldw, a.bias = W.data_ptr(), W.stride(0), bias.data_ptr()
        a.C, a.ldc, a.N, a.ldc_zero_to = Cg.data_ptr(), ldc, N, ldc
        a.batch, a.rows_per_sample, a.pro_mode = B, rps, 1
        a.sc, a.sh, a.ld_scsh = sc.data_ptr(), sh.data_ptr(), sc.stride(0)
        if add is not None:
            a.add, a.ld_add = add.data_ptr(), add.stride(0)
        a.stats, a.use_tf32 = stats.data_ptr(), 1
        a.tail_rows, a.T, a.ldt = src.data_ptr(), table.data_ptr(), table.stride(0)
        a.T2, a.ldt2, a.t_split, a.k_pro = geo.data_ptr(), 12, Cp, k_pro
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        rc = cuda_lib.pdr_gemm_fused(ctypes.c_void_p(ctypes.addressof(a)), stream)
        torch.cuda.synchronize()
        a.use_tf32 = 0
        rc32 = cuda_lib.pdr_gemm_fused(ctypes.c_void_p(ctypes.addressof(a)), stream)
        return rc, rc32, Cg, stats.view(B, tiles, N, 4).sum(1)

    rc, rc32, C1, st1 = run()
    assert rc == 0 and rc32 == -4, cuda_lib.pdr_last_error_string()
    torch.testing.assert_close(C1[:, :N].double().cpu(), y64, rtol=5e-3, atol=5e-3 * float(y64.abs().max()) / 4)
    assert C1[:, N:].abs().sum() == 0
    own = C1[:, :N].double().view(B, rps, N)
    l1 = own.abs().sum(1)
    assert ((st1[..., 0].double() - own.sum(1)).abs() <= 1e-5 * l1 + 1e-2).all()
    _, _, C2, _ = run()
    assert torch.equal(C1, C2)
