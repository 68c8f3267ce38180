"""Two gathered-A GEMMs of the last grouped stage (2 M rows: 32 samples x 2048 points x 32 neighbours, table = 32 feature columns
of the 65536 level-0 points) launched 3 times each, for ncu (capture the third launch):  N = 76 with statistics (first | res | key),
N = 32 without.  PDR_GEMM_TMA_GATHER=0/1 selects cp.async pieces / TMA gather4 for the table chunk."""
import ctypes
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from point_diffusion_refinement_b200 import _lib
from point_diffusion_refinement_b200.fused import GemmArgs, tf32_round

lib = _lib.lib()
dev = torch.device("cuda:0")
B, P, Kn, Cp = 32, 2048, 32, 32
rps, M, K = P * Kn, B * P * Kn, Cp + 12
g = torch.Generator().manual_seed(0)
table = tf32_round(torch.randn(B * P, Cp, generator=g)).to(dev)
# neighbours of a point: a window of nearby indices inside the same sample (ball query after FPS ordering is not local; use random)
src = (torch.randint(0, P, (B, P, Kn), generator=g) + torch.arange(B).view(B, 1, 1) * P).to(torch.int32).reshape(-1).to(dev)
geo = torch.zeros(M, 12, device=dev)
geo[:, :9] = torch.randn(M, 9, device=dev)
timings = {}
for N, want_stats in ((76, True), (32, False)):
    W = tf32_round(torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    ldc = (N + 3) // 4 * 4
    C = torch.empty(M, ldc, device=dev)
    tiles = rps // 128
    stats = torch.zeros(B * tiles, N, 4, device=dev)
    a = GemmArgs()
    a.A, a.lda, a.K = table.data_ptr(), table.stride(0), K
    a.W, a.ldw, a.bias = W.data_ptr(), W.stride(0), bias.data_ptr()
    a.C, a.ldc, a.N, a.ldc_zero_to = C.data_ptr(), ldc, N, ldc
    a.batch, a.rows_per_sample, a.pro_mode, a.use_tf32, a.w_static = B, rps, 0, 1, 1
    a.stats = stats.data_ptr() if want_stats else None
    a.a_rows, a.A2, a.lda2, a.k_split, a.table_rows = src.data_ptr(), geo.data_ptr(), 12, Cp, B * P
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(3):
        e0.record()
        assert lib.pdr_gemm_fused(ctypes.c_void_p(ctypes.addressof(a)), stream) == 0, lib.pdr_last_error_string()
        e1.record()
    torch.cuda.synchronize()
    print("N=%d: %.1f us" % (N, e0.elapsed_time(e1) * 1e3))
