"""Run a few representative fused-GEMM shapes (for ncu) and print CUDA-event timings."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from point_diffusion_refinement_b200 import _lib
from tests.test_gemm_gpu import _run
lib = _lib.lib()
dev = "cuda"
shapes = [  # B, rps, K, N, pro, label
    (32, 65536, 16, 80, 0, "enc_map0.gemm1"), (32, 65536, 32, 32, 1, "enc_map0.second"),
    (32, 65536, 52, 116, 0, "dec_map0.gemm1"), (32, 32768, 44, 140, 0, "sa0.gemm1"),
    (32, 32768, 64, 64, 1, "sa0.v"), (32, 16384, 184, 440, 0, "fp0.mlp1.gemm1"), (32, 16384, 128, 128, 1, "fp0.v"),
]
g = torch.Generator().manual_seed(0)
for (B, rps, K, N, pro, label) in shapes:
    M = B * rps
    A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) / K ** 0.5; bias = torch.randn(N, device=dev)
    sc = torch.ones(B, K, device=dev); sh = torch.zeros(B, K, device=dev)
    for tf32 in (1, 0):
        _run(lib, A, W, bias, B, rps, N, pro, sc, sh, None, None, None, 0, tf32)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(3):
            _run(lib, A, W, bias, B, rps, N, pro, sc, sh, None, None, None, 0, tf32)
        e.record(); torch.cuda.synchronize()
        # _run allocates C/stats each call; subtract nothing, report upper bound
        ms = s.elapsed_time(e) / 3
        byts = 4 * (M * K + M * ((N + 3) // 4 * 4))
        print("%-18s tf32=%d M=%d K=%d N=%d  %.3f ms  %.0f GB/s (incl. alloc+fill of C)" % (label, tf32, M, K, N, ms, byts / ms / 1e6))
    del A, W
