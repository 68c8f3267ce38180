import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from point_diffusion_refinement_b200 import _lib
from point_diffusion_refinement_b200.fused import tf32_round
from tests.test_gemm_gpu import _run
lib = _lib.lib(); dev = "cuda"
for (B, rps, K, N, pro) in [(32, 16384, 172, 128, 1), (32, 16384, 172, 428, 0)]:
    M = B * rps
    A = torch.randn(M, K, device=dev); W = tf32_round(torch.randn(N, K, device=dev) / K ** 0.5); bias = torch.randn(N, device=dev)
    sc = torch.ones(B, K, device=dev); sh = torch.zeros(B, K, device=dev)
    for it in range(3):
        _run(lib, A, W, bias, B, rps, N, pro, sc, sh, None, None, None, 0, 1)
    del A, W
