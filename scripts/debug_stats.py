"""Prints the TF32-vs-fp32 statistics deviations of the fused GEMM for a few test cases (debug aid)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import test_gemm_gpu as T
from point_diffusion_refinement_b200 import _lib
lib = _lib.lib()
DEV = "cuda"
for ci in (8, 5, 9, 11):
    B, rps, K, N, pro, use_add, use_R, div = T.CASES[ci]
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
    C0, st0 = T._run(lib, A, W, bias, B, rps, N, pro, sc, sh, add, R, rowadd, div, use_tf32=False)
    C1, st1 = T._run(lib, A, W, bias, B, rps, N, pro, sc, sh, add, R, rowadd, div, use_tf32=True)
    l1 = C0[:, :N].abs().view(B, rps, N).sum(1)
    d0 = (st1[..., 0] - st0[..., 0]).abs()
    # stats recomputed from the stored C1: separates "stats kernel wrong" from "TF32 drift"
    re0 = C1[:, :N].double().view(B, rps, N).sum(1)
    print(ci, "max|dsum|", d0.max().item(), "max ratio to L1", (d0 / l1).max().item(),
          "stats-vs-own-C", (st1[..., 0].double() - re0).abs().max().item(),
          "maxabs C diff", (C1[:, :N] - C0[:, :N]).abs().max().item(),
          "mean signed C diff", (C1[:, :N] - C0[:, :N]).mean().item())
    idx = (d0 / l1).argmax().item()
    b, n = divmod(idx, N)
    print("   worst at b,n", b, n, "st1", st1[b, n].tolist(), "st0", st0[b, n].tolist())
