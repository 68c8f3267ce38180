"""Times pdr_gemm_fused on short GEMMs (few row tiles) as a function of K, N and the row count: is the floor the launch,
the K loop (per-chunk latency) or the epilogue?  Usage (GPU box): python scripts/gemm_small_probe.py"""
import ctypes
import torch
from point_diffusion_refinement_b200 import _lib
from point_diffusion_refinement_b200.fused import GemmArgs

lib = _lib.lib()
dev = torch.device("cuda:0")


def time_gemm(B, rps, K, N, reps=50):
    M = B * rps
    A = torch.randn(M, K, device=dev)
    W = torch.randn(N, K, device=dev)
    bias = torch.randn(N, device=dev)
    C = torch.empty(M, N, device=dev)
    g = GemmArgs()
    p = lambda t: t.data_ptr()
    g.A, g.lda, g.W, g.ldw, g.bias = p(A), K, p(W), K, p(bias)
    g.C, g.ldc, g.N, g.ldc_zero_to, g.K = p(C), N, N, N, K
    g.batch, g.rows_per_sample, g.use_tf32, g.w_static = B, rps, 1, 1
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(5):
        assert lib.pdr_gemm_fused(ctypes.byref(g), ctypes.c_void_p(s)) == 0, lib.pdr_last_error_string()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(reps):
            lib.pdr_gemm_fused(ctypes.byref(g), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); graph.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for B, rps, K, N in [(1, 32, 512, 1248), (1, 32, 64, 1248), (1, 32, 512, 32), (1, 128, 512, 1248), (1, 32, 512, 256),
                     (32, 16, 512, 512), (32, 64, 256, 256), (32, 256, 256, 256), (32, 512, 512, 512), (32, 512, 844, 512),
                     (32, 2048, 256, 256), (32, 2048, 332, 588)]:
    print("B=%d rows/sample=%d K=%d N=%d: %.1f us per GEMM (back to back in a graph)" % (B, rps, K, N, time_gemm(B, rps, K, N)))
