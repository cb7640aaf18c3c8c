"""Experiment: the grouped tcgen05 GEMM (k_gemm_tc) and cuBLAS (torch.matmul) at the shapes of a materialised-G backward
(dU = G V, dV = G^T U: M = 4096, N = 768, K = 4096), graph-replayed, one L2-warm number each."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dclip_b200 import assembly as A

dev = torch.device("cuda:0")

def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

for n, k, m in ((4096, 4096, 768), (8192, 4096, 768), (4096, 768, 4096)):
    x = torch.randn(n, k, device=dev).half()
    w = torch.randn(m, k, device=dev).half()
    wt = w.t().contiguous()
    t_ours = timed(lambda: A.linear_rows(x, w))
    t_cublas = timed(lambda: torch.matmul(x, wt))
    fl = 2.0 * n * k * m
    print(f"n={n} k={k} m={m}: k_gemm_tc {t_ours:7.1f} us ({fl / t_ours / 1e6:6.1f} TFLOP/s)   cuBLAS {t_cublas:7.1f} us ({fl / t_cublas / 1e6:6.1f} TFLOP/s)", flush=True)
