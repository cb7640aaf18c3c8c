"""Times the grouped tcgen05 GEMM (k_gemm_tc through dclip_linear_rows) at a few shapes: asymptotic rate vs the
teacher-step shapes.  Usage: python tools/bench_gemm.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dclip_b200 import assembly as A

dev = torch.device("cuda:0")
for n, k, m in ((3616, 768, 768), (3616, 768, 2304), (2464, 768, 1536), (16384, 768, 2304), (16384, 4096, 4096), (65536, 768, 768)):
    x = torch.randn(n, k, device=dev).bfloat16()
    w = torch.randn(m, k, device=dev).bfloat16()
    for _ in range(3):
        y = A.linear_rows(x, w)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        y = A.linear_rows(x, w)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    ref = (x.float() @ w.float().T)
    rel = float((y.float() - ref).norm() / ref.norm())
    tiles = ((n + 127) // 128) * ((m + 255) // 256)
    print(f"n={n} k={k} m={m}: {us:8.1f} us  {2.0 * n * k * m / us / 1e6:7.1f} TFLOP/s  tiles={tiles}  rel_err={rel:.2e}", flush=True)
