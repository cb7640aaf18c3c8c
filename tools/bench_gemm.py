"""Times the grouped tcgen05 GEMM (k_gemm_tc through dclip_linear_rows) at a few shapes, replayed from a CUDA graph
(no host launch overhead): (a) back to back, (b) alternating with a kernel that uses no dynamic shared memory
(does a shared-memory carve-out switch between launches cost device time?)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dclip_b200 import assembly as A

dev = torch.device("cuda:0")
small = torch.zeros(1024, device=dev)
for n, k, m in ((3616, 768, 768), (3616, 768, 2304), (16384, 768, 2304), (16384, 4096, 4096)):
    x = torch.randn(n, k, device=dev).bfloat16()
    w = torch.randn(m, k, device=dev).bfloat16()
    for _ in range(3):
        y = A.linear_rows(x, w)
    torch.cuda.synchronize()
    res = []
    for alt in (False, True):
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            A.linear_rows(x, w)
            small.add_(1.0)
        torch.cuda.current_stream().wait_stream(s)
        reps = 20
        with torch.cuda.graph(g):
            for _ in range(reps):
                y = A.linear_rows(x, w)
                if alt:
                    small.add_(1.0)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / reps * 1e3)
    tiles = ((n + 127) // 128) * ((m + 255) // 256)
    print(f"n={n} k={k} m={m} tiles={tiles}: back-to-back {res[0]:7.1f} us ({2.0 * n * k * m / res[0] / 1e6:6.1f} TFLOP/s)   alternating with a tiny elementwise kernel {res[1]:7.1f} us per pair", flush=True)
