"""Raw pinned host-to-device bandwidth at the e2e step's copy sizes, next to HostDistillStep's step time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dclip_b200 import losses as L
dev = torch.device("cuda:0")
B, d = 4096, 768
hs = [torch.randn(B, d).bfloat16().pin_memory() for _ in range(4)]
ds = [torch.empty(B, d, dtype=torch.bfloat16, device=dev) for _ in range(4)]
big_h = torch.empty(4 * B * d, dtype=torch.bfloat16).pin_memory(); big_d = torch.empty(4 * B * d, dtype=torch.bfloat16, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def timeit(fn, n=30):
    fn(); torch.cuda.synchronize()
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
t4 = timeit(lambda: [d_.copy_(h_, non_blocking=True) for d_, h_ in zip(ds, hs)])
t1 = timeit(lambda: big_d.copy_(big_h, non_blocking=True))
mb = 4 * B * d * 2 / 1e6
print(f"4 x {mb/4:.1f} MB copies: {t4*1e3:.1f} us = {mb/t4/1e3:.1f} GB/s;  one {mb:.1f} MB copy: {t1*1e3:.1f} us = {mb/t1/1e3:.1f} GB/s")
step = L.HostDistillStep(B, d, torch.bfloat16, dev, path="tensor")
for _ in range(3): step(*hs)
t0 = time.perf_counter()
for _ in range(30): step(*hs)
print(f"HostDistillStep: {(time.perf_counter() - t0) / 30 * 1e6:.1f} us/step (wall, includes the per-step host sync)")
