"""Timeline of HostDistillStep (host buffers -> losses on the host): copies, kernels and the gaps between them."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from dclip_b200 import losses as L
dev = torch.device("cuda:0")
B, d = 4096, 768
hs = [torch.randn(B, d).bfloat16().pin_memory() for _ in range(4)]
step = L.HostDistillStep(B, d, torch.bfloat16, dev, path="tensor")
for _ in range(5): step(*hs)
t0 = time.perf_counter()
for _ in range(50): step(*hs)
print(f"{(time.perf_counter() - t0) / 50 * 1e6:.1f} us/step wall")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step(*hs)
evs = sorted((e for e in prof.events() if "CUDA" in str(getattr(e, "device_type", "")) and e.device_time_total > 0), key=lambda e: e.time_range.start)
n = len(evs) // 3
last = evs[-n:]
t0 = last[0].time_range.start
prev_first = evs[-2 * n].time_range.start
print(f"# one step: device activities (start_us dur_us name); previous step started {t0 - prev_first:.1f} us earlier")
for e in last:
    print(f"{e.time_range.start - t0:8.1f} {e.device_time_total:8.2f}  {e.name[:80]}")
