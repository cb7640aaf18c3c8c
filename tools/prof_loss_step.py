"""Per-kernel WARM device times of the single-GPU distill-loss step (CUPTI via torch.profiler), L2 flushed before each step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from dclip_b200 import losses as L
import workloads as synthetic
dev = torch.device("cuda:0")
B, d = 4096, 768
x = [t.to(torch.bfloat16).to(dev) for t in synthetic.loss_inputs(B, d, seed=42)]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
flush_rd = torch.ones(64 * 1024 * 1024, dtype=torch.float32, device=dev)      # read after the write: L2 ends up cold AND clean (as in bench.py)
for _ in range(5):
    L.distill_loss_fwd_bwd(*x, path="tensor")
torch.cuda.synchronize()
n = 10
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(n):
        flush.zero_()
        flush_rd.sum()
        L.distill_loss_fwd_bwd(*x, path="tensor")
    torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0 and "Fill" not in e.key and "Memset" not in e.key and "reduce_kernel" not in e.key]
rows.sort(key=lambda r: -r[2])
tot = sum(r[2] for r in rows)
print(f"# warm per-kernel device time, distill-loss step B={B} d={d} bf16, L2 cold and clean before every step; sum of kernels {tot / n:.1f} us/step")
print("share%  us/step  launches/step  avg_us  kernel")
for k, c, t in rows:
    print(f"{100 * t / tot:6.1f} {t / n:8.1f} {c / n:8.1f} {t / c:8.2f}  {k[:110]}")
