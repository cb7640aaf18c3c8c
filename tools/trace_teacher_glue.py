"""Which torch ops launch the small at:: kernels between the loss kernel and k_global_bwd of the meta-teacher step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from dclip_b200 import losses as L
from dclip_b200.patch_text_aggregation import PatchTextAggregation
dev = torch.device("cuda:0")
B, T, R, d = 32, 77, 36, 768
m = PatchTextAggregation(embed_dim=d, num_heads=8).to(dev).to(torch.bfloat16)
g = torch.Generator().manual_seed(0)
tg = torch.randn(B, T, d, generator=g).to(dev).bfloat16(); ig = torch.randn(B, R, d, generator=g).to(dev).bfloat16()
xg = torch.randn(B, d, generator=g).to(dev).bfloat16()
def step():
    for p in m.parameters():
        p.grad = None
    L.compute_contrastive_loss(m.global_embedding(tg, ig), xg).backward()
for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    step()
    torch.cuda.synchronize()
evs = sorted(prof.events(), key=lambda e: e.time_range.start)
for e in evs:
    if str(e.device_type).endswith("CPU") and (e.name.startswith("aten::") or "Backward" in e.name or "Fn" in e.name):
        print(f"{e.time_range.start:12.1f} {e.name[:60]:60s} {str(e.input_shapes)[:80]}")
