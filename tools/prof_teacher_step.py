"""Per-kernel WARM device times of the meta-teacher step (CUPTI through torch.profiler; eager launches, steady state).
Complements the ncu launch list (cold cache, serialised).  Usage: python tools/prof_teacher_step.py [bf16|fp32] [B T R d]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from dclip_b200 import losses as L
from dclip_b200.patch_text_aggregation import PatchTextAggregation
import workloads as synthetic

dtype = torch.bfloat16 if (len(sys.argv) < 2 or sys.argv[1] == "bf16") else torch.float32
B, T, R, d = (int(x) for x in sys.argv[2:6]) if len(sys.argv) >= 6 else (32, 77, 36, 768)
dev = torch.device("cuda:0")
m = PatchTextAggregation(embed_dim=d, num_heads=8)
m.cross_modal_attention.load_state_dict(synthetic.perturbed_cross_attention_params(d, 8, 3))
m = m.to(dev).to(dtype)
text, image = synthetic.xattn_inputs(B, T, R, d, 3, ragged=True)
txt = torch.randn(B, d, generator=torch.Generator().manual_seed(1))
tg, ig, xg = text.to(dev).to(dtype), image.to(dev).to(dtype), txt.to(dev).to(dtype)

def step():
    for p in m.parameters():
        p.grad = None
    loss = L.compute_contrastive_loss(m.global_embedding(tg, ig), xg, path="fp32")
    loss.backward()

for _ in range(5):
    step()
torch.cuda.synchronize()
n = 10
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(n):
        step()
    torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0]
rows.sort(key=lambda r: -r[2])
tot = sum(r[2] for r in rows)
print(f"# warm per-kernel device time, teacher step B={B} T={T} R={R} d={d} {dtype}; {n} steps; sum of kernels {tot / n:.1f} us/step")
print("share%  us/step  launches/step  avg_us  kernel")
for k, c, t in rows:
    print(f"{100 * t / tot:6.1f} {t / n:8.1f} {c / n:8.1f} {t / c:8.2f}  {k[:110]}")

# one step in launch order (the last of the profiled steps): which of the same-named launches is the slow one
evs = sorted((e for e in prof.events() if e.device_time_total > 0 and getattr(e, "device_type", None) is not None and "CUDA" in str(e.device_type)),
             key=lambda e: e.time_range.start)
per = len(evs) // n
print(f"# launch order, last step ({per} launches)")
for e in evs[-per:]:
    print(f"{e.device_time_total:8.2f} us  {e.name[:100]}")
