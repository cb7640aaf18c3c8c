"""Graph-replayed meta-teacher step (configs[2] shape): back-to-back replays (warm) vs. L2 flushed before each, and the per-kernel
device times inside the replay."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from dclip_b200 import losses as L
from dclip_b200.patch_text_aggregation import GraphedTeacherStep, PatchTextAggregation
import workloads as synthetic
dev = torch.device("cuda:0")
fp32 = len(sys.argv) > 1 and sys.argv[1] == "fp32"
B, T, R, d, dtype = (32, 77, 8, 512, torch.float32) if fp32 else (32, 77, 36, 768, torch.bfloat16)
mod = PatchTextAggregation(embed_dim=d, num_heads=8)
mod.cross_modal_attention.load_state_dict(synthetic.perturbed_cross_attention_params(d, 8, 42))
mod = mod.to(dev).to(dtype)
text, image = synthetic.xattn_inputs(B, T, R, d, 42, ragged=True)
txt = torch.randn(B, d, generator=torch.Generator().manual_seed(43))
inputs = [text.to(dev).to(dtype), image.to(dev).to(dtype), txt.to(dev).to(dtype)]
gs = GraphedTeacherStep(mod, lambda m, t, i, x: L.compute_contrastive_loss(m.global_embedding(t, i), x, temperature=0.05), inputs)
for _ in range(5):
    gs(*inputs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    gs(*inputs)
e1.record(); torch.cuda.synchronize()
print(f"back-to-back replays: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us/step")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
flush_rd = torch.zeros(32 * 1024 * 1024, dtype=torch.int64, device=dev)
tot = 0.0
for _ in range(20):
    flush.zero_(); flush_rd.sum(); torch.cuda.synchronize()
    e0.record(); gs(*inputs); e1.record(); torch.cuda.synchronize()
    tot += e0.elapsed_time(e1)
print(f"L2 flushed before each replay: {tot / 20 * 1e3:.1f} us/step")
for cold in (False, True):
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            if cold:
                flush.zero_(); flush_rd.sum()
            gs(*inputs)
        torch.cuda.synchronize()
    evs = sorted((e for e in prof.events() if e.device_time_total > 0 and "CUDA" in str(getattr(e, "device_type", ""))), key=lambda e: e.time_range.start)
    evs = [e for e in evs if "FillFunctor<unsigned char>" not in e.name and "reduce_kernel" not in e.name]
    per = len(evs) // 5
    last = evs[-per:]
    t0 = last[0].time_range.start
    print(f"# {'cold' if cold else 'warm'} replay, kernels in order: start_us dur_us name   (span {last[-1].time_range.end - t0:.1f} us, sum {sum(e.device_time_total for e in last):.1f} us)")
    for e in last:
        print(f"{e.time_range.start - t0:8.1f} {e.device_time_total:7.2f}  {e.name[:70]}")
