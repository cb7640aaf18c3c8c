"""Runs N eager meta-teacher steps (configs[2] shape by default, `N fp32` = configs[0]) - the target of ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dclip_b200 import losses as L
from dclip_b200.patch_text_aggregation import PatchTextAggregation
import workloads as synthetic

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
fp32 = len(sys.argv) > 2 and sys.argv[2] == "fp32"            # fp32: BASELINE configs[0] (B=32, T=77, R=8, d=512)
dtype = torch.float32 if fp32 else torch.bfloat16
B, T, R, d = (32, 77, 8, 512) if fp32 else (32, 77, 36, 768)
dev = torch.device("cuda:0")
m = PatchTextAggregation(embed_dim=d, num_heads=8)
m.cross_modal_attention.load_state_dict(synthetic.perturbed_cross_attention_params(d, 8, 3))
m = m.to(dev).to(dtype)
text, image = synthetic.xattn_inputs(B, T, R, d, 3, ragged=True)
txt = torch.randn(B, d, generator=torch.Generator().manual_seed(1))
tg, ig, xg = text.to(dev).to(dtype), image.to(dev).to(dtype), txt.to(dev).to(dtype)
for _ in range(steps):
    for p in m.parameters():
        p.grad = None
    loss = L.compute_contrastive_loss(m.global_embedding(tg, ig), xg, path="fp32")
    loss.backward()
torch.cuda.synchronize()
print("loss", float(loss))
