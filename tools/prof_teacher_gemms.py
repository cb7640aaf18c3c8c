"""Phase clocks (CTA 0) of the four k_gemm_tc launches of one meta-teacher step (stall-accounting build:
make -C dclip_b200/csrc prof; DCLIP_B200_LIB=build/prof/libdclip_b200_prof.so)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dclip_b200 import _lib, losses as L
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
    L.compute_contrastive_loss(m.global_embedding(tg, ig), xg, path="fp32").backward()
for _ in range(4):        # 16 launches: the ring holds the last two steps
    step()
torch.cuda.synchronize()
lib = _lib.load_library()
buf = (ctypes.c_longlong * 128)()
lib.dclip_debug_gemm_prof_ring.restype = ctypes.c_int
assert lib.dclip_debug_gemm_prof_ring(buf) == 0
names = ["setup", "first TMA issued", "first stage landed", "last stage of tile 0", "acc ready", "tile 0 stored", "exit"]
what = ["in-projections (261 tiles)", "out-projections + residual, fp32 Z (87 tiles)", "dO (2 problems)", "dWq, dWkv, dWo (6 problems)"]
for s in range(4):
    t = list(buf[s * 16:(s + 1) * 16])
    print(f"launch {s}: {what[s]}")
    print("   " + "  ".join(f"{names[i]} +{t[i + 1] - t[i]}" for i in range(7)) + f"   total {t[7] - t[0]} cycles")
