"""Phase clocks of k_distill_small (stall-accounting build: make -C dclip_b200/csrc prof; DCLIP_B200_LIB=build/prof/libdclip_b200_prof.so)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dclip_b200 import _lib, losses as L
dev = torch.device("cuda:0")
B, d = 32, 768
g = torch.randn(B, d, device=dev).bfloat16(); x = torch.randn(B, d, device=dev).bfloat16()
for _ in range(5):
    L.distill_loss_fwd_bwd(g, x, path="fp32")
torch.cuda.synchronize()
lib = _lib.load_library()
buf = (ctypes.c_longlong * 16)()
lib.dclip_debug_small_prof.restype = ctypes.c_int
lib.dclip_debug_small_prof(buf)
t = list(buf)
names = ["load", "partial S + norms", "cluster.sync 1", "exchange 1", "inv + logits", "lse", "losses", "G + normalise", "grad slice", "proj dots", "cluster.sync 2", "exchange 2 + store", "cluster.sync 3"]
for i in range(13):
    print(f"{names[i]:22s} {t[i + 1] - t[i]:8d} cycles")
print("total", t[13] - t[0])

