"""Stall accounting of the two streaming kernels of the single-GPU loss step (prof build: make -C dclip_b200/csrc prof):
cycles CTA (0,0,0) spends per role, forward (mode 1) and gradient-coefficient (mode 3) kernels."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("DCLIP_B200_LIB", os.path.join(ROOT, "build", "prof", "libdclip_b200_prof.so"))
import torch
from dclip_b200 import _lib, losses as L
import workloads as synthetic
lib = _lib.load_library()
dev = torch.device("cuda:0")
B, d = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 768)
x = [t.to(torch.bfloat16).to(dev) for t in synthetic.loss_inputs(B, d, seed=42)]
for _ in range(3):
    L.distill_loss_fwd_bwd(*x, path="tensor")
buf = (ctypes.c_ulonglong * 16)()
lib.dclip_debug_stream_prof(buf, 1)
L.distill_loss_fwd_bwd(*x, path="tensor")
lib.dclip_debug_stream_prof(buf, 1)
v = list(buf)
names = ["MMA issuer: whole loop", "MMA issuer: wait s_empty (epilogue behind)", "MMA issuer: wait full (TMA behind)", "epilogue warp 0: wait s_full (MMA behind)",
         "epilogue warp 0: tile work", "producer 0: wait empty", "setup (barriers, TMEM alloc, cluster sync)", "kernel entry -> exit"]
for m, off in (("forward, mode 1", 0), ("gradient coefficients, mode 3", 8)):
    print(f"# {m}, B={B} d={d}, CTA (0,0,0)")
    for i, nm in enumerate(names):
        print(f"  {nm:48s} {v[off + i]:>9d} cyc")
