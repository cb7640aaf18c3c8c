"""Small driver for profiling the tensor retrieval path (used under ncu)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dclip_b200 import retrieval as R

ng = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
dev = torch.device("cuda:0")
g = torch.randn(ng, 768, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
g = (g / g.norm(dim=1, keepdim=True)).half()
q = g[torch.randperm(ng, device=dev)[:nq]].clone()
gal = R.TensorGallery(g)
for _ in range(2):
    s, i = R.topk_ip(q, g, 10, path="tensor", gallery=gal)
torch.cuda.synchronize()
t0 = time.perf_counter()
s, i = R.topk_ip(q, g, 10, path="tensor", gallery=gal)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"ng={ng} nq={nq}: {dt*1e3:.2f} ms, {2*nq*ng*768/dt/1e12:.1f} TFLOP/s")

from dclip_b200 import _lib
lib = _lib.load_library()
if hasattr(lib, "dclip_debug_stream_prof"):
    import ctypes
    buf = (ctypes.c_ulonglong * 16)()
    lib.dclip_debug_stream_prof(buf, 1)
    s, i = R.topk_ip(q, g, 10, path="tensor", gallery=gal)
    torch.cuda.synchronize()
    lib.dclip_debug_stream_prof(buf, 1)
    v = list(buf)
    print(f"CTA(0,0,0): mma total {v[0]}, wait s_empty {v[1]}, wait full {v[2]} | epilogue warp0 lane0: wait s_full {v[3]}, work {v[4]} | producer0 wait empty {v[5]}")
