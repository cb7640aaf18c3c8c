"""Phase clocks of k_gemm_tc, CTA 0 (stall-accounting build)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dclip_b200 import _lib, assembly as A
dev = torch.device("cuda:0")
lib = _lib.load_library()
lib.dclip_debug_gemm_prof.restype = ctypes.c_int
for n, k, m in ((3616, 768, 768), (3616, 768, 2304)):
    x = torch.randn(n, k, device=dev).bfloat16(); w = torch.randn(m, k, device=dev).bfloat16()
    for _ in range(5):
        A.linear_rows(x, w)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 16)()
    lib.dclip_debug_gemm_prof(buf)
    t = list(buf)
    names = ["setup (barriers, TMEM alloc, sync)", "producer: first TMA issued", "MMA: first stage landed", "MMA: last stage of tile 0 landed", "epilogue: accumulator ready", "epilogue: tile 0 stored", "teardown"]
    print(f"n={n} k={k} m={m}")
    for i in range(7):
        print(f"  {names[i]:40s} +{t[i + 1] - t[i]:8d}  (at {t[i + 1] - t[0]})")
