"""Accuracy of an fp32 GEMM emulated on tcgen05 by bf16 operand splitting (x = h + m + l, six products concatenated along K),
against an fp64 matmul; torch fp32 (TF32 off) beside it.  Stall-accounting build (dclip_debug_gemm16)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dclip_b200 import _lib
dev = torch.device("cuda:0")
lib = _lib.load_library()
fn = lib.dclip_debug_gemm16
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 8 + [ctypes.c_float] + [ctypes.c_int] * 3 + [ctypes.c_void_p]
torch.backends.cuda.matmul.allow_tf32 = False

def split3(x):
    h = x.bfloat16(); r = x - h.float(); m = r.bfloat16(); l = (r - m.float()).bfloat16()
    return h, m, l

def run(M, N, K, scale_a=1.0, scale_b=1.0, order="small_first", nseg=6):
    g = torch.Generator(device="cpu").manual_seed(1)
    a = (torch.randn(M, K, generator=g) * scale_a).to(dev); b = (torch.randn(N, K, generator=g) * scale_b).to(dev)
    ah, am, al = split3(a); bh, bm, bl = split3(b)
    segs = [(ah, bl), (al, bh), (am, bm), (ah, bm), (am, bh), (ah, bh)]
    if nseg == 3: segs = [(ah, bm), (am, bh), (ah, bh)]
    if order != "small_first": segs = segs[::-1]
    A = torch.cat([s[0] for s in segs], 1).contiguous(); B = torch.cat([s[1] for s in segs], 1).contiguous()
    C = torch.empty(M, N, device=dev)
    KK = A.shape[1]
    e = fn(A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, KK, KK, KK, N, 0, 0, 1.0, 0, 1, 1, torch.cuda.current_stream().cuda_stream)
    assert e == 0, e
    torch.cuda.synchronize()
    ref = a.double() @ b.double().T
    rel = lambda x: float((x.double() - ref).norm() / ref.norm())
    print(f"M={M} N={N} K={K} segs={nseg} {order:11s} scale_a={scale_a:g}: split-bf16 {rel(C):.2e}   torch fp32 {rel(a @ b.T):.2e}   max-elem {float((C.double() - ref).abs().max() / ref.abs().max()):.2e}")

for K in (512, 768, 2464):
    run(2464, 768, K); run(2464, 768, K, order="big_first"); run(2464, 768, K, nseg=3)
run(512, 512, 2464, scale_a=1e-7)
