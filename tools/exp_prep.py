"""k_loss_prep_pair_nc alone (B=4096, d=768, bf16 -> fp16 operands; 25 MB read, 12.6 MB written): L2-warm back to back, cold after the
bench's L2 flush (GPU idle in front of it), and cold but busy (rotating over input sets larger than L2, no idle gap) - which part of
the in-step 17 us is the kernel's and which is the cold start of a step.  A same-traffic torch copy is timed the same three ways."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dclip_b200 import _lib
from dclip_b200._lib import check, dtype_code, ptr, stream_ptr

dev = torch.device("cuda:0")
lib = _lib.load_library()
B, d = 4096, 768
NSETS = 12      # 12 x 25 MB = 300 MB of inputs > 126 MB of L2
sets = [[torch.randn(B, d, device=dev).bfloat16() for _ in range(4)] for _ in range(NSETS)]
u = [torch.empty(B, d, dtype=torch.float16, device=dev) for _ in range(NSETS)]
v = [torch.empty(B, d, dtype=torch.float16, device=dev) for _ in range(NSETS)]
vec = [torch.empty(B, dtype=torch.float32, device=dev) for _ in range(5)]
csrc = [torch.empty(19 * 2 ** 20 // 2, dtype=torch.float16, device=dev) for _ in range(NSETS)]     # 19 MB read + 19 MB written
cdst = [torch.empty_like(c) for c in csrc]
flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device=dev)
flush_rd = torch.ones(64 * 2 ** 20, dtype=torch.float32, device=dev)
st = stream_ptr(dev)


def prep(i):
    s = sets[i]
    check(lib.dclip_sharded_loss_prep(ptr(s[0]), ptr(s[1]), ptr(s[2]), ptr(s[3]), B, d, dtype_code(s[0]), 0.05, ptr(u[i]), ptr(v[i]), 0,
                                      ptr(vec[0]), ptr(vec[1]), ptr(vec[2]), ptr(vec[3]), ptr(vec[4]), None, st), "prep")


def copy(i):
    cdst[i].copy_(csrc[i])


def timed(fn, mode, n=48):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for i in range(n):
        k = 0 if mode == "warm" else i % NSETS
        if mode == "flushed":
            flush.zero_(); flush_rd.sum(); torch.cuda.synchronize()
        ev[i][0].record(); fn(k); ev[i][1].record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) * 1e3 for a, b in ev[8:])
    return t[len(t) // 2], t[0]


for name, fn in (("k_loss_prep_pair_nc", prep), ("torch copy 19 MB -> 19 MB", copy)):
    for mode in ("warm", "rotating", "flushed"):
        for _ in range(3):
            fn(0)
        med, best = timed(fn, mode)
        print(f"{name:28s} {mode:9s} median {med:6.2f} us  best {best:6.2f} us  ({38.0e6 / med / 1e6:.2f} TB/s)")
