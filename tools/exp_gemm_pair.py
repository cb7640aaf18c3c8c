"""Experiment (prof build: make -C dclip_b200/csrc prof): the CTA-pair 256 x 384 GEMM against the single-CTA 128 x 256 grouped
GEMM at the shapes of the materialised-G backward; K sweep separates the fixed cost from the per-k-block cost."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("DCLIP_B200_LIB", os.path.join(ROOT, "build", "prof", "libdclip_b200_prof.so"))
import torch
from dclip_b200 import _lib
from dclip_b200._lib import ptr, stream_ptr
lib = _lib.load_library()
fn = lib.dclip_debug_gemm16
fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 8 + [ctypes.c_float] + [ctypes.c_int] * 3 + [ctypes.c_void_p]
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timed(call, reps=10):
    """device time per launch: `reps` launches replayed from one CUDA graph (no host gaps)"""
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                call()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps

M, N = 4096, 768
for K in (512, 1024, 2048, 4096):
    G = torch.randn(max(M, K), max(M, K), device=dev).half()
    V = torch.randn(K, N, device=dev).half()
    C = torch.empty(2, M, N, device=dev)
    for a_mn in (0, 1):
        row = []
        for kernel in (0, 1):
            for count in (1, 2):
                def call():
                    r = fn(ptr(G), ptr(V), ptr(C), M, N, K, G.shape[1], N, N, a_mn, 1, 1.0, kernel, 2, count, stream_ptr(dev))
                    assert r == 0, r
                t = timed(call)
                row.append(f"{'pair' if kernel else 'tc  '} x{count}: {t:6.1f} us ({2.0 * M * N * K * count / t / 1e6:6.0f} TF)")
        print(f"K={K:5d} a_mn={a_mn}  " + "   ".join(row), flush=True)
    if K == 4096:
        ref = (G[:M, :K].float() @ V.float())
        fn(ptr(G), ptr(V), ptr(C), M, N, K, G.shape[1], N, N, 0, 1, 1.0, 1, 2, 1, stream_ptr(dev))
        torch.cuda.synchronize()
        print("pair a_mn=0 rel err", float((C[0] - ref).norm() / ref.norm()))

# clock stamps of CTA 0 (leader of pair 0), K = 4096, two problems
prof = lib.dclip_debug_gpair_prof
buf = (ctypes.c_longlong * 16)()
fn(ptr(G), ptr(V), ptr(C), M, N, 4096, G.shape[1], N, N, 0, 1, 1.0, 1, 2, 2, stream_ptr(dev))
prof(buf)
v = list(buf)
names = ["entry", "setup done", "first stage full", "last stage full", "accumulator complete (epilogue sees it)", "epilogue done", "exit"]
for i in range(1, 7):
    print(f"  {names[i]:42s} +{v[i] - v[0]:8d} cyc  (step {v[i] - v[i - 1]:8d})")
print(f"  CTA 0 lifetime {v[14] - v[8]} ns for {v[6] - v[0]} cycles -> effective SM clock {(v[6] - v[0]) / max(1, v[14] - v[8]):.3f} GHz")
