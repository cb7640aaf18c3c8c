"""Times the InfoNCE backward kernel alone (both directions, one launch) at a given B, d."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dclip_b200 import _lib, losses as L
from dclip_b200._lib import check, ptr, stream_ptr, workspace

def run(B, d, n=20):
    lib = _lib.load_library()
    dev = torch.device("cuda:0")
    ops = L.CudaLossOps("tensor")
    g = torch.Generator(device="cpu").manual_seed(1)
    a = torch.randn(B, d, generator=g).to(dev)
    b = (0.5 * a.cpu() + torch.randn(B, d, generator=g)).to(dev)
    u = torch.empty((B, d), dtype=ops.op_torch_dtype, device=dev)
    v = torch.empty((B, d), dtype=ops.op_torch_dtype, device=dev)
    ops.prep(a, None, u); ops.prep(b, None, v)
    lse_u, _ = ops.rowlse(u, v, 0.05, 0, True)
    lse_v, _ = ops.rowlse(v, u, 0.05, 0, False)
    du = torch.empty((B, d), dtype=torch.float32, device=dev)
    dv = torch.empty((B, d), dtype=torch.float32, device=dev)
    pc = _lib.path_code("tensor")
    ws = workspace(lib.dclip_infonce_grad_pair_workspace_bytes(B, d, pc), dev)
    def launch():
        check(lib.dclip_infonce_grad_pair(ptr(u), ptr(v), B, d, ops.op_dtype, 0.05, ptr(lse_u), ptr(lse_v), 1.0 / (2 * B),
                                          ptr(du), ptr(dv), ptr(ws), ws.numel(), pc, stream_ptr(dev)), "grad_pair")
    for _ in range(3): launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): launch()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    if hasattr(lib, "dclip_debug_pair_prof"):
        import ctypes
        buf = (ctypes.c_ulonglong * 32)()
        lib.dclip_debug_pair_prof(buf, 1)
        launch(); torch.cuda.synchronize()
        lib.dclip_debug_pair_prof(buf, 1)
        v = list(buf)
        names = ["mma total", "mma wait x_full", "mma wait s_empty", "mma wait full(P1)", "mma wait g_full", "mma wait full(P2)",
                 "producer0 wait empty (leader)", "producer wait empty (peer)"]
        for i, nm in enumerate(names): print(f"  {nm:34s} {v[i]:>10d} cyc")
        print(f"  prologue {v[28]} cyc (2 CTAs), drain {v[29]} cyc (2 CTAs), kernel entry->exit {v[27]} cyc; leader fence+arrive local {v[30]} remote {v[31]}")
        for cr in range(2):
            for hh in range(2):
                b = 8 + 5 * (cr * 2 + hh)
                print(f"  epilogue cta{cr} half{hh} ({'local' if cr == hh else 'remote'} store): wait s_full {v[b]}, ld {v[b+1]}, compute {v[b+2]}, "
                      f"wait g_empty {v[b+3]}, store+fence {v[b+4]}")
    print(f"bwd B={B} d={d} slots={os.environ.get('DCLIP_PAIR_SLOTS','auto')} single={os.environ.get('DCLIP_BWD_SINGLE','0')}: "
          f"{us:.1f} us  algorithmic {4.0*B*B*d/us/1e6:.0f} TFLOP/s  err={lib.dclip_debug_pipeline_error(stream_ptr(dev))}", flush=True)

if __name__ == "__main__":
    run(int(sys.argv[1]), int(sys.argv[2]))
