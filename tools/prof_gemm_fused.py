"""Clock stamps of CTA 0 of the gradient GEMM (k_gemm_pair) as the loss step launches it, with and without the fused tail
(prof build: make -C dclip_b200/csrc prof)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("DCLIP_B200_LIB", os.path.join(ROOT, "build", "prof", "libdclip_b200_prof.so"))
import torch
from dclip_b200 import _lib, losses as L
from dclip_b200._lib import check, dtype_code, ptr, stream_ptr
import workloads as synthetic
lib = _lib.load_library()
dev = torch.device("cuda:0")
B, d = 4096, 768
s_img, s_txt, t_img, t_txt = (t.to(torch.bfloat16).to(dev) for t in synthetic.loss_inputs(B, d, seed=5))
ops = L.CudaLossOps("tensor")
u = torch.empty((B, d), dtype=torch.float16, device=dev); v = torch.empty((B, d), dtype=torch.float16, device=dev)
inv_u, cos_i = ops.prep(s_img, t_img, u); inv_v, cos_t = ops.prep(s_txt, t_txt, v)
lse_u, _ = ops.rowlse(u, v, 0.05, 0, True); lse_v, _ = ops.rowlse(v, u, 0.05, 0, False)
coeffs = torch.empty(lib.dclip_infonce_coeffs_bytes(B), dtype=torch.uint8, device=dev)
dots = torch.empty(lib.dclip_infonce_coeffs_dot_bytes(B) // 4, dtype=torch.float32, device=dev)
check(lib.dclip_infonce_coeffs(ptr(u), ptr(v), B, d, 0.05, ptr(lse_u), ptr(lse_v), ptr(coeffs), ptr(dots), stream_ptr(dev)), "coeffs")
du = torch.empty((B, d), dtype=torch.float32, device=dev); dv = torch.empty((B, d), dtype=torch.float32, device=dev)
tinv_i = (1.0 / t_img.float().norm(dim=1).clamp_min(1e-12)).contiguous(); tinv_t = (1.0 / t_txt.float().norm(dim=1).clamp_min(1e-12)).contiguous()
class Tail(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("s_img", "s_txt", "t_img", "t_txt", "inv_u", "inv_v", "cos_img", "cos_txt", "tinv_img", "tinv_txt")] + \
               [("w_img_over_B", ctypes.c_float), ("w_txt_over_B", ctypes.c_float), ("grad_scale", ctypes.c_float), ("dtype", ctypes.c_int),
                ("d_s_img", ctypes.c_void_p), ("d_s_txt", ctypes.c_void_p)]
d_img, d_txt = torch.empty_like(s_img), torch.empty_like(s_txt)
tail = Tail(ptr(s_img), ptr(s_txt), ptr(t_img), ptr(t_txt), ptr(inv_u), ptr(inv_v), ptr(cos_i), ptr(cos_t), ptr(tinv_i), ptr(tinv_t),
            1.0 / B, 1.0 / B, 1.0, dtype_code(s_img), ptr(d_img), ptr(d_txt))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
flush_rd = torch.ones(64 << 20, dtype=torch.float32, device=dev)
names = ["entry", "setup done", "first stage full", "last stage full", "accumulator complete", "epilogue done (warp 6)", "exit"]
for fused in (0, 1):
    for cold in (0, 1):
        def run():
            check(lib.dclip_infonce_grad_from_coeffs(ptr(coeffs), ptr(dots), ptr(u), ptr(v), B, d, 0.05, 1.0 / (2 * B), ptr(du) if not fused else None,
                                                     ptr(dv) if not fused else None, ctypes.addressof(tail) if fused else None, stream_ptr(dev)), "gemms")
        for _ in range(3):
            run()
        if cold:
            flush.zero_(); flush_rd.sum()
            check(lib.dclip_infonce_coeffs(ptr(u), ptr(v), B, d, 0.05, ptr(lse_u), ptr(lse_v), ptr(coeffs), ptr(dots), stream_ptr(dev)), "coeffs")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        buf = (ctypes.c_longlong * 16)()
        lib.dclip_debug_gpair_prof(buf)
        p = list(buf)
        print(f"# fused_tail={fused} {'after flush + coefficient kernel' if cold else 'L2-warm repeat'}: {e0.elapsed_time(e1) * 1e3:.1f} us by events; "
              f"CTA 0: {p[14] - p[8]} ns, {p[6] - p[0]} cycles -> {(p[6] - p[0]) / max(1, p[14] - p[8]):.3f} GHz")
        print("   " + "  ".join(f"{names[i]} +{p[i] - p[i - 1]}" for i in range(1, 7)))
