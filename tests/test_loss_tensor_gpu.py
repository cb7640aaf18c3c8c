"""tcgen05 tensor path of the distillation objective vs the oracle (1e-3 tier, north_star)."""
import numpy as np
import pytest
import torch

from oracle import reference_math as om
from oracle import synthetic

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _pipeline_ok(lib):
    from dclip_b200 import _lib
    return lib.dclip_debug_pipeline_error(_lib.stream_ptr(torch.device("cuda:0"))) == 0


@pytest.mark.parametrize("B,d,dtype", [(128, 128, torch.float32), (64, 768, torch.bfloat16), (256, 768, torch.bfloat16),
                                        (1024, 512, torch.float16), (200, 256, torch.bfloat16), (4096, 768, torch.bfloat16),
                                        (4096, 768, torch.float32), (4096, 768, torch.float16)])
def test_tensor_path_matches_oracle(built_library, B, d, dtype):
    """Losses and gradients of the tcgen05 path within 1e-3 of the fp64 oracle (north_star's bf16/tf32 tier).

    fp32 I/O (operands still fp16, fp32 accumulate) is held to 1e-3 outright - including the headline shape
    B=4096, d=768.  With 16-bit OUTPUTS the final rounding alone costs e_r = |round(oracle) - oracle| / |oracle|
    (~1.7e-3 for bf16), which no implementation can avoid: the error of the tensor MATH is what is left after
    removing it in quadrature (independent errors), and that is what is held to 1e-3; both are printed."""
    from dclip_b200 import losses as L
    dev = torch.device("cuda:0")
    s_img, s_txt, t_img, t_txt = (t.to(dtype) for t in synthetic.loss_inputs(B, d, seed=B + d))
    # fp16 gradients of magnitude 1e-5 are subnormal: use a loss scale like AMP does (part of the API)
    gs = 1024.0 if dtype == torch.float16 else 1.0
    losses, d_img, d_txt = L.distill_loss_fwd_bwd(s_img.to(dev), s_txt.to(dev), t_img.to(dev), t_txt.to(dev), path="tensor",
                                                  grad_scale=gs)
    torch.cuda.synchronize()
    assert _pipeline_ok(built_library), "tensor-path pipeline timed out"
    (tot, li, lt, lc), g_img, g_txt = om.distill_total_with_grads(s_img.float(), s_txt.float(), t_img.float(), t_txt.float(),
                                                                   dtype=torch.float64)
    want = torch.stack([tot, li, lt, lc])
    got = losses.cpu().double()
    assert torch.allclose(got, want, rtol=TOL, atol=1e-5), (got, want)
    for name, ours, ref in (("d_img", d_img, gs * g_img), ("d_txt", d_txt, gs * g_txt)):
        total = _rel(ours, ref)
        rounding = _rel(ref.to(dtype), ref)                       # the oracle's own gradient rounded to the output dtype
        math_err = max(0.0, total * total - rounding * rounding) ** 0.5
        print(f"[tensor-path parity] B={B} d={d} {dtype} {name}: total {total:.3e}  output-rounding {rounding:.3e}  "
              f"tensor-math {math_err:.3e}  (bound {TOL:.0e})")
        assert math_err < TOL, (name, total, rounding, math_err)
        assert total < TOL + rounding, (name, total, rounding)


def test_tensor_building_blocks_and_splits(built_library):
    """Sharded building blocks on the tensor path (m < n forces column splits + partial combines)."""
    from dclip_b200 import losses as L
    dev = torch.device("cuda:0")
    P, m, d = 4, 128, 256
    B = P * m
    s_img, s_txt, t_img, t_txt = (t.to(dev) for t in synthetic.loss_inputs(B, d, seed=21))
    ref_losses, ref_di, ref_dt = L.distill_loss_fwd_bwd(s_img, s_txt, t_img, t_txt, path="fp32")
    ops = L.CudaLossOps("tensor")
    u_all = torch.empty(B, d, device=dev, dtype=torch.float16)
    v_all = torch.empty(B, d, device=dev, dtype=torch.float16)
    inv_u, cos_i = ops.prep(s_img, t_img, u_all)
    inv_v, cos_t = ops.prep(s_txt, t_txt, v_all)
    lse_u = torch.empty(B, device=dev)
    lse_v = torch.empty(B, device=dev)
    diag = torch.empty(B, device=dev)
    for r in range(P):
        sl = slice(r * m, (r + 1) * m)
        lse_u[sl], diag[sl] = ops.rowlse(u_all[sl], v_all, 0.05, r * m, True)
        lse_v[sl], _ = ops.rowlse(v_all[sl], u_all, 0.05, r * m, False)
    sums = ops.partial_sums(lse_u, lse_v, diag, cos_i, cos_t)
    l_con = sums[0] / (2 * B)
    assert abs(float(l_con) - float(ref_losses[3])) < TOL * float(ref_losses[3]) + 1e-5
    for r in range(P):
        sl = slice(r * m, (r + 1) * m)
        du = ops.grad(u_all[sl], v_all, 0.05, r * m, lse_u[sl].contiguous(), lse_v, 1.0 / (2 * B))
        di = ops.tail(s_img[sl].contiguous(), t_img[sl].contiguous(), inv_u[sl].contiguous(), du, 1.0 / B, 1.0)
        assert _rel(di, ref_di[sl]) < TOL
    torch.cuda.synchronize()
    assert _pipeline_ok(built_library)


def test_tensor_path_rejects_unsupported(built_library):
    from dclip_b200 import losses as L
    dev = torch.device("cuda:0")
    x = torch.randn(64, 192, device=dev)
    with pytest.raises(RuntimeError, match="tensor path needs"):
        L.distill_loss_fwd_bwd(x, x.clone(), path="tensor")
    y = torch.randn(64, 256, device=dev)
    with pytest.raises(RuntimeError, match="temperature"):
        L.distill_loss_fwd_bwd(y, y.clone(), temperature=0.01, path="tensor")


def test_pair_backward_is_bit_reproducible(built_library):
    """The CTA-pair backward publishes its G tiles across the pair without GPU-scope membars (ordering by
    construction, csrc/loss_tc_pair.cu).  A lost ordering would show up as run-to-run differences: every run
    must be bit-identical (fixed reduction order everywhere)."""
    from dclip_b200 import losses as L
    dev = torch.device("cuda:0")
    for B, d in ((1024, 768), (4096, 256), (640, 512)):
        s_img, s_txt, t_img, t_txt = (t.to(torch.bfloat16).to(dev) for t in synthetic.loss_inputs(B, d, seed=7))
        ref = None
        for _ in range(25):
            losses, d_img, d_txt = L.distill_loss_fwd_bwd(s_img, s_txt, t_img, t_txt, path="tensor")
            cur = (losses.clone(), d_img.clone(), d_txt.clone())
            if ref is None:
                ref = cur
            else:
                for a, b in zip(ref, cur):
                    assert torch.equal(a, b)
    torch.cuda.synchronize()
    assert _pipeline_ok(built_library)


@pytest.mark.parametrize("B,d,dtype,path", [(512, 768, torch.bfloat16, "tensor"), (96, 256, torch.float32, "fp32"),
                                            (4096, 768, torch.bfloat16, "tensor"), (250, 256, torch.float16, "tensor"),
                                            (101, 512, torch.float32, "fp32"), (203, 384, torch.bfloat16, "fp32")])
def test_host_step_split_matches_composite(built_library, B, d, dtype, path):
    """HostDistillStep (students phase, then the teachers phase per side; the last block of the text side's gradient tail stores the
    losses into the pinned host buffer the call waits on) against the one-call composite, incl. batches that do not fill the tail's
    row blocks."""
    from dclip_b200 import losses as L
    dev = torch.device("cuda:0")
    host = [t.to(dtype).contiguous().pin_memory() for t in synthetic.loss_inputs(B, d, seed=3)]
    gs = 1024.0 if dtype == torch.float16 else 1.0      # fp16 gradients of magnitude 1e-5 are subnormal: loss scale, as AMP does
    step = L.HostDistillStep(B, d, dtype, dev, path=path, grad_scale=gs)
    for i in range(4):
        lh, di, dt = step(*host)
        assert torch.equal(lh, step.losses.cpu())        # the host copy is the device's result of THIS step
        if i == 1:      # other inputs in between: every call must deliver its own losses (sequence word), not a stale buffer
            other = [t.to(dtype).contiguous().pin_memory() for t in synthetic.loss_inputs(B, d, seed=4)]
            lo, _, _ = step(*other)
            ref_o, _, _ = L.distill_loss_fwd_bwd(*(t.to(dev) for t in other), path=path)
            assert torch.allclose(lo, ref_o.cpu(), rtol=1e-5, atol=1e-6), (lo, ref_o)
    ref_l, ref_di, ref_dt = L.distill_loss_fwd_bwd(*(t.to(dev) for t in host), path=path, grad_scale=gs)
    torch.cuda.synchronize()
    assert _pipeline_ok(built_library)
    assert torch.allclose(lh, ref_l.cpu(), rtol=1e-5, atol=1e-6), (lh, ref_l)
    # (the composite finishes its gradients in the GEMM epilogue, the split step in the tail kernel: same fp32 math in another order,
    # so 16-bit outputs may round the other way here and there)
    slack = {torch.bfloat16: 2 ** -8, torch.float16: 2 ** -11}.get(dtype, 0)
    assert _rel(di, ref_di) < 1e-5 + slack
    assert _rel(dt, ref_dt) < 1e-5 + slack


@pytest.mark.parametrize("B,d,dtype,path", [(1000, 512, torch.bfloat16, "tensor"), (77, 256, torch.float32, "fp32")])
def test_split_objective_through_the_c_abi(built_library, B, d, dtype, path):
    """dclip_distill_loss_students + dclip_distill_loss_teachers (one tail launch for both sides, its last block finishes the losses)
    against the one-call composite; twice over the same workspace (the block counter must be left clear), and without gradients."""
    from dclip_b200 import _lib, losses as L
    from dclip_b200._lib import check, dtype_code, ptr, stream_ptr
    lib = built_library
    dev = torch.device("cuda:0")
    s_img, s_txt, t_img, t_txt = (t.to(dtype).to(dev) for t in synthetic.loss_inputs(B, d, seed=9))
    ref_l, ref_di, ref_dt = L.distill_loss_fwd_bwd(s_img, s_txt, t_img, t_txt, w_img=0.7, w_txt=1.3, w_con=0.9, path=path)
    pc = _lib.resolve_loss_path(path, s_img, 0.05)
    ws = _lib.workspace(lib.dclip_distill_loss_workspace_bytes(B, d, pc), dev)
    ws.fill_(0xA5)                                   # no reliance on a zeroed workspace
    losses = torch.empty(4, dtype=torch.float32, device=dev)
    di, dt = torch.empty_like(s_img), torch.empty_like(s_txt)
    st = stream_ptr(dev)
    for rep in range(2):
        losses.fill_(-1.0)
        check(lib.dclip_distill_loss_students(ptr(s_img), ptr(s_txt), B, d, dtype_code(s_img), 0.05, 0.9, ptr(ws), ws.numel(), pc, st), "students")
        check(lib.dclip_distill_loss_teachers(ptr(s_img), ptr(s_txt), ptr(t_img), ptr(t_txt), B, d, dtype_code(s_img), 0.05, 0.7, 1.3, 0.9, 1.0,
                                              ptr(losses), ptr(di), ptr(dt), ptr(ws), ws.numel(), pc, st), "teachers")
        torch.cuda.synchronize()
        assert torch.allclose(losses, ref_l, rtol=1e-5, atol=1e-6), (rep, losses, ref_l)
        tol = 1e-5 + (2 ** -8 if dtype == torch.bfloat16 else 0)
        assert _rel(di, ref_di) < tol and _rel(dt, ref_dt) < tol
    losses.fill_(-1.0)          # loss only: cosine rows + loss sums kernels
    check(lib.dclip_distill_loss_students(ptr(s_img), ptr(s_txt), B, d, dtype_code(s_img), 0.05, 0.9, ptr(ws), ws.numel(), pc, st), "students")
    check(lib.dclip_distill_loss_teachers(ptr(s_img), ptr(s_txt), ptr(t_img), ptr(t_txt), B, d, dtype_code(s_img), 0.05, 0.7, 1.3, 0.9, 1.0,
                                          ptr(losses), None, None, ptr(ws), ws.numel(), pc, st), "teachers (loss only)")
    torch.cuda.synchronize()
    assert torch.allclose(losses, ref_l, rtol=1e-5, atol=1e-6), (losses, ref_l)
    assert _pipeline_ok(lib)


@pytest.mark.parametrize("B,d", [(4096, 768), (1000, 512), (264, 256)])
def test_materialised_g_backward_matches_fused_pair_kernel(built_library, B, d, monkeypatch):
    """Single-GPU backward through the materialised fp16 coefficient matrix (stream mode 3 + two GEMMs, the default)
    against the fused CTA-pair kernel (DCLIP_BWD_GMAT=0, the kernel of the sharded path): same G values, fp32
    accumulation in a different order."""
    from dclip_b200 import losses as L
    dev = torch.device("cuda:0")
    s_img, s_txt, t_img, t_txt = (t.to(dev) for t in synthetic.loss_inputs(B, d, seed=11))
    monkeypatch.setenv("DCLIP_BWD_GMAT", "0")
    l0, di0, dt0 = L.distill_loss_fwd_bwd(s_img, s_txt, t_img, t_txt, path="tensor")
    monkeypatch.delenv("DCLIP_BWD_GMAT")
    l1, di1, dt1 = L.distill_loss_fwd_bwd(s_img, s_txt, t_img, t_txt, path="tensor")
    torch.cuda.synchronize()
    assert _pipeline_ok(built_library)
    assert torch.equal(l0, l1)
    print(f"[gmat vs fused] B={B} d={d}: d_img {_rel(di1, di0):.3e} d_txt {_rel(dt1, dt0):.3e}")
    assert _rel(di1, di0) < 5e-5 and _rel(dt1, dt0) < 5e-5      # (the fused tail projects with the fp16 operand rows)


def test_two_stage_building_blocks_match_composite(built_library):
    """dclip_infonce_coeffs + dclip_infonce_grad_from_coeffs (the two launches of the single-GPU backward, as bench.py times
    them) reproduce the composite's gradients bit for bit, with the fused tail and with fp32 dX_hat outputs + the tail kernel."""
    import ctypes
    from dclip_b200 import _lib, losses as L
    from dclip_b200._lib import check, dtype_code, ptr, stream_ptr
    lib = built_library
    dev = torch.device("cuda:0")
    B, d = 1152, 768
    s_img, s_txt, t_img, t_txt = (t.to(torch.bfloat16).to(dev) for t in synthetic.loss_inputs(B, d, seed=5))
    _, ref_di, ref_dt = L.distill_loss_fwd_bwd(s_img, s_txt, t_img, t_txt, path="tensor")
    ops = L.CudaLossOps("tensor")
    u = torch.empty((B, d), dtype=torch.float16, device=dev)
    v = torch.empty((B, d), dtype=torch.float16, device=dev)
    inv_u, cos_i = ops.prep(s_img, t_img, u)
    inv_v, cos_t = ops.prep(s_txt, t_txt, v)
    lse_u, _ = ops.rowlse(u, v, 0.05, 0, True)
    lse_v, _ = ops.rowlse(v, u, 0.05, 0, False)
    coeffs = torch.empty(lib.dclip_infonce_coeffs_bytes(B), dtype=torch.uint8, device=dev)
    dots = torch.empty(lib.dclip_infonce_coeffs_dot_bytes(B) // 4, dtype=torch.float32, device=dev)
    check(lib.dclip_infonce_coeffs(ptr(u), ptr(v), B, d, 0.05, ptr(lse_u), ptr(lse_v), ptr(coeffs), ptr(dots), stream_ptr(dev)), "coeffs")
    # fp32 dX_hat + tail kernel
    du = torch.empty((B, d), dtype=torch.float32, device=dev)
    dv = torch.empty((B, d), dtype=torch.float32, device=dev)
    check(lib.dclip_infonce_grad_from_coeffs(ptr(coeffs), ptr(dots), ptr(u), ptr(v), B, d, 0.05, 1.0 / (2 * B), ptr(du), ptr(dv), None,
                                             stream_ptr(dev)), "grad_from_coeffs")
    di = ops.tail(s_img, t_img, inv_u, du, 1.0 / B, 1.0)
    dt = ops.tail(s_txt, t_txt, inv_v, dv, 1.0 / B, 1.0)
    assert _rel(di, ref_di) < 5e-5 + 2 ** -8 and _rel(dt, ref_dt) < 5e-5 + 2 ** -8
    # fused tail
    tinv_i = (1.0 / t_img.float().norm(dim=1).clamp_min(1e-12)).contiguous()
    tinv_t = (1.0 / t_txt.float().norm(dim=1).clamp_min(1e-12)).contiguous()

    class Tail(ctypes.Structure):
        _fields_ = [(n, ctypes.c_void_p) for n in ("s_img", "s_txt", "t_img", "t_txt", "inv_u", "inv_v", "cos_img", "cos_txt", "tinv_img", "tinv_txt")] + \
                   [("w_img_over_B", ctypes.c_float), ("w_txt_over_B", ctypes.c_float), ("grad_scale", ctypes.c_float), ("dtype", ctypes.c_int),
                    ("d_s_img", ctypes.c_void_p), ("d_s_txt", ctypes.c_void_p)]
    d_img, d_txt = torch.empty_like(s_img), torch.empty_like(s_txt)
    tail = Tail(ptr(s_img), ptr(s_txt), ptr(t_img), ptr(t_txt), ptr(inv_u), ptr(inv_v), ptr(cos_i), ptr(cos_t), ptr(tinv_i), ptr(tinv_t),
                1.0 / B, 1.0 / B, 1.0, dtype_code(s_img), ptr(d_img), ptr(d_txt))
    check(lib.dclip_infonce_grad_from_coeffs(ptr(coeffs), ptr(dots), ptr(u), ptr(v), B, d, 0.05, 1.0 / (2 * B), None, None,
                                             ctypes.addressof(tail), stream_ptr(dev)), "grad_from_coeffs(tail)")
    torch.cuda.synchronize()
    assert _pipeline_ok(lib)
    # the teachers' inverse norms come from torch here and from the prep kernel in the composite: equal to a rounding
    assert _rel(d_img, ref_di) < 1e-5 + 2 ** -8 and _rel(d_txt, ref_dt) < 1e-5 + 2 ** -8
