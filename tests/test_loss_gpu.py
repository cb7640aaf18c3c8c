"""Parity of the CUDA distillation objective (through the C ABI) against the oracle and the
reference's golden outputs.  Tolerances (north_star): 1e-5 relative in fp32, 1e-3 for 16-bit."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import reference_math as om
from oracle import synthetic

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-5
TENSOR_TOL = 1e-3


def _t(a):
    return torch.from_numpy(np.asarray(a))


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _run(s_img, s_txt, t_img, t_txt, temperature=0.05, path="fp32", **kw):
    from dclip_b200 import losses as L
    dev = torch.device("cuda:0")
    args = [None if t is None else t.to(dev) for t in (s_img, s_txt, t_img, t_txt)]
    out = L.distill_loss_fwd_bwd(*args, temperature=temperature, path=path, **kw)
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("name", ["loss_small", "loss_b1", "loss_zero_row", "loss_temp", "loss_cfg1"])
def test_golden_fp32(built_library, name):
    g = load_golden(name)
    losses, d_img, d_txt = _run(_t(g["s_img"]), _t(g["s_txt"]), _t(g["t_img"]), _t(g["t_txt"]), float(g["temperature"]))
    assert np.allclose(losses.cpu().numpy(), g["losses"], rtol=FP32_TOL, atol=2e-6)
    assert _rel(d_img, _t(g["d_s_img"])) < FP32_TOL
    assert _rel(d_txt, _t(g["d_s_txt"])) < FP32_TOL


@pytest.mark.parametrize("B,d", [(32, 512), (100, 768), (257, 1024), (7, 8), (512, 256)])
def test_oracle_fp32_shapes(built_library, B, d):
    s_img, s_txt, t_img, t_txt = synthetic.loss_inputs(B, d, seed=B + d)
    losses, d_img, d_txt = _run(s_img, s_txt, t_img, t_txt)
    (tot, li, lt, lc), g_img, g_txt = om.distill_total_with_grads(s_img, s_txt, t_img, t_txt, dtype=torch.float64)
    want = torch.stack([tot, li, lt, lc])
    assert torch.allclose(losses.cpu().double(), want, rtol=FP32_TOL, atol=2e-6)
    assert _rel(d_img, g_img) < FP32_TOL and _rel(d_txt, g_txt) < FP32_TOL


def test_pure_infonce_and_autograd_wrappers(built_library):
    from dclip_b200 import losses as L
    dev = torch.device("cuda:0")
    s_img, s_txt, t_img, t_txt = synthetic.loss_inputs(48, 128, seed=5)
    a = s_img.to(dev).requires_grad_(True)
    b = s_txt.to(dev).requires_grad_(True)
    loss = L.compute_contrastive_loss(a, b, temperature=0.05)
    (3.0 * loss).backward()
    ref, ga, gb = om.contrastive_with_grads(s_img, s_txt, 0.05, torch.float64)
    assert abs(float(loss) - float(ref)) < FP32_TOL * abs(float(ref)) + 1e-7
    assert _rel(a.grad, 3.0 * ga) < FP32_TOL and _rel(b.grad, 3.0 * gb) < FP32_TOL

    # 16-bit gradients scaled by a device-side upstream factor (dclip_scale_grad): exact 0.5 x, so bit-identical halves
    a16 = s_img.to(dev).bfloat16().requires_grad_(True)
    b16 = s_txt.to(dev).bfloat16().requires_grad_(True)
    L.compute_contrastive_loss(a16, b16, temperature=0.05, path="fp32").backward()
    g1 = a16.grad.clone()
    a16.grad = None
    (0.5 * L.compute_contrastive_loss(a16, b16, temperature=0.05, path="fp32")).backward()
    assert torch.equal(a16.grad.float(), 0.5 * g1.float())

    a2 = s_img.to(dev).requires_grad_(True)
    cl = L.cosine_distillation_loss(a2, t_img.to(dev))
    cl.backward()
    x = s_img.double().requires_grad_(True)
    ref2 = om.cosine_distillation_loss(x, t_img.double())
    ref2.backward()
    assert abs(float(cl) - float(ref2)) < FP32_TOL
    assert _rel(a2.grad, x.grad) < FP32_TOL

    mod = L.DistillationLoss()
    a3 = s_img.to(dev).requires_grad_(True)
    b3 = s_txt.to(dev).requires_grad_(True)
    total = mod(a3, b3, t_img.to(dev), t_txt.to(dev))
    total.backward()
    (tot, li, lt, lc), g_img, g_txt = om.distill_total_with_grads(s_img, s_txt, t_img, t_txt, dtype=torch.float64)
    assert abs(float(total) - float(tot)) < FP32_TOL * float(tot)
    assert torch.allclose(mod.last_components.cpu().double(), torch.stack([tot, li, lt, lc]), rtol=FP32_TOL)
    assert _rel(a3.grad, g_img) < FP32_TOL and _rel(b3.grad, g_txt) < FP32_TOL


def test_known_answers(built_library):
    # identical rows everywhere -> uniform logits -> L_con = log B exactly
    B, d = 16, 64
    x = torch.ones(B, d)
    losses, _, _ = _run(x, x.clone(), None, None, need_grad=False)
    assert abs(float(losses[3]) - np.log(B)) < 1e-5
    # orthogonal one-hot rows at tiny temperature -> loss ~ 0
    e = torch.eye(B, d)
    losses, _, _ = _run(e, e.clone(), e.clone(), e.clone(), temperature=0.01, need_grad=False)
    assert float(losses[3]) < 1e-6 and abs(float(losses[1])) < 1e-6


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_16bit_inputs_fp32_path(built_library, dtype):
    s_img, s_txt, t_img, t_txt = (t.to(dtype) for t in synthetic.loss_inputs(96, 256, seed=11))
    losses, d_img, d_txt = _run(s_img, s_txt, t_img, t_txt)
    (tot, li, lt, lc), g_img, g_txt = om.distill_total_with_grads(s_img.float(), s_txt.float(), t_img.float(),
                                                                   t_txt.float(), dtype=torch.float64)
    assert torch.allclose(losses.cpu().double(), torch.stack([tot, li, lt, lc]), rtol=1e-5, atol=2e-6)
    # gradients are rounded once to the 16-bit output dtype
    tol = 2 ** -8 if dtype == torch.bfloat16 else 2 ** -10
    assert _rel(d_img, g_img) < tol and _rel(d_txt, g_txt) < tol


def test_simulated_sharding_equals_single_gpu(built_library):
    """Run the per-rank building blocks P times on one device: must reproduce the
    single-GPU composite."""
    from dclip_b200 import losses as L
    dev = torch.device("cuda:0")
    P, m, d = 4, 24, 128
    s_img, s_txt, t_img, t_txt = (t.to(dev) for t in synthetic.loss_inputs(P * m, d, seed=3))
    ref_losses, ref_di, ref_dt = L.distill_loss_fwd_bwd(s_img, s_txt, t_img, t_txt, path="fp32")
    ops = L.CudaLossOps("fp32")
    u_all = torch.empty(P * m, d, device=dev)
    v_all = torch.empty(P * m, d, device=dev)
    pre = []
    for r in range(P):
        sl = slice(r * m, (r + 1) * m)
        pre.append((ops.prep(s_img[sl].contiguous(), t_img[sl].contiguous(), u_all[sl]),
                    ops.prep(s_txt[sl].contiguous(), t_txt[sl].contiguous(), v_all[sl])))
    lse_u = torch.empty(P * m, device=dev)
    lse_v = torch.empty(P * m, device=dev)
    sums = torch.zeros(3, device=dev)
    for r in range(P):
        sl = slice(r * m, (r + 1) * m)
        lu, dg = ops.rowlse(u_all[sl], v_all, 0.05, r * m, True)
        lv, _ = ops.rowlse(v_all[sl], u_all, 0.05, r * m, False)
        lse_u[sl], lse_v[sl] = lu, lv
        sums += ops.partial_sums(lu, lv, dg, pre[r][0][1], pre[r][1][1])
    B = P * m
    l_con, l_img, l_txt = sums[0] / (2 * B), sums[1] / B, sums[2] / B
    got = torch.stack([l_img + l_txt + l_con, l_img, l_txt, l_con])
    assert torch.allclose(got, ref_losses, rtol=1e-6, atol=1e-6)
    for r in range(P):
        sl = slice(r * m, (r + 1) * m)
        du = ops.grad(u_all[sl], v_all, 0.05, r * m, lse_u[sl].contiguous(), lse_v, 1.0 / (2 * B))
        di = ops.tail(s_img[sl].contiguous(), t_img[sl].contiguous(), pre[r][0][0], du, 1.0 / B, 1.0)
        assert torch.allclose(di, ref_di[sl], rtol=1e-5, atol=1e-9)
    torch.cuda.synchronize()


def test_errors_are_loud(built_library):
    from dclip_b200 import losses as L
    dev = torch.device("cuda:0")
    with pytest.raises(RuntimeError, match="multiple of 8"):
        L.distill_loss_fwd_bwd(torch.randn(4, 7, device=dev), torch.randn(4, 7, device=dev))
    with pytest.raises(ValueError):
        L.distill_loss_fwd_bwd(torch.randn(4, 8, device=dev), torch.randn(5, 8, device=dev))
