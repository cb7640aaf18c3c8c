"""Edge cases on the GPU: empty inputs, tiny shapes, the CUDA-graph wrapper, zero-shot top-k."""
import numpy as np
import pytest
import torch

from oracle import reference_math as om
from oracle import retrieval as orc
from oracle import synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_empty_queries_and_empty_gallery(built_library):
    from dclip_b200 import retrieval as R
    g = torch.randn(100, 64, device=DEV)
    s, i = R.topk_ip(torch.empty(0, 64, device=DEV), g, 5, path="fp32")
    assert s.shape == (0, 5) and i.shape == (0, 5)
    s, i = R.topk_ip(torch.randn(3, 64, device=DEV), torch.empty(0, 64, device=DEV), 4, path="fp32")
    assert torch.all(i == -1) and torch.all(torch.isinf(s))
    index = R.IndexFlatIP(64)
    D, I = index.search(np.zeros((2, 64), np.float32), 3)         # search before any add, like an empty faiss index
    assert (I == -1).all() and index.ntotal == 0


def test_single_pair_batch(built_library):
    from dclip_b200 import losses as L
    a = torch.randn(1, 64, device=DEV)
    b = torch.randn(1, 64, device=DEV)
    losses, da, db = L.distill_loss_fwd_bwd(a, b, path="fp32")
    assert abs(float(losses[3])) < 1e-6          # one pair: both cross-entropies are log(1)
    assert float(da.abs().max()) < 1e-6 and float(db.abs().max()) < 1e-6


def test_graphed_step_matches_eager(built_library):
    from dclip_b200 import losses as L
    ins = [t.to(torch.bfloat16).to(DEV) for t in synthetic.loss_inputs(256, 256, seed=4)]
    fn = lambda a, b, c, d: L.distill_loss_fwd_bwd(a, b, c, d, path="tensor")
    ref = fn(*ins)
    g = L.GraphedStep(fn, ins)
    for _ in range(2):
        out = g(*ins)
    torch.cuda.synchronize()
    for x, y in zip(out, ref):
        assert torch.equal(x, y)
    ins2 = [t.to(torch.bfloat16).to(DEV) for t in synthetic.loss_inputs(256, 256, seed=5)]
    out2 = g(*ins2)
    ref2 = fn(*ins2)
    torch.cuda.synchronize()
    assert torch.equal(out2[0], ref2[0]) and torch.equal(out2[1], ref2[1])


def test_zero_shot_top5_matches_reference_semantics(built_library):
    """eval_scripts/test_zero_shot_ImageNet.py:79-85: 100 * (i_hat @ t_hat^T) then topk(5): the index order
    equals the inner-product top-5 of the normalised features."""
    from dclip_b200 import retrieval as R
    g = torch.Generator().manual_seed(3)
    img = torch.randn(64, 512, generator=g)
    txt = torch.randn(1000, 512, generator=g)
    vals, idx = om.zero_shot_topk(img, txt, 5)
    ih = R.normalize_rows(img.to(DEV))
    th = R.normalize_rows(txt.to(DEV))
    s, i = R.topk_ip(ih, th, 5, path="fp32")
    assert torch.equal(i.cpu(), idx)
    assert torch.allclose(100.0 * s.cpu(), vals, rtol=1e-5, atol=1e-4)


def test_grad_pair_entry_matches_composite(built_library):
    from dclip_b200 import _lib
    from dclip_b200 import losses as L
    from dclip_b200._lib import check, ptr, stream_ptr, workspace
    lib = built_library
    dev = torch.device(DEV)
    B, d = 192, 256
    s_img, s_txt, _, _ = (t.to(dev) for t in synthetic.loss_inputs(B, d, seed=9))
    ops = L.CudaLossOps("tensor")
    u = torch.empty(B, d, dtype=torch.float16, device=dev)
    v = torch.empty(B, d, dtype=torch.float16, device=dev)
    ops.prep(s_img, None, u)
    ops.prep(s_txt, None, v)
    lu, _ = ops.rowlse(u, v, 0.05, 0, True)
    lv, _ = ops.rowlse(v, u, 0.05, 0, False)
    du = torch.empty(B, d, device=dev)
    dv = torch.empty(B, d, device=dev)
    ws = workspace(lib.dclip_infonce_grad_pair_workspace_bytes(B, d, _lib.PATH_TENSOR), dev)
    check(lib.dclip_infonce_grad_pair(ptr(u), ptr(v), B, d, _lib.DCLIP_F16, 0.05, ptr(lu), ptr(lv), 1.0 / (2 * B), ptr(du), ptr(dv),
                                      ptr(ws), ws.numel(), _lib.PATH_TENSOR, stream_ptr(dev)), "grad_pair")
    want_u = ops.grad(u, v, 0.05, 0, lu, lv, 1.0 / (2 * B))
    want_v = ops.grad(v, u, 0.05, 0, lv, lu, 1.0 / (2 * B))
    torch.cuda.synchronize()
    assert torch.allclose(du, want_u, rtol=1e-5, atol=1e-9) and torch.allclose(dv, want_v, rtol=1e-5, atol=1e-9)
