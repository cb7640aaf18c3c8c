"""World-size-2 gloo tests of the multi-GPU host logic (no GPU): the sharded loss and retrieval
choreography in dclip_b200 is driven with oracle-backed stand-ins for the CUDA ops and must
reproduce the single-process oracle."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import reference_math as om
from oracle import retrieval as orc
from oracle import synthetic


class OracleLossOps:
    """CPU stand-in with the CudaLossOps interface (restates the C-ABI building blocks)."""
    op_torch_dtype = torch.float64

    def prep(self, x, teacher, out_hat):
        x = x.double()
        nx = x.norm(dim=1, keepdim=True).clamp_min(1e-12)
        out_hat.copy_(x / nx)
        cos = torch.zeros(x.shape[0], dtype=torch.float32)
        if teacher is not None:
            cos = (out_hat * om.normalize(teacher.double())).sum(1).float()
        return (1.0 / nx).squeeze(1).float(), cos

    def rowlse(self, x_hat, y_all, temperature, row_offset, want_diag, out=None):
        s = x_hat @ y_all.T / temperature
        diag = s[torch.arange(s.shape[0]), row_offset + torch.arange(s.shape[0])].float() if want_diag else None
        lse = torch.logsumexp(s, dim=1).float()
        if out is not None:
            out.copy_(lse)
            lse = out
        return lse, diag

    def grad(self, x_hat, y_all, temperature, row_offset, lse_x, lse_y_all, coef):
        s = x_hat @ y_all.T / temperature
        g = torch.exp(s - lse_x.double()[:, None]) + torch.exp(s - lse_y_all.double()[None, :])
        g[torch.arange(s.shape[0]), row_offset + torch.arange(s.shape[0])] -= 2.0
        return (coef / temperature * g @ y_all).float()

    def tail(self, x, teacher, inv, dx_hat, w_cos_over_B, grad_scale):
        g = dx_hat.double() if dx_hat is not None else torch.zeros_like(x, dtype=torch.float64)
        if teacher is not None:
            g = g - w_cos_over_B * om.normalize(teacher.double())
        xh = x.double() * inv.double()[:, None]
        return ((g - xh * (xh * g).sum(1, keepdim=True)) * inv.double()[:, None] * grad_scale).to(x.dtype)

    def partial_sums(self, lse_u, lse_v, diag, cos_img, cos_txt, out=None):
        out = out if out is not None else torch.zeros(3)
        out.zero_()
        out[0] = (lse_u + lse_v - 2 * diag).sum()
        if cos_img is not None:
            out[1] = (1 - cos_img).sum()
        if cos_txt is not None:
            out[2] = (1 - cos_txt).sum()
        return out


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _loss_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dclip_b200 import losses as L
    m, d = 6, 32
    s_img, s_txt, t_img, t_txt = synthetic.loss_inputs(world * m, d, seed=9)
    sl = slice(rank * m, (rank + 1) * m)
    losses, d_img, d_txt = L.sharded_distill_loss_fwd_bwd(s_img[sl], s_txt[sl], t_img[sl], t_txt[sl], ops=OracleLossOps())
    ret[rank] = (losses.numpy(), d_img.numpy(), d_txt.numpy())
    dist.destroy_process_group()


def test_sharded_loss_choreography_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_loss_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    m, d = 6, 32
    s_img, s_txt, t_img, t_txt = synthetic.loss_inputs(world * m, d, seed=9)
    (tot, li, lt, lc), g_img, g_txt = om.distill_total_with_grads(s_img, s_txt, t_img, t_txt, dtype=torch.float64)
    want = np.array([float(tot), float(li), float(lt), float(lc)])
    for r in range(world):
        losses, d_img, d_txt = ret[r]
        assert np.allclose(losses, want, rtol=1e-5)
        sl = slice(r * m, (r + 1) * m)
        assert np.allclose(d_img, g_img[sl].numpy(), rtol=1e-4, atol=1e-7)
        assert np.allclose(d_txt, g_txt[sl].numpy(), rtol=1e-4, atol=1e-7)


def _topk_cpu(q, g, k, off, path):
    s, i = orc.c_topk(q.numpy(), g.numpy(), k)
    i = np.where(i >= 0, i + off, -1)
    return torch.from_numpy(s), torch.from_numpy(i)


def _merge_cpu(all_s, all_i):
    nl, nq, k = all_s.shape
    s = all_s.permute(1, 0, 2).reshape(nq, nl * k).numpy()
    i = all_i.permute(1, 0, 2).reshape(nq, nl * k).numpy()
    out_s = np.empty((nq, k), np.float32)
    out_i = np.empty((nq, k), np.int64)
    for r in range(nq):
        valid = i[r] >= 0
        order = sorted(np.flatnonzero(valid), key=lambda c: (-s[r, c], i[r, c]))[:k]
        out_s[r], out_i[r] = -np.inf, -1
        out_s[r, :len(order)], out_i[r, :len(order)] = s[r, order], i[r, order]
    return torch.from_numpy(out_s), torch.from_numpy(out_i)


def _ranks_cpu(q, g, ts, tgt, qidx, off, path):
    s = orc.c_scores(q.numpy(), g.numpy())
    j = np.arange(g.shape[0])[None, :] + off
    t, ti = ts.numpy()[:, None], tgt.numpy()[:, None]
    return torch.from_numpy(((s > t) | ((s == t) & (j < ti))).sum(1).astype(np.int32))


def _retrieval_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dclip_b200 import retrieval as R
    gal, q, planted = synthetic.gallery_inputs(301, 9, 16, seed=4)
    bounds = [0, 150, 301]
    a, b = bounds[rank], bounds[rank + 1]
    s, i = R.sharded_topk_ip(torch.from_numpy(q), torch.from_numpy(gal[a:b]), 10, a, topk_fn=_topk_cpu, merge_fn=_merge_cpu)
    tgt = torch.from_numpy(planted.astype(np.int64))
    ts = torch.from_numpy(orc.c_scores(q, gal)[np.arange(9), planted])
    r = R.sharded_retrieval_ranks(torch.from_numpy(q), torch.from_numpy(gal[a:b]), ts, tgt, a, rank_fn=_ranks_cpu)
    ret[rank] = (s.numpy(), i.numpy(), r.numpy())
    dist.destroy_process_group()


def test_sharded_retrieval_choreography_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_retrieval_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    gal, q, planted = synthetic.gallery_inputs(301, 9, 16, seed=4)
    ws, wi = orc.c_topk(q, gal, 10)
    wr = orc.c_ranks(q, gal, np.arange(9), planted)
    for r in range(world):
        s, i, rk = ret[r]
        assert np.array_equal(i, wi) and np.array_equal(s, ws)
        assert np.array_equal(rk, wr)


def test_build_eval_index_matches_reference_bookkeeping():
    from dclip_b200.retrieval import build_eval_index
    image_ids = ["a", "b", "a", "c"]            # duplicate id: list.index -> first match
    cap_ids = ["b", "a", "a", "zz"[:0] or "c"]
    gt, pair_img, pair_cap = build_eval_index(image_ids, cap_ids)
    assert gt.tolist() == [1, 0, 0, 3]
    assert list(zip(pair_img.tolist(), pair_cap.tolist())) == [(0, 1), (0, 2), (1, 0), (2, 1), (2, 2), (3, 3)]
    import pytest
    with pytest.raises(ValueError):
        build_eval_index(["a"], ["q"])
