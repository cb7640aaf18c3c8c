"""SURVEY.md 8(f) rows 2-3 on the GPU: ragged batch assembly (bit-exact data movement, projection GEMM at the 1e-5 /
1e-3 tiers) against the line-for-line oracle loops, and the fused Adam / AdamW + global-norm clip against the very
optimizers the reference constructs (torch.optim.Adam / AdamW + clip_grad_norm_, run on the CPU)."""
import numpy as np
import pytest
import torch

from oracle import assembly as oa

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ragged(B, max_n, d, seed, allow_empty=True):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(0 if allow_empty else 1, max_n + 1, (B,), generator=g).tolist()
    lens[0] = max_n
    return [torch.randn(n, d, generator=g) for n in lens]


@pytest.mark.parametrize("B,max_n,d,dtype", [(32, 36, 768, torch.float32), (7, 5, 64, torch.float32), (32, 8, 512, torch.bfloat16),
                                              (3, 1, 96, torch.float16), (5, 9, 20, torch.float32)])
def test_pad_ragged_bit_exact(built_library, B, max_n, d, dtype):
    from dclip_b200 import assembly as A
    parts = [p.to(dtype) for p in _ragged(B, max_n, d, seed=B + d)]
    want = oa.pad_patches(parts, d)
    got = A.pad_ragged([p.to(DEV) for p in parts])
    assert got.shape == want.shape and got.dtype == dtype
    assert torch.equal(got.cpu(), want)
    # truncation to a smaller maximum (reference :571-572) and padding to a larger one
    got2 = A.pad_ragged([p.to(DEV) for p in parts], max_len=max(1, max_n - 2))
    assert torch.equal(got2.cpu(), want[:, :max(1, max_n - 2)])
    got3 = A.pad_ragged([p.to(DEV) for p in parts], max_len=max_n + 3)
    assert torch.equal(got3[:, :max_n].cpu(), want) and float(got3[:, max_n:].abs().max()) == 0.0


def test_pad_text_nan_guard(built_library):
    from dclip_b200 import assembly as A
    parts = _ragged(6, 11, 64, seed=5, allow_empty=False)
    parts[2] = torch.randn(4, 64)
    parts[2][1, 3] = float("nan")
    parts[4][0, 0] = float("inf")
    want = oa.pad_text(parts, 64, nan_guard=True)
    got, flags = A.pad_ragged([p.to(DEV) for p in parts], nan_guard=True)
    assert torch.equal(got.cpu(), want)
    assert flags.cpu().tolist() == [0, 0, 1, 0, 1, 0]


def _token_batch(B, L, seed):
    g = torch.Generator().manual_seed(seed)
    ids = torch.zeros((B, L), dtype=torch.int64)
    mask = torch.zeros((B, L), dtype=torch.int64)
    for b in range(B):
        n = int(torch.randint(2, L + 1, (1,), generator=g))       # BOS + tokens + EOS, then padding (id 0, mask 0)
        ids[b, 0] = 49406
        ids[b, 1:n - 1] = torch.randint(1, 49405, (n - 2,), generator=g)
        ids[b, n - 1] = 49407
        mask[b, :n] = 1
    ids[1, :] = 0; ids[1, 0] = 49406; ids[1, 1] = 49407; mask[1, :] = 0; mask[1, :2] = 1      # empty caption: fallback row
    ids[B - 1, :] = torch.randint(1, 49405, (L,), generator=g); mask[B - 1, :] = 1               # full-length caption
    return ids, mask


@pytest.mark.parametrize("B,L,width,d,dtype,tol", [(8, 77, 512, 512, torch.float32, 1e-5), (32, 77, 768, 768, torch.bfloat16, 1e-3),
                                                    (4, 20, 64, 32, torch.float32, 1e-5), (16, 77, 768, 768, torch.float16, 1e-3)])
def test_project_text_tokens(built_library, B, L, width, d, dtype, tol):
    """Token filter + per-token text_projection loop (text_tokenizer.py:196-213) + padding (:602-616) vs the oracle loop."""
    from dclip_b200 import assembly as A
    g = torch.Generator().manual_seed(B + L)
    hidden = torch.randn(B, L, width, generator=g).to(dtype)
    W = (torch.randn(d, width, generator=g) / width ** 0.5).to(dtype)
    pooled = torch.randn(B, d, generator=g).to(dtype)
    ids, mask = _token_batch(B, L, seed=L)
    want = oa.batch_text(hidden.double(), ids, mask, W.double(), pooled.double())
    got = A.project_text_tokens(hidden.to(DEV), ids.to(DEV), mask.to(DEV), W.to(DEV), pooled=pooled.to(DEV))
    assert got.shape == want.shape, (got.shape, want.shape)
    gotc = got.double().cpu()
    # structure is exact: padding rows are exact zeros, the fallback row is the pooled embedding bit for bit
    lens, pos = A.select_tokens(ids.to(DEV), mask.to(DEV))
    lens = lens.cpu().tolist()
    for b in range(B):
        n = max(lens[b], 1)
        assert float(gotc[b, n:].abs().max() if n < gotc.shape[1] else 0.0) == 0.0
        if lens[b] == 0:
            assert torch.equal(got[b, 0].cpu(), pooled[b])
        assert pos[b, :lens[b]].cpu().tolist() == [i for i in range(1, L - 1) if ids[b, i] > 0 and mask[b, i] == 1]
    rel = float((gotc - want).norm() / want.norm())
    rounding = float((want.to(dtype).double() - want).norm() / want.norm())
    assert max(0.0, rel * rel - rounding * rounding) ** 0.5 < tol, (rel, rounding)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_linear_rows_odd_shapes(built_library, dtype):
    """The grouped tcgen05 GEMM behind dclip_linear_rows at shapes that do not fill its 128 x 256 tiles, 64-wide k-blocks or
    the 64- / 32-column chunks of its specialised epilogue (16-bit output + fp32 bias), against an fp64 product of the same
    16-bit operands: only the output rounding separates the two."""
    from dclip_b200 import assembly as A
    g = torch.Generator().manual_seed(7)
    for n, k, m in ((1, 8, 8), (33, 72, 24), (129, 200, 40), (300, 64, 264), (257, 136, 328), (128, 768, 256)):
        x = torch.randn(n, k, generator=g).to(dtype)
        w = (torch.randn(m, k, generator=g) / k ** 0.5).to(dtype)
        b = torch.randn(m, generator=g)
        for bias in (None, b):
            got = A.linear_rows(x.to(DEV), w.to(DEV), None if bias is None else bias.to(DEV)).double().cpu()
            want = x.double() @ w.double().T + (0 if bias is None else bias.double())
            rounded = want.to(dtype).double()
            # fp32 accumulation: the result is the correctly rounded fp64 value except where fp32 summation error crosses a rounding boundary
            err = float((got - want).abs().max() / want.abs().max())
            rnd = float((rounded - want).abs().max() / want.abs().max())
            # (an element whose fp32 sum lands across a rounding boundary is off by one ulp instead of half: 2 x, never more)
            assert got.shape == want.shape and err <= 2.1 * rnd + 1e-6, (n, k, m, bias is not None, err, rnd)


def _clone_params(shapes, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(s, generator=g).to(dtype) for s in shapes]


TEACHER_SHAPES = [(3 * 64, 64), (3 * 64,), (64, 64), (64,), (3 * 64, 64), (3 * 64,), (64, 64), (64,), (64,), (64,), (64,), (64,)]


@pytest.mark.parametrize("kind,kwargs,clip", [("adam", dict(lr=1e-5), None),                                  # train_contrastive_teacher.py:245
                                               ("adamw", dict(lr=1e-3, weight_decay=1e-2), 0.5),               # CLIP_image_distillation.py:680 + clip 0.5
                                               ("adam", dict(lr=3e-3, weight_decay=0.1, betas=(0.8, 0.95), eps=1e-6), 0.05)])
def test_fused_adam_matches_torch(built_library, kind, kwargs, clip):
    """The 12 teacher tensors (and odd sizes) over 5 steps against torch.optim on the CPU, the reference's own optimizer."""
    from dclip_b200.optim import FusedAdam
    shapes = TEACHER_SHAPES + [(8193,), (5, 7, 3), (1,)]
    ref_p = [torch.nn.Parameter(t.clone()) for t in _clone_params(shapes, torch.float32, 1)]
    our_p = [torch.nn.Parameter(t.detach().clone().to(DEV)) for t in ref_p]
    ref_opt = (torch.optim.AdamW if kind == "adamw" else torch.optim.Adam)(ref_p, **kwargs)
    ours = FusedAdam(our_p, adamw=(kind == "adamw"), max_grad_norm=clip, **kwargs)
    for step in range(5):
        grads = _clone_params(shapes, torch.float32, 100 + step)
        scale = [1.0, 0.01, 30.0, 1e-3, 1.0][step]
        for p, q, g in zip(ref_p, our_p, grads):
            p.grad = g.clone() * scale
            q.grad = (g * scale).to(DEV)
        if clip is not None:
            norm = torch.nn.utils.clip_grad_norm_(ref_p, clip)
        ref_opt.step()
        ours.step()
        if clip is not None:
            st = ours.read_state()
            assert abs(st["grad_norm"] - float(norm)) < 1e-5 * float(norm)
            assert st["step"] == step + 1
        for p, q in zip(ref_p, our_p):
            assert torch.allclose(q.detach().cpu(), p.detach(), rtol=1e-6, atol=1e-7), (kind, step, float((q.detach().cpu() - p.detach()).abs().max()))
    for m, v, p in zip(ours.exp_avg, ours.exp_avg_sq, ref_p):
        stt = ref_opt.state[p]
        # moments: relative to the tensor's norm (single elements of m can be differences of nearly equal terms)
        assert float((m.cpu() - stt["exp_avg"]).norm()) <= 1e-6 * float(stt["exp_avg"].norm()) + 1e-12
        assert float((v.cpu() - stt["exp_avg_sq"]).norm()) <= 1e-6 * float(stt["exp_avg_sq"].norm()) + 1e-12


def test_fused_adam_bf16_params(built_library):
    from dclip_b200.optim import FusedAdam
    shapes = [(256, 96), (96,)]
    base = _clone_params(shapes, torch.bfloat16, 3)
    our_p = [torch.nn.Parameter(t.clone().to(DEV)) for t in base]
    ref_p = [t.float().clone() for t in base]
    ours = FusedAdam(our_p, lr=1e-2)
    m = [torch.zeros_like(t) for t in ref_p]
    v = [torch.zeros_like(t) for t in ref_p]
    for step in range(1, 4):
        grads = _clone_params(shapes, torch.bfloat16, 50 + step)
        for q, g in zip(our_p, grads):
            q.grad = g.to(DEV)
        ours.step()
        for i, g in enumerate(grads):      # fp32 math on the bf16 values, parameter rounded to bf16 after every step
            g = g.float()
            m[i] = 0.9 * m[i] + 0.1 * g
            v[i] = 0.999 * v[i] + 0.001 * g * g
            ref_p[i] = (ref_p[i] - (1e-2 / (1 - 0.9 ** step)) * m[i] / (v[i].sqrt() / (1 - 0.999 ** step) ** 0.5 + 1e-8)).to(torch.bfloat16).float()
    for q, p in zip(our_p, ref_p):
        assert torch.allclose(q.detach().float().cpu(), p, rtol=2 ** -7, atol=1e-6)


def test_graphed_teacher_step_with_optimizer(built_library):
    """forward + backward + FusedAdam.step captured in one CUDA graph == the eager sequence, step after step."""
    from dclip_b200 import losses as L
    from dclip_b200.optim import FusedAdam
    from dclip_b200.patch_text_aggregation import GraphedTeacherStep, PatchTextAggregation
    from oracle import reference_math as om
    from oracle import synthetic
    B, T, R, d = 8, 12, 5, 128
    params = om.perturbed_cross_attention_params(d, 8, 3)
    text, image = synthetic.xattn_inputs(B, T, R, d, 3, ragged=True)
    txt = torch.randn(B, d, generator=torch.Generator().manual_seed(1))

    def make():
        m = PatchTextAggregation(embed_dim=d, num_heads=8)
        m.cross_modal_attention.load_state_dict(params)
        return m.to(DEV)

    def loss_fn(mod, t, i, x):
        return L.compute_contrastive_loss(mod.global_embedding(t, i), x, path="fp32")

    tg, ig, xg = text.to(DEV), image.to(DEV), txt.to(DEV)
    eager, graphed = make(), make()
    opt_e = FusedAdam(eager.parameters(), lr=1e-3)
    opt_g = FusedAdam(graphed.parameters(), lr=1e-3)
    gs = GraphedTeacherStep(graphed, loss_fn, [tg, ig, xg], optimizer=opt_g)
    for _ in range(4):
        opt_e.zero_grad()
        le = loss_fn(eager, tg, ig, xg)
        le.backward()
        opt_e.step()
        lg = gs(tg, ig, xg)
        assert float(le) == float(lg)
    assert opt_g.read_state()["step"] == 4
    for a, b in zip(eager.parameters(), graphed.parameters()):
        assert torch.equal(a, b)
