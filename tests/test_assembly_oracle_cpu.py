"""CPU checks of the assembly oracle (line-for-line restatement of the reference loops) and of the host-side
argument handling of the fused optimizer / assembly wrappers (no GPU compute)."""
import pytest
import torch

from oracle import assembly as oa


def test_pad_patches_matches_reference_semantics():
    d = 8
    parts = [torch.ones(3, d), torch.zeros(0, d), 2 * torch.ones(1, d)]
    out = oa.pad_patches(parts, d)
    assert out.shape == (3, 3, d)
    assert torch.equal(out[0], torch.ones(3, d))
    assert float(out[1].abs().sum()) == 0.0                          # no boxes -> one zero row + zero padding
    assert torch.equal(out[2, 0], 2 * torch.ones(d)) and float(out[2, 1:].abs().sum()) == 0.0


def test_pad_text_nan_guard():
    d = 4
    a = torch.ones(2, d)
    b = torch.ones(3, d); b[1, 1] = float("nan")
    out = oa.pad_text([a, b], d)
    assert torch.equal(out[0, :2], a) and float(out[0, 2].abs().sum()) == 0.0
    assert float(out[1].abs().sum()) == 0.0                          # whole sample zeroed


def test_token_embeddings_filter_and_fallback():
    L, w, d = 6, 4, 3
    hidden = torch.arange(L * w, dtype=torch.float32).reshape(L, w)
    W = torch.eye(d, w)
    ids = torch.tensor([49406, 5, 9, 49407, 0, 0])
    mask = torch.tensor([1, 1, 1, 1, 0, 0])
    toks = oa.token_embeddings(hidden, ids, mask, W, torch.zeros(d))
    # i = 0 is BOS (skipped by i > 0); EOS at i = 3 passes the reference's filter (token_id > 0, not the LAST position)
    assert len(toks) == 3 and torch.equal(toks[0], hidden[1, :d])
    none = oa.token_embeddings(hidden, torch.tensor([49406, 0, 0, 0, 0, 0]), torch.tensor([1, 0, 0, 0, 0, 0]), W, torch.full((d,), 7.0))
    assert len(none) == 1 and torch.equal(none[0], torch.full((d,), 7.0))


def test_wrappers_reject_cpu_tensors(built_library):
    from dclip_b200 import assembly as A
    from dclip_b200.optim import FusedAdam
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        A.pad_ragged([torch.zeros(2, 8)])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        FusedAdam([torch.nn.Parameter(torch.zeros(4))])
