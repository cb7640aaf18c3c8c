"""Ragged batch assembly feeding the cross-attention — the step immediately upstream of the hot path.

Mirrors the per-sample Python loops of the reference as varlen device ops:

  pad_ragged(list_of_[n_i, d]) -> [B, max_n, d]
      the zero-pad / truncate / stack of region embeddings (training/patch_text_aggregation.py:555-581, :600) and of
      token embeddings (:602-616): rows beyond n_i are EXACT zeros (and stay attended: the reference has no mask),
      a sample with no rows becomes all zeros (= the reference's single zero row + padding, :489-491)
  project_text_tokens(last_hidden_state, input_ids, attention_mask, text_projection_weight, pooled=None)
      the token filter + per-token `text_projection` loop of training/text_tokenizer.py:196-213 for a whole batch:
      ONE GEMM over all B x L hidden rows, one filter kernel, one gather that writes the padded [B, T, d] batch
      (T = the batch maximum of kept tokens); `pooled` [B, d] is the sentence-embedding fallback for samples with no
      kept token; nan_guard zeroes a sample whose embedding holds a NaN / Inf (patch_text_aggregation.py:542-544).

Everything runs in the CUDA library (dclip_pad_gather, dclip_token_select, dclip_cu_seqlens, dclip_linear_rows).
"""
import torch

from . import _lib
from ._lib import check, dtype_code, ptr, require_cuda, stream_ptr

__all__ = ["pad_ragged", "pad_flat", "linear_rows", "select_tokens", "project_text_tokens"]


def pad_flat(flat, lengths, max_len=None, nan_guard=False):
    """flat [sum(lengths), d] + per-sample lengths (list / int tensor) -> ([B, max_len, d], sample_flags or None)."""
    dev = require_cuda(flat)
    lib = _lib.load_library()
    if not isinstance(lengths, torch.Tensor):
        lengths_host = [int(x) for x in lengths]
        lengths = torch.tensor(lengths_host, dtype=torch.int32)
    else:
        lengths_host = None
    lengths = lengths.to(device=dev, dtype=torch.int32).contiguous()
    B = lengths.shape[0]
    d = flat.shape[1]
    if max_len is None:      # the reference takes the batch maximum on the host (patch_text_aggregation.py:602)
        max_len = max(lengths_host) if lengths_host is not None else int(lengths.max())
    max_len = max(int(max_len), 1)
    cu = torch.empty(B + 1, dtype=torch.int32, device=dev)
    out = torch.empty((B, max_len, d), dtype=flat.dtype, device=dev)
    flags = torch.empty(B, dtype=torch.int32, device=dev) if nan_guard else None
    with torch.cuda.device(dev):
        st = stream_ptr(dev)
        check(lib.dclip_cu_seqlens(ptr(lengths), B, 0, ptr(cu), None, st), "cu_seqlens")
        check(lib.dclip_pad_gather(ptr(flat), dtype_code(flat), ptr(cu), ptr(lengths), None, 0, None, 1 if nan_guard else 0, ptr(flags),
                                   B, max_len, d, ptr(out), st), "pad_gather")
    return (out, flags) if nan_guard else out


def pad_ragged(tensors, max_len=None, nan_guard=False):
    """List of per-sample [n_i, d] CUDA tensors (n_i may be 0) -> zero-padded [B, max_len, d]."""
    tensors = [t if t.dim() == 2 else t.reshape(-1, t.shape[-1]) for t in tensors]
    flat = torch.cat(tensors, dim=0).contiguous()
    return pad_flat(flat, [t.shape[0] for t in tensors], max_len, nan_guard)


def linear_rows(x, weight, bias=None):
    """x [n, k] @ weight[m, k].T (+ bias) as one GEMM (tcgen05 for 16-bit inputs, exact fp32 otherwise)."""
    dev = require_cuda(x, weight, bias)
    if x.dtype != weight.dtype:
        raise TypeError("linear_rows: x and weight must share a dtype")
    n, k = x.shape
    m = weight.shape[0]
    out = torch.empty((n, m), dtype=x.dtype, device=dev)
    b32 = None if bias is None else bias.float().contiguous()
    with torch.cuda.device(dev):
        check(_lib.load_library().dclip_linear_rows(ptr(x), ptr(weight), ptr(b32), n, k, m, dtype_code(x), ptr(out), stream_ptr(dev)),
              "linear_rows")
    return out


def select_tokens(input_ids, attention_mask):
    """-> (lengths int32[B], positions int32[B, L]) of the tokens text_tokenizer.py:199-201 keeps."""
    dev = require_cuda(input_ids, attention_mask)
    B, L = input_ids.shape
    ids = input_ids.to(torch.int64).contiguous()
    mask = attention_mask.to(torch.int64).contiguous()
    lengths = torch.empty(B, dtype=torch.int32, device=dev)
    positions = torch.empty((B, L), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.load_library().dclip_token_select(ptr(ids), ptr(mask), B, L, ptr(lengths), ptr(positions), stream_ptr(dev)), "token_select")
    return lengths, positions


def project_text_tokens(last_hidden_state, input_ids, attention_mask, text_projection_weight, pooled=None, max_tokens=None,
                        nan_guard=False):
    """[B, L, width] hidden states -> padded batch_text [B, T, d] of projected, filtered tokens.
    T = batch maximum of kept tokens (one host read, like the reference's max(...) :602) unless max_tokens is given."""
    dev = require_cuda(last_hidden_state, input_ids, attention_mask, text_projection_weight, pooled)
    B, L, width = last_hidden_state.shape
    lib = _lib.load_library()
    proj = linear_rows(last_hidden_state.reshape(B * L, width).contiguous(), text_projection_weight.contiguous())      # [B*L, d]
    d = proj.shape[1]
    lengths, positions = select_tokens(input_ids, attention_mask)
    eff = torch.clamp(lengths, min=1) if pooled is not None else lengths
    T = int(max_tokens) if max_tokens is not None else max(int(eff.max()), 1)
    out = torch.empty((B, T, d), dtype=proj.dtype, device=dev)
    flags = torch.empty(B, dtype=torch.int32, device=dev) if nan_guard else None
    fb = None if pooled is None else pooled.to(proj.dtype).contiguous()
    with torch.cuda.device(dev):
        check(lib.dclip_pad_gather(ptr(proj), dtype_code(proj), None, ptr(lengths), ptr(positions), L, ptr(fb), 1 if nan_guard else 0,
                                   ptr(flags), B, T, d, ptr(out), stream_ptr(dev)), "pad_gather")
    return out
