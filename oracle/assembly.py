"""CPU restatement of the reference's batch-assembly loops — TEST INFRASTRUCTURE ONLY.

Line-for-line Python loops (small cases only), each citing the reference lines it follows.  The product never
imports this module.
"""
import torch


def pad_patches(patch_embeddings_list, embed_dim):
    """training/patch_text_aggregation.py:555-581, :600: zero-pad every [n_i, d] to the batch maximum and stack.
    (max_patches is the batch maximum computed at :493-497; a sample without boxes contributes one zero row :489-491.)"""
    plist = [p if p.shape[0] > 0 else torch.zeros((1, embed_dim), dtype=p.dtype) for p in patch_embeddings_list]
    max_patches = max(p.shape[0] for p in plist)
    padded = []
    for p in plist:
        num_patches = p.shape[0]
        if num_patches < max_patches:
            pad = torch.zeros((max_patches - num_patches, embed_dim), dtype=p.dtype)
            p = torch.cat([p, pad], dim=0)
        elif num_patches > max_patches:
            p = p[:max_patches]
        padded.append(p)
    return torch.stack(padded)


def pad_text(text_embeddings_list, embed_dim, nan_guard=True):
    """:542-544 (a sample holding NaN / Inf becomes zeros) and :602-616 (zero-pad tokens to the batch maximum, stack)."""
    tl = []
    for t in text_embeddings_list:
        if nan_guard and (torch.isnan(t).any() or torch.isinf(t).any()):
            t = torch.zeros_like(t)
        tl.append(t)
    max_tokens = max(t.shape[0] for t in tl)
    padded = []
    for t in tl:
        num_tokens = t.shape[0]
        if num_tokens < max_tokens:
            t = torch.cat([t, torch.zeros((max_tokens - num_tokens, embed_dim), dtype=t.dtype)], dim=0)
        elif num_tokens > max_tokens:
            t = t[:max_tokens]
        padded.append(t)
    return torch.stack(padded)


def token_embeddings(last_hidden_state, input_ids, attention_mask, text_projection_weight, pooled):
    """training/text_tokenizer.py:196-213 for ONE sample (batch element 0 of the tokenizer output): keep token i iff
    token_id > 0 and 0 < i < len - 1 and attention_mask == 1, project each kept token with text_projection
    (nn.Linear without bias); no token kept -> [pooled sentence embedding]."""
    token_embeddings_ = last_hidden_state
    masked = []
    for i, token_id in enumerate(input_ids.tolist()):
        if token_id > 0 and i > 0 and i < len(token_embeddings_) - 1 and attention_mask[i] == 1:
            masked.append(torch.nn.functional.linear(token_embeddings_[i].unsqueeze(0), text_projection_weight).squeeze(0))
    if not masked:
        return [pooled]
    return masked


def batch_text(last_hidden_state, input_ids, attention_mask, text_projection_weight, pooled):
    """The per-sample loop of patch_text_aggregation.py:500-553 reduced to its tensor work: token embeddings per
    sample (text_tokenizer.get_embeddings), stacked per sample, then padded by pad_text."""
    d = text_projection_weight.shape[0]
    per_sample = []
    for b in range(last_hidden_state.shape[0]):
        toks = token_embeddings(last_hidden_state[b], input_ids[b], attention_mask[b], text_projection_weight, pooled[b])
        per_sample.append(torch.stack(toks))
    return pad_text(per_sample, d, nan_guard=False)
