"""Seeded synthetic inputs of the BASELINE.json shapes — test/bench infrastructure.
Pure torch-CPU / numpy generators so the same tensors are rebuilt bit-identically in the build
container (golden generation) and on the GPU box (parity tests, bench)."""
import numpy as np
import torch


def loss_inputs(B, d, seed=42, dtype=torch.float32):
    """Student/teacher embeddings: S_txt = 0.5*S_img + N(0,1) gives the diagonal a signal."""
    g = torch.Generator().manual_seed(seed)
    s_img = torch.randn(B, d, generator=g)
    s_txt = 0.5 * s_img + torch.randn(B, d, generator=g)
    t_img = torch.randn(B, d, generator=g)
    t_txt = torch.randn(B, d, generator=g)
    return tuple(t.to(dtype) for t in (s_img, s_txt, t_img, t_txt))


def xattn_inputs(B, T, R, d, seed=42, ragged=False):
    """text [B,T,d], image [B,R,d] ~ N(0,1); ragged: rows beyond n_tok_i / n_reg_i are exact zeros
    (the reference zero-pads and attends the padding, patch_text_aggregation.py:555-620)."""
    g = torch.Generator().manual_seed(seed)
    text = torch.randn(B, T, d, generator=g)
    image = torch.randn(B, R, d, generator=g)
    if ragged:
        n_tok = torch.randint(min(5, T), T + 1, (B,), generator=g)
        n_reg = torch.randint(1, R + 1, (B,), generator=g)
        for b in range(B):
            text[b, int(n_tok[b]):] = 0
            image[b, int(n_reg[b]):] = 0
    return text, image


def retrieval_inputs(n_img, caps_per_img, d, sigma=1.0, seed=42):
    """Flickr-shaped eval: C_{5i+k} = I_i + sigma*N(0,1) (cfg2)."""
    rng = np.random.default_rng(seed)
    img = rng.standard_normal((n_img, d)).astype(np.float32)
    cap = (np.repeat(img, caps_per_img, axis=0)
           + sigma * rng.standard_normal((n_img * caps_per_img, d))).astype(np.float32)
    image_ids = list(range(n_img))
    caption_image_ids = [i // caps_per_img for i in range(n_img * caps_per_img)]
    return cap, img, image_ids, caption_image_ids


def gallery_inputs(n_gallery, n_query, d, noise=0.1, seed=42):
    """Normalised gallery + queries planted next to known gallery rows (cfg5 recipe, small scale)."""
    rng = np.random.default_rng(seed)
    gal = rng.standard_normal((n_gallery, d)).astype(np.float32)
    gal /= np.linalg.norm(gal, axis=1, keepdims=True)
    planted = rng.permutation(n_gallery)[:n_query]
    q = gal[planted] + noise * rng.standard_normal((n_query, d)).astype(np.float32) / np.sqrt(d)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return gal.astype(np.float32), q.astype(np.float32), planted
