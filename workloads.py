"""Seeded synthetic workloads of the BASELINE.json shapes: inputs and parameter initialisations only - no
arithmetic of the hot path lives here.  Neutral ground shared by bench.py (GPU arm and CPU arm), the tests, the oracle's golden
generator and the tools, so that the product side of the bench never imports the checker (`oracle/`).  Pure torch-CPU / numpy
generators: the same tensors are rebuilt bit-identically in the build container (golden generation) and on the GPU box."""
import math

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


def init_cross_attention_params(embed_dim, num_heads=8, seed=42, dtype=torch.float32):
    """State dict with the reference's 12 keys and its initialisation
    (training/patch_text_aggregation.py:14-19 -> nn.MultiheadAttention._reset_parameters:
    xavier_uniform in_proj, zero biases; out_proj = nn.Linear default with zero bias;
    LayerNorm ones/zeros).  Same distributions as the reference's initialisation, drawn from a
    private generator (fixtures store or re-derive the exact tensors, so no RNG-stream
    compatibility with the reference is needed)."""
    del num_heads
    g = torch.Generator().manual_seed(seed)
    d = embed_dim
    sd = {}
    for name in ("text_to_image", "image_to_text"):
        bound = math.sqrt(6.0 / (d + 3 * d))
        sd[f"{name}.in_proj_weight"] = (torch.rand(3 * d, d, generator=g) * 2 - 1) * bound
        sd[f"{name}.in_proj_bias"] = torch.zeros(3 * d)
        kb = 1.0 / math.sqrt(d)
        sd[f"{name}.out_proj.weight"] = (torch.rand(d, d, generator=g) * 2 - 1) * kb
        sd[f"{name}.out_proj.bias"] = torch.zeros(d)
    for name in ("norm_text", "norm_image"):
        sd[f"{name}.weight"] = torch.ones(d)
        sd[f"{name}.bias"] = torch.zeros(d)
    return {k: v.to(dtype) for k, v in sd.items()}


def perturbed_cross_attention_params(embed_dim, num_heads=8, seed=42, dtype=torch.float32):
    """init_cross_attention_params plus N(0, 0.1) noise on biases and LayerNorm affine, so every one
    of the 12 tensors influences the output (fixtures and parity tests)."""
    sd = init_cross_attention_params(embed_dim, num_heads, seed, torch.float32)
    g = torch.Generator().manual_seed(seed + 1000)
    for k in sd:
        if "bias" in k or "norm" in k:
            sd[k] = sd[k] + 0.1 * torch.randn(sd[k].shape, generator=g)
    return {k: v.to(dtype) for k, v in sd.items()}
