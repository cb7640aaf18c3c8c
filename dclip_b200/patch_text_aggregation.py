"""Meta-teacher cross-attention — drop-in for the hot-path classes of
training/patch_text_aggregation.py.

  CrossModalAttention(embed_dim, num_heads)          reference :13-46
      forward(text_embedding[B,T,d], image_embedding[B,R,d]) -> (text_out, image_out)
      Holds two `nn.MultiheadAttention` and two `nn.LayerNorm` purely as PARAMETER CONTAINERS, so
      `state_dict()` has exactly the reference's 12 keys/shapes/init and teacher checkpoints load
      unchanged (they are loaded with strict=False, so a renamed key would be dropped silently).
      Their torch forward is never called: the math runs in the CUDA library.
  PatchTextAggregation                                reference :49-99, :227-265, :632-656
      .cross_attention(text, patch), .aggregation(attended, temperature=2.0),
      .global_embedding(text, patch) = the tail of compute_global_embedding_batch.
      The per-sample batch assembly (PIL, YOLO crops, frozen CLIP encoders, :268-630) is out of scope
      (SURVEY.md §8): callers hand the padded [B,T,d] / [B,R,d] tensors it produces.
"""
import ctypes
import threading

import torch
import torch.nn as nn

from . import _lib
from ._lib import check, dtype_code, ptr, require_cuda, stream_ptr, workspace

__all__ = ["CrossModalAttention", "PatchTextAggregation", "aggregation", "combine_global", "fused_global_embedding", "PARAM_ORDER",
           "GraphedTeacherStep"]

PARAM_ORDER = (
    "text_to_image.in_proj_weight", "text_to_image.in_proj_bias", "text_to_image.out_proj.weight",
    "text_to_image.out_proj.bias", "image_to_text.in_proj_weight", "image_to_text.in_proj_bias",
    "image_to_text.out_proj.weight", "image_to_text.out_proj.bias", "norm_text.weight", "norm_text.bias",
    "norm_image.weight", "norm_image.bias",
)


def _ptr_array(tensors):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


class _CrossAttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, text, image, num_heads, *params):
        dev = require_cuda(text, image, *params)
        lib = _lib.load_library()
        B, T, d = text.shape
        R = image.shape[1]
        dt = dtype_code(text)
        text_out = torch.empty_like(text)
        image_out = torch.empty_like(image)
        saved = workspace(lib.dclip_xattn_saved_bytes(B, T, R, d, dt), dev)
        with torch.cuda.device(dev):
            check(lib.dclip_xattn_fwd(ptr(text), ptr(image), B, T, R, d, num_heads, dt, _ptr_array(params), ptr(text_out),
                                      ptr(image_out), ptr(saved), saved.numel(), stream_ptr(dev)), "xattn_fwd")
        ctx.save_for_backward(text, image, saved, *params)
        ctx.num_heads = num_heads
        ctx.need_input_grad = text.requires_grad or image.requires_grad
        return text_out, image_out

    @staticmethod
    def backward(ctx, d_text_out, d_image_out):
        text, image, saved, *params = ctx.saved_tensors
        dev = text.device
        lib = _lib.load_library()
        B, T, d = text.shape
        R = image.shape[1]
        dt = dtype_code(text)
        d_text_out = d_text_out.contiguous()
        d_image_out = d_image_out.contiguous()
        grads = [torch.empty_like(p) for p in params]
        d_text = torch.empty_like(text) if ctx.need_input_grad else None
        d_image = torch.empty_like(image) if ctx.need_input_grad else None
        scratch = workspace(lib.dclip_xattn_bwd_scratch_bytes(B, T, R, d, dt), dev)
        with torch.cuda.device(dev):
            check(lib.dclip_xattn_bwd(ptr(text), ptr(image), B, T, R, d, ctx.num_heads, dt, _ptr_array(params), ptr(saved),
                                      ptr(d_text_out), ptr(d_image_out), _ptr_array(grads), ptr(d_text), ptr(d_image),
                                      ptr(scratch), scratch.numel(), stream_ptr(dev)), "xattn_bwd")
        return (d_text, d_image, None, *grads)


class _AggregationFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, temperature):
        dev = require_cuda(x)
        B, L, d = x.shape
        out = torch.empty((B, d), dtype=x.dtype, device=dev)
        w = torch.empty((B, L), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(_lib.load_library().dclip_aggregate_fwd(ptr(x), B, L, d, dtype_code(x), float(temperature), ptr(out),
                                                          ptr(w), stream_ptr(dev)), "aggregate_fwd")
        ctx.save_for_backward(x, w)
        ctx.temperature = float(temperature)
        return out

    @staticmethod
    def backward(ctx, d_out):
        x, w = ctx.saved_tensors
        B, L, d = x.shape
        dx = torch.empty_like(x)
        d_out = d_out.contiguous()
        with torch.cuda.device(x.device):
            check(_lib.load_library().dclip_aggregate_bwd(ptr(x), ptr(w), ptr(d_out), B, L, d, dtype_code(x),
                                                          ctx.temperature, ptr(dx), stream_ptr(x.device)), "aggregate_bwd")
        return dx, None


class _CombineGuardFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        dev = require_cuda(a, b)
        out = torch.empty_like(a)
        flag = torch.empty(1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            check(_lib.load_library().dclip_combine_guard_fwd(ptr(a), ptr(b), a.numel(), dtype_code(a), ptr(out), ptr(flag),
                                                              stream_ptr(dev)), "combine_guard_fwd")
        ctx.save_for_backward(flag)
        return out

    @staticmethod
    def backward(ctx, d_out):
        (flag,) = ctx.saved_tensors
        d_out = d_out.contiguous()
        da = torch.empty_like(d_out)
        db = torch.empty_like(d_out)
        with torch.cuda.device(d_out.device):
            check(_lib.load_library().dclip_combine_guard_bwd(ptr(d_out), d_out.numel(), dtype_code(d_out), ptr(flag),
                                                              ptr(da), ptr(db), stream_ptr(d_out.device)),
                  "combine_guard_bwd")
        return da, db


_COUNTERS = {}
_CAPTURE = threading.local()


def _zeroed_counter(dev):
    """A persistent zero-initialised device word: the completion counter of the fused global-embedding kernel, which every
    launch leaves zero again.  One per (device, stream) - stream order keeps the launches that share it apart; while a CUDA
    graph is being captured, the word its owner registered (GraphedTeacherStep), else None (the library then clears a
    per-call word with a memset node)."""
    if torch.cuda.is_current_stream_capturing():
        return getattr(_CAPTURE, "counter", None)
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    c = _COUNTERS.get(key)
    if c is None:
        c = _COUNTERS[key] = torch.zeros(1, dtype=torch.int32, device=dev)
    return c


class _GlobalEmbeddingFn(torch.autograd.Function):
    """aggregation(text) + aggregation(image) + 0.5 / 0.5 combine + whole-tensor NaN / Inf guard: one launch each way."""

    @staticmethod
    def forward(ctx, attended_text, attended_image, temperature):
        dev = require_cuda(attended_text, attended_image)
        lib = _lib.load_library()
        B, T, d = attended_text.shape
        R = attended_image.shape[1]
        out = torch.empty((B, d), dtype=attended_text.dtype, device=dev)
        w_t = torch.empty((B, T), dtype=torch.float32, device=dev)
        w_i = torch.empty((B, R), dtype=torch.float32, device=dev)
        flags = torch.empty(B + 1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            check(lib.dclip_global_embedding_fwd(ptr(attended_text), ptr(attended_image), B, T, R, d, dtype_code(attended_text),
                                                 float(temperature), ptr(out), ptr(w_t), ptr(w_i), ptr(flags), ptr(_zeroed_counter(dev)),
                                                 stream_ptr(dev)),
                  "global_embedding_fwd")
        ctx.save_for_backward(attended_text, attended_image, w_t, w_i, flags)
        ctx.temperature = float(temperature)
        return out

    @staticmethod
    def backward(ctx, d_out):
        attended_text, attended_image, w_t, w_i, flags = ctx.saved_tensors
        dev = attended_text.device
        B, T, d = attended_text.shape
        R = attended_image.shape[1]
        d_out = d_out.contiguous()
        d_t = torch.empty_like(attended_text)
        d_i = torch.empty_like(attended_image)
        with torch.cuda.device(dev):
            check(_lib.load_library().dclip_global_embedding_bwd(ptr(attended_text), ptr(attended_image), ptr(w_t), ptr(w_i), ptr(d_out),
                                                                 ptr(flags), B, T, R, d, dtype_code(attended_text), ctx.temperature,
                                                                 ptr(d_t), ptr(d_i), stream_ptr(dev)), "global_embedding_bwd")
        return d_t, d_i, None


def fused_global_embedding(attended_text, attended_image, temperature=2.0):
    """0.5 * aggregation(attended_text) + 0.5 * aggregation(attended_image) with the reference's whole-tensor NaN / Inf
    guard (patch_text_aggregation.py:640-656) as one fused op (embed_dim % 8 == 0, <= 1024)."""
    return _GlobalEmbeddingFn.apply(attended_text.contiguous(), attended_image.contiguous(), temperature)


def aggregation(attended_text, temperature=2.0):
    """Temperature-scaled aggregation (reference :243-265): softmax over cos(x_l, mean) / temperature."""
    if attended_text.dim() != 3:
        raise ValueError(f"aggregation: expected [B, L, d], got {tuple(attended_text.shape)}")
    return _AggregationFn.apply(attended_text.contiguous(), temperature)


def combine_global(text_global, image_global):
    """0.5 * text_global + 0.5 * image_global with the reference's whole-tensor NaN/Inf guard (:647-651),
    decided on the device (the reference's `.any()` forces a host sync)."""
    return _CombineGuardFn.apply(text_global.contiguous(), image_global.contiguous())


class CrossModalAttention(nn.Module):
    def __init__(self, embed_dim, num_heads):
        super().__init__()
        if embed_dim % num_heads != 0:
            raise ValueError("embed_dim must be divisible by num_heads")
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        # parameter containers only (same names, shapes and initialisation as the reference :14-19)
        self.text_to_image = nn.MultiheadAttention(embed_dim=embed_dim, num_heads=num_heads)
        self.image_to_text = nn.MultiheadAttention(embed_dim=embed_dim, num_heads=num_heads)
        self.norm_text = nn.LayerNorm(embed_dim)
        self.norm_image = nn.LayerNorm(embed_dim)

    def _params(self):
        sd = dict(self.named_parameters())
        return [sd[k] for k in PARAM_ORDER]

    def forward(self, text_embedding, image_embedding):
        """Bidirectional attention between text and image tokens; both directions read the ORIGINAL
        inputs and padded (all-zero) rows are attended, exactly like the reference (no masks)."""
        if text_embedding.dim() != 3 or image_embedding.dim() != 3 or text_embedding.shape[0] != image_embedding.shape[0] \
                or text_embedding.shape[2] != self.embed_dim or image_embedding.shape[2] != self.embed_dim:
            raise ValueError(f"CrossModalAttention: expected [B,T,{self.embed_dim}] and [B,R,{self.embed_dim}], got "
                             f"{tuple(text_embedding.shape)} and {tuple(image_embedding.shape)}")
        params = self._params()
        dtype = params[0].dtype
        return _CrossAttnFn.apply(text_embedding.to(dtype).contiguous(), image_embedding.to(dtype).contiguous(),
                                  self.num_heads, *params)


class PatchTextAggregation(nn.Module):
    """Hot-path part of the reference class: the trainable cross-modal attention block and the
    aggregation / combine math.  The tokenizers and the per-sample assembly loop stay with the
    caller (out of scope)."""

    def __init__(self, embed_dim=512, num_heads=8, similarity_threshold=0.85, projection_model_path=None,
                 faiss_index_path=None, embeddings_json_path=None):
        super().__init__()
        self.embed_dim = embed_dim
        self.similarity_threshold = similarity_threshold
        self.use_knn_projection = all([projection_model_path, faiss_index_path, embeddings_json_path])
        self.cross_modal_attention = CrossModalAttention(embed_dim, num_heads)

    def cross_attention(self, text_embedding, patch_embedding):
        attended_text, attended_image = self.cross_modal_attention(text_embedding, patch_embedding)
        return attended_text, attended_image

    def aggregation(self, attended_text, temperature=2.0):
        return aggregation(attended_text, temperature)

    def global_embedding(self, batch_text, batch_patches, temperature=2.0):
        """Tail of compute_global_embedding_batch (:632-656) on already assembled padded batches."""
        attended_text, attended_image = self.cross_attention(batch_text, batch_patches)
        if self.embed_dim % 8 == 0 and self.embed_dim <= 1024:
            return fused_global_embedding(attended_text, attended_image, temperature)
        return combine_global(self.aggregation(attended_text, temperature), self.aggregation(attended_image, temperature))


class GraphedTeacherStep:
    """Forward + backward of one meta-teacher step (reference training/train_contrastive_teacher.py:347-359:
    global embedding -> contrastive loss -> backward) captured in ONE CUDA graph.

    At the reference shapes (B = 32, 36-77 tokens) the step is ~60 kernels of a few microseconds each: launched one
    by one from Python it is launch-latency bound; replayed as a graph the launches cost nothing.  The graph owns
    static copies of the inputs and writes the parameter gradients into static `.grad` tensors, so the caller's
    loop becomes ``loss = step(text, patches, txt_emb); optimizer.step()`` - do not call
    ``zero_grad(set_to_none=True)`` afterwards (every replay overwrites the gradients anyway).

    loss_fn(module, *inputs) -> scalar loss tensor; shapes and dtypes are fixed at construction.
    ``step.static_in`` are the graph's input buffers: a caller that fills them in place (the target of its host-to-device
    copies) and calls ``step(*step.static_in)`` skips the device-to-device staging copies of ``step(text, patches, txt)``.
    optimizer: an optional dclip_b200.optim.FusedAdam over the same parameters; its step (two launches for all 12
    tensors, step counter and clip coefficient in device memory) is captured at the end of the graph, so one replay
    is the reference's whole ``backward(); optimizer.step()`` (train_contrastive_teacher.py:361-362).  The warm-up
    iterations do not step the optimizer.
    """

    def __init__(self, module, loss_fn, example_inputs, warmup=3, optimizer=None):
        self.module = module
        self.optimizer = optimizer
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.static_in = [t.detach().clone() for t in example_inputs]
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                for p in self.params:
                    p.grad = None
                loss_fn(module, *self.static_in).backward()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        for p in self.params:
            p.grad = None
        dev = self.static_in[0].device
        self._counter = torch.zeros(1, dtype=torch.int32, device=dev)      # the graph's own completion counter (see _zeroed_counter)
        self._one = torch.ones((), dtype=torch.float32, device=dev)          # root gradient: no ones_like fill kernel per replay
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        _CAPTURE.counter = self._counter
        try:
            with torch.cuda.graph(self.graph):
                self.static_loss = loss_fn(module, *self.static_in)
                root = self._one if self.static_loss.dtype == torch.float32 and self.static_loss.dim() == 0 else None
                self.static_loss.backward(gradient=root)
                if optimizer is not None:
                    optimizer.step()
        finally:
            _CAPTURE.counter = None

    def __call__(self, *inputs):
        for dst, src in zip(self.static_in, inputs):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_loss
