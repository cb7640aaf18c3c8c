"""Fused optimizer for the teacher / distillation step — drop-in for the optimizers the reference constructs.

  FusedAdam(params, lr=1e-5)                          torch.optim.Adam at training/train_contrastive_teacher.py:245-248
  FusedAdam(params, lr, weight_decay=1e-2, adamw=True, max_grad_norm=0.5)
                                                      AdamW at training/CLIP_image_distillation.py:679-682 with Lightning's
                                                      gradient_clip_val=0.5 (training/CLIP_image_distill_training.py:41-42)

``step()`` is two kernel launches for all parameter tensors together (one without clipping) instead of one
multi-kernel foreach group per tensor: a global-norm pass that leaves ``clip_coef`` in device memory, and the
Adam update that applies it.  The step counter also lives on the device, so ``step()`` can be captured into a CUDA
graph together with the forward/backward (``GraphedTeacherStep(..., optimizer=opt)``).  Same defaults and update rule
as torch (betas=(0.9, 0.999), eps=1e-8, amsgrad=False); moments are fp32.  There is no CPU fallback.
"""
import ctypes

import torch

from . import _lib
from ._lib import check, dtype_code, require_cuda, stream_ptr

__all__ = ["FusedAdam", "FusedAdamW"]

_MAX_TENSORS = 64


def _ptrs(tensors):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


class FusedAdam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, adamw=False, max_grad_norm=None):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FusedAdam: no trainable parameters")
        dev = require_cuda(*self.params)
        dtypes = {p.dtype for p in self.params}
        if len(dtypes) != 1:
            raise ValueError(f"FusedAdam: parameters must share one dtype, got {dtypes}")
        self.device = dev
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.adamw = bool(adamw)
        self.max_grad_norm = None if max_grad_norm is None else float(max_grad_norm)
        self.lib = _lib.load_library()
        self.exp_avg = [torch.zeros(p.shape, dtype=torch.float32, device=dev) for p in self.params]
        self.exp_avg_sq = [torch.zeros(p.shape, dtype=torch.float32, device=dev) for p in self.params]
        self.lr_dev = None                     # optional device scalar (schedulers inside a captured graph)
        numel = [p.numel() for p in self.params]
        self._groups = [range(i, min(i + _MAX_TENSORS, len(numel))) for i in range(0, len(numel), _MAX_TENSORS)]
        self._numel = (ctypes.c_int64 * len(numel))(*numel)
        with torch.cuda.device(dev):
            self.state = torch.zeros(self.lib.dclip_optim_state_bytes() * len(self._groups), dtype=torch.uint8, device=dev)
            self.ws = torch.empty(max(self.lib.dclip_optim_workspace_bytes(self._numel, len(numel)), 256), dtype=torch.uint8, device=dev)
            for gi in range(len(self._groups)):
                check(self.lib.dclip_optim_init_state(self._state_ptr(gi), stream_ptr(dev)), "optim_init_state")
        if len(self._groups) > 1 and self.max_grad_norm is not None:
            raise NotImplementedError("FusedAdam: global-norm clipping supports up to 64 tensors")

    def _state_ptr(self, gi):
        return ctypes.c_void_p(self.state.data_ptr() + gi * self.lib.dclip_optim_state_bytes())

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def set_lr(self, lr):
        """Host-side schedule (e.g. the reference's linear warm-up): takes effect at the next eager step()."""
        self.lr = float(lr)
        if self.lr_dev is not None:
            self.lr_dev.fill_(self.lr)

    def use_device_lr(self):
        """Keep the learning rate in a device scalar so that a captured graph sees schedule updates."""
        if self.lr_dev is None:
            self.lr_dev = torch.full((1,), self.lr, dtype=torch.float32, device=self.device)
        return self.lr_dev

    @torch.no_grad()
    def step(self):
        grads = []
        for p in self.params:
            if p.grad is None:
                raise RuntimeError("FusedAdam.step: a parameter has no gradient (all tensors are updated together)")
            g = p.grad
            if g.dtype != p.dtype or not g.is_contiguous():
                raise RuntimeError("FusedAdam.step: gradients must be contiguous and of the parameter dtype")
            grads.append(g)
        dev = self.device
        dt = dtype_code(self.params[0])
        with torch.cuda.device(dev):
            st = stream_ptr(dev)
            for gi, idx in enumerate(self._groups):
                ps = [self.params[i] for i in idx]
                gs = [grads[i] for i in idx]
                n = len(ps)
                numel = (ctypes.c_int64 * n)(*[p.numel() for p in ps])
                if self.max_grad_norm is not None:
                    check(self.lib.dclip_grad_clip_coef(_ptrs(gs), numel, n, dt, self.max_grad_norm, self._state_ptr(gi),
                                                        ctypes.c_void_p(self.ws.data_ptr()), self.ws.numel(), st), "grad_clip_coef")
                check(self.lib.dclip_adam_step(_ptrs(ps), _ptrs(gs), _ptrs([self.exp_avg[i] for i in idx]),
                                               _ptrs([self.exp_avg_sq[i] for i in idx]), numel, n, dt, self.lr,
                                               None if self.lr_dev is None else ctypes.c_void_p(self.lr_dev.data_ptr()),
                                               self.betas[0], self.betas[1], self.eps, self.weight_decay, 1 if self.adamw else 0,
                                               self._state_ptr(gi), st), "adam_step")

    def read_state(self):
        """{grad_norm, clip_coef, step} of the last step (one host synchronisation; logging / tests)."""
        gn, cc, stp = ctypes.c_float(), ctypes.c_float(), ctypes.c_int()
        with torch.cuda.device(self.device):
            check(self.lib.dclip_optim_read_state(self._state_ptr(0), ctypes.byref(gn), ctypes.byref(cc), ctypes.byref(stp),
                                                  stream_ptr(self.device)), "optim_read_state")
        return {"grad_norm": gn.value, "clip_coef": cc.value, "step": stp.value}


class FusedAdamW(FusedAdam):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=None):
        super().__init__(params, lr, betas, eps, weight_decay, adamw=True, max_grad_norm=max_grad_norm)
