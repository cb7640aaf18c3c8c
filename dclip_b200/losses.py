"""Distillation objective — drop-in for the reference's loss functions.

Mirrors (same names, argument meaning and defaults):
  compute_contrastive_loss(image_embeddings, text_embeddings, temperature=0.05)
      training/CLIP_image_distillation.py:532-562 == training/train_contrastive_teacher.py:251-261
  cosine_distillation_loss(student_embeddings, teacher_embeddings)
      training/CLIP_image_distillation.py:564-576
  DistillationLoss          the combine of training_step, CLIP_image_distillation.py:594-634

All arithmetic runs in the CUDA library behind include/dclip_b200.h; the B x B logits are
never materialised.  `distill_loss_fwd_bwd` is the raw fused call (forward + backward in
one pass over the library), the autograd wrappers are thin shells around it.
"""
import ctypes

import torch
import torch.nn as nn

from . import _lib
from ._lib import check, dtype_code, path_code, ptr, require_cuda, stream_ptr, workspace

__all__ = [
    "compute_contrastive_loss", "cosine_distillation_loss", "DistillationLoss", "distill_loss_fwd_bwd",
    "sharded_distill_loss_fwd_bwd", "ShardedDistillationLoss", "CudaLossOps", "GraphedStep", "close_peer_buffers",
]


def _check_pair(a, b, what):
    if a.dim() != 2 or a.shape != b.shape or a.dtype != b.dtype:
        raise ValueError(f"{what}: expected two [B, d] tensors of equal shape and dtype, got "
                         f"{tuple(a.shape)} {a.dtype} and {tuple(b.shape)} {b.dtype}")


def distill_loss_fwd_bwd(s_img, s_txt, t_img=None, t_txt=None, temperature=0.05, w_img=1.0, w_txt=1.0, w_con=1.0,
                         grad_scale=1.0, need_grad=True, path="auto", return_stats=False):
    """Fused forward+backward of L = w_img*L_img + w_txt*L_txt + w_con*L_con on one GPU.

    Returns (losses[4] = {total, L_img, L_txt, L_con} (float32, on device), d_s_img, d_s_txt)
    with the gradients in the input dtype (None when need_grad is False).  No host sync.
    """
    _check_pair(s_img, s_txt, "distill_loss")
    if (t_img is None) != (t_txt is None):
        raise ValueError("distill_loss: give both teacher tensors or neither")
    if t_img is not None:
        _check_pair(s_img, t_img, "distill_loss")
        _check_pair(s_txt, t_txt, "distill_loss")
    dev = require_cuda(s_img, s_txt, t_img, t_txt)
    lib = _lib.load_library()
    B, d = s_img.shape
    pc = _lib.resolve_loss_path(path, s_img, temperature)      # auto: tcgen05 path for 16-bit inputs, FP32 otherwise
    losses = torch.empty(4, dtype=torch.float32, device=dev)
    d_img = torch.empty_like(s_img) if need_grad else None
    d_txt = torch.empty_like(s_txt) if need_grad else None
    stats = torch.empty(2 * B, dtype=torch.float32, device=dev) if return_stats else None
    ws = workspace(lib.dclip_distill_loss_workspace_bytes(B, d, pc), dev)
    with torch.cuda.device(dev):
        check(lib.dclip_distill_loss_fwd_bwd(ptr(s_img), ptr(s_txt), ptr(t_img), ptr(t_txt), B, d, dtype_code(s_img),
                                             float(temperature), float(w_img), float(w_txt), float(w_con),
                                             float(grad_scale), ptr(losses), ptr(d_img), ptr(d_txt), ptr(stats),
                                             ptr(ws), ws.numel(), pc, stream_ptr(dev)), "distill_loss_fwd_bwd")
    if return_stats:
        return losses, d_img, d_txt, stats
    return losses, d_img, d_txt


class HostDistillStep:
    """Distillation objective for inputs that live in (pinned) HOST memory: the four [B, d] tensors are copied
    to the device on a copy stream, students first, and the objective is split where the teachers are first
    needed (``dclip_distill_loss_students`` / ``_teachers``): the contrastive part - normalise, forward tiles,
    log-sum-exp, backward tiles, 85 % of the step - runs while the teachers are still crossing PCIe.

        step = HostDistillStep(B, d, torch.bfloat16, device)
        losses_host, d_img, d_txt = step(s_img_h, s_txt_h, t_img_h, t_txt_h)     # losses: CPU tensor [4]

    The gradients stay on the device (valid until the next call).  One host synchronisation per call (the read of
    the losses), as in the reference's ``loss.item()``."""

    def __init__(self, B, d, dtype, device, temperature=0.05, w_img=1.0, w_txt=1.0, w_con=1.0, grad_scale=1.0, path="auto"):
        self.lib = _lib.load_library()
        self.dev = torch.device(device)
        self.B, self.d, self.dtype = int(B), int(d), dtype
        self.hyper = (float(temperature), float(w_img), float(w_txt), float(w_con), float(grad_scale))
        self.bufs = [torch.empty((B, d), dtype=dtype, device=self.dev) for _ in range(4)]
        self.pc = _lib.resolve_loss_path(path, self.bufs[0], temperature)
        self.losses = torch.empty(4, dtype=torch.float32, device=self.dev)
        # pinned float[8]: the device stores the four losses and then a sequence word [4] through the buffer's mapped address
        self._losses_host8 = torch.zeros(8, dtype=torch.float32).pin_memory()
        self.losses_host = self._losses_host8[:4]
        self.d_img = torch.empty((B, d), dtype=dtype, device=self.dev)
        self.d_txt = torch.empty((B, d), dtype=dtype, device=self.dev)
        self.ws = workspace(self.lib.dclip_distill_loss_workspace_bytes(B, d, self.pc), self.dev)
        self.copy_stream = torch.cuda.Stream(self.dev)
        self.ev_students = torch.cuda.Event()
        self.ev_teacher_img = torch.cuda.Event()
        self.ev_teachers = torch.cuda.Event()
        self.ev_done = torch.cuda.Event()
        with torch.cuda.device(self.dev):
            for ev in (self.ev_students, self.ev_teacher_img, self.ev_teachers, self.ev_done):      # creates the CUDA events (torch makes them lazily)
                ev.record(torch.cuda.current_stream(self.dev))
        self._d_in = (ctypes.c_void_p * 4)(*[t.data_ptr() for t in self.bufs])
        self._size = torch.Size((self.B, self.d))
        # everything but the four host pointers and the current stream is fixed: converted once (the per-call Python sits between
        # two steps, where nothing overlaps it)
        self._fn = self.lib.dclip_host_distill_step
        self._fixed = (self._d_in, self.B, self.d, dtype_code(self.bufs[0]), *self.hyper, ptr(self.losses),
                       ctypes.c_void_p(self._losses_host8.data_ptr()), ptr(self.d_img), ptr(self.d_txt), ptr(self.ws), self.ws.numel(), self.pc)
        self._sync = tuple(ctypes.c_void_p(h) for h in (self.copy_stream.cuda_stream, self.ev_students.cuda_event, self.ev_teacher_img.cuda_event,
                                                        self.ev_teachers.cuda_event, self.ev_done.cuda_event))
        self._index = self.dev.index if self.dev.index is not None else torch.cuda.current_device()

    def __call__(self, s_img, s_txt, t_img, t_txt):
        """One step = ONE library call (dclip_host_distill_step): the four copies on the copy stream, students first; the contrastive
        part on the current stream behind the students' event, overlapped with the teachers' copies; the image side's gradient tail
        behind the image teacher, the text side's behind the text teacher - its last block stores the four losses into the pinned
        host buffer, which is the one host wait of the step."""
        size, dtype = self._size, self.dtype
        for t in (s_img, s_txt, t_img, t_txt):
            if t.shape != size or t.dtype is not dtype or t.is_cuda or not t.is_contiguous():
                raise ValueError(f"HostDistillStep: expected four contiguous host [{self.B}, {self.d}] {self.dtype} tensors")
        if torch.cuda.current_device() != self._index:
            with torch.cuda.device(self._index):
                return self.__call__(s_img, s_txt, t_img, t_txt)
        status = self._fn(s_img.data_ptr(), s_txt.data_ptr(), t_img.data_ptr(), t_txt.data_ptr(), *self._fixed,
                          torch.cuda.current_stream(self._index).cuda_stream, *self._sync)
        if status:
            check(status, "host_distill_step")
        return self.losses_host, self.d_img, self.d_txt


class _DistillLossFn(torch.autograd.Function):
    """Autograd shell: the library computes loss and gradients in one call; backward only
    scales the saved gradients by the upstream scalar."""

    @staticmethod
    def forward(ctx, s_img, s_txt, t_img, t_txt, temperature, w_img, w_txt, w_con, path):
        need = s_img.requires_grad or s_txt.requires_grad
        losses, d_img, d_txt = distill_loss_fwd_bwd(s_img.detach(), s_txt.detach(),
                                                    None if t_img is None else t_img.detach(),
                                                    None if t_txt is None else t_txt.detach(),
                                                    temperature, w_img, w_txt, w_con, 1.0, need, path)
        if need:
            ctx.save_for_backward(d_img, d_txt)
        # (total, {total, L_img, L_txt, L_con}): the 0-dim total is an output of its own, so that autograd does not have to
        # scatter its gradient back into a [4] tensor (a fill + a copy kernel per step); the components are for logging
        ctx.mark_non_differentiable(losses)
        ctx.set_materialize_grads(False)      # no zero tensor for the logging output's gradient (a fill kernel per step)
        return losses[0], losses

    @staticmethod
    def backward(ctx, g, _unused):
        d_img, d_txt = ctx.saved_tensors
        return (_scale_grad(d_img, g) if ctx.needs_input_grad[0] else None, _scale_grad(d_txt, g) if ctx.needs_input_grad[1] else None,
                None, None, None, None, None, None, None)


def _scale_grad(x, g):
    """x * g for the 0-dim upstream gradient g, on the device: one vectorised launch (torch's type-promoting
    bf16 x fp32-scalar multiply is a generic elementwise kernel, 5 us at [32, 768])."""
    if g is None:
        return None
    if x.numel() % 8 or g.numel() != 1 or not x.is_contiguous():
        return x * g
    g32 = g.detach().reshape(1)
    if g32.dtype != torch.float32:
        g32 = g32.float()
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(_lib.load_library().dclip_scale_grad(ptr(x), x.numel(), dtype_code(x), ptr(g32), ptr(out), stream_ptr(x.device)), "scale_grad")
    return out


def _contig(t):
    return t if t is None or t.is_contiguous() else t.contiguous()


def compute_contrastive_loss(image_embeddings, text_embeddings, temperature=0.05, path="auto"):
    """Symmetric InfoNCE over the B x B image-text logits (un-normalised [B, d] inputs).
    Returns a 0-dim tensor with an autograd graph, like the reference."""
    total, _ = _DistillLossFn.apply(_contig(image_embeddings), _contig(text_embeddings), None, None, temperature,
                                    0.0, 0.0, 1.0, path)
    return total


class _CosineLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, student, teacher):
        dev = require_cuda(student, teacher)
        lib = _lib.load_library()
        m, d = student.shape
        xhat = torch.empty((m, d), dtype=torch.float32, device=dev)
        inv = torch.empty(m, dtype=torch.float32, device=dev)
        cos = torch.empty(m, dtype=torch.float32, device=dev)
        out3 = torch.empty(3, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            st = stream_ptr(dev)
            check(lib.dclip_loss_prep(ptr(student), ptr(teacher), m, d, dtype_code(student), ptr(xhat), _lib.DCLIP_F32,
                                      ptr(inv), ptr(cos), st), "loss_prep")
            check(lib.dclip_loss_partial_sums(None, None, None, ptr(cos), None, m, ptr(out3), st), "loss_partial_sums")
        ctx.save_for_backward(student, teacher, inv)
        return out3[1] / m

    @staticmethod
    def backward(ctx, g):
        student, teacher, inv = ctx.saved_tensors
        lib = _lib.load_library()
        m, d = student.shape
        dx = torch.empty_like(student)
        with torch.cuda.device(student.device):
            check(lib.dclip_loss_grad_tail(ptr(student), ptr(teacher), ptr(inv), None, m, d, dtype_code(student),
                                           1.0 / m, 1.0, ptr(dx), stream_ptr(student.device)), "loss_grad_tail")
        return dx * g.to(dx.dtype), None


def cosine_distillation_loss(student_embeddings, teacher_embeddings):
    """mean_i (1 - cos(student_i, teacher_i)); gradient flows to the student only (the
    reference's teacher tensors are built under no_grad, CLIP_image_distillation.py:597-608)."""
    _check_pair(student_embeddings, teacher_embeddings, "cosine_distillation_loss")
    return _CosineLossFn.apply(_contig(student_embeddings), _contig(teacher_embeddings.detach()))


class DistillationLoss(nn.Module):
    """loss = L_img + L_txt + contrastive_weight * L_con  (CLIP_image_distillation.py:628).

    forward returns the total (0-dim, differentiable w.r.t. the student embeddings);
    ``last_components`` holds the device tensor {total, L_img, L_txt, L_con} so callers can log
    all three terms with one host read instead of the reference's three ``.item()`` syncs."""

    def __init__(self, temperature=0.05, contrastive_weight=1.0, image_weight=1.0, text_weight=1.0, path="auto"):
        super().__init__()
        self.temperature = temperature
        self.contrastive_weight = contrastive_weight
        self.image_weight = image_weight
        self.text_weight = text_weight
        self.path = path
        self.last_components = None

    def forward(self, student_image_embeddings, student_text_embeddings, teacher_image_embeddings,
                teacher_text_embeddings):
        total, losses = _DistillLossFn.apply(_contig(student_image_embeddings), _contig(student_text_embeddings),
                                             _contig(teacher_image_embeddings.detach()),
                                             _contig(teacher_text_embeddings.detach()), self.temperature,
                                             self.image_weight, self.text_weight, self.contrastive_weight, self.path)
        self.last_components = losses.detach()
        return total


# ------------------------------------------------------------------------------------------------
# data-parallel (sharded global batch) — SURVEY.md §8(e)
# ------------------------------------------------------------------------------------------------
class CudaLossOps:
    """The per-rank compute steps of the sharded objective, bound to the C ABI.  The
    choreography in `sharded_distill_loss_fwd_bwd` only sees this interface, so the CPU
    (gloo) tests can drive the same choreography with a stand-in."""

    def __init__(self, path="auto"):
        self.lib = _lib.load_library()
        self.path = path_code(path)
        # operands of the FP32 path are float32; the tensor path uses fp16 normalised rows
        self.op_torch_dtype = torch.float32 if self.path != _lib.PATH_TENSOR else torch.float16
        self.op_dtype = _lib.DCLIP_F32 if self.path != _lib.PATH_TENSOR else _lib.DCLIP_F16

    def prep(self, x, teacher, out_hat):
        dev = require_cuda(x, teacher, out_hat)
        m, d = x.shape
        inv = torch.empty(m, dtype=torch.float32, device=dev)
        cos = torch.empty(m, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(self.lib.dclip_loss_prep(ptr(x), ptr(teacher), m, d, dtype_code(x), ptr(out_hat), self.op_dtype,
                                           ptr(inv), ptr(cos), stream_ptr(dev)), "loss_prep")
        return inv, cos

    def rowlse(self, x_hat, y_all, temperature, row_offset, want_diag, out=None):
        dev = require_cuda(x_hat, y_all)
        m, d = x_hat.shape
        n = y_all.shape[0]
        lse = out if out is not None else torch.empty(m, dtype=torch.float32, device=dev)
        diag = torch.empty(m, dtype=torch.float32, device=dev) if want_diag else None
        ws = workspace(self.lib.dclip_infonce_rowlse_workspace_bytes(m, n, d, self.path), dev)
        with torch.cuda.device(dev):
            check(self.lib.dclip_infonce_rowlse(ptr(x_hat), m, ptr(y_all), n, d, self.op_dtype, float(temperature),
                                                int(row_offset), ptr(lse), ptr(diag), ptr(ws), ws.numel(), self.path,
                                                stream_ptr(dev)), "infonce_rowlse")
        return lse, diag

    def grad(self, x_hat, y_all, temperature, row_offset, lse_x, lse_y_all, coef):
        dev = require_cuda(x_hat, y_all, lse_x, lse_y_all)
        m, d = x_hat.shape
        n = y_all.shape[0]
        dx = torch.empty((m, d), dtype=torch.float32, device=dev)
        ws = workspace(self.lib.dclip_infonce_grad_workspace_bytes(m, n, d, self.path), dev)
        with torch.cuda.device(dev):
            check(self.lib.dclip_infonce_grad(ptr(x_hat), m, ptr(y_all), n, d, self.op_dtype, float(temperature),
                                              int(row_offset), ptr(lse_x), ptr(lse_y_all), float(coef), ptr(dx),
                                              ptr(ws), ws.numel(), self.path, stream_ptr(dev)), "infonce_grad")
        return dx

    def tail(self, x, teacher, inv, dx_hat, w_cos_over_B, grad_scale):
        dev = require_cuda(x, teacher, inv, dx_hat)
        m, d = x.shape
        out = torch.empty_like(x)
        with torch.cuda.device(dev):
            check(self.lib.dclip_loss_grad_tail(ptr(x), ptr(teacher), ptr(inv), ptr(dx_hat), m, d, dtype_code(x),
                                                float(w_cos_over_B), float(grad_scale), ptr(out), stream_ptr(dev)),
                  "loss_grad_tail")
        return out

    def partial_sums(self, lse_u, lse_v, diag, cos_img, cos_txt, out=None):
        dev = require_cuda(lse_u, lse_v, diag, cos_img, cos_txt)
        out = out if out is not None else torch.empty(3, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(self.lib.dclip_loss_partial_sums(ptr(lse_u), ptr(lse_v), ptr(diag), ptr(cos_img), ptr(cos_txt),
                                                   lse_u.shape[0], ptr(out), stream_ptr(dev)), "loss_partial_sums")
        return out


def sharded_distill_loss_fwd_bwd(s_img, s_txt, t_img=None, t_txt=None, temperature=0.05, w_img=1.0, w_txt=1.0,
                                 w_con=1.0, grad_scale=1.0, need_grad=True, group=None, ops=None, path="auto", comm="auto"):
    """Global-batch objective with the batch sharded over the ranks of `group`.

    Every rank passes its own [m, d] rows (equal m on all ranks; rank r owns global rows
    [r*m, (r+1)*m)).  Returns (losses[4] of the GLOBAL batch, d_s_img, d_s_txt for the local
    rows) — identical on every rank to what one process would compute on the concatenated
    batch.  Collectives: all-gather of the normalised operands (2 x [m, d]), all-gather of
    the two log-sum-exp vectors (2 x [m] floats), all-reduce of 3 floats.  Each rank
    recomputes the column block it owns, so no [B, d] gradient reduce-scatter is needed.

    comm: "p2p" = the fused phases exchange through symmetric NVLink peer memory (dclip_b200.peer: push kernels +
    flag waits, no NCCL call in the step); "nccl" = torch.distributed all-gathers; "auto" = p2p on the fused
    tensor path when the peer mapping can be set up (env DCLIP_SHARDED_COMM overrides), else nccl.
    """
    import torch.distributed as dist

    if ops is None and _fused_sharded_ok(s_img, temperature, path, group):
        return _sharded_fused(s_img, s_txt, t_img, t_txt, temperature, w_img, w_txt, w_con, grad_scale, need_grad, group, comm)
    ops = ops or CudaLossOps(path)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    m, d = s_img.shape
    B = m * world
    row_offset = rank * m
    dev = s_img.device

    # (1) normalise local rows straight into this rank's slot of the gathered operand buffers
    ops_all = torch.empty((2, B, d), dtype=ops.op_torch_dtype, device=dev)
    u_all, v_all = ops_all[0], ops_all[1]
    u_loc = u_all[row_offset:row_offset + m]
    v_loc = v_all[row_offset:row_offset + m]
    inv_u, cos_img = ops.prep(s_img, t_img, u_loc)
    inv_v, cos_txt = ops.prep(s_txt, t_txt, v_loc)
    if world > 1:
        # in-place all-gathers: every rank's input is its own slice of the output buffer
        dist.all_gather_into_tensor(u_all, u_loc, group=group)
        dist.all_gather_into_tensor(v_all, v_loc, group=group)

    # (2) forward: log-sum-exp of the local row block of S (image side) and of S^T (text side); the two
    #     lse vectors and the three partial loss sums travel in ONE small all-gather
    stat = torch.zeros((world, 2 * m + 4), dtype=torch.float32, device=dev)
    mine = stat[rank]
    lse_u, diag = ops.rowlse(u_loc, v_all, temperature, row_offset, True, out=mine[0:m])
    lse_v, _ = ops.rowlse(v_loc, u_all, temperature, row_offset, False, out=mine[m:2 * m])
    ops.partial_sums(lse_u, lse_v, diag, cos_img if t_img is not None else None, cos_txt if t_txt is not None else None,
                     out=mine[2 * m:2 * m + 3])
    if world > 1:
        dist.all_gather_into_tensor(stat.view(-1), mine.clone(), group=group)
        lse_u_all = stat[:, 0:m].reshape(B)
        lse_v_all = stat[:, m:2 * m].reshape(B)
    else:
        lse_u_all, lse_v_all = lse_u, lse_v
    sums = stat[:, 2 * m:2 * m + 3].sum(dim=0)          # fixed rank order -> same value on every rank
    l_con = sums[0] / (2.0 * B)
    l_img = sums[1] / B if t_img is not None else torch.zeros((), device=dev)
    l_txt = sums[2] / B if t_txt is not None else torch.zeros((), device=dev)
    losses = torch.stack([w_img * l_img + w_txt * l_txt + w_con * l_con, l_img, l_txt, l_con]).float()
    if not need_grad:
        return losses, None, None

    # (3) backward: score tiles recomputed; gradients of the local rows only
    coef = w_con / (2.0 * B)
    du_hat = ops.grad(u_loc, v_all, temperature, row_offset, lse_u.contiguous(), lse_v_all, coef) if w_con != 0 else None
    dv_hat = ops.grad(v_loc, u_all, temperature, row_offset, lse_v.contiguous(), lse_u_all, coef) if w_con != 0 else None
    d_img = ops.tail(s_img, t_img, inv_u, du_hat, w_img / B, grad_scale)
    d_txt = ops.tail(s_txt, t_txt, inv_v, dv_hat, w_txt / B, grad_scale)
    return losses, d_img, d_txt


def _fused_sharded_ok(x, temperature, path, group=None):
    """The fused 7-launch per-rank path needs the tensor kernels: 16-bit inputs (or path="tensor"),
    d % 128 == 0, d <= 768, temperature >= 0.025, an sm_100 device."""
    if not x.is_cuda:
        return False
    if path == "tensor":      # explicit request: any dtype, the shape rules are enforced by the library
        d = x.shape[1]
        return (d % 128 == 0 and 128 <= d <= 768 and temperature >= 0.025
                and _lib.load_library().dclip_device_supports_tensor_path(x.device.index or 0) == 1)
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    return _lib.resolve_loss_path(path, x, temperature, global_batch=x.shape[0] * world) == _lib.PATH_TENSOR


_PEER_BUFFERS = {}        # (process group object, m, n, d, device index) -> (PeerBuffer | None, ops_bytes)


def _peer_buffer(group, m, n, d, world, dev):
    """Symmetric buffer of one (group, shape): u_all | v_all [n, d] fp16, stat [world, 2m + 4] fp32, flags.
    Returns (None, ops_bytes) - on EVERY rank, and remembers it - when the peer mapping cannot be established
    anywhere: PeerBuffer's constructor agrees on success collectively after each phase, so a failure on some
    ranks never leaves the others inside a mismatched collective."""
    import warnings

    import torch.distributed as dist
    pg = group if group is not None else dist.group.WORLD
    key = (pg, m, n, d, dev.index)          # the group OBJECT (kept alive by the cache), not its id()
    if key not in _PEER_BUFFERS:
        from .peer import PeerBuffer, PeerSetupError
        ops_bytes = (n * d * 2 + 255) // 256 * 256
        stat_bytes = world * (2 * m + 4) * 4
        try:
            buf = PeerBuffer(2 * ops_bytes + stat_bytes, dev, group)
        except PeerSetupError as e:          # raised on all ranks together (e.g. cudaIpc not permitted in this container)
            warnings.warn(f"dclip_b200: peer-memory exchange unavailable ({e}); using NCCL all-gathers")
            buf = None
        _PEER_BUFFERS[key] = (buf, ops_bytes)
    return _PEER_BUFFERS[key]


def close_peer_buffers():
    """Unmap and free every cached symmetric buffer (collective: call on all ranks, before destroy_process_group)."""
    for key in sorted(_PEER_BUFFERS, key=lambda k: k[1:]):
        buf, _ = _PEER_BUFFERS.pop(key)
        if buf is not None:
            buf.close()


def _sharded_fused(s_img, s_txt, t_img, t_txt, temperature, w_img, w_txt, w_con, grad_scale, need_grad, group, comm="auto"):
    """Per-rank step on the fused C-ABI phases: prep -> [all-gather u, v] -> forward + local statistics ->
    [all-gather stats] -> global losses + backward + gradient tail.  The two exchanges are either pushes through
    NVLink peer memory (comm = "p2p": dclip_p2p_push, 3 extra small kernels, no collective) or
    NCCL all-gathers."""
    import os

    import torch.distributed as dist

    lib = _lib.load_library()
    dev = require_cuda(s_img, s_txt, t_img, t_txt)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    m, d = s_img.shape
    n = m * world
    row_offset = rank * m
    f32 = dict(dtype=torch.float32, device=dev)
    if comm == "auto":
        comm = os.environ.get("DCLIP_SHARDED_COMM", "p2p")
    pb = None
    if world > 1 and comm == "p2p" and m % 2 == 0:
        pb, ops_bytes = _peer_buffer(group, m, n, d, world, dev)
    stride = 2 * m + 4
    if pb is not None:
        import ctypes
        from .peer import KIND_DONE, KIND_OPERANDS, KIND_STATS
        p_u, p_v, p_stat = pb.ptr(0), pb.ptr(ops_bytes), pb.ptr(2 * ops_bytes)
        p_stat_row = pb.ptr(2 * ops_bytes + rank * stride * 4)
    else:
        ops_all = torch.empty((2, n, d), dtype=torch.float16, device=dev)
        u_all, v_all = ops_all[0], ops_all[1]
        stat = torch.empty((world, stride), **f32)
        p_u, p_v, p_stat, p_stat_row = ptr(u_all), ptr(v_all), ptr(stat), ptr(stat[rank])
    vecs = torch.empty((5, m), **f32)                       # inv_u, inv_v, cos_img, cos_txt, diag
    losses = torch.empty(4, **f32)
    d_img = torch.empty_like(s_img) if need_grad else None
    d_txt = torch.empty_like(s_txt) if need_grad else None
    ws = workspace(lib.dclip_sharded_loss_workspace_bytes(m, n, d), dev)
    has_t = t_img is not None
    # in-kernel exchange: the prep / finalize / tail kernels write to the peers and signal, their consumers wait in their
    # prologues - no push kernels, nothing between the three calls (16-bit inputs, gradients requested; DCLIP_SHARDED_PUSH=1
    # keeps the separate push kernels)
    fused_x = (pb is not None and need_grad and s_img.dtype != torch.float32 and d <= 1024
               and os.environ.get("DCLIP_SHARDED_PUSH", "0") != "1")
    peers = pb.peers_arg() if fused_x else None
    with torch.cuda.device(dev):
        st = stream_ptr(dev)
        check(lib.dclip_sharded_loss_prep(ptr(s_img), ptr(s_txt), ptr(t_img), ptr(t_txt), m, d, dtype_code(s_img),
                                          float(temperature), p_u, p_v, row_offset, ptr(vecs[0]), ptr(vecs[1]),
                                          ptr(vecs[2]), ptr(vecs[3]), ptr(vecs[4]), peers, st), "sharded_loss_prep")
        if fused_x:
            pass
        elif pb is not None:
            rows = (row_offset * d * 2, m * d * 2)
            # my rows -> every peer (after the peers have finished reading the previous step's operands)
            pb.push([rows, (ops_bytes + rows[0], rows[1])], KIND_DONE, KIND_OPERANDS, wait_after=True)
        elif world > 1:
            dist.all_gather_into_tensor(u_all, u_all[row_offset:row_offset + m], group=group)
            dist.all_gather_into_tensor(v_all, v_all[row_offset:row_offset + m], group=group)
        check(lib.dclip_sharded_loss_fwd(p_u, p_v, m, n, d, float(temperature), row_offset,
                                         ptr(vecs[2]) if has_t else None, ptr(vecs[3]) if has_t else None, ptr(vecs[4]),
                                         p_stat_row, ptr(ws), ws.numel(), peers, st), "sharded_loss_fwd")
        if fused_x:
            pass
        elif pb is not None:
            pb.push([(2 * ops_bytes + rank * stride * 4, stride * 4)], -1, KIND_STATS, wait_after=True)
        elif world > 1:
            dist.all_gather_into_tensor(stat.view(-1), stat[rank], group=group)
        check(lib.dclip_sharded_loss_bwd(ptr(s_img), ptr(s_txt), ptr(t_img), ptr(t_txt), p_u, p_v, m, n, d,
                                         dtype_code(s_img), float(temperature), row_offset, rank, world, p_stat, ptr(vecs[0]),
                                         ptr(vecs[1]), float(w_img), float(w_txt), float(w_con), float(grad_scale), ptr(losses),
                                         ptr(d_img), ptr(d_txt), ptr(ws), ws.numel(), peers, st), "sharded_loss_bwd")
        if pb is not None and not fused_x:
            pb.push([], -1, KIND_DONE)           # my reads of this step's operands / statistics are stream-ordered before this
    return losses, d_img, d_txt


class GraphedStep:
    """Capture a fixed-shape step (the sharded objective including its NCCL collectives, or the
    single-GPU composite) into one CUDA graph and replay it: the data-parallel step at B = 4096 is
    launch-latency bound (a few us of work per kernel), so the ~25 host-side launches and
    collective enqueues are replaced by a single graph launch.  Inputs are copied into static
    buffers; outputs are the graph's static outputs (valid until the next call)."""

    def __init__(self, fn, example_inputs, warmup=3):
        self.fn = fn
        self.static_in = [None if t is None else torch.empty_like(t) for t in example_inputs]
        for dst, src in zip(self.static_in, example_inputs):
            if dst is not None:
                dst.copy_(src)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = fn(*self.static_in)

    def __call__(self, *inputs):
        for dst, src in zip(self.static_in, inputs):
            if dst is not None:
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_out


class _ShardedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, s_img, s_txt, t_img, t_txt, temperature, w_img, w_txt, w_con, group, path):
        need = s_img.requires_grad or s_txt.requires_grad
        losses, d_img, d_txt = sharded_distill_loss_fwd_bwd(s_img.detach(), s_txt.detach(), t_img, t_txt, temperature,
                                                            w_img, w_txt, w_con, 1.0, need, group, None, path)
        if need:
            ctx.save_for_backward(d_img, d_txt)
        return losses

    @staticmethod
    def backward(ctx, g):
        d_img, d_txt = ctx.saved_tensors
        return (d_img * g[0].to(d_img.dtype), d_txt * g[0].to(d_txt.dtype)) + (None,) * 8


class ShardedDistillationLoss(DistillationLoss):
    """DistillationLoss over a global batch sharded across the ranks of a process group."""

    def __init__(self, *args, group=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.group = group

    def forward(self, s_img, s_txt, t_img, t_txt):
        losses = _ShardedFn.apply(_contig(s_img), _contig(s_txt), _contig(t_img.detach()), _contig(t_txt.detach()),
                                  self.temperature, self.image_weight, self.text_weight, self.contrastive_weight,
                                  self.group, self.path)
        self.last_components = losses.detach()
        return losses[0]
