"""ctypes binding of the C ABI declared in include/dclip_b200.h.

PyTorch is used for device memory and streams only: every op receives raw device
pointers (``tensor.data_ptr()``), sizes and ``torch.cuda.current_stream().cuda_stream``.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# enums of include/dclip_b200.h
DCLIP_F32, DCLIP_BF16, DCLIP_F16 = 0, 1, 2
PATH_AUTO, PATH_FP32, PATH_TENSOR = 0, 1, 2
_PATHS = {"auto": PATH_AUTO, "fp32": PATH_FP32, "tensor": PATH_TENSOR}

_vp, _i64, _i32, _f32, _sz = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_size_t

# name -> (restype, argtypes); the list is also what tests check against the header
SIGNATURES = {
    "dclip_version": (_i32, []),
    "dclip_last_error": (_i32, [ctypes.c_char_p, _sz]),
    "dclip_device_supports_tensor_path": (_i32, [_i32]),
    "dclip_distill_loss_workspace_bytes": (_sz, [_i64, _i64, _i32]),
    "dclip_distill_loss_fwd_bwd": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _i32, _f32, _f32, _f32, _f32, _f32, _vp, _vp,
                                          _vp, _vp, _vp, _sz, _i32, _vp]),
    "dclip_distill_loss_students": (_i32, [_vp, _vp, _i64, _i64, _i32, _f32, _f32, _vp, _sz, _i32, _vp]),
    "dclip_distill_loss_teachers": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _i32, _f32, _f32, _f32, _f32, _f32, _vp, _vp, _vp, _vp, _sz,
                                           _i32, _vp]),
    "dclip_distill_loss_launch_count": (_i32, [_i64, _i64, _i64, _i32, _i32]),
    "dclip_tensor_path_ready": (_i32, []),
    "dclip_resolve_loss_path": (_i32, [_i32, _i32, _i64, _i64, _f32]),
    "dclip_debug_pipeline_error": (_i32, [_vp]),
    "dclip_loss_prep": (_i32, [_vp, _vp, _i64, _i64, _i32, _vp, _i32, _vp, _vp, _vp]),
    "dclip_infonce_rowlse_workspace_bytes": (_sz, [_i64, _i64, _i64, _i32]),
    "dclip_infonce_rowlse": (_i32, [_vp, _i64, _vp, _i64, _i64, _i32, _f32, _i64, _vp, _vp, _vp, _sz, _i32, _vp]),
    "dclip_infonce_grad_workspace_bytes": (_sz, [_i64, _i64, _i64, _i32]),
    "dclip_infonce_grad": (_i32, [_vp, _i64, _vp, _i64, _i64, _i32, _f32, _i64, _vp, _vp, _f32, _vp, _vp, _sz, _i32,
                                  _vp]),
    "dclip_infonce_grad_pair_workspace_bytes": (_sz, [_i64, _i64, _i32]),
    "dclip_infonce_coeffs_bytes": (_sz, [_i64]),
    "dclip_infonce_coeffs_ld": (_i64, [_i64]),
    "dclip_infonce_coeffs_dot_bytes": (_sz, [_i64]),
    "dclip_infonce_coeffs": (_i32, [_vp, _vp, _i64, _i64, _f32, _vp, _vp, _vp, _vp, _vp]),
    "dclip_infonce_grad_from_coeffs": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _f32, _f32, _vp, _vp, _vp, _vp]),
    "dclip_infonce_grad_pair": (_i32, [_vp, _vp, _i64, _i64, _i32, _f32, _vp, _vp, _f32, _vp, _vp, _vp, _sz, _i32, _vp]),
    "dclip_loss_grad_tail": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _i32, _f32, _f32, _vp, _vp]),
    "dclip_loss_partial_sums": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    "dclip_xattn_saved_bytes": (_sz, [_i64, _i64, _i64, _i64, _i32]),
    "dclip_xattn_bwd_scratch_bytes": (_sz, [_i64, _i64, _i64, _i64, _i32]),
    "dclip_xattn_fwd": (_i32, [_vp, _vp, _i64, _i64, _i64, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "dclip_xattn_bwd": (_i32, [_vp, _vp, _i64, _i64, _i64, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz,
                               _vp]),
    "dclip_aggregate_fwd": (_i32, [_vp, _i64, _i64, _i64, _i32, _f32, _vp, _vp, _vp]),
    "dclip_aggregate_bwd": (_i32, [_vp, _vp, _vp, _i64, _i64, _i64, _i32, _f32, _vp, _vp]),
    "dclip_combine_guard_fwd": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    "dclip_combine_guard_bwd": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp, _vp]),
    "dclip_scale_grad": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp]),
    "dclip_host_distill_step": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i32, _f32, _f32, _f32, _f32, _f32, _vp, _vp, _vp, _vp, _vp, _sz, _i32,
                                       _vp, _vp, _vp, _vp, _vp, _vp]),
    "dclip_sharded_loss_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "dclip_sharded_loss_prep": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _i32, _f32, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dclip_sharded_loss_fwd": (_i32, [_vp, _vp, _i64, _i64, _i64, _f32, _i64, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "dclip_sharded_loss_bwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i32, _f32, _i64, _i32, _i32, _vp, _vp, _vp,
                                      _f32, _f32, _f32, _f32, _vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "dclip_p2p_flags_bytes": (_sz, []),
    "dclip_p2p_alloc": (_i32, [_sz, ctypes.POINTER(ctypes.c_void_p), ctypes.c_char_p]),
    "dclip_p2p_init_flags": (_i32, [_vp, _sz, _i32]),
    "dclip_p2p_open": (_i32, [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]),
    "dclip_p2p_close": (_i32, [_vp]),
    "dclip_p2p_free": (_i32, [_vp]),
    "dclip_p2p_push": (_i32, [ctypes.POINTER(ctypes.c_void_p), _i32, _i32, ctypes.POINTER(_sz), ctypes.POINTER(_sz), _i32, _sz, _i32,
                              _i32, _i32, _vp]),
    "dclip_p2p_wait": (_i32, [_vp, _sz, _i32, _i32, _i32, _vp]),
    "dclip_normalize_rows": (_i32, [_vp, _i64, _i64, _i32, _vp, _vp]),
    "dclip_pair_scores": (_i32, [_vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp]),
    "dclip_rank_counts_workspace_bytes": (_sz, [_i64, _i64, _i64, _i32]),
    "dclip_rank_counts": (_i32, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _sz, _i32, _vp]),
    "dclip_rank_counts_multi": (_i32, [_vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _i64, _i32, _i64, _vp, _vp]),
    "dclip_matrix_ranks": (_i32, [_vp, _i32, _i64, _i64, _vp, _vp, _i64, _i32, _vp, _vp]),
    "dclip_recall_metrics": (_i32, [_vp, _vp, _i64, _i64, _vp, _vp, _sz, _vp]),
    "dclip_topk_ip_workspace_bytes": (_sz, [_i64, _i64, _i64, _i64, _i32, _i32, _i32]),
    "dclip_topk_ip": (_i32, [_vp, _vp, _i64, _i64, _i64, _i64, _i32, _i32, _i64, _vp, _vp, _vp, _sz, _i32, _vp]),
    "dclip_max_row_norm": (_i32, [_vp, _i64, _i64, _i32, _vp, _vp]),
    "dclip_topk_ip_tc_workspace_bytes": (_sz, [_i64, _i64, _i64, _i32]),
    "dclip_topk_ip_tc": (_i32, [_vp, _i32, _vp, _i32, _vp, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _sz,
                                _vp]),
    "dclip_topk_merge": (_i32, [_vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp]),
    "dclip_global_embedding_flags_bytes": (_sz, [_i64]),
    "dclip_global_embedding_fwd": (_i32, [_vp, _vp, _i64, _i64, _i64, _i64, _i32, _f32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dclip_global_embedding_bwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i32, _f32, _vp, _vp, _vp]),
    "dclip_token_select": (_i32, [_vp, _vp, _i64, _i64, _vp, _vp, _vp]),
    "dclip_cu_seqlens": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp]),
    "dclip_pad_gather": (_i32, [_vp, _i32, _vp, _vp, _vp, _i64, _vp, _i32, _vp, _i64, _i64, _i64, _vp, _vp]),
    "dclip_linear_rows": (_i32, [_vp, _vp, _vp, _i64, _i64, _i64, _i32, _vp, _vp]),
    "dclip_optim_state_bytes": (_sz, []),
    "dclip_optim_workspace_bytes": (_sz, [ctypes.POINTER(_i64), _i32]),
    "dclip_optim_init_state": (_i32, [_vp, _vp]),
    "dclip_grad_clip_coef": (_i32, [ctypes.POINTER(_vp), ctypes.POINTER(_i64), _i32, _i32, _f32, _vp, _vp, _sz, _vp]),
    "dclip_adam_step": (_i32, [ctypes.POINTER(_vp), ctypes.POINTER(_vp), ctypes.POINTER(_vp), ctypes.POINTER(_vp), ctypes.POINTER(_i64),
                               _i32, _i32, ctypes.c_double, _vp, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, _i32, _vp, _vp]),
    "dclip_optim_read_state": (_i32, [_vp, ctypes.POINTER(_f32), ctypes.POINTER(_f32), ctypes.POINTER(_i32), _vp]),
}


def lib_path():
    return os.environ.get("DCLIP_B200_LIB", os.path.join(_HERE, "lib", "libdclip_b200.so"))


def load_library():
    """Load libdclip_b200.so and declare every C-ABI signature.  Raises if it is not built."""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise RuntimeError(
                f"dclip_b200: native library not found at {path}. Build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` (or `make -C dclip_b200/csrc`). "
                "There is no CPU / eager fallback.")
        lib = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)   # AttributeError here == header/library mismatch: fail loudly
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
    return _LIB


def version():
    return load_library().dclip_version()


def last_error():
    buf = ctypes.create_string_buffer(512)
    load_library().dclip_last_error(buf, 512)
    return buf.value.decode("utf-8", "replace")


def check(status, what):
    if status != 0:
        raise RuntimeError(f"dclip_b200.{what} failed (status {status}): {last_error()}")


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return DCLIP_F32
    if t.dtype == torch.bfloat16:
        return DCLIP_BF16
    if t.dtype == torch.float16:
        return DCLIP_F16
    raise TypeError(f"dclip_b200: unsupported dtype {t.dtype} (float32 / bfloat16 / float16 only)")


def path_code(path) -> int:
    if isinstance(path, int):
        return path
    try:
        return _PATHS[path]
    except KeyError:
        raise ValueError(f"dclip_b200: unknown path {path!r} (auto / fp32 / tensor)") from None


def resolve_loss_path(path, x, temperature, global_batch=None) -> int:
    """DCLIP_PATH_AUTO -> tensor / fp32 for the objective on `x` ([m, d], any supported dtype): decided by the
    library (dclip_resolve_loss_path) so that the single-GPU and the sharded entry points agree."""
    B = int(x.shape[0] if global_batch is None else global_batch)
    with torch.cuda.device(x.device):
        return load_library().dclip_resolve_loss_path(path_code(path), dtype_code(x), B, int(x.shape[1]), float(temperature))


def require_cuda(*tensors):
    """No CPU fallback: every tensor handed to an op must be a contiguous CUDA tensor on one device."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise RuntimeError("dclip_b200 ops run on a B200 only: got a non-CUDA tensor (there is no CPU fallback)")
        if not t.is_contiguous():
            raise RuntimeError("dclip_b200: tensors must be contiguous")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError("dclip_b200: all tensors of one call must live on the same device")
    return dev


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def workspace(nbytes, device):
    """Scratch owned by torch's caching allocator (stream-ordered reuse, no cudaMalloc per call)."""
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
