"""The C-ABI library loads and exports every symbol include/dclip_b200.h declares (no compute)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "dclip_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dclip_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_entry_points():
    syms = header_symbols()
    assert "dclip_distill_loss_fwd_bwd" in syms and "dclip_topk_ip" in syms and len(syms) >= 15


def test_library_exports_every_header_symbol(built_library):
    for name in header_symbols():
        assert hasattr(built_library, name), f"{name} declared in the header but not exported"


def test_binding_covers_header(built_library):
    from dclip_b200 import _lib
    assert sorted(_lib.SIGNATURES) == header_symbols()


def test_version_and_error_string(built_library):
    assert built_library.dclip_version() == 100
    buf = ctypes.create_string_buffer(64)
    assert built_library.dclip_last_error(buf, 64) == 0


def test_workspace_queries_need_no_gpu(built_library):
    # pure host arithmetic: safe on a CPU-only machine (sm_count falls back to 148)
    assert built_library.dclip_distill_loss_workspace_bytes(4096, 768, 1) > 4096 * 768 * 4 * 4
    assert built_library.dclip_topk_ip_workspace_bytes(100, 100000, 768, 10, 0, 0, 1) > 0


def test_cpu_tensors_are_rejected(built_library):
    import torch
    from dclip_b200 import losses, retrieval
    a = torch.randn(4, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        losses.distill_loss_fwd_bwd(a, a)
    with pytest.raises(RuntimeError):
        retrieval.topk_ip(a, a, 2)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from dclip_b200 import _lib
    monkeypatch.setenv("DCLIP_B200_LIB", str(tmp_path / "nope.so"))
    monkeypatch.setattr(_lib, "_LIB", None)
    with pytest.raises(RuntimeError, match="native library not found"):
        _lib.load_library()
