"""Import the UNMODIFIED DCLIP reference modules from /root/reference on CPU.

Build-container only (the GPU box has no /root/reference).  Used by make_golden.py and
by the opt-in tests that compare the oracle restatement with the real reference.

Recipe: pre-seed ``sys.modules`` with tiny stubs for packages that are
absent here (ultralytics, faiss, gensim, pytorch_lightning) and for ``image_tokenizer`` /
``text_tokenizer`` (they load HF weights at construction; CLIP_image_distillation.py:30
even instantiates one at import time).  The hot-path classes/functions themselves are the
reference's own code, executed as-is.
"""
import importlib
import os
import sys
import types

import torch.nn as nn

REFERENCE_ROOT = os.environ.get("DCLIP_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "training"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _Dummy:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        return _Dummy()


def install_stubs():
    if "ultralytics" not in sys.modules:
        _stub("ultralytics", YOLO=_Dummy)
    if "faiss" not in sys.modules:
        _stub("faiss", IndexFlatIP=_Dummy, read_index=_Dummy, write_index=_Dummy)
    if "gensim" not in sys.modules:
        g = _stub("gensim")
        g.downloader = _stub("gensim.downloader", load=_Dummy)
    if "pytorch_lightning" not in sys.modules:
        class LightningModule(nn.Module):
            def save_hyperparameters(self, *a, **k):
                pass

            def log(self, *a, **k):
                pass
        pl = _stub("pytorch_lightning", LightningModule=LightningModule,
                   seed_everything=lambda s: None, Trainer=_Dummy)
        pl.core = _stub("pytorch_lightning.core", LightningModule=LightningModule)
        pl.callbacks = _stub("pytorch_lightning.callbacks", ModelCheckpoint=_Dummy)
    # frozen third-party encoders: out of scope, replaced by inert stand-ins
    _stub("image_tokenizer", CLIPPatchTokenizer=_Dummy, TokenizerWithKNN=_Dummy,
          ImageTokenizer=_Dummy)
    _stub("text_tokenizer", CLIPTextTokenizer=_Dummy)


def import_reference():
    """Returns (patch_text_aggregation, CLIP_image_distillation, flickr30k_eval) modules."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    install_stubs()
    for sub in ("training", "eval_scripts"):
        p = os.path.join(REFERENCE_ROOT, sub)
        if p not in sys.path:
            sys.path.insert(0, p)
    pta = importlib.import_module("patch_text_aggregation")
    dist = importlib.import_module("CLIP_image_distillation")
    ev = importlib.import_module("flickr30k_eval")
    return pta, dist, ev
