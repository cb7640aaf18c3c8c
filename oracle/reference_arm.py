"""The CPU arm of the benchmark: the reference's OWN code for each measured step — TEST / BENCH INFRASTRUCTURE.

`load()` imports the staged, unmodified reference modules (oracle/_ref, see stage_reference.py; or
/root/reference directly in the build container) through the stub recipe of import_reference.py and returns an
arm whose steps call the reference's functions as the training / eval scripts do (``kind = "reference"``).
Where no copy of the reference is reachable it returns the oracle port (``kind = "port"``), which
tests/golden pins to the reference.
"""
import contextlib
import io
import os
import types

import numpy as np
import torch

from . import reference_math as om
from . import retrieval as orc
from .import_reference import import_reference, reference_available

HERE = os.path.dirname(os.path.abspath(__file__))


class ReferenceArm:
    kind = "reference"

    def __init__(self, pta, dist, ev):
        self.pta, self.dist, self.ev = pta, dist, ev
        self.ns = types.SimpleNamespace(device="cpu")
        self._teachers = {}

    def loss_step(self, s_img, s_txt, t_img, t_txt, temperature=0.05):
        """training_step's objective + backward (CLIP_image_distillation.py:594-634)."""
        a = s_img.clone().requires_grad_(True)
        b = s_txt.clone().requires_grad_(True)
        C = self.dist.CLIPImageDistillation
        l_img = C.cosine_distillation_loss(self.ns, a, t_img)
        l_txt = C.cosine_distillation_loss(self.ns, b, t_txt)
        l_con = C.compute_contrastive_loss(self.ns, a, b, temperature)
        total = l_img + l_txt + 1.0 * l_con
        total.backward()
        return total.detach(), a.grad, b.grad

    def teacher_step(self, params, text, image, txt_emb, num_heads=8):
        """Meta-teacher step (train_contrastive_teacher.py:347-361): cross-attention -> aggregation x2 -> combine
        -> contrastive loss -> backward, on the reference's module."""
        key = (text.shape[-1], num_heads)
        if key not in self._teachers:
            self._teachers[key] = self.pta.CrossModalAttention(text.shape[-1], num_heads)
        mod = self._teachers[key]
        mod.load_state_dict(params, strict=True)
        for p in mod.parameters():
            p.grad = None
        t_out, i_out = mod(text, image)
        agg = self.pta.PatchTextAggregation.aggregation
        glob = 0.5 * agg(None, t_out) + 0.5 * agg(None, i_out)
        loss = self.dist.CLIPImageDistillation.compute_contrastive_loss(self.ns, glob, txt_emb)
        loss.backward()
        return loss.detach()

    def retrieval_eval(self, cap, img, image_ids, cap_ids):
        """flickr30k_eval.py:249-269: normalise, chunked similarity into a float64 buffer, calculate_retrieval_metrics."""
        ie = img / np.linalg.norm(img, axis=1, keepdims=True)
        ce = cap / np.linalg.norm(cap, axis=1, keepdims=True)
        sim = np.zeros((len(ce), len(ie)))
        chunk = 1000
        for i in range(0, len(ce), chunk):
            for j in range(0, len(ie), chunk):
                sim[i:i + chunk, j:j + chunk] = np.matmul(ce[i:i + chunk], ie[j:j + chunk].T)
        with contextlib.redirect_stdout(io.StringIO()):
            return self.ev.calculate_retrieval_metrics(sim, image_ids, cap_ids)


class PortArm:
    kind = "port"

    def loss_step(self, s_img, s_txt, t_img, t_txt, temperature=0.05):
        (tot, _, _, _), g_img, g_txt = om.distill_total_with_grads(s_img, s_txt, t_img, t_txt, temperature)
        return tot, g_img, g_txt

    def teacher_step(self, params, text, image, txt_emb, num_heads=8):
        return om.teacher_step(params, text, image, txt_emb, num_heads)[0]

    def retrieval_eval(self, cap, img, image_ids, cap_ids):
        return orc.calculate_retrieval_metrics(orc.numpy_similarity(cap, img), image_ids, cap_ids)


def topk_stand_in(q, g, k):
    """faiss is absent: exact inner product + top-k with torch on the host cores, the stand-in
    BASELINE.md names for IndexFlatIP.search.  Same for both arms."""
    s = torch.from_numpy(q) @ torch.from_numpy(g).T
    return torch.topk(s, k, dim=1)


def load():
    staged = os.path.join(HERE, "_ref")
    if not reference_available() and os.path.isdir(os.path.join(staged, "training")):
        os.environ["DCLIP_REFERENCE_ROOT"] = staged
        from . import import_reference as ir
        ir.REFERENCE_ROOT = staged
    if reference_available():
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                return ReferenceArm(*import_reference())
        except Exception as e:          # noqa: BLE001 - any import problem of the third-party stack: say so, use the port
            print(f"reference_arm: importing the staged reference failed ({type(e).__name__}: {e}); using the oracle port", flush=True)
    return PortArm()
