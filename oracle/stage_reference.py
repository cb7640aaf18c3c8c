"""Stage the UNMODIFIED reference hot-path modules for the CPU baseline — build-time recipe.

    python -m oracle.stage_reference          (also run by __graft_entry__.build())

The reference is pure Python: nothing to compile, so "building" it means copying the files whose functions
the hot path consists of, byte for byte, from /root/reference into the git-ignored ``oracle/_ref/`` (same
sub-directories).  ``oracle/_ref/`` is NOT gpurun-ignored: it travels to the GPU box with the snapshot, where
``bench.py --impl reference`` / ``cpu_baseline`` import it through the stub recipe of import_reference.py and
time the reference's own code on the box's host cores (``kind: "reference"``).  When the directory is absent
(fresh clone without /root/reference) those legs fall back to the oracle port and say ``kind: "port"``.
The sources are never committed; the manifest written next to them records their sha256.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("DCLIP_REFERENCE_ROOT", "/root/reference")
STAGE = os.path.join(HERE, "_ref")

# the modules that hold the path's functions; their third-party imports are stubbed at import time
FILES = (
    "training/patch_text_aggregation.py",       # CrossModalAttention, PatchTextAggregation.aggregation
    "training/CLIP_image_distillation.py",      # compute_contrastive_loss, cosine_distillation_loss
    "eval_scripts/flickr30k_eval.py",           # calculate_retrieval_metrics
    "eval_scripts/test_zero_shot_ImageNet.py",  # zero-shot scoring lines (read only; a script, not imported)
    "training/train_contrastive_teacher.py",    # teacher-step closure (read only; nested in main)
    "training/compute_faiss.py",                # faiss call sites (read only; a script)
)


def stage(verbose=True):
    if not os.path.isdir(os.path.join(REF_ROOT, "training")):
        if verbose:
            print(f"stage_reference: {REF_ROOT} not present, nothing staged")
        return False
    manifest = {}
    for rel in FILES:
        src = os.path.join(REF_ROOT, rel)
        dst = os.path.join(STAGE, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    with open(os.path.join(STAGE, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF_ROOT, "sha256": manifest}, f, indent=1)
    if verbose:
        print(f"stage_reference: {len(manifest)} files -> {STAGE}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
