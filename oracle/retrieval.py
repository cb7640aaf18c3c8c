"""numpy / C restatement of the DCLIP retrieval path — TEST INFRASTRUCTURE ONLY.

Two layers:

* ``numpy_*``: the reference's own arithmetic restated with numpy (BLAS sgemm + argsort),
  following eval_scripts/flickr30k_eval.py:16-88 and :243-266.  Used for the CPU baseline
  timing and for fixtures that are tie-free (where any summation order gives the same
  ranks).  ``calculate_retrieval_metrics`` here uses the *stable* argsort; the reference's
  default ``np.argsort`` is unstable on ties, so fixtures are tie-free.
* ``c_*``: the fully specified contract in ``retrieval_oracle.c`` (sequential-fmaf fp32
  scores, ties by ascending index) which the CUDA path must match bit-for-bit, ties
  included.

faiss (compute_faiss.py:27,79; image_tokenizer.py:262) is third-party, unpinned by the
reference and absent here: PARITY UNPINNED at that boundary; ``IndexFlatIP`` semantics are
restated as exact inner product + top-k by (-score, +index).
"""
import ctypes
import os
import subprocess
from collections import defaultdict

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_c_oracle():
    """gcc-compile retrieval_oracle.c -> oracle/_build/libdclip_oracle.so (idempotent)."""
    so = os.path.join(_HERE, "_build", "libdclip_oracle.so")
    src = os.path.join(_HERE, "retrieval_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build_c_oracle())
        f32p = ctypes.POINTER(ctypes.c_float)
        i64p = ctypes.POINTER(ctypes.c_int64)
        i32p = ctypes.POINTER(ctypes.c_int32)
        i64 = ctypes.c_int64
        lib.dclip_oracle_normalize_rows.argtypes = [f32p, f32p, i64, i64]
        lib.dclip_oracle_scores.argtypes = [f32p, f32p, i64, i64, i64, f32p]
        lib.dclip_oracle_ranks.argtypes = [f32p, f32p, i64, i64, i64p, i64p, i64, i32p]
        lib.dclip_oracle_topk.argtypes = [f32p, f32p, i64, i64, i64, i64, f32p, i64p]
        for fn in (lib.dclip_oracle_normalize_rows, lib.dclip_oracle_scores,
                   lib.dclip_oracle_ranks, lib.dclip_oracle_topk):
            fn.restype = None
        _LIB = lib
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


# ----------------------------------------------------------------------------- C contract
def c_normalize_rows(x):
    x = _f32(x)
    out = np.empty_like(x)
    _lib().dclip_oracle_normalize_rows(_p(x, ctypes.c_float), _p(out, ctypes.c_float), x.shape[0], x.shape[1])
    return out


def c_scores(q, g):
    q, g = _f32(q), _f32(g)
    s = np.empty((q.shape[0], g.shape[0]), dtype=np.float32)
    _lib().dclip_oracle_scores(_p(q, ctypes.c_float), _p(g, ctypes.c_float), q.shape[0], g.shape[0], q.shape[1],
                               _p(s, ctypes.c_float))
    return s


def c_ranks(q, g, qidx, target):
    q, g = _f32(q), _f32(g)
    qidx = np.ascontiguousarray(qidx, dtype=np.int64)
    target = np.ascontiguousarray(target, dtype=np.int64)
    rank = np.empty(len(qidx), dtype=np.int32)
    _lib().dclip_oracle_ranks(_p(q, ctypes.c_float), _p(g, ctypes.c_float), g.shape[0], q.shape[1],
                              _p(qidx, ctypes.c_int64), _p(target, ctypes.c_int64), len(qidx),
                              _p(rank, ctypes.c_int32))
    return rank


def c_topk(q, g, k):
    q, g = _f32(q), _f32(g)
    s = np.empty((q.shape[0], k), dtype=np.float32)
    i = np.empty((q.shape[0], k), dtype=np.int64)
    _lib().dclip_oracle_topk(_p(q, ctypes.c_float), _p(g, ctypes.c_float), q.shape[0], g.shape[0], q.shape[1], k,
                             _p(s, ctypes.c_float), _p(i, ctypes.c_int64))
    return s, i


def metrics_from_ranks(t2i_ranks, i2t_ranks):
    """R@K / "MAP" (really MRR) exactly as flickr30k_eval.py:67-86."""
    def recall_at_k(ranks, k):
        return len([r for r in ranks if r < k]) / len(ranks)

    def mean_ap(ranks):
        return np.mean([1.0 / (r + 1) for r in ranks])

    return {
        "t2i": {"R@1": recall_at_k(t2i_ranks, 1), "R@5": recall_at_k(t2i_ranks, 5),
                "R@10": recall_at_k(t2i_ranks, 10), "MAP": mean_ap(t2i_ranks)},
        "i2t": {"R@1": recall_at_k(i2t_ranks, 1), "R@5": recall_at_k(i2t_ranks, 5),
                "R@10": recall_at_k(i2t_ranks, 10), "MAP": mean_ap(i2t_ranks)},
    }


def ranks_from_embeddings_c(caption_emb, image_emb, image_ids, caption_image_ids):
    """(t2i_ranks, i2t_ranks) under the C contract, from UN-normalised embeddings."""
    c = c_normalize_rows(caption_emb)
    im = c_normalize_rows(image_emb)
    first = {}
    for idx, iid in enumerate(image_ids):
        first.setdefault(iid, idx)                      # list.index() == first match (:34)
    gt_img = np.array([first[cid] for cid in caption_image_ids], dtype=np.int64)
    t2i = c_ranks(c, im, np.arange(len(caption_image_ids)), gt_img)
    img_to_caps = defaultdict(list)
    for ci, cid in enumerate(caption_image_ids):
        img_to_caps[cid].append(ci)
    qidx, tgt, owner = [], [], []
    for ii, iid in enumerate(image_ids):
        for ci in img_to_caps.get(iid, []):
            qidx.append(ii); tgt.append(ci); owner.append(ii)
    pair_ranks = c_ranks(im, c, np.array(qidx, dtype=np.int64), np.array(tgt, dtype=np.int64))
    best = {}
    for o, r in zip(owner, pair_ranks):
        best[o] = min(best.get(o, 1 << 30), int(r))
    i2t = [best[ii] for ii in range(len(image_ids)) if ii in best]   # images w/o captions skipped (:55-56)
    return [int(r) for r in t2i], i2t


# ------------------------------------------------------------------------ numpy restatement
def numpy_similarity(caption_emb, image_emb, chunk_size=1000):
    """flickr30k_eval.py:243-266: fp32 normalise (no eps), fp32 matmul chunks into a float64 buffer."""
    image_emb = np.array(image_emb)
    caption_emb = np.array(caption_emb)
    image_emb = image_emb / np.linalg.norm(image_emb, axis=1, keepdims=True)
    caption_emb = caption_emb / np.linalg.norm(caption_emb, axis=1, keepdims=True)
    sim = np.zeros((len(caption_emb), len(image_emb)))
    for i in range(0, len(caption_emb), chunk_size):
        for j in range(0, len(image_emb), chunk_size):
            sim[i:i + chunk_size, j:j + chunk_size] = np.matmul(caption_emb[i:i + chunk_size],
                                                                image_emb[j:j + chunk_size].T)
    return sim


def numpy_ranks(similarity_matrix, image_ids, caption_image_ids):
    """Ranks as flickr30k_eval.py:16-64 but with the stable argsort (ties by index)."""
    img_to_caps = defaultdict(list)
    for i, iid in enumerate(caption_image_ids):
        img_to_caps[iid].append(i)
    t2i = []
    for ci in range(len(caption_image_ids)):
        gt = image_ids.index(caption_image_ids[ci])
        order = np.argsort(-similarity_matrix[ci], kind="stable")
        t2i.append(int(np.where(order == gt)[0][0]))
    i2t = []
    for ii in range(len(image_ids)):
        gts = img_to_caps.get(image_ids[ii], [])
        if not gts:
            continue
        order = np.argsort(-similarity_matrix[:, ii], kind="stable")
        pos = np.empty(len(order), dtype=np.int64)
        pos[order] = np.arange(len(order))
        i2t.append(int(min(pos[g] for g in gts)))
    return t2i, i2t


def calculate_retrieval_metrics(similarity_matrix, image_ids, caption_image_ids):
    """Restated flickr30k_eval.calculate_retrieval_metrics (:16-88)."""
    return metrics_from_ranks(*numpy_ranks(similarity_matrix, list(image_ids), list(caption_image_ids)))


def numpy_topk(q, g, k):
    """Exact IP top-k by (-score, +index) on BLAS scores (for tie-free checks / CPU timing)."""
    s = np.asarray(q, dtype=np.float32) @ np.asarray(g, dtype=np.float32).T
    idx = np.argsort(-s, axis=1, kind="stable")[:, :k]
    return np.take_along_axis(s, idx, axis=1), idx.astype(np.int64)


def min_rank_gap(similarity_matrix):
    """Smallest gap between distinct sorted scores of any row/column: fixtures assert this
    is >> fp32 rounding so that ranks do not depend on summation order."""
    s = np.sort(np.asarray(similarity_matrix, dtype=np.float64), axis=1)
    g1 = np.diff(s, axis=1).min()
    s = np.sort(np.asarray(similarity_matrix, dtype=np.float64), axis=0)
    g0 = np.diff(s, axis=0).min()
    return float(min(g0, g1))


def zero_shot_counts(image_features, text_features, labels, k=5):
    """Restated eval_scripts/test_zero_shot_ImageNet.py:79-108 (torch on CPU, line for line): normalise the image
    features, 100 * (img @ text.T), topk(k), top-1 hits and the per-sample `label in indices[i]` loop."""
    import torch
    import torch.nn.functional as F
    img = F.normalize(torch.as_tensor(np.asarray(image_features, dtype=np.float32)), dim=1)
    txt = torch.as_tensor(np.asarray(text_features, dtype=np.float32))
    lab = torch.as_tensor(np.asarray(labels, dtype=np.int64))
    similarity = 100.0 * (img @ txt.T)
    values, indices = similarity.topk(min(k, txt.shape[0]), dim=1)
    top1 = int((indices[:, 0] == lab).sum().item())
    topk = 0
    for i, label in enumerate(lab):
        if label in indices[i]:
            topk += 1
    return top1, topk, values.numpy(), indices.numpy()
