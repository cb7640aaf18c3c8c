"""Cosine retrieval — drop-in for the reference's evaluation / KNN functions.

Mirrors:
  calculate_retrieval_metrics(similarity_matrix, image_ids, caption_image_ids)
      eval_scripts/flickr30k_eval.py:16-88  (kept for compatibility: takes a materialised matrix)
  retrieval_metrics_from_embeddings(caption_emb, image_emb, image_ids, caption_image_ids)
      the fused form of flickr30k_eval.py:243-269: normalise -> scores -> ranks -> R@K without
      ever building the [n_cap, n_img] similarity matrix
  IndexFlatIP(d).add(x) / .search(q, k) / .ntotal
      the faiss surface used at training/compute_faiss.py:27,79 and training/image_tokenizer.py:262

Score contract (bit-exact against oracle/retrieval_oracle.c): fp32 fmaf chain over k ascending,
order (-score, +index).  Rank of a target = #{s_j > s_t} + #{j < t : s_j == s_t}, i.e. the
position in ``np.argsort(-s, kind="stable")``; partial ranks over gallery shards add up.
"""
from collections import defaultdict

import numpy as np
import torch

from . import _lib
from ._lib import check, dtype_code, path_code, ptr, require_cuda, stream_ptr, workspace

__all__ = ["calculate_retrieval_metrics", "retrieval_metrics_from_embeddings", "retrieval_ranks", "normalize_rows",
           "pair_scores", "rank_counts", "matrix_ranks", "recall_metrics", "topk_ip", "topk_merge", "IndexFlatIP",
           "TensorGallery", "tensor_path_ok",
           "build_eval_index", "device_eval_index", "sharded_topk_ip", "sharded_retrieval_ranks"]


def _as_cuda(x, device=None, dtype=None):
    """numpy / CPU tensors are accepted at this boundary (the reference hands numpy arrays)
    and moved to the GPU; compute never happens on the host."""
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(x)
    if dtype is not None and x.dtype != dtype:
        x = x.to(dtype)
    if not x.is_cuda:
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("dclip_b200.retrieval needs a CUDA device (there is no CPU fallback)")
            device = torch.device("cuda", torch.cuda.current_device())
        x = x.to(device, non_blocking=True)
    return x.contiguous()


# ------------------------------------------------------------------------------------------------
# thin wrappers over the C ABI
# ------------------------------------------------------------------------------------------------
def normalize_rows(x):
    """x / ||x|| per row with the contract's arithmetic (numpy semantics: no eps). -> float32."""
    dev = require_cuda(x)
    n, d = x.shape
    out = torch.empty((n, d), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.load_library().dclip_normalize_rows(ptr(x), n, d, dtype_code(x), ptr(out), stream_ptr(dev)),
              "normalize_rows")
    return out


def pair_scores(q, g, gidx, qidx=None):
    dev = require_cuda(q, g, gidx, qidx)
    n = gidx.shape[0]
    out = torch.empty(n, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.load_library().dclip_pair_scores(ptr(q), ptr(g), q.shape[1], ptr(qidx), ptr(gidx), n, ptr(out),
                                                    stream_ptr(dev)), "pair_scores")
    return out


def rank_counts(q, g, target_score, target_index, qidx=None, g_index_offset=0, path="fp32"):
    """Partial ranks of the targets over gallery shard `g` (int32[npairs])."""
    dev = require_cuda(q, g, target_score, target_index, qidx)
    lib = _lib.load_library()
    npairs = target_index.shape[0]
    ng, d = g.shape
    rank = torch.empty(npairs, dtype=torch.int32, device=dev)
    pc = path_code(path)
    ws = workspace(lib.dclip_rank_counts_workspace_bytes(npairs, ng, d, pc), dev)
    with torch.cuda.device(dev):
        check(lib.dclip_rank_counts(ptr(q), ptr(g), ng, d, ptr(qidx), ptr(target_score), ptr(target_index), npairs,
                                    int(g_index_offset), ptr(rank), ptr(ws), ws.numel(), pc, stream_ptr(dev)),
              "rank_counts")
    return rank


def rank_counts_multi(q, g, target_ptr, target_score, target_index, max_targets, g_index_offset=0):
    """Partial ranks for query rows with several targets each (CSR target_ptr [nq + 1]); int32[npairs]."""
    dev = require_cuda(q, g, target_ptr, target_score, target_index)
    npairs = target_index.shape[0]
    rank = torch.empty(npairs, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.load_library().dclip_rank_counts_multi(ptr(q), ptr(g), q.shape[0], g.shape[0], g.shape[1], ptr(target_ptr),
                                                          ptr(target_score), ptr(target_index), npairs, int(max_targets),
                                                          int(g_index_offset), ptr(rank), stream_ptr(dev)), "rank_counts_multi")
    return rank


def recall_metrics(rank, group=None, n_groups=0):
    """-> float64[4] device tensor {R@1, R@5, R@10, MRR}."""
    dev = require_cuda(rank, group)
    lib = _lib.load_library()
    out = torch.empty(4, dtype=torch.float64, device=dev)
    ws = workspace(4 * max(int(n_groups), 1) + 64, dev)
    with torch.cuda.device(dev):
        check(lib.dclip_recall_metrics(ptr(rank), ptr(group), rank.shape[0], int(n_groups), ptr(out), ptr(ws),
                                       ws.numel(), stream_ptr(dev)), "recall_metrics")
    return out


class TensorGallery:
    """What the tensor path needs besides the stored gallery: its float16 image for the tcgen05 pass
    (the gallery itself when it is stored in float16) and the largest row norm for the exactness
    certificate.  Built once per index / shard and reused by every search."""

    def __init__(self, g):
        dev = require_cuda(g)
        self.g = g
        self.g16 = g if g.dtype == torch.float16 else g.to(torch.float16)
        self.gmax = torch.empty(1, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(_lib.load_library().dclip_max_row_norm(ptr(g), g.shape[0], g.shape[1], dtype_code(g), ptr(self.gmax),
                                                         stream_ptr(dev)), "max_row_norm")


def tensor_path_ok(q, g, k):
    """Shapes the tcgen05 retrieval kernel supports (otherwise the exact FP32 path is used)."""
    d = g.shape[1]
    return (d % 8 == 0 and 64 <= d <= 1024 and 1 <= k <= 16 and 0 < g.shape[0] < (1 << 26)
            and q.dtype in (torch.float32, torch.float16) and g.dtype in (torch.float32, torch.float16)
            and _lib.load_library().dclip_device_supports_tensor_path(g.device.index or 0) == 1)


def _topk_fp32(q, g, k, g_index_offset):
    dev = require_cuda(q, g)
    lib = _lib.load_library()
    nq, d = q.shape
    ng = g.shape[0]
    out_s = torch.empty((nq, k), dtype=torch.float32, device=dev)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=dev)
    ws = workspace(lib.dclip_topk_ip_workspace_bytes(nq, ng, d, k, dtype_code(q), dtype_code(g), _lib.PATH_FP32), dev)
    with torch.cuda.device(dev):
        check(lib.dclip_topk_ip(ptr(q), ptr(g), nq, ng, d, k, dtype_code(q), dtype_code(g), int(g_index_offset),
                                ptr(out_s), ptr(out_i), ptr(ws), ws.numel(), _lib.PATH_FP32, stream_ptr(dev)), "topk_ip")
    return out_s, out_i


def _topk_tensor(q, gallery, k, g_index_offset):
    g = gallery.g
    dev = require_cuda(q, g)
    lib = _lib.load_library()
    nq, d = q.shape
    ng = g.shape[0]
    out_s = torch.empty((nq, k), dtype=torch.float32, device=dev)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=dev)
    flags = torch.empty(nq, dtype=torch.int32, device=dev)
    n_flagged = torch.empty(1, dtype=torch.int32, device=dev)
    ws = workspace(lib.dclip_topk_ip_tc_workspace_bytes(nq, ng, d, dtype_code(q)), dev)
    with torch.cuda.device(dev):
        check(lib.dclip_topk_ip_tc(ptr(q), dtype_code(q), ptr(g), dtype_code(g), ptr(gallery.g16), nq, ng, d, k,
                                   int(g_index_offset), ptr(gallery.gmax), ptr(out_s), ptr(out_i), ptr(flags),
                                   ptr(n_flagged), ptr(ws), ws.numel(), stream_ptr(dev)), "topk_ip_tc")
    # the one host read of a search: queries whose exactness certificate failed (more than 32 near-ties)
    # are re-run on the exact FP32 path, so the result is bit-exact in every case
    if int(n_flagged.item()) > 0:
        idx = torch.nonzero(flags, as_tuple=False).flatten()
        s2, i2 = _topk_fp32(q.index_select(0, idx).contiguous(), g, k, g_index_offset)
        out_s.index_copy_(0, idx, s2)
        out_i.index_copy_(0, idx, i2)
    return out_s, out_i


def topk_ip(q, g, k, g_index_offset=0, path="auto", gallery=None):
    """Exact inner-product top-k: (scores float32[nq,k] descending, index int64[nq,k]), ties by ascending
    index, bit-exact against the oracle contract on every path.
      path "fp32"   CUDA-core kernel, every score is the contract's fmaf chain
      path "tensor" tcgen05 candidate generation + exact re-scoring + certificate (large galleries)
      path "auto"   tensor when the shape is supported and the gallery has >= 16384 rows
    `gallery`: an optional TensorGallery(g) to reuse across calls (IndexFlatIP keeps one)."""
    if isinstance(g, TensorGallery):
        gallery, g = g, g.g
    require_cuda(q, g)
    if g.shape[1] != q.shape[1]:
        raise ValueError(f"topk_ip: dim mismatch {tuple(q.shape)} vs {tuple(g.shape)}")
    if path not in ("auto", "fp32", "tensor"):
        raise ValueError(f"topk_ip: unknown path {path!r}")
    if q.shape[0] == 0 or g.shape[0] == 0:      # nothing to score: faiss-style (-inf, -1) padding
        return (torch.full((q.shape[0], k), float("-inf"), dtype=torch.float32, device=q.device),
                torch.full((q.shape[0], k), -1, dtype=torch.int64, device=q.device))
    use_tensor = path == "tensor" or (path == "auto" and g.shape[0] >= 16384 and tensor_path_ok(q, g, k))
    if use_tensor:
        if not tensor_path_ok(q, g, k):
            raise RuntimeError("topk_ip: tensor path needs d % 8 == 0, 64 <= d <= 1024, k <= 16, float16/float32 "
                               f"inputs and an sm_100 device (got q {tuple(q.shape)} {q.dtype}, g {tuple(g.shape)} {g.dtype}, k={k})")
        return _topk_tensor(q, gallery or TensorGallery(g), k, g_index_offset)
    return _topk_fp32(q, g, k, g_index_offset)


def zero_shot_topk_accuracy(image_features, text_features, labels, k=5, logit_scale=100.0):
    """Zero-shot classification scoring (reference eval_scripts/test_zero_shot_ImageNet.py:79-108 and
    CIFAR_zeroshot.py:93-122): `F.normalize(image_features)`, `100 * (img @ text.T)`,
    `topk(5)`, top-1 / top-k hit counts - one top-k kernel call and a device-side comparison instead of the
    matmul + topk + per-sample Python `label in indices[i]` loop.

    image_features [n, d] un-normalised, text_features [classes, d] as the caller holds them (already
    normalised in the reference), labels int64 [n].  Returns {"top1", "top5" (= top-k), "correct_top1",
    "correct_topk", "values" [n,k] = logit_scale * score, "indices" [n,k]}; ties go to the lower class index."""
    dev = require_cuda(image_features, text_features, labels)
    n = image_features.shape[0]
    k = min(int(k), text_features.shape[0])
    img = torch.nn.functional.normalize(image_features.float(), dim=1)        # eps 1e-12, as the reference
    scores, indices = topk_ip(img.contiguous(), text_features.float().contiguous(), k, path="fp32")
    lab = labels.to(dev).view(-1, 1)
    hit = indices == lab
    c1 = int(hit[:, 0].sum())
    ck = int(hit.any(dim=1).sum())
    return {"top1": c1 / n if n else 0, "top5": ck / n if n else 0, "correct_top1": c1, "correct_topk": ck,
            "values": scores * float(logit_scale), "indices": indices}


def topk_merge(scores, index):
    """[nlists, nq, k] per-shard results -> [nq, k] by (-score, +index)."""
    dev = require_cuda(scores, index)
    nlists, nq, k = scores.shape
    out_s = torch.empty((nq, k), dtype=torch.float32, device=dev)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        check(_lib.load_library().dclip_topk_merge(ptr(scores), ptr(index), nlists, nq, k, ptr(out_s), ptr(out_i),
                                                   stream_ptr(dev)), "topk_merge")
    return out_s, out_i


# ------------------------------------------------------------------------------------------------
# Recall@K evaluation
# ------------------------------------------------------------------------------------------------
def build_eval_index(image_ids, caption_image_ids):
    """Host-side bookkeeping of calculate_retrieval_metrics (flickr30k_eval.py:18-21, :33-34, :52-56):
    gt image index per caption (first match, like list.index) and the (image, caption) pairs of
    the i2t direction grouped by image."""
    first = {}
    for idx, iid in enumerate(image_ids):
        first.setdefault(iid, idx)
    try:
        gt_img = np.fromiter((first[c] for c in caption_image_ids), dtype=np.int64, count=len(caption_image_ids))
    except KeyError as e:      # the reference raises ValueError from list.index here
        raise ValueError(f"{e.args[0]!r} is not in list") from None
    img_to_caps = defaultdict(list)
    for ci, cid in enumerate(caption_image_ids):
        img_to_caps[cid].append(ci)
    pair_img, pair_cap = [], []
    for ii, iid in enumerate(image_ids):
        for ci in img_to_caps.get(iid, ()):
            pair_img.append(ii)
            pair_cap.append(ci)
    return gt_img, np.asarray(pair_img, dtype=np.int64), np.asarray(pair_cap, dtype=np.int64)


_EVAL_INDEX_CACHE = {}


def device_eval_index(image_ids, caption_image_ids, device):
    """build_eval_index on the device, cached: an evaluation loop passes the same two id lists every time (the Karpathy
    split does not change between checkpoints), and the Python bookkeeping over 6000 entries + three host->device copies
    cost more than the kernels.  Keyed by the lists' contents."""
    key = (hash(tuple(image_ids)), hash(tuple(caption_image_ids)), len(image_ids), len(caption_image_ids), str(device))
    hit = _EVAL_INDEX_CACHE.get(key)
    if hit is None:
        gt_img, pair_img, pair_cap = build_eval_index(list(image_ids), list(caption_image_ids))
        # CSR over the pairs (grouped by image, ascending): image r owns pairs ptr[r] .. ptr[r + 1]
        counts = np.bincount(pair_img, minlength=len(image_ids)) if len(pair_img) else np.zeros(len(image_ids), dtype=np.int64)
        ptr_h = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        hit = (torch.from_numpy(gt_img).to(device), torch.from_numpy(pair_img).to(device), torch.from_numpy(pair_cap).to(device),
               torch.from_numpy(ptr_h).to(device), int(counts.max()) if len(counts) else 0)
        if len(_EVAL_INDEX_CACHE) >= 8:
            _EVAL_INDEX_CACHE.pop(next(iter(_EVAL_INDEX_CACHE)))
        _EVAL_INDEX_CACHE[key] = hit
    return hit


def retrieval_ranks(caption_emb, image_emb, image_ids, caption_image_ids, path="fp32"):
    """(t2i_rank int32[n_cap], i2t_pair_rank int32[n_pairs], pair_img int64[n_pairs]) on the device.
    Embeddings are un-normalised [n, d]; normalisation is part of the op (flickr30k_eval.py:249-250)."""
    cap = _as_cuda(caption_emb)
    img = _as_cuda(image_emb, cap.device)
    dev = cap.device
    gt_img_d, pair_img_d, pair_cap_d, pair_ptr_d, max_targets = device_eval_index(image_ids, caption_image_ids, dev)
    c_hat = normalize_rows(cap)
    i_hat = normalize_rows(img)
    # text -> image: every caption against all images, target = its ground-truth image
    ts = pair_scores(c_hat, i_hat, gt_img_d)
    t2i = rank_counts(c_hat, i_hat, ts, gt_img_d, path=path)
    # image -> text: one (image, gt caption) pair per caption; best rank per image taken later
    ts2 = pair_scores(i_hat, c_hat, pair_cap_d, qidx=pair_img_d)
    if path == "fp32" and max_targets <= 8:
        # every image's column of S is scored once and ranked against all of its captions (same contract, 1/5 of the FLOPs)
        i2t_pairs = rank_counts_multi(i_hat, c_hat, pair_ptr_d, ts2, pair_cap_d, max_targets)
    else:
        i2t_pairs = rank_counts(i_hat, c_hat, ts2, pair_cap_d, qidx=pair_img_d, path=path)
    return t2i, i2t_pairs, pair_img_d


def _metrics_dict(t2i_m, i2t_m):
    both = torch.stack([t2i_m, i2t_m]).tolist()      # the one device->host read of the evaluation
    t, i = both[0], both[1]
    keys = ("R@1", "R@5", "R@10", "MAP")
    return {"t2i": dict(zip(keys, t)), "i2t": dict(zip(keys, i))}


def retrieval_metrics_from_embeddings(caption_emb, image_emb, image_ids, caption_image_ids, path="fp32"):
    """Fused evaluation: same dict as calculate_retrieval_metrics, S never materialised."""
    t2i, i2t_pairs, pair_img = retrieval_ranks(caption_emb, image_emb, image_ids, caption_image_ids, path)
    t2i_m = recall_metrics(t2i)
    i2t_m = recall_metrics(i2t_pairs, group=pair_img, n_groups=len(image_ids))
    return _metrics_dict(t2i_m, i2t_m)


def matrix_ranks(sim, line_index, target_index, along_rows):
    """Ranks inside a materialised matrix: pair p looks at row (along_rows) or column
    `line_index[p]` of `sim` and ranks element `target_index[p]` of that line."""
    dev = require_cuda(sim, line_index, target_index)
    n = line_index.shape[0]
    rank = torch.empty(n, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.load_library().dclip_matrix_ranks(ptr(sim), 1 if sim.dtype == torch.float64 else 0, sim.shape[0],
                                                     sim.shape[1], ptr(line_index), ptr(target_index), n,
                                                     1 if along_rows else 0, ptr(rank), stream_ptr(dev)),
              "matrix_ranks")
    return rank


def calculate_retrieval_metrics(similarity_matrix, image_ids, caption_image_ids):
    """Signature-compatible with flickr30k_eval.calculate_retrieval_metrics (:16): takes the
    materialised [n_cap, n_img] matrix (numpy float64/float32 or a tensor).  Ranks are computed
    on the GPU by comparison counting (stable-argsort semantics), not by sorting.

    The matrix form exists for compatibility with callers that already built S; the eval script
    should call `retrieval_metrics_from_embeddings` instead, which never builds it."""
    sim = _as_cuda(similarity_matrix)
    dev = sim.device
    if sim.dtype not in (torch.float32, torch.float64):
        sim = sim.float()
    gt_img, pair_img, pair_cap = build_eval_index(list(image_ids), list(caption_image_ids))
    n_cap, n_img = sim.shape
    # rank counting over a given matrix is an HBM-bound compare-and-count (dclip_matrix_ranks)
    gt_img_d = torch.from_numpy(gt_img).to(dev)
    pair_img_d = torch.from_numpy(pair_img).to(dev)
    pair_cap_d = torch.from_numpy(pair_cap).to(dev)
    t2i = matrix_ranks(sim, torch.arange(n_cap, device=dev), gt_img_d, along_rows=True)
    i2t_pairs = matrix_ranks(sim, pair_img_d, pair_cap_d, along_rows=False)
    t2i_m = recall_metrics(t2i)
    i2t_m = recall_metrics(i2t_pairs, group=pair_img_d, n_groups=len(image_ids))
    return _metrics_dict(t2i_m, i2t_m)


# ------------------------------------------------------------------------------------------------
# faiss-style flat inner-product index
# ------------------------------------------------------------------------------------------------
class IndexFlatIP:
    """Exact inner-product index with the faiss surface the reference uses
    (compute_faiss.py:27 `IndexFlatIP(d)`, :79 `.add`, image_tokenizer.py:262 `.search`).

    Vectors live in HBM (float32, or float16 with ``storage="float16"`` for 10 M-row galleries);
    `search` returns numpy arrays for numpy queries (faiss convention) and CUDA tensors for CUDA
    queries.  Results are ordered by (-score, +index); missing slots are (-inf, -1)."""

    def __init__(self, d, device=None, storage="float32", path="auto"):
        self.d = int(d)
        self.device = torch.device(device) if device is not None else None
        self.storage = torch.float32 if storage == "float32" else torch.float16
        self.path = path
        self._chunks = []
        self._data = None
        self._tensor_gallery = None
        self.index_offset = 0     # global index of row 0 (gallery sharding)

    @property
    def ntotal(self):
        return sum(c.shape[0] for c in self._chunks) + (0 if self._data is None else self._data.shape[0])

    def add(self, x):
        x = _as_cuda(x, self.device, None)
        if x.dim() != 2 or x.shape[1] != self.d:
            raise ValueError(f"IndexFlatIP.add: expected [n, {self.d}], got {tuple(x.shape)}")
        if self.device is None:
            self.device = x.device
        self._chunks.append(x.to(self.storage))

    def _matrix(self):
        if self._chunks:
            parts = ([self._data] if self._data is not None else []) + self._chunks
            self._data = parts[0] if len(parts) == 1 else torch.cat(parts, dim=0)
            self._chunks = []
            self._tensor_gallery = None
        if self._data is None:
            dev = self.device or torch.device("cuda", torch.cuda.current_device())
            self._data = torch.empty((0, self.d), dtype=self.storage, device=dev)
        return self._data

    def search(self, q, k):
        as_numpy = isinstance(q, np.ndarray)
        g = self._matrix()
        qd = _as_cuda(q, g.device)
        if qd.dtype not in (torch.float32, torch.float16):
            qd = qd.float()
        if g.dtype == torch.float32 and qd.dtype != torch.float32:
            qd = qd.float()
        gallery = None
        if self.path != "fp32" and g.shape[0] >= 16384 and tensor_path_ok(qd, g, int(k)):
            if self._tensor_gallery is None:
                self._tensor_gallery = TensorGallery(g)
            gallery = self._tensor_gallery
        scores, index = topk_ip(qd, g, int(k), self.index_offset, self.path, gallery)
        if as_numpy:
            return scores.cpu().numpy(), index.cpu().numpy()
        return scores, index

    def reset(self):
        self._chunks, self._data, self._tensor_gallery = [], None, None


# ------------------------------------------------------------------------------------------------
# gallery sharding across ranks — SURVEY.md §8(e)
# ------------------------------------------------------------------------------------------------
def sharded_topk_ip(q, g_shard, k, g_index_offset, group=None, path="auto", topk_fn=None, merge_fn=None):
    """Every rank holds all queries and one contiguous gallery shard (global row index =
    g_index_offset + local row).  Per-rank top-k, all-gather of the [nq, k] (score, index)
    lists, merge by (-score, +index).  Returns the same global result on every rank."""
    import torch.distributed as dist

    topk_fn = topk_fn or topk_ip
    merge_fn = merge_fn or topk_merge
    s, i = topk_fn(q, g_shard, k, g_index_offset, path)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return s, i
    nq = s.shape[0]
    all_s = torch.empty((world * nq, k), dtype=s.dtype, device=s.device)     # concat along dim 0
    all_i = torch.empty((world * nq, k), dtype=i.dtype, device=i.device)
    dist.all_gather_into_tensor(all_s, s.contiguous(), group=group)
    dist.all_gather_into_tensor(all_i, i.contiguous(), group=group)
    return merge_fn(all_s.view(world, nq, k), all_i.view(world, nq, k))


def sharded_retrieval_ranks(q_hat, g_hat_shard, target_score, target_index, g_index_offset, qidx=None, group=None,
                            path="fp32", rank_fn=None):
    """Partial ranks over this rank's gallery shard, summed over ranks (integer all-reduce:
    exact and order-independent).  `target_score` must be identical on all ranks (computed by
    the owner of the target row with `pair_scores` and broadcast, or recomputed from replicated
    rows — the contract's arithmetic is deterministic)."""
    import torch.distributed as dist

    rank_fn = rank_fn or rank_counts
    r = rank_fn(q_hat, g_hat_shard, target_score, target_index, qidx, g_index_offset, path)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(r, group=group)
    return r
