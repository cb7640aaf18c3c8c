"""On-disk format of a flat inner-product index (`faiss.write_index` /
`faiss.read_index`, reference training/compute_faiss.py:97, training/image_tokenizer.py:222).

PARITY UNPINNED: faiss is a third-party dependency the reference does not vendor or pin and that is absent from
this image; the layout below restates faiss's published `index_write.cpp` / `index_read.cpp` for `IndexFlatIP`
(fourcc "IxFI", the common index header, then the vector payload - `WRITEVECTOR(xb)` in older releases and
`WRITEXBVECTOR(codes)` in newer ones produce the same bytes).  Tests cover the round trip and the byte layout
only; a file written by a real faiss has not been read here.

    header   uint32 fourcc "IxFI" | int32 d | int64 ntotal | int64 dummy (1 << 20) | int64 dummy (1 << 20)
             | uint8 is_trained | int32 metric_type (0 = inner product)
    payload  uint64 count (= ntotal * d floats) | float32[count], row-major
All little-endian.
"""
import json
import struct

import numpy as np

FOURCC_FLAT_IP = b"IxFI"
FOURCC_FLAT_L2 = b"IxF2"
_HEADER = "<4siqqqBi"          # fourcc, d, ntotal, dummy, dummy, is_trained, metric_type
METRIC_INNER_PRODUCT, METRIC_L2 = 0, 1


def write_index_arrays(path, xb):
    """Write float32 vectors [ntotal, d] as a flat inner-product index file."""
    xb = np.ascontiguousarray(xb, dtype=np.float32)
    if xb.ndim != 2:
        raise ValueError(f"write_index_arrays: expected [ntotal, d], got {xb.shape}")
    n, d = xb.shape
    with open(path, "wb") as f:
        f.write(struct.pack(_HEADER, FOURCC_FLAT_IP, d, n, 1 << 20, 1 << 20, 1, METRIC_INNER_PRODUCT))
        f.write(struct.pack("<Q", n * d))
        f.write(xb.tobytes(order="C"))


def read_index_arrays(path):
    """-> (d, ntotal, metric_type, xb float32 [ntotal, d]) of a flat index file (IxFI / IxF2)."""
    with open(path, "rb") as f:
        head = f.read(struct.calcsize(_HEADER))
        if len(head) != struct.calcsize(_HEADER):
            raise ValueError("read_index_arrays: truncated header")
        fourcc, d, n, _d0, _d1, _trained, metric = struct.unpack(_HEADER, head)
        if fourcc not in (FOURCC_FLAT_IP, FOURCC_FLAT_L2):
            raise ValueError(f"read_index_arrays: not a flat index (fourcc {fourcc!r}); only IndexFlatIP / IndexFlatL2 files are supported")
        if metric > 1:
            f.read(4)                         # metric_arg (float) follows for the exotic metrics
        (count,) = struct.unpack("<Q", f.read(8))
        if count != n * d:
            raise ValueError(f"read_index_arrays: payload of {count} floats does not match ntotal * d = {n} * {d}")
        xb = np.frombuffer(f.read(count * 4), dtype="<f4")
        if xb.size != count:
            raise ValueError("read_index_arrays: truncated payload")
    return d, n, metric, xb.reshape(n, d).copy()


def write_index(index, path):
    """`faiss.write_index(index, path)` for dclip_b200.retrieval.IndexFlatIP (vectors are fetched from HBM)."""
    g = index._matrix()
    write_index_arrays(path, g.float().cpu().numpy())


def read_index(path, device=None, storage="float32"):
    """`faiss.read_index(path)` -> dclip_b200.retrieval.IndexFlatIP holding the file's vectors in HBM."""
    from .retrieval import IndexFlatIP
    d, n, metric, xb = read_index_arrays(path)
    if metric != METRIC_INNER_PRODUCT:
        raise ValueError("read_index: only inner-product flat indexes are supported (the reference builds IndexFlatIP)")
    index = IndexFlatIP(d, device=device, storage=storage)
    if n:
        index.add(xb)
    return index


# ------------------------------------------------------------------------------------------------
# embeddings JSON written next to the index (training/compute_faiss.py:83-91, :100-102; read at
# training/image_tokenizer.py:225-227): {patch_id: {"embedding": [[d floats]], "position": [x1, y1, x2, y2]}}
# (the embedding is a [1, d] array's tolist(), hence the nesting; entries whose position could not be stored keep
# the bare nested list).  Row i of the index is the i-th key of the JSON, in insertion order.
# ------------------------------------------------------------------------------------------------
def write_embeddings_json(path, patch_ids, embeddings, positions=None):
    """Write the reference's embeddings JSON: `embeddings` [n, d] float32, `positions` [n, 4] (normalised boxes) or None."""
    embeddings = np.asarray(embeddings, dtype=np.float32)
    if embeddings.ndim != 2 or len(patch_ids) != embeddings.shape[0]:
        raise ValueError("write_embeddings_json: expected one [d] embedding per patch id")
    out = {}
    for i, pid in enumerate(patch_ids):
        emb = embeddings[i:i + 1].tolist()                    # [1, d] nesting, as the reference writes it
        out[pid] = {"embedding": emb, "position": [float(x) for x in positions[i]]} if positions is not None else emb
    with open(path, "w") as f:
        json.dump(out, f)


def read_embeddings_json(path):
    """-> (patch_ids list, embeddings float32 [n, d], positions float32 [n, 4] with NaN where absent), in file order
    (`self.patch_ids = list(self.embedding_dict.keys())`, image_tokenizer.py:227)."""
    with open(path) as f:
        data = json.load(f)
    ids = list(data.keys())
    embs, pos = [], []
    for pid in ids:
        e = data[pid]
        p = None
        if isinstance(e, dict):
            p = e.get("position")
            e = e["embedding"]
        v = np.asarray(e, dtype=np.float32).reshape(-1)
        embs.append(v)
        pos.append(np.asarray(p, dtype=np.float32) if p is not None else np.full(4, np.nan, dtype=np.float32))
    if not embs:
        return ids, np.zeros((0, 0), dtype=np.float32), np.zeros((0, 4), dtype=np.float32)
    d = embs[0].size
    if any(v.size != d for v in embs):
        raise ValueError("read_embeddings_json: embeddings of different lengths")
    return ids, np.stack(embs), np.stack(pos)


def index_from_embeddings_json(path, device=None, storage="float32"):
    """Rebuild the flat inner-product index from the JSON alone (the two files hold the same vectors in the same order).
    -> (IndexFlatIP, patch_ids)."""
    from .retrieval import IndexFlatIP
    ids, embs, _ = read_embeddings_json(path)
    index = IndexFlatIP(embs.shape[1] if embs.size else 0, device=device, storage=storage)
    if embs.shape[0]:
        index.add(embs)
    return index, ids


def knn_lookup(index, patch_ids, embeddings, query, similarity_threshold=0.85, k=3):
    """The KNN branch of ImageTokenizer.get_patch_embedding (training/image_tokenizer.py:259-282): search k = 3, use the
    stored embedding of the best match when its inner product reaches the threshold.
    -> (embedding [d] or None, "knn" | "miss", max_similarity)."""
    D, I = index.search(np.asarray(query, dtype=np.float32).reshape(1, -1), k)
    max_similarity = float(D[0][0])
    if max_similarity >= similarity_threshold and I[0][0] >= 0:
        return np.asarray(embeddings[int(I[0][0])]), "knn", max_similarity
    return None, "miss", max_similarity
