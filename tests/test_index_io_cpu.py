"""Flat-index file format: round trip and byte layout.  Parity with a real faiss file is unpinned."""
import struct

import numpy as np
import pytest

from dclip_b200 import index_io


def test_round_trip_and_layout(tmp_path):
    rng = np.random.default_rng(0)
    xb = rng.standard_normal((37, 512)).astype(np.float32)
    p = tmp_path / "patches.index"
    index_io.write_index_arrays(p, xb)
    raw = p.read_bytes()
    assert raw[:4] == b"IxFI"
    d, n = struct.unpack("<iq", raw[4:16])
    assert (d, n) == (512, 37)
    assert struct.unpack("<qq", raw[16:32]) == (1 << 20, 1 << 20)
    assert raw[32] == 1 and struct.unpack("<i", raw[33:37]) == (0,)          # is_trained, METRIC_INNER_PRODUCT
    assert struct.unpack("<Q", raw[37:45]) == (37 * 512,)
    assert len(raw) == 45 + 37 * 512 * 4
    d2, n2, metric, got = index_io.read_index_arrays(p)
    assert (d2, n2, metric) == (512, 37, 0) and np.array_equal(got, xb)


def test_empty_and_errors(tmp_path):
    p = tmp_path / "empty.index"
    index_io.write_index_arrays(p, np.zeros((0, 64), dtype=np.float32))
    d, n, metric, got = index_io.read_index_arrays(p)
    assert (d, n) == (64, 0) and got.shape == (0, 64)
    bad = tmp_path / "bad.index"
    bad.write_bytes(b"IxHN" + b"\0" * 60)
    with pytest.raises(ValueError, match="not a flat index"):
        index_io.read_index_arrays(bad)
    trunc = tmp_path / "trunc.index"
    trunc.write_bytes(p.read_bytes()[:20])
    with pytest.raises(ValueError, match="truncated"):
        index_io.read_index_arrays(trunc)


def test_embeddings_json_round_trip(tmp_path):
    """The reference's schema (compute_faiss.py:83-91): nested [1, d] embedding lists, optional positions, key order = row order."""
    import json
    rng = np.random.default_rng(1)
    embs = rng.standard_normal((5, 16)).astype(np.float32)
    pos = rng.uniform(0, 1, size=(5, 4)).astype(np.float32)
    ids = [f"img{i // 2}.jpg_patch{i % 2}" for i in range(5)]
    p = tmp_path / "embeddings.json"
    index_io.write_embeddings_json(p, ids, embs, pos)
    raw = json.load(open(p))
    assert list(raw.keys()) == ids
    first = raw[ids[0]]
    assert set(first) == {"embedding", "position"} and len(first["embedding"]) == 1 and len(first["embedding"][0]) == 16
    ids2, embs2, pos2 = index_io.read_embeddings_json(p)
    assert ids2 == ids and np.array_equal(embs2, embs) and np.allclose(pos2, pos)
    # the older bare-list form of an entry (compute_faiss.py:84) is read too
    raw[ids[1]] = raw[ids[1]]["embedding"]
    json.dump(raw, open(p, "w"))
    ids3, embs3, pos3 = index_io.read_embeddings_json(p)
    assert np.array_equal(embs3, embs) and np.isnan(pos3[1]).all() and not np.isnan(pos3[0]).any()
