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
