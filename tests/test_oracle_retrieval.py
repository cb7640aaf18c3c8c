"""Known-answer and property tests of the retrieval oracle itself (C contract vs numpy)."""
import numpy as np

from oracle import retrieval as orc
from oracle import synthetic


def test_c_scores_match_blas_within_rounding():
    rng = np.random.default_rng(0)
    q = rng.standard_normal((9, 40)).astype(np.float32)
    g = rng.standard_normal((33, 40)).astype(np.float32)
    assert np.allclose(orc.c_scores(q, g), q @ g.T, rtol=1e-5, atol=1e-5)


def test_topk_ties_break_by_index():
    g = np.zeros((10, 8), dtype=np.float32)
    g[[2, 5, 7], 0] = 1.0            # three identical best rows
    g[4, 0] = 0.5
    q = np.zeros((1, 8), dtype=np.float32)
    q[0, 0] = 1.0
    s, i = orc.c_topk(q, g, 5)
    assert i.tolist() == [[2, 5, 7, 4, 0]]
    assert s.tolist() == [[1.0, 1.0, 1.0, 0.5, 0.0]]


def test_topk_short_gallery_pads_with_minus_one():
    g = np.eye(3, 4, dtype=np.float32)
    q = np.ones((2, 4), dtype=np.float32)
    s, i = orc.c_topk(q, g, 5)
    assert (i[:, 3:] == -1).all() and np.isinf(s[:, 3:]).all()
    assert i[:, :3].tolist() == [[0, 1, 2], [0, 1, 2]]


def test_topk_agrees_with_numpy_on_tie_free_data():
    gal, q, planted = synthetic.gallery_inputs(2000, 16, 64, seed=3)
    s, i = orc.c_topk(q, gal, 10)
    s2, i2 = orc.numpy_topk(q, gal, 10)
    assert np.array_equal(i, i2)
    assert np.array_equal(i[:, 0], planted)          # planted neighbour is the nearest (built-in KAT)
    assert np.allclose(s, s2, atol=1e-6)


def test_ranks_equal_stable_argsort_with_ties():
    rng = np.random.default_rng(1)
    g = rng.integers(-1, 2, size=(40, 6)).astype(np.float32)     # many exact ties
    q = rng.integers(-1, 2, size=(7, 6)).astype(np.float32)
    s = orc.c_scores(q, g)
    tgt = rng.integers(0, 40, size=7)
    r = orc.c_ranks(q, g, np.arange(7), tgt)
    for a in range(7):
        order = np.argsort(-s[a], kind="stable")
        assert r[a] == int(np.where(order == tgt[a])[0][0])


def test_identity_embeddings_give_perfect_recall():
    img = np.eye(16, 32, dtype=np.float32) * 3
    cap = np.repeat(img, 2, axis=0)
    m = orc.metrics_from_ranks(*orc.ranks_from_embeddings_c(cap, img, list(range(16)), [i // 2 for i in range(32)]))
    assert m["t2i"]["R@1"] == 1.0 and m["i2t"]["R@1"] == 1.0 and m["t2i"]["MAP"] == 1.0


def test_zero_shot_counts_restatement():
    """oracle.retrieval.zero_shot_counts (test_zero_shot_ImageNet.py:79-108 restated) against plain numpy."""
    from oracle import retrieval as orc
    rng = np.random.default_rng(0)
    txt = rng.standard_normal((50, 64)).astype(np.float32)
    txt /= np.linalg.norm(txt, axis=1, keepdims=True)
    labels = rng.integers(0, 50, size=40)
    img = (txt[labels] + 0.35 * rng.standard_normal((40, 64))).astype(np.float32) * 3.0
    top1, top5, values, indices = orc.zero_shot_counts(img, txt, labels, 5)
    s = (img / np.linalg.norm(img, axis=1, keepdims=True)) @ txt.T
    order = np.argsort(-s, axis=1, kind="stable")[:, :5]
    assert np.array_equal(indices, order)
    assert top1 == int((order[:, 0] == labels).sum())
    assert top5 == int((order == labels[:, None]).any(axis=1).sum())
    assert np.allclose(values, 100.0 * np.take_along_axis(s, order, axis=1), rtol=1e-5)
