"""Parity of the CUDA retrieval path against the C oracle (bit-exact, ties included) and the
reference's golden metrics."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import retrieval as orc
from oracle import synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RANK_PATHS = ("fp32", "tensor")     # tensor: tcgen05 score tiles + exact re-scoring of the band around every target score


def _d(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.mark.parametrize("name", ["retrieval_small", "retrieval_mid", "retrieval_dropped"])
def test_golden_metrics(built_library, name):
    from dclip_b200 import retrieval as R
    g = load_golden(name)
    image_ids = [int(x) for x in g["image_ids"]]
    cap_ids = [int(x) for x in g["caption_image_ids"]]
    m = R.retrieval_metrics_from_embeddings(g["cap"], g["img"], image_ids, cap_ids)
    got = np.array([m[dn][k] for dn in ("t2i", "i2t") for k in ("R@1", "R@5", "R@10", "MAP")])
    want = g["metrics"]
    assert np.array_equal(got[[0, 1, 2, 4, 5, 6]], want[[0, 1, 2, 4, 5, 6]])      # Recall@K bit-exact
    assert np.allclose(got[[3, 7]], want[[3, 7]], rtol=0, atol=1e-12)
    # compatibility form on the materialised matrix (float64 buffer of fp32 products, like the reference)
    sim = orc.numpy_similarity(g["cap"], g["img"])
    m2 = R.calculate_retrieval_metrics(sim, image_ids, cap_ids)
    got2 = np.array([m2[dn][k] for dn in ("t2i", "i2t") for k in ("R@1", "R@5", "R@10", "MAP")])
    assert np.array_equal(got2[[0, 1, 2, 4, 5, 6]], want[[0, 1, 2, 4, 5, 6]])
    assert np.allclose(got2[[3, 7]], want[[3, 7]], rtol=0, atol=1e-12)


def test_golden_metrics_cfg2_shape(built_library):
    """BASELINE configs[1] shape against the REFERENCE's metrics (tests/golden/retrieval_cfg2.npz, t2i R@1 = 0.489):
    Recall@K bit-exact on both the fused form and the materialised-matrix form, on the FP32 and the tcgen05 path."""
    from dclip_b200 import retrieval as R
    from test_oracle_golden import _cfg2_inputs
    g = load_golden("retrieval_cfg2")
    cap, img, image_ids, cap_ids = _cfg2_inputs(g)
    want = g["metrics"]
    for path in RANK_PATHS:
        m = R.retrieval_metrics_from_embeddings(cap, img, image_ids, cap_ids, path=path)
        got = np.array([m[dn][k] for dn in ("t2i", "i2t") for k in ("R@1", "R@5", "R@10", "MAP")])
        assert np.array_equal(got[[0, 1, 2, 4, 5, 6]], want[[0, 1, 2, 4, 5, 6]]), (path, got, want)
        assert np.allclose(got[[3, 7]], want[[3, 7]], rtol=0, atol=1e-6), path
    m2 = R.calculate_retrieval_metrics(orc.numpy_similarity(cap, img), image_ids, cap_ids)
    got2 = np.array([m2[dn][k] for dn in ("t2i", "i2t") for k in ("R@1", "R@5", "R@10", "MAP")])
    assert np.array_equal(got2[[0, 1, 2, 4, 5, 6]], want[[0, 1, 2, 4, 5, 6]])


def test_normalize_and_scores_bit_exact(built_library):
    from dclip_b200 import retrieval as R
    rng = np.random.default_rng(0)
    x = rng.standard_normal((77, 200)).astype(np.float32)
    xh = R.normalize_rows(_d(x))
    assert np.array_equal(xh.cpu().numpy(), orc.c_normalize_rows(x))
    g = rng.standard_normal((301, 200)).astype(np.float32)
    gi = rng.integers(0, 301, size=77)
    ps = R.pair_scores(_d(x), _d(g), _d(gi))
    want = orc.c_scores(x, g)[np.arange(77), gi]
    assert np.array_equal(ps.cpu().numpy(), want)


@pytest.mark.parametrize("n_img,cpi,d,path", [(1000, 5, 512, "fp32"), (37, 3, 40, "fp32"), (130, 1, 768, "fp32"),
                                                (1000, 5, 512, "tensor"), (37, 3, 64, "tensor"), (130, 1, 768, "tensor"), (301, 4, 256, "tensor")])
def test_ranks_bit_exact(built_library, n_img, cpi, d, path):
    from dclip_b200 import retrieval as R
    cap, img, image_ids, cap_ids = synthetic.retrieval_inputs(n_img, cpi, d, 1.0, seed=n_img)
    t2i, i2t_pairs, pair_img = R.retrieval_ranks(cap, img, image_ids, cap_ids, path=path)
    want_t2i, want_i2t = orc.ranks_from_embeddings_c(cap, img, image_ids, cap_ids)
    assert t2i.cpu().tolist() == want_t2i
    best = {}
    for o, r in zip(pair_img.cpu().tolist(), i2t_pairs.cpu().tolist()):
        best[o] = min(best.get(o, 1 << 30), r)
    assert [best[i] for i in range(n_img)] == want_i2t


def test_i2t_many_captions_falls_back_and_matches(built_library):
    """Images with more than 8 captions take the per-pair kernel, fewer take the one-tile-per-image kernel: same ranks."""
    from dclip_b200 import retrieval as R
    for cpi in (9, 8, 2):
        cap, img, image_ids, cap_ids = synthetic.retrieval_inputs(23, cpi, 64, 1.0, seed=cpi)
        keep = [c for c in range(len(cap_ids)) if not (cap_ids[c] == 3 or (cap_ids[c] == 7 and c % 2))]     # an image without captions, a ragged one
        cap, cap_ids = cap[keep], [cap_ids[c] for c in keep]
        t2i, i2t_pairs, pair_img = R.retrieval_ranks(cap, img, image_ids, cap_ids)
        want_t2i, want_i2t = orc.ranks_from_embeddings_c(cap, img, image_ids, cap_ids)
        assert t2i.cpu().tolist() == want_t2i
        best = {}
        for o, r in zip(pair_img.cpu().tolist(), i2t_pairs.cpu().tolist()):
            best[o] = min(best.get(o, 1 << 30), r)
        assert [best[i] for i in sorted(best)] == want_i2t


def test_ranks_with_exact_ties(built_library):
    from dclip_b200 import retrieval as R
    rng = np.random.default_rng(2)
    g = rng.integers(-1, 2, size=(150, 8)).astype(np.float32)
    q = rng.integers(-1, 2, size=(70, 8)).astype(np.float32)
    tgt = rng.integers(0, 150, size=70)
    ts = R.pair_scores(_d(q), _d(g), _d(tgt))
    r = R.rank_counts(_d(q), _d(g), ts, _d(tgt))
    assert r.cpu().tolist() == orc.c_ranks(q, g, np.arange(70), tgt).tolist()
    # gallery sharding: partial ranks add up
    parts = [R.rank_counts(_d(q), _d(g[a:b]), ts, _d(tgt), g_index_offset=a) for a, b in ((0, 64), (64, 65), (65, 150))]
    assert (parts[0] + parts[1] + parts[2]).cpu().tolist() == r.cpu().tolist()


def test_tensor_ranks_with_ties_shards_and_band_overflow(built_library):
    """Tensor-path rank counting stays bit-exact where the fp16 score tiles cannot decide: exact ties (small-integer rows: every
    score is a tie candidate and lands in the band), gallery shards with an index offset, and a gallery of identical rows, whose
    band list overflows and hands the whole problem to the gated exact kernel."""
    from dclip_b200 import retrieval as R
    rng = np.random.default_rng(5)
    g = rng.integers(-1, 2, size=(700, 64)).astype(np.float32)
    q = rng.integers(-1, 2, size=(300, 64)).astype(np.float32)
    tgt = rng.integers(0, 700, size=300)
    ts = R.pair_scores(_d(q), _d(g), _d(tgt))
    want = orc.c_ranks(q, g, np.arange(300), tgt).tolist()
    r = R.rank_counts(_d(q), _d(g), ts, _d(tgt), path="tensor")
    assert r.cpu().tolist() == want
    parts = [R.rank_counts(_d(q), _d(g[a:b]), ts, _d(tgt), g_index_offset=a, path="tensor") for a, b in ((0, 256), (256, 257), (257, 700))]
    assert (parts[0] + parts[1] + parts[2]).cpu().tolist() == want
    # unit rows, every gallery row identical: all 40000 x 129 scores sit inside the band -> overflow -> exact fallback
    row = rng.standard_normal(128).astype(np.float32)
    row /= np.linalg.norm(row)
    g2 = np.tile(row, (40000, 1)).astype(np.float32)
    q2 = rng.standard_normal((129, 128)).astype(np.float32)
    q2 /= np.linalg.norm(q2, axis=1, keepdims=True)
    tgt2 = rng.integers(0, 40000, size=129)
    ts2 = R.pair_scores(_d(q2), _d(g2), _d(tgt2))
    r2 = R.rank_counts(_d(q2), _d(g2), ts2, _d(tgt2), path="tensor")
    assert r2.cpu().tolist() == R.rank_counts(_d(q2), _d(g2), ts2, _d(tgt2), path="fp32").cpu().tolist() == tgt2.tolist()


@pytest.mark.parametrize("ng,nq,d,k", [(5000, 33, 128, 10), (64, 1, 512, 3), (100000, 16, 64, 10), (5, 4, 16, 10), (777, 70, 768, 32)])
def test_topk_bit_exact(built_library, ng, nq, d, k):
    from dclip_b200 import retrieval as R
    gal, q, planted = synthetic.gallery_inputs(ng, nq, d, seed=ng + k)
    s, i = R.topk_ip(_d(q), _d(gal), k)
    ws, wi = orc.c_topk(q, gal, k)
    assert np.array_equal(i.cpu().numpy(), wi)
    assert np.array_equal(s.cpu().numpy(), ws)
    assert np.array_equal(i[:, 0].cpu().numpy(), planted)


def test_topk_ties_and_sharded_merge(built_library):
    from dclip_b200 import retrieval as R
    rng = np.random.default_rng(5)
    g = rng.integers(-1, 2, size=(900, 8)).astype(np.float32)      # heavy exact ties
    q = rng.integers(-1, 2, size=(20, 8)).astype(np.float32)
    ws, wi = orc.c_topk(q, g, 10)
    s, i = R.topk_ip(_d(q), _d(g), 10)
    assert np.array_equal(i.cpu().numpy(), wi) and np.array_equal(s.cpu().numpy(), ws)
    bounds = [0, 300, 301, 640, 900]
    parts = [R.topk_ip(_d(q), _d(g[a:b]), 10, g_index_offset=a) for a, b in zip(bounds[:-1], bounds[1:])]
    ms, mi = R.topk_merge(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]))
    assert np.array_equal(mi.cpu().numpy(), wi) and np.array_equal(ms.cpu().numpy(), ws)


def test_fp16_gallery(built_library):
    from dclip_b200 import retrieval as R
    gal, q, planted = synthetic.gallery_inputs(3000, 12, 96, seed=8)
    gal16 = gal.astype(np.float16)
    s, i = R.topk_ip(_d(q), _d(gal16), 10)
    ws, wi = orc.c_topk(q, gal16.astype(np.float32), 10)     # contract: products of the converted values
    assert np.array_equal(i.cpu().numpy(), wi) and np.array_equal(s.cpu().numpy(), ws)


def test_index_flat_ip_surface(built_library):
    from dclip_b200 import retrieval as R
    gal, q, planted = synthetic.gallery_inputs(500, 3, 64, seed=9)
    index = R.IndexFlatIP(64)
    for row in gal[:10]:
        index.add(row[None, :])                 # the reference adds one [1, d] row at a time (compute_faiss.py:79)
    index.add(gal[10:])
    assert index.ntotal == 500
    D, I = index.search(q[:1], 3)               # image_tokenizer.py:262
    assert isinstance(D, np.ndarray) and D.shape == (1, 3) and I.dtype == np.int64
    ws, wi = orc.c_topk(q[:1], gal, 3)
    assert np.array_equal(I, wi) and np.array_equal(D, ws)
    assert I[0, 0] == planted[0] and D[0, 0] >= 0.85


@pytest.mark.parametrize("n,classes,d", [(257, 1000, 512), (64, 10, 512), (100, 100, 768)])
def test_zero_shot_topk_accuracy_matches_reference_loop(built_library, n, classes, d):
    """zero-shot top-1 / top-5 (test_zero_shot_ImageNet.py:79-108) on tie-free synthetic features."""
    from dclip_b200 import retrieval as R
    rng = np.random.default_rng(n + classes)
    txt = rng.standard_normal((classes, d)).astype(np.float32)
    txt /= np.linalg.norm(txt, axis=1, keepdims=True)
    labels = rng.integers(0, classes, size=n)
    other = rng.integers(0, classes, size=n)                      # a distractor class of varying strength: some top-1 misses
    a = rng.uniform(0.4, 1.6, size=(n, 1)).astype(np.float32)
    img = ((txt[labels] + a * txt[other] + 0.5 * rng.standard_normal((n, d)).astype(np.float32) / np.sqrt(d)) * 7.5).astype(np.float32)
    got = R.zero_shot_topk_accuracy(torch.from_numpy(img).cuda(), torch.from_numpy(txt).cuda(), torch.from_numpy(labels).cuda(), k=5)
    top1, top5, values, indices = orc.zero_shot_counts(img, txt, labels, 5)
    assert got["correct_top1"] == top1 and got["correct_topk"] == top5
    assert np.array_equal(got["indices"].cpu().numpy(), indices)
    assert np.allclose(got["values"].cpu().numpy(), values, rtol=1e-5, atol=1e-4)
    assert 0 < top1 < n          # the fixture is neither trivial nor hopeless
