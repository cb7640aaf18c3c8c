"""Tensor-path retrieval (tcgen05 candidates + exact re-scoring + certificate): results must be
bit-identical to the C oracle / the FP32 CUDA path, ties and certificate failures included."""
import numpy as np
import pytest
import torch

from oracle import retrieval as orc
from oracle import synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _d(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def _pipeline_ok(lib):
    from dclip_b200 import _lib
    return lib.dclip_debug_pipeline_error(_lib.stream_ptr(torch.device(DEV))) == 0


@pytest.mark.parametrize("ng,nq,d,k,qdt,gdt", [
    (20000, 20, 256, 10, np.float32, np.float32),
    (30001, 33, 128, 10, np.float32, np.float16),
    (16500, 1, 768, 3, np.float16, np.float16),
    (50000, 16, 512, 16, np.float16, np.float16),
])
def test_tensor_topk_matches_oracle(built_library, ng, nq, d, k, qdt, gdt):
    from dclip_b200 import retrieval as R
    gal, q, planted = synthetic.gallery_inputs(ng, nq, d, seed=ng + k)
    gal, q = gal.astype(gdt), q.astype(qdt)
    s, i = R.topk_ip(_d(q), _d(gal), k, path="tensor")
    torch.cuda.synchronize()
    assert _pipeline_ok(built_library)
    ws, wi = orc.c_topk(q.astype(np.float32), gal.astype(np.float32), k)
    assert np.array_equal(i.cpu().numpy(), wi)
    assert np.array_equal(s.cpu().numpy(), ws)


def test_tensor_topk_large_matches_fp32_path(built_library):
    from dclip_b200 import retrieval as R
    g = torch.randn(300000, 768, device=DEV, generator=torch.Generator(device=DEV).manual_seed(1))
    g = (g / g.norm(dim=1, keepdim=True)).half()
    idx = torch.randperm(300000, device=DEV)[:200]
    q = (g[idx].float() + 0.01 * torch.randn(200, 768, device=DEV)).half()
    s1, i1 = R.topk_ip(q, g, 10, g_index_offset=12345, path="tensor")
    s2, i2 = R.topk_ip(q, g, 10, g_index_offset=12345, path="fp32")
    assert torch.equal(i1, i2) and torch.equal(s1, s2)
    assert torch.equal(i1[:, 0], idx + 12345)
    torch.cuda.synchronize()
    assert _pipeline_ok(built_library)


def test_certificate_fallback_on_massive_ties(built_library):
    """Thousands of exactly tied rows defeat any 32-candidate list: the certificate must flag those
    queries and the FP32 re-run must still give the oracle's index-ordered answer."""
    from dclip_b200 import retrieval as R
    rng = np.random.default_rng(3)
    gal = rng.integers(-1, 2, size=(20000, 128)).astype(np.float32)
    gal[::7] = gal[0]                       # ~2857 identical rows
    q = rng.integers(-1, 2, size=(5, 128)).astype(np.float32)
    q[0] = gal[0]
    s, i = R.topk_ip(_d(q), _d(gal), 10, path="tensor")
    ws, wi = orc.c_topk(q, gal, 10)
    assert np.array_equal(i.cpu().numpy(), wi) and np.array_equal(s.cpu().numpy(), ws)


def test_index_flat_ip_auto_uses_tensor_and_stays_exact(built_library):
    from dclip_b200 import retrieval as R
    gal, q, planted = synthetic.gallery_inputs(40000, 8, 512, seed=2)
    index = R.IndexFlatIP(512)
    index.add(gal)
    D, I = index.search(q, 3)
    assert index._tensor_gallery is not None          # the tensor path really ran
    ws, wi = orc.c_topk(q, gal, 3)
    assert np.array_equal(I, wi) and np.array_equal(D, ws)


def test_pair_kernel_odd_row_blocks_matches_oracle(built_library):
    """More than 128 queries run on CTA pairs (cta_group::2); 300 queries = 3 row blocks = a padded second pair,
    and a gallery size that is not a multiple of the 256-row tile."""
    from dclip_b200 import retrieval as R
    gal, q, planted = synthetic.gallery_inputs(20001, 300, 128, seed=11)
    gal, q = gal.astype(np.float16), q.astype(np.float16)
    s, i = R.topk_ip(_d(q), _d(gal), 10, path="tensor")
    torch.cuda.synchronize()
    assert _pipeline_ok(built_library)
    ws, wi = orc.c_topk(q.astype(np.float32), gal.astype(np.float32), 10)
    assert np.array_equal(i.cpu().numpy(), wi)
    assert np.array_equal(s.cpu().numpy(), ws)


def test_lockstep_streaming_matches_fp32_path(built_library):
    """Gallery larger than L2 (107 MB) and more than two waves of CTAs: the streaming kernel runs with its soft
    lockstep barrier.  The barrier is a performance device only - results stay bit-identical to the FP32 path."""
    from dclip_b200 import retrieval as R
    ng, nq, d = 70000, 19000, 768
    g = torch.randn(ng, d, device=DEV, generator=torch.Generator(device=DEV).manual_seed(5))
    g = (g / g.norm(dim=1, keepdim=True)).half()
    idx = torch.randint(0, ng, (nq,), device=DEV, generator=torch.Generator(device=DEV).manual_seed(6))
    q = (g[idx].float() + 0.02 * torch.randn(nq, d, device=DEV)).half()
    lib = built_library
    assert lib.dclip_topk_ip_tc_workspace_bytes(nq, ng, d, 2) > 0
    s1, i1 = R.topk_ip(q, g, 10, path="tensor")
    s2, i2 = R.topk_ip(q, g, 10, path="fp32")
    torch.cuda.synchronize()
    assert _pipeline_ok(built_library)
    assert torch.equal(i1, i2) and torch.equal(s1, s2)
