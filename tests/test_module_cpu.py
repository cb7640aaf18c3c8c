"""Host-side checks of the drop-in modules that need no GPU."""
import pytest
import torch

from oracle import reference_math as om


def test_cross_modal_attention_state_dict_matches_reference_keys():
    from dclip_b200.patch_text_aggregation import PARAM_ORDER, CrossModalAttention
    m = CrossModalAttention(64, 8)
    sd = m.state_dict()
    assert list(sd.keys()) == list(om.PARAM_KEYS) == list(PARAM_ORDER)
    assert sd["text_to_image.in_proj_weight"].shape == (192, 64)
    assert sd["image_to_text.out_proj.weight"].shape == (64, 64)
    # reference checkpoints load with strict=True
    m.load_state_dict(om.init_cross_attention_params(64), strict=True)


def test_modules_refuse_cpu_tensors():
    from dclip_b200.patch_text_aggregation import PatchTextAggregation
    m = PatchTextAggregation(embed_dim=64, num_heads=8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.cross_attention(torch.randn(2, 5, 64), torch.randn(2, 3, 64))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.aggregation(torch.randn(2, 5, 64))


def test_shape_errors():
    from dclip_b200.patch_text_aggregation import CrossModalAttention
    m = CrossModalAttention(64, 8)
    with pytest.raises(ValueError):
        m(torch.randn(2, 5, 32), torch.randn(2, 3, 64))
    with pytest.raises(ValueError):
        CrossModalAttention(65, 8)
