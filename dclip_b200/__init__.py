"""dclip_b200 — B200-native (sm_100a) implementation of DCLIP's data-parallel hot path.

Drop-in Python surface (same names / signatures as the reference modules):

  dclip_b200.patch_text_aggregation   CrossModalAttention, PatchTextAggregation
  dclip_b200.losses                   compute_contrastive_loss, cosine_distillation_loss,
                                      DistillationLoss (+ sharded variant)
  dclip_b200.retrieval                calculate_retrieval_metrics,
                                      retrieval_metrics_from_embeddings, IndexFlatIP, topk_ip

Everything underneath is hand-written CUDA behind the C ABI in ``include/dclip_b200.h``
(``dclip_b200/lib/libdclip_b200.so``).  There is no CPU or eager-PyTorch fallback: if the
library is missing, or an op is called with CPU tensors, it raises.
"""
from ._lib import lib_path, load_library, version  # noqa: F401

__all__ = ["lib_path", "load_library", "version"]
