"""Seeded synthetic inputs of the BASELINE.json shapes - kept under this name for the tests and the golden generator; the recipes
themselves live in the neutral top-level module `workloads.py` (bench.py's GPU arm must not import the checker)."""
from workloads import gallery_inputs, loss_inputs, retrieval_inputs, xattn_inputs  # noqa: F401
