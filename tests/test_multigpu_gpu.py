"""Multi-GPU (NCCL) run of the sharded loss / retrieval paths; skipped on boxes with < 2 GPUs."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_paths_match_single_gpu(built_library, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + world), os.path.join(ROOT, "tests", "mp", "sharded_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(out.stdout[-6000:])                       # the worker's parity lines: visible with pytest -s (kept under profiles/)
    assert out.returncode == 0 and f"SHARDED_OK world={world}" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
