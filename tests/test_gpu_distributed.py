"""Multi-GPU parity (NCCL): runs tests/run_distributed_gpu.py under torchrun on all visible GPUs
(2, 4 or 8).  Skipped on a single-GPU box -- the same exchange logic is covered on the CPU by
tests/test_distributed_gloo.py."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_distributed_grid_and_fill_over_nccl():
    n = torch.cuda.device_count()
    n = 8 if n >= 8 else 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tests", "run_distributed_gpu.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert res.returncode == 0 and lines, res.stdout[-2000:] + res.stderr[-2000:]
    out = json.loads(lines[-1])
    assert out["world"] == n and out["failures"] == []
