"""GPU (needs >= 2 devices): sharded evaluation and sharded PNCG over NCCL equal the single-GPU path."""

import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_sharding_matches_single_gpu(native_lib):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29517", str(ROOT / "tools" / "dist_check.py")]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0 and "DIST_CHECK PASS" in proc.stdout, proc.stdout[-3000:] + proc.stderr[-3000:]
