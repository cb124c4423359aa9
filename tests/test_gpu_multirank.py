"""Row-sharded path on >= 2 real GPUs (one process per GPU, NCCL all-reduce of [g ‖ loss] and of the Gram)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_parity():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (the 1-GPU shard-additivity test in test_gpu_invariants.py covers the math)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29577", os.path.join(ROOT, "tests", "mr_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MULTIRANK_OK world=2" in r.stdout
