"""Image-sharded Path B on >= 2 GPUs: runs tests/dist_reg_check.py under torchrun with 2 ranks and the library's own
NCCL communicator; skipped on single-GPU boxes. The ownership rule itself is covered on CPU in tests/test_dist_gloo.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_registration_matches_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(ROOT, "tests", "dist_reg_check.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=280)
    assert p.returncode == 0 and "DIST_REG_OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]
