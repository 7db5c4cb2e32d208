"""Row-partitioned multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 3])
def test_row_partitioned_solvers_match_oracle(world):
    """with fewer GPUs than ranks the ranks share a device (see the worker): the 3-rank case has a
    middle rank with a two-sided halo and unequal halo sizes"""
    import torch
    if torch.cuda.device_count() < 1:
        pytest.skip("needs a GPU")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world,
           "--master-addr", "127.0.0.1", "--master-port", "29621",
           os.path.join(ROOT, "tests", "_dist_gpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-5000:]
    for r in range(world):
        assert "rank %d ok" % r in out.stdout
