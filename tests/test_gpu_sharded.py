"""Sharded (multi-GPU) bake parity: needs at least 2 CUDA devices on one box."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shards_per_rank", [1, 2, 3])
def test_two_rank_sharded_bake_matches_single_gpu(product_lib, shards_per_rank):
    """One contiguous run of work items per rank, and two / three runs dealt in boustrophedon order (OMM_B200_SHARDS_PER_RANK)."""
    n = product_lib.dll.ommB200GetDeviceCount()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29611",
           os.path.join(ROOT, "tests", "sharded_worker.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900, env=dict(os.environ, OMM_B200_SHARDS_PER_RANK=str(shards_per_rank)))
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]
    assert "SHARDED_OK" in r.stdout
