"""N > 1 on CPU (gloo, world_size 2): the host-side pieces of the sharded bake -- the work-item partition every rank
computes (ommB200ComputeShardBounds is the same function the device kernel runs), the NCCL-id style byte broadcast the
launcher does, max-over-ranks timing, and the reference arm's "rank 0 only" behaviour under torch.distributed.run."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import ctypes as C, os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["OMM_ROOT"])
from omm_b200 import capi
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
lib = capi.OmmLib(capi.PRODUCT_LIB)
# the launcher's id exchange: rank 0 owns 128 bytes, everybody must end up with the same bytes
idbuf = torch.arange(128, dtype=torch.uint8) if rank == 0 else torch.zeros(128, dtype=torch.uint8)
dist.broadcast(idbuf, 0)
assert idbuf.tolist() == list(range(128))
# every rank derives the same partition from the same (replicated) work-item list
rng = np.random.default_rng(7)
levels = rng.integers(0, 9, size=5000)
units = np.maximum((4 ** levels) // 32, 1).astype(np.uint64)
prefix = np.concatenate([[0], np.cumsum(units)]).astype(np.uint64)
first = np.zeros(world + 1, dtype=np.uint32)
assert lib.dll.ommB200ComputeShardBounds(prefix.ctypes.data, prefix.size, world, first.ctypes.data) == capi.SUCCESS
mine = torch.tensor([int(first[rank]), int(first[rank + 1])], dtype=torch.int64)
allr = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
dist.all_gather(allr, mine)
ranges = [tuple(t.tolist()) for t in allr]
assert ranges[0][0] == 0 and ranges[-1][1] == levels.size, ranges
for a, b in zip(ranges, ranges[1:]):
    assert a[1] == b[0], ranges                      # contiguous, disjoint, covering
work = [int(prefix[e] - prefix[s]) for s, e in ranges]
assert max(work) - min(work) <= int(units.max()), work   # balanced to within one work item
# timing reduction used by bench.py: the slowest rank defines the step time
t = torch.tensor([float(rank + 1)])
dist.all_reduce(t, op=dist.ReduceOp.MAX)
assert t.item() == float(world)
dist.barrier()
if rank == 0:
    print("MULTI_RANK_OK", ranges)
dist.destroy_process_group()
'''


def _torchrun(args, env_extra=None, timeout=300):
    env = dict(os.environ, OMM_ROOT=ROOT, OMP_NUM_THREADS="2")
    env.update(env_extra or {})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29577"] + args
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)


def test_shard_partition_single_process():
    from omm_b200 import capi
    lib = capi.OmmLib(capi.PRODUCT_LIB)
    units = np.array([128] * 10 + [1] * 7 + [2048] * 3, dtype=np.uint64)
    prefix = np.concatenate([[0], np.cumsum(units)]).astype(np.uint64)
    for world in (1, 2, 3, 4, 8):
        first = np.zeros(world + 1, dtype=np.uint32)
        assert lib.dll.ommB200ComputeShardBounds(prefix.ctypes.data, prefix.size, world, first.ctypes.data) == capi.SUCCESS
        assert first[0] == 0 and first[-1] == units.size
        assert np.all(np.diff(first.astype(np.int64)) >= 0)


def test_runs_are_dealt_evenly_and_folded():
    """Sharded bakes cut the work items into world x shardsPerRank runs; every rank gets the same number of runs, each run exactly one
    owner, and consecutive passes run in opposite directions (so that a cost trend along the mesh averages out)."""
    from omm_b200 import capi
    lib = capi.OmmLib(capi.PRODUCT_LIB)
    assert lib.dll.ommB200ShardsPerRank(1) == 1
    for world in (2, 3, 4, 8, 16, 64):
        per = lib.dll.ommB200ShardsPerRank(world)
        assert 1 <= per <= 4 and world * per <= 64
        for per_rank in (1, 2, 3, 4):
            owners = [lib.dll.ommB200ShardOwner(s, world) for s in range(world * per_rank)]
            assert sorted(owners) == sorted(list(range(world)) * per_rank)
            for p in range(per_rank):
                run = owners[p * world:(p + 1) * world]
                assert run == (list(range(world)) if p % 2 == 0 else list(range(world - 1, -1, -1)))
    assert lib.dll.ommB200ShardOwner(-1, 2) == -1 and lib.dll.ommB200ShardOwner(0, 0) == -1


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    r = _torchrun([str(script)])
    assert r.returncode == 0, r.stdout + r.stderr
    assert "MULTI_RANK_OK" in r.stdout


def test_reference_arm_runs_on_rank0_only():
    r = _torchrun(["bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--tris", "4096", "--tex", "256", "--level", "4",
                   "--cpu-sample-tris", "512"])
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port")
