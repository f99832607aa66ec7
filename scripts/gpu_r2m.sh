#!/bin/bash
# round 2, 2-GPU pass: cost-balanced shards (A/B against equal micro-triangle counts), NUMA-interleaved window, event pool
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x 2>&1 | tail -4
run() {
  env $2 OMM_B200_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --no-secondary > gpurun_out/r2m_$1.json 2> gpurun_out/r2m_$1.err
  grep -v "^\[omm-b200 trace\]" gpurun_out/r2m_$1.err | grep -v "^\*\*\*\|OMP_NUM\|^$" | tail -3
  grep "mbind" gpurun_out/r2m_$1.err | head -1
  python - $1 <<'PY'
import json,sys
j=json.loads([l for l in open(f'gpurun_out/r2m_{sys.argv[1]}.json') if l.startswith('{')][-1])
c=j['config']
print(sys.argv[1], 'N=2 step', round(j['ms_per_step'],3), c['step_ms'], 'e2e', round(j['e2e']['ms_per_step'],2), j['e2e']['last_step_breakdown'])
print('  ', {k:round(c[k],3) for k in ('setup_ms','classify_ms','post_ms','item_post_ms','gather_ms')}, j['parity'].get('matches_golden'))
PY
}
run cost ""
run count "OMM_B200_NO_COST_BALANCE=1"
run cost1 "OMM_B200_SHARDS_PER_RANK=1"
run nonuma "OMM_B200_NO_NUMA_INTERLEAVE=1"
