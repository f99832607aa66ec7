#!/bin/bash
# round 2, last 2-GPU sanity of the final build: sharded parity (both result modes), bench at N = 2
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 > gpurun_out/r2q_n2.json 2> gpurun_out/r2q_n2.err
grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r2q_n2.err | tail -3
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r2q_n2.json') if l.startswith('{')][-1])
c=j['config']
print('N=2 step', round(j['ms_per_step'],3), 'e2e', round(j['e2e']['ms_per_step'],2), j['e2e']['last_step_breakdown'], j['parity'].get('matches_golden'))
for k,v in c['secondary'].items(): print('  ', k, round(v['ms_per_step'],3), round(v['e2e_ms_per_step'],3), v.get('matches_golden'))
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --impl reference --steps 1 --warmup 0 --ref-budget-s 10 | cut -c1-300
