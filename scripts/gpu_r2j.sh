#!/bin/bash
# round 2, 2-GPU pass: result modes (replicated / rank 0), staging through rank 0; sharded parity and bench at N = 2
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_validation.py -m gpu -q -x 2>&1 | tail -6
for mode in rank0 replicated; do
OMM_B200_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 4 --result-mode $mode > gpurun_out/r2j_n2_$mode.json 2> gpurun_out/r2j_n2_$mode.err
grep -v "^\[omm-b200 trace\]" gpurun_out/r2j_n2_$mode.err | grep -v "^\*\*\*\|OMP_NUM\|^$" | tail -5
python - $mode <<'PY'
import json,sys
j=json.loads([l for l in open(f'gpurun_out/r2j_n2_{sys.argv[1]}.json') if l.startswith('{')][-1])
c=j['config']
print(sys.argv[1], 'N=2 step', j['ms_per_step'], c['step_ms'], 'e2e', j['e2e']['ms_per_step'], j['e2e']['last_step_breakdown'])
print('  ', {k:c[k] for k in ('setup_ms','classify_ms','post_ms','item_post_ms','gather_ms')}, j['parity'])
for k,v in c['secondary'].items(): print('  ', k, v['ms_per_step'], v['e2e_ms_per_step'], v.get('matches_golden'), v['ranks_identical'])
PY
done
