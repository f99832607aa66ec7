#!/bin/bash
# round 2, third pass (2 GPUs): streamed ommCpuBake at N=1, shared window at N=2
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_full_size.py::test_config3_full_size_is_byte_identical_with_the_sdk_bake --deselect tests/test_gpu_sdk_suite.py 2>&1 | tail -8 | tee gpurun_out/r2c_pytest.txt
OMM_B200_TRACE=1 timeout 600 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2c_bench_n1.json'))
print('N=1 step', j['ms_per_step'], 'e2e', j['e2e']['ms_per_step'], 'pageable', j['e2e']['pageable_ms_per_step'], j['e2e']['last_step_breakdown'], j['parity'].get('matches_golden'))
for k,v in j['config']['secondary'].items(): print(k, v['ms_per_step'], v['e2e_ms_per_step'], v.get('matches_golden'))
PY
grep -B30 "array data on the host (streamed)" gpurun_out/r2c_bench_n1.err | tail -45
for div in 1 2 8; do OMM_B200_STREAM_DIV=$div timeout 300 python bench.py --no-cpu-baseline --no-secondary --steps 3 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('div', $div, 'step', j['ms_per_step'], 'e2e', j['e2e']['ms_per_step'])"; done
OMM_B200_NO_STREAMING=1 timeout 300 python bench.py --no-cpu-baseline --no-secondary --steps 3 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('no streaming: step', j['ms_per_step'], 'e2e', j['e2e']['ms_per_step'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 > gpurun_out/r2c_bench_n2.json 2> gpurun_out/r2c_bench_n2.err
grep -v "^\[omm-b200 trace\]" gpurun_out/r2c_bench_n2.err | grep -v "^\*\*\*\|OMP_NUM" | tail -5
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2c_bench_n2.json'))
print('N=2 step', j['ms_per_step'], 'e2e', j['e2e']['ms_per_step'], j['e2e']['last_step_breakdown'], j['parity'])
c=j['config']; print({k:c[k] for k in ('setup_ms','classify_ms','post_ms','item_post_ms','gather_ms')})
for k,v in c['secondary'].items(): print(k, v['ms_per_step'], v['e2e_ms_per_step'], v.get('matches_golden'), v['ranks_identical'])
PY
