#!/bin/bash
# round 2, fourth pass (1 GPU): streamed ommCpuBake with copy-engine forwarding
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_full_size.py::test_config3_full_size_is_byte_identical_with_the_sdk_bake --deselect tests/test_gpu_sdk_suite.py 2>&1 | tail -8 | tee gpurun_out/r2d_pytest.txt
OMM_B200_TRACE=1 timeout 600 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.err
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r2d_bench_n1.json') if l.startswith('{')][-1])
print('N=1 step', j['ms_per_step'], 'e2e', j['e2e']['ms_per_step'], 'pageable', j['e2e']['pageable_ms_per_step'], j['e2e']['last_step_breakdown'], j['parity'].get('matches_golden'))
for k,v in j['config']['secondary'].items(): print(k, v['ms_per_step'], v['e2e_ms_per_step'], v.get('matches_golden'))
PY
grep -B34 "staged inputs freed" gpurun_out/r2d_bench_n1.err | grep -A34 "ommCpuBake entry" | tail -40
for div in 1 2 8; do OMM_B200_STREAM_DIV=$div timeout 300 python bench.py --no-cpu-baseline --no-secondary --steps 3 2>/dev/null | python -c "import json,sys; j=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('div', $div, 'step', j['ms_per_step'], 'e2e', j['e2e']['ms_per_step'], j['e2e']['last_step_breakdown'])"; done
OMM_B200_NO_STREAMING=1 timeout 300 python bench.py --no-cpu-baseline --no-secondary --steps 3 2>/dev/null | python -c "import json,sys; j=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('no streaming: step', j['ms_per_step'], 'e2e', j['e2e']['ms_per_step'])"
