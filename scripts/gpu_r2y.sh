#!/bin/bash
# round 2 (1 GPU): producer / chain-warp digest kernel by default -- the GPU suite without the two long tests, then the bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --durations=5 --deselect tests/test_gpu_full_size.py::test_config3_full_size_is_byte_identical_with_the_sdk_bake --deselect tests/test_gpu_sdk_suite.py > gpurun_out/r2y_pytest.txt 2>&1; tail -12 gpurun_out/r2y_pytest.txt | cut -c1-300
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r2y_bench.json') if l.startswith('{')][-1])
c=j['config']
print('N=1 step', j['ms_per_step'], c['step_ms'], 'classify', c['classify_ms'], 'post', c['post_ms'], 'e2e', j['e2e']['ms_per_step'], 'pageable', j['e2e']['pageable_ms_per_step'], j['parity'].get('matches_golden'), 'launches', j['gpu_launches'])
for k,v in c['secondary'].items(): print(k, round(v['ms_per_step'],3), 'item_post', round(v['item_post_ms'],3), 'e2e', round(v['e2e_ms_per_step'],3), v.get('matches_golden'))
PY
tail -2 gpurun_out/r2y_bench.err
