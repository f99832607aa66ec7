#!/bin/bash
# round 2, last validation (1 GPU): every -m gpu test, smoke(), bench without the 3-minute CPU leg
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/r2o_pytest.txt 2>&1; tail -12 gpurun_out/r2o_pytest.txt | cut -c1-400
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r2o_bench.json') if l.startswith('{')][-1])
c=j['config']
print('N=1 step', j['ms_per_step'], c['step_ms'], 'e2e', j['e2e']['ms_per_step'], 'pageable', j['e2e']['pageable_ms_per_step'], j['parity'].get('matches_golden'))
for k,v in c['secondary'].items(): print(k, round(v['ms_per_step'],3), 'e2e', round(v['e2e_ms_per_step'],3), v.get('matches_golden'))
PY
tail -2 gpurun_out/r2o_bench.err
