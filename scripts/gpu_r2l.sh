#!/bin/bash
# round 2, validation pass (1 GPU): what the driver runs at round end -- every -m gpu test, smoke(), the default bench command
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x --durations=6 > gpurun_out/r2l_pytest.txt 2>&1; tail -14 gpurun_out/r2l_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r2l_smoke.txt
( time timeout 1500 python bench.py > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err ) 2>&1 | tail -3
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r2l_bench.json') if l.startswith('{')][-1])
c=j['config']
print('N=1 step', j['ms_per_step'], c['step_ms'], 'e2e', j['e2e']['ms_per_step'], 'pageable', j['e2e']['pageable_ms_per_step'])
print({k:c[k] for k in ('setup_ms','classify_ms','post_ms','item_post_ms')}, 'launches', j['gpu_launches'], j['clocks'])
print('parity', {k:v for k,v in j['parity'].items() if k not in ('compared','golden_source')})
print('cpu', j.get('cpu_baseline'))
print('roofline', {k:v for k,v in j['roofline'].items() if k not in ('note','issue','kernel','algorithmic_bytes_def')}, 'issue' , {k:v for k,v in (j['roofline']['issue'] or {}).items() if k not in ('kernels','source','note')})
for k,v in c['secondary'].items(): print(k, v)
PY
tail -3 gpurun_out/r2l_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-700
