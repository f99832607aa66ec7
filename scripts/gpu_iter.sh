#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --tris 400000 > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_iter.json'))
print('value %.3e  ms/step %.1f  e2e %.3e (%.1f ms)'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))
print('cfg', {k:d['config'][k] for k in ('setup_ms','classify_ms','post_ms')})
print('e2e breakdown', d['e2e'].get('last_step_breakdown'))
print('clocks', d['clocks'])
PY
tail -3 gpurun_out/bench_iter.err
