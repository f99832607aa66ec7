#!/bin/bash
# round-1d check: all gpu tests, smoke, bench (+ host trace of the e2e call)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_all.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for rep in 1 2; do
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /tmp/b.json 2>/tmp/b.err || { echo FAILED; tail -3 /tmp/b.err; continue; }
  python - <<'PY'
import json,sys
d=json.load(open('/tmp/b.json'))
print(f"classify {d['config']['classify_ms']:8.2f} ms  step {d['ms_per_step']:8.2f} ms  e2e {d['e2e']['ms_per_step']:8.2f} ms", d['e2e'].get('last_step_breakdown'))
PY
done
OMM_B200_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_trace.json 2> gpurun_out/bench_trace.err
tail -n 36 gpurun_out/bench_trace.err
