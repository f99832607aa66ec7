#!/bin/bash
# round 2 (1 GPU): page-locked descriptor / index host arrays: validation tests, parity, e2e
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_validation.py tests/test_gpu_parity.py tests/test_gpu_serialize.py -m gpu -q -x 2>&1 | tail -3
for i in 1 2; do timeout 600 python bench.py --no-cpu-baseline --no-secondary --steps 6 2>/dev/null | python -c "
import json,sys
j=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('step', round(j['ms_per_step'],3), 'e2e', round(j['e2e']['ms_per_step'],3), 'pageable', round(j['e2e']['pageable_ms_per_step'],3), j['e2e']['last_step_breakdown'], j['parity'].get('matches_golden'))"; done
