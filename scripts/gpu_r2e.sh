#!/bin/bash
# round 2, fifth pass (1 GPU): device post passes (a17 / a18), streamed ommCpuBake
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -X faulthandler -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "optional or neardup or compress" > gpurun_out/r2e_pytest_opt.txt 2>&1; tail -5 gpurun_out/r2e_pytest_opt.txt
timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_full_size.py::test_config3_full_size_is_byte_identical_with_the_sdk_bake > gpurun_out/r2e_pytest.txt 2>&1; tail -5 gpurun_out/r2e_pytest.txt
for div in 4 2 8; do OMM_B200_TRACE=1 OMM_B200_STREAM_DIV=$div timeout 300 python bench.py --no-cpu-baseline --no-secondary --steps 4 2>gpurun_out/r2e_div$div.err | python -c "import json,sys; j=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('div', $div, 'step', j['ms_per_step'], j['config']['step_ms'], 'e2e', j['e2e']['ms_per_step'], 'pageable', j['e2e']['pageable_ms_per_step'], j['e2e']['last_step_breakdown'], j['parity'].get('matches_golden'))"; done
OMM_B200_NO_STREAMING=1 timeout 300 python bench.py --no-cpu-baseline --no-secondary --steps 4 2>/dev/null | python -c "import json,sys; j=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('no streaming: step', j['ms_per_step'], 'e2e', j['e2e']['ms_per_step'])"
