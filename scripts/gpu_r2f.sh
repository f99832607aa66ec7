#!/bin/bash
# round 2, sixth pass (1 GPU): device post passes (a17 / a18) + opt-in streamed download
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -X faulthandler -m pytest tests -m gpu -q -x --deselect tests/test_gpu_full_size.py::test_config3_full_size_is_byte_identical_with_the_sdk_bake > gpurun_out/r2f_pytest.txt 2>&1; head -30 gpurun_out/r2f_pytest.txt | cut -c1-300; tail -5 gpurun_out/r2f_pytest.txt
OMM_B200_TRACE=1 OMM_B200_STREAMED_DOWNLOAD=1 timeout 300 python bench.py --no-cpu-baseline --no-secondary --steps 3 2>gpurun_out/r2f_streamed.err | python -c "import json,sys; j=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('streamed: step', j['ms_per_step'], 'e2e', j['e2e']['ms_per_step'], j['parity'].get('matches_golden'))"
grep "falls back" gpurun_out/r2f_streamed.err | head -2
