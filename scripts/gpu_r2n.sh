#!/bin/bash
# round 2 (1 GPU): host-side latency work (scratch blocks, page-locked read-back block): parity subset, sanitizer, bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -X faulthandler -m pytest tests -m gpu -q -x --deselect tests/test_gpu_full_size.py::test_config3_full_size_is_byte_identical_with_the_sdk_bake --deselect tests/test_gpu_sdk_suite.py --deselect tests/test_gpu_parity.py::test_streamed_download_matches_checker > gpurun_out/r2n_pytest.txt 2>&1; head -20 gpurun_out/r2n_pytest.txt | cut -c1-200; tail -4 gpurun_out/r2n_pytest.txt
timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_cases.py > gpurun_out/r2n_sanitizer_memcheck.txt 2>&1; tail -3 gpurun_out/r2n_sanitizer_memcheck.txt
OMM_B200_TRACE=1 timeout 600 python bench.py --no-cpu-baseline --steps 5 > gpurun_out/r2n_bench_n1.json 2> gpurun_out/r2n_bench_n1.err
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r2n_bench_n1.json') if l.startswith('{')][-1])
c=j['config']
print('N=1 step', j['ms_per_step'], c['step_ms'], 'e2e', j['e2e']['ms_per_step'], 'pageable', j['e2e']['pageable_ms_per_step'], j['parity'].get('matches_golden'))
print({k:round(c[k],3) for k in ('setup_ms','classify_ms','post_ms','item_post_ms')}, j['e2e']['last_step_breakdown'])
for k,v in c['secondary'].items(): print(k, round(v['ms_per_step'],3), 'e2e', round(v['e2e_ms_per_step'],3), v.get('matches_golden'), 'setup', round(v['setup_ms'],3))
PY
LN=$(grep -n "ommB200BakeResident entry" gpurun_out/r2n_bench_n1.err | sed -n 6p | cut -d: -f1); sed -n "${LN},$((LN+9))p" gpurun_out/r2n_bench_n1.err
