#!/bin/bash
# round 2, seventh pass (1 GPU): big-block hashing, evidence (launch list with issue counters, full capture of HierLeaves with source), sanitizer
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -X faulthandler -m pytest tests -m gpu -q -x --deselect tests/test_gpu_full_size.py::test_config3_full_size_is_byte_identical_with_the_sdk_bake --deselect tests/test_gpu_sdk_suite.py > gpurun_out/r2g_pytest.txt 2>&1; head -30 gpurun_out/r2g_pytest.txt | cut -c1-300; tail -5 gpurun_out/r2g_pytest.txt
timeout 600 python bench.py --no-cpu-baseline --steps 5 > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.err
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r2g_bench_n1.json') if l.startswith('{')][-1])
print('N=1 step', j['ms_per_step'], j['config']['step_ms'], 'e2e', j['e2e']['ms_per_step'], 'pageable', j['e2e']['pageable_ms_per_step'], j['parity'].get('matches_golden'))
for k,v in j['config']['secondary'].items(): print(k, v['ms_per_step'], 'classify', v['classify_ms'], 'item_post', v['item_post_ms'], 'e2e', v['e2e_ms_per_step'], v.get('matches_golden'))
PY
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -c 420 --csv --log-file gpurun_out/r2g_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2g_ncu_launches.log 2>&1
tail -2 gpurun_out/r2g_ncu_launches.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:HierLeaves -s 2 -c 1 -o gpurun_out/r2g_HierLeaves -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/r2g_ncu_full.log 2>&1
tail -2 gpurun_out/r2g_ncu_full.log
for tool in memcheck racecheck; do timeout 900 compute-sanitizer --tool $tool python scripts/sanitize_cases.py > gpurun_out/r2g_sanitizer_$tool.txt 2>&1; tail -4 gpurun_out/r2g_sanitizer_$tool.txt; done
ls -la gpurun_out | grep r2g
