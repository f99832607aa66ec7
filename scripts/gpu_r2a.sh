#!/bin/bash
# round 2, first pass: every -m gpu test (full-size parity, randomized campaign, threads), smoke, the default bench command, launch list with issue counters
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L | head -3; nproc; free -g | head -2
timeout 2400 python -m pytest tests -m gpu -q -x --durations=8 2>&1 | tail -25 | tee gpurun_out/r2a_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r2a_smoke.txt
timeout 1200 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -c 6000 gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2a_ncu_launches.log 2>&1
tail -2 gpurun_out/r2a_ncu_launches.log
