#!/bin/bash
# round 2 (1 GPU): launch list with issue counters of the FINAL build (128 blocks per SM)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -c 420 --csv --log-file gpurun_out/r2s_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2s_ncu_launches.log 2>&1
tail -1 gpurun_out/r2s_ncu_launches.log | cut -c1-200; ls -la gpurun_out/r2s_launches.csv
