#!/bin/bash
# per-kernel time / instruction / issue metrics of the hierarchical classifier kernels on a C3 slice
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,l1tex__t_sector_hit_rate.pct \
  --clock-control none -k regex:Hier -c 12 --csv --log-file gpurun_out/hier_metrics.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --tris ${TRIS:-200000} > gpurun_out/hier_metrics.log 2>&1
tail -2 gpurun_out/hier_metrics.log
