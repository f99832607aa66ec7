#!/bin/bash
# round 2 (1 GPU): `ncu --set full` of the two hottest kernels of the FINAL build (128 blocks per SM), with source
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:HierLeaves -s 2 -c 1 -o gpurun_out/r2t_HierLeaves -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/r2t_ncu_full_leaves.log 2>&1
tail -1 gpurun_out/r2t_ncu_full_leaves.log | cut -c1-160
timeout 300 ncu --set full --clock-control none --import-source on -k regex:HierTestList -s 5 -c 1 -o gpurun_out/r2t_HierTestList -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/r2t_ncu_full_list.log 2>&1
tail -1 gpurun_out/r2t_ncu_full_list.log | cut -c1-160
ls -la gpurun_out | grep r2t
