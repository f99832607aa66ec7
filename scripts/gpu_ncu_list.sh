#!/bin/bash
# full ncu capture of one HierTestList launch (second-largest kernel of the classification stage)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 170 ncu --set full --clock-control none --import-source on -k regex:HierTestList -s 3 -c 1 -o gpurun_out/prof_HierTestList -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log; ls -la gpurun_out | grep HierTestList
