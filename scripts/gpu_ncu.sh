#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:ClassifyKernel -s 1 -c 1 -o gpurun_out/prof_classify -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --tris ${TRIS:-200000} > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/
