#!/bin/bash
# full ncu capture of one kernel: KERNEL=regex TRIS=n
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:${KERNEL:-HierLeaves} -s ${SKIP:-0} -c 1 -o gpurun_out/prof_${KERNEL:-HierLeaves} -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --tris ${TRIS:-200000} > gpurun_out/ncu_k.log 2>&1
tail -2 gpurun_out/ncu_k.log
