#!/bin/bash
# full GPU check: all gpu tests, smoke, default bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L | head -3; nproc
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_all.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 3500 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
