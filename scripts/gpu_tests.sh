#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu_all.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
