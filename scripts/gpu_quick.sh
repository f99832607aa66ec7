#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_oracle_units.py tests/test_gpu_validation.py -m gpu -q 2>&1 | tail -25
