#!/bin/bash
# round 2 (1 GPU): grid size of the classifier kernels, same-box A/B (parity-neutral switches OMM_B200_{LIST,INIT,LEAF}_GRID_MULT = blocks per SM)
cd "$GRAFT_REPO_ROOT" || exit 1
run() {
  env $2 timeout 300 python bench.py --no-cpu-baseline --no-secondary --steps 6 2>/dev/null | python -c "
import json,sys
j=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$1', 'step', round(j['ms_per_step'],3), 'classify', round(j['config']['classify_ms'],3), j['parity'].get('matches_golden'))"
}
run "all 128 (default)     " "A=1"
run "init 16, rest 128     " "OMM_B200_INIT_GRID_MULT=16"
run "init 512, rest 128    " "OMM_B200_INIT_GRID_MULT=512"
run "leaf 64, rest 128     " "OMM_B200_LEAF_GRID_MULT=64"
run "leaf 256, rest 128    " "OMM_B200_LEAF_GRID_MULT=256"
run "leaf 512, rest 128    " "OMM_B200_LEAF_GRID_MULT=512"
run "list 64, init+leaf 128" "OMM_B200_LIST_GRID_MULT=64 OMM_B200_INIT_GRID_MULT=128 OMM_B200_LEAF_GRID_MULT=128"
run "list 256, init+leaf 128" "OMM_B200_LIST_GRID_MULT=256 OMM_B200_INIT_GRID_MULT=128 OMM_B200_LEAF_GRID_MULT=128"
