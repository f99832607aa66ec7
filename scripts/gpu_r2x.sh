#!/bin/bash
# round 2 (1 GPU): the XXH64 chain of the big-block digest kernel -- forms of the step and the producer / chain-warp kernel, config 5 (one
# level-12 block), one process; digest checked per setting
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
OMM_SWEEP_CONFIG=C5 timeout 300 python scripts/sweep_lanes.py 4 "plain:OMM_B200_BIG_HASH=plain" "funnel (default):" "split:OMM_B200_BIG_HASH=split" "pipe1:OMM_B200_BIG_HASH=pipe1" "pipe2:OMM_B200_BIG_HASH=pipe2" "pipe3:OMM_B200_BIG_HASH=pipe3" "funnel again:" > gpurun_out/r2x_sweep.jsonl 2> gpurun_out/r2x_sweep.err
tail -3 gpurun_out/r2x_sweep.err
python - <<'PY'
import json
for l in open('gpurun_out/r2x_sweep.jsonl'):
    if l.startswith('{'):
        r = json.loads(l)
        print(f"{r['name']:28s} step {r['step_ms']:7.3f} (min {r['min_ms']:7.3f}) classify {r['classify_ms']:7.3f} item_post {r['item_post_ms']:6.3f} post {r['post_ms']:6.3f} e2e {r['e2e_ms']:7.3f} launches {r['launches']:4d} golden {r['matches_golden']}")
PY
