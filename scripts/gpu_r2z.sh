#!/bin/bash
# round 2 (1 GPU, the last seconds of the budget): producer with / without prefetch in the two-warp digest kernel, config 5, digest checked
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
OMM_SWEEP_CONFIG=C5 timeout 60 python scripts/sweep_lanes.py 3 "pipe64 (default):" "pipe128 + prefetch:OMM_B200_BIG_HASH=pipe128" "pipe64 again:" "pipe128 again:OMM_B200_BIG_HASH=pipe128" > gpurun_out/r2z_sweep.jsonl 2> gpurun_out/r2z_sweep.err
tail -3 gpurun_out/r2z_sweep.err
python - <<'PY'
import json
for l in open('gpurun_out/r2z_sweep.jsonl'):
    if l.startswith('{'):
        r = json.loads(l)
        print(f"{r['name']:28s} step {r['step_ms']:7.3f} (min {r['min_ms']:7.3f}) item_post {r['item_post_ms']:6.3f} post {r['post_ms']:6.3f} e2e {r['e2e_ms']:7.3f} golden {r['matches_golden']}")
PY
