#!/bin/bash
# round 2 (1 GPU): classifier chunks side by side on several streams (OMM_B200_CHUNK_LANES) x chunk size x grid shape, one process, digest checked per setting
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 400 python scripts/sweep_lanes.py 4 > gpurun_out/r2u_sweep.jsonl 2> gpurun_out/r2u_sweep.err
tail -3 gpurun_out/r2u_sweep.err
python - <<'PY'
import json
for l in open('gpurun_out/r2u_sweep.jsonl'):
    if l.startswith('{'):
        r = json.loads(l)
        print(f"{r['name']:46s} step {r['step_ms']:7.3f} (min {r['min_ms']:7.3f}) classify {r['classify_ms']:7.3f} item_post {r['item_post_ms']:6.3f} post {r['post_ms']:6.3f} e2e {r['e2e_ms']:7.3f} launches {r['launches']:4d} golden {r['matches_golden']}")
PY
