#!/bin/bash
# round 2 (1 GPU): second sweep of chunk lanes -- few big chunks, all of them in flight
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
M=1048576
timeout 400 python scripts/sweep_lanes.py 5 \
  "default:" \
  "2 lanes 32M:OMM_B200_CHUNK_LANES=2" \
  "2 lanes 32M grid 64:OMM_B200_CHUNK_LANES=2,OMM_B200_LIST_GRID_MULT=64" \
  "2 lanes 32M grid 256:OMM_B200_CHUNK_LANES=2,OMM_B200_LIST_GRID_MULT=256" \
  "4 lanes 16M grid 64:OMM_B200_CHUNK_LANES=4,OMM_B200_CHUNK_REGIONS=$((16*M)),OMM_B200_LIST_GRID_MULT=64" \
  "4 lanes 16M grid 128:OMM_B200_CHUNK_LANES=4,OMM_B200_CHUNK_REGIONS=$((16*M))" \
  "2 lanes 16M grid 64:OMM_B200_CHUNK_LANES=2,OMM_B200_CHUNK_REGIONS=$((16*M)),OMM_B200_LIST_GRID_MULT=64" \
  "3 lanes 22M grid 96:OMM_B200_CHUNK_LANES=3,OMM_B200_CHUNK_REGIONS=22369622,OMM_B200_LIST_GRID_MULT=96" \
  "1 lane 64M grid 128:OMM_B200_CHUNK_REGIONS=$((64*M))" \
  "1 lane 64M grid 256:OMM_B200_CHUNK_REGIONS=$((64*M)),OMM_B200_LIST_GRID_MULT=256" \
  "2 lanes 32M leaf 256:OMM_B200_CHUNK_LANES=2,OMM_B200_LEAF_GRID_MULT=256" \
  "2 lanes 32M again:OMM_B200_CHUNK_LANES=2" \
  "default again:" > gpurun_out/r2v_sweep.jsonl 2> gpurun_out/r2v_sweep.err
tail -3 gpurun_out/r2v_sweep.err
python - <<'PY'
import json
for l in open('gpurun_out/r2v_sweep.jsonl'):
    if l.startswith('{'):
        r = json.loads(l)
        print(f"{r['name']:28s} step {r['step_ms']:7.3f} (min {r['min_ms']:7.3f}) classify {r['classify_ms']:7.3f} item_post {r['item_post_ms']:6.3f} post {r['post_ms']:6.3f} e2e {r['e2e_ms']:7.3f} launches {r['launches']:4d} golden {r['matches_golden']}")
PY
