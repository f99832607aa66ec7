#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
for lib in omm_b200/lib/libomm-b200.so omm_b200/lib/variants/*.so; do
  OMM_B200_LIB=$PWD/$lib timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --tris ${TRIS:-400000} > /tmp/b.json 2>/tmp/b.err || { echo "$lib FAILED"; tail -3 /tmp/b.err; continue; }
  python - "$lib" <<'PY'
import json,sys
d=json.load(open('/tmp/b.json'))
print(f"{sys.argv[1]:45s} classify {d['config']['classify_ms']:8.2f} ms  step {d['ms_per_step']:8.2f} ms  e2e {d['e2e']['ms_per_step']:8.2f} ms")
PY
done
