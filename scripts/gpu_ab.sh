#!/bin/bash
# same-box A/B of an environment toggle: AB_VAR=NAME
cd "$GRAFT_REPO_ROOT" || exit 1
for rep in 1 2; do
for v in "" "1"; do
  if [ -n "$v" ]; then export ${AB_VAR}=1; else unset ${AB_VAR}; fi
  timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --tris ${TRIS:-1000000} > /tmp/b.json 2>/tmp/b.err || { echo FAILED; tail -3 /tmp/b.err; continue; }
  python - "${AB_VAR}=${v:-unset}" <<'PY'
import json,sys
d=json.load(open('/tmp/b.json'))
print(f"{sys.argv[1]:40s} classify {d['config']['classify_ms']:8.2f} ms  step {d['ms_per_step']:8.2f} ms  e2e {d['e2e']['ms_per_step']:8.2f} ms")
PY
done
done
