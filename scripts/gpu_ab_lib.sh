#!/bin/bash
# same-box A/B of builds of the library: omm_b200/lib/libomm-b200.so (current) against omm_b200/lib/variant_<name>.so for each name in $VARIANTS
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
run() {
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /tmp/b.json 2>/tmp/b.err || { echo FAILED; tail -3 /tmp/b.err; return; }
  python - "$1" <<'PY'
import json,sys
d=json.load(open('/tmp/b.json'))
print(f"{sys.argv[1]:24s} classify {d['config']['classify_ms']:8.2f} ms  step {d['ms_per_step']:8.2f} ms  e2e {d['e2e']['ms_per_step']:8.2f} ms")
PY
}
cp omm_b200/lib/libomm-b200.so /tmp/current.so
for rep in 1 2; do
  cp /tmp/current.so omm_b200/lib/libomm-b200.so; run current
  for v in $VARIANTS; do cp omm_b200/lib/variant_$v.so omm_b200/lib/libomm-b200.so; run "$v"; done
done
cp /tmp/current.so omm_b200/lib/libomm-b200.so
