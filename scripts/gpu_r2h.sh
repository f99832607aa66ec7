#!/bin/bash
# round 2, eighth pass (1 GPU): warp-cooperative leaf walk: parity, then same-box A/B against the per-lane walk and a 6-blocks/SM build
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -X faulthandler -m pytest tests -m gpu -q -x --deselect tests/test_gpu_full_size.py::test_config3_full_size_is_byte_identical_with_the_sdk_bake --deselect tests/test_gpu_sdk_suite.py > gpurun_out/r2h_pytest.txt 2>&1; head -30 gpurun_out/r2h_pytest.txt | cut -c1-300; tail -5 gpurun_out/r2h_pytest.txt
run() {
  OMM_B200_LIB=$2 timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-secondary 2>/tmp/b.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print(f\"$1 classify {d['config']['classify_ms']:8.3f} ms  step {d['ms_per_step']:8.3f} ms  e2e {d['e2e']['ms_per_step']:8.2f} ms  golden {d['parity'].get('matches_golden')}\")"
}
for rep in 1 2; do
  run "warp leaf (current)   " $PWD/omm_b200/lib/libomm-b200.so
  run "per-lane leaf (old)   " $PWD/omm_b200/lib/variant_leafold.so
  run "warp leaf, 6 blocks/SM" $PWD/omm_b200/lib/variant_leaf6.so
done
