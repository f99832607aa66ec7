#!/bin/bash
# 2-GPU check: sharded parity (1, 2, 3 runs per rank) + bench at N=2
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_scale_n2.json 2> gpurun_out/bench_scale_n2.err
python - <<'PY'
import json
txt=[l for l in open('gpurun_out/bench_scale_n2.json') if l.startswith('{')][-1]
d=json.loads(txt); c=d['config']
print(f"N=2: value {d['value']:.3e} ({d['ms_per_step']:.2f} ms)  e2e {d['e2e']['value']:.3e} ({d['e2e']['ms_per_step']:.1f} ms) classify {c['classify_ms']:.2f} itempost {c['item_post_ms']:.2f} gather {c['gather_ms']:.2f} post {c['post_ms']:.2f} setup {c['setup_ms']:.2f}")
print(c.get('step_ms'), d['clocks'])
PY
