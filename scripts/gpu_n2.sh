#!/bin/bash
# 2-GPU check: sharded parity (default shards per rank and 4), bench at N=2 for 1 / 2 / 4 shards per rank
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x 2>&1 | tail -3
OMM_B200_SHARDS_PER_RANK=4 timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x 2>&1 | tail -1
for R in 1 2 4; do
OMM_B200_SHARDS_PER_RANK=$R timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$R bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_scale_n2_r$R.json 2> gpurun_out/bench_scale_n2_r$R.err
python - $R <<'PY'
import json,sys
R=sys.argv[1]
try:
    txt=[l for l in open(f'gpurun_out/bench_scale_n2_r{R}.json') if l.startswith('{')][-1]
    d=json.loads(txt); c=d['config']
    print(f"N=2 R={R}: value {d['value']:.3e} ({d['ms_per_step']:.2f} ms)  e2e {d['e2e']['value']:.3e} ({d['e2e']['ms_per_step']:.1f} ms) classify {c['classify_ms']:.2f} itempost {c['item_post_ms']:.2f} gather {c['gather_ms']:.2f} post {c['post_ms']:.2f} setup {c['setup_ms']:.2f}")
except Exception as e:
    print(R, 'ERR', e); print(open(f'gpurun_out/bench_scale_n2_r{R}.err').read()[-1500:])
PY
done
