#!/bin/bash
# 4-GPU: bench at N=4 for 1 / 2 shards per rank
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for R in 1 2; do
OMM_B200_SHARDS_PER_RANK=$R timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2954$R bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_scale_n4_r$R.json 2> gpurun_out/bench_scale_n4_r$R.err
python - $R <<'PY'
import json,sys
R=sys.argv[1]
try:
    txt=[l for l in open(f'gpurun_out/bench_scale_n4_r{R}.json') if l.startswith('{')][-1]
    d=json.loads(txt); c=d['config']
    print(f"N=4 R={R}: value {d['value']:.3e} ({d['ms_per_step']:.2f} ms)  e2e {d['e2e']['value']:.3e} ({d['e2e']['ms_per_step']:.1f} ms) classify {c['classify_ms']:.2f} itempost {c['item_post_ms']:.2f} gather {c['gather_ms']:.2f} post {c['post_ms']:.2f} setup {c['setup_ms']:.2f}")
except Exception as e:
    print(R, 'ERR', e); print(open(f'gpurun_out/bench_scale_n4_r{R}.err').read()[-1500:])
PY
done
