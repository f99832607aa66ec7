#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${N:-2}
timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_scale_n1.json 2> gpurun_out/bench_scale_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/bench_scale_n$N.json 2> gpurun_out/bench_scale_n$N.err
tail -3 gpurun_out/bench_scale_n$N.err
python - <<PY
import json
for n in (1,$N):
    try:
        txt=[l for l in open(f'gpurun_out/bench_scale_n{n}.json') if l.startswith('{')][-1]
        d=json.loads(txt)
        c=d['config']
        print(f"N={n}: value {d['value']:.3e} ({d['ms_per_step']:.1f} ms)  e2e {d['e2e']['value']:.3e} ({d['e2e']['ms_per_step']:.1f} ms) classify {c['classify_ms']:.1f} itempost {c['item_post_ms']:.2f} gather {c['gather_ms']:.2f} post {c['post_ms']:.1f} setup {c['setup_ms']:.2f}")
    except Exception as e:
        print(n, 'ERR', e)
PY
