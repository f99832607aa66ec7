#!/bin/bash
# round 2, 8-GPU pass: scaling on ONE box at N = 8, 4, 2, 1 (rank-0 result mode), sharded parity worker with world = 8
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 tests/sharded_worker.py 2>&1 | grep -E "SHARDED_OK|Error|error|mismatch|differs" | cut -c1-200 | head -5
for N in 8 4 2 1; do
  if [ $N -eq 1 ]; then
    timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2k_n$N.json 2> gpurun_out/r2k_n$N.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2k_n$N.json 2> gpurun_out/r2k_n$N.err
  fi
  python - $N <<'PY'
import json,sys
n=sys.argv[1]
try:
    txt=[l for l in open(f'gpurun_out/r2k_n{n}.json') if l.startswith('{')][-1]
    d=json.loads(txt); c=d['config']
    print(f"N={n}: value {d['value']:.3e} ({d['ms_per_step']:.2f} ms)  e2e {d['e2e']['value']:.3e} ({d['e2e']['ms_per_step']:.2f} ms) classify {c['classify_ms']:.2f} itempost {c['item_post_ms']:.2f} gather {c['gather_ms']:.2f} post {c['post_ms']:.2f} setup {c['setup_ms']:.2f}")
    print('   e2e breakdown', d['e2e']['last_step_breakdown'], 'golden', d['parity'].get('matches_golden'))
    for k,v in c['secondary'].items(): print('  ', k, round(v['ms_per_step'],3), 'e2e', round(v['e2e_ms_per_step'],3), v.get('matches_golden'))
except Exception as e:
    print(n, 'ERR', e)
PY
  grep -v "omm-b200 trace" gpurun_out/r2k_n$N.err | grep -v "^\*\*\*\|OMP_NUM\|^$" | tail -3
done
