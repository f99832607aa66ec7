#!/bin/bash
# A/B of the classifier's chunk size (initial regions per chunk): time and DRAM traffic
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for v in 33554432 8388608 4194304 2097152 1048576; do
  OMM_B200_CHUNK_REGIONS=$v timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > /tmp/b.json 2>/tmp/b.err || { echo FAILED; tail -3 /tmp/b.err; continue; }
  python - "CHUNK_REGIONS=$v" <<'PY'
import json,sys
d=json.load(open('/tmp/b.json'))
print(f"{sys.argv[1]:28s} classify {d['config']['classify_ms']:8.2f} ms  step {d['ms_per_step']:8.2f} ms  e2e {d['e2e']['ms_per_step']:8.2f} ms launches {d['gpu_launches']}")
PY
done
for v in 33554432 2097152; do
OMM_B200_CHUNK_REGIONS=$v timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1200 --csv --log-file gpurun_out/chunks_$v.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python - gpurun_out/chunks_$v.csv <<'PY'
import csv,sys,collections
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]; iK=hdr.index('Kernel Name'); iM=hdr.index('Metric Name'); iV=hdr.index('Metric Value'); iI=hdr.index('ID')
agg=collections.defaultdict(lambda: collections.defaultdict(float)); cnt=collections.Counter()
for r in rows[1:]:
    k=r[iK].split('(')[0][-40:]
    v=float(r[iV].replace(',',''))
    agg[k][r[iM]]+=v
    if r[iM]=='gpu__time_duration.sum': cnt[k]+=1
for k,m in sorted(agg.items(), key=lambda kv:-kv[1]['gpu__time_duration.sum'])[:8]:
    print(f"  {k:42s} n={cnt[k]:4d} t={m['gpu__time_duration.sum']/1e6:8.3f} ms rd={m['dram__bytes_read.sum']/1e6:9.1f} MB wr={m['dram__bytes_write.sum']/1e6:9.1f} MB")
h=[m for k,m in agg.items() if 'Hier' in k]
print(sys.argv[1], 'Hier* total: t=%.2f ms rd=%.0f MB wr=%.0f MB (all captured bakes)'%(sum(m['gpu__time_duration.sum'] for m in h)/1e6, sum(m['dram__bytes_read.sum'] for m in h)/1e6, sum(m['dram__bytes_write.sum'] for m in h)/1e6))
PY
done
