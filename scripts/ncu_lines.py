#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line (instructions executed)."""
import csv, collections, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
rows = list(csv.reader(open(path)))
sec = None; hdr = None
per = collections.defaultdict(lambda: [0, 0]); src = {}
for r in rows:
    if not r: continue
    if r[0] == 'File Path': sec = r[1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None: continue
    try: ln = int(r[0])
    except ValueError: continue
    d = dict(zip(hdr, r))
    if d.get('Address', '') not in ('', '-'): continue   # SASS rows carry an address; keep only the per-source-line rows
    try: ie = int(d.get('Instructions Executed', '0') or 0); sm = int(d.get('# Samples', '0') or 0)
    except ValueError: continue
    per[(sec, ln)][0] += ie; per[(sec, ln)][1] += sm; src[(sec, ln)] = r[1]
tot = sum(v[0] for v in per.values()); tots = sum(v[1] for v in per.values())
print('total warp-instructions', tot, 'samples', tots)
for (f, l), (ie, sm) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*ie/tot:6.2f}% inst {100*sm/max(tots,1):6.2f}% smp  {f.split('/')[-1]}:{l}: {src[(f,l)].strip()[:120]}")
