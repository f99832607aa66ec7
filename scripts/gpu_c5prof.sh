#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
cat > /tmp/c5.py <<'PY'
import sys
sys.path.insert(0, '.')
from omm_b200 import load_product_library, Baker, workloads as W
lib = load_product_library()
wl = W.config5()
import numpy as np
print("levels hist", np.bincount(wl.subdivision_levels, minlength=13))
with Baker(lib) as b:
    inp, tex = W.make_input(b, wl)
    for it in range(2):
        r = b.bake(inp)
    print("classify", r.timings.classifyMs)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/c5_launches.csv python /tmp/c5.py 2>&1 | tail -3
