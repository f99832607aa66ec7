#!/bin/bash
# BASELINE configs 2 and 5 at full size: timing of the resident bake + size-independent checks; parity on reduced copies
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python - <<'PY' 2>&1 | tee gpurun_out/configs.txt
import sys, time, os
sys.path.insert(0, '.')
import numpy as np
from omm_b200 import load_product_library, Baker, capi, workloads as W
lib = load_product_library()
ref = capi.OmmLib(os.path.join('oracle', '_ref', 'libomm-lib.so'))
def bake(l, wl, **kw):
    with Baker(l) as b:
        inp, tex = W.make_input(b, wl, **kw)
        t = time.time(); r = b.bake(inp); dt = time.time() - t
        tex.destroy()
    return r, dt
for name, wl in (("C2 10k tris/1024^2/L4", W.config2()), ("C5 small parity", W.config5(num_tris=20000, tex_size=1024, distinct=512, flat_tris=2000, max_level=9)),
                 ("C5 1M tris mixed L0-12", W.config5())):
    for it in range(2):
        r, dt = bake(lib, wl)
    tm = r.timings
    print(f"{name}: wall {dt*1e3:.1f} ms device {tm.totalDeviceMs:.2f} ms (setup {tm.setupMs:.2f} classify {tm.classifyMs:.2f} post {tm.postMs:.2f}) utris {tm.microTriangles} items {tm.workItems} "
          f"array {tm.arrayDataBytes} descs {tm.descCount} launches {tm.kernelLaunches} -> {tm.microTriangles/tm.classifyMs/1e6:.1f} Gutri/s classify")
    if wl.num_triangles <= 20000:
        o, dto = bake(ref, wl, bake_flags=capi.BAKE_ENABLE_INTERNAL_THREADS)
        d = r.diff(o)
        print(f"   vs SDK build ({dto*1e3:.0f} ms): {'IDENTICAL' if not d else d}")
PY
