#!/bin/bash
# first GPU contact: parity suite + a quick timing of a C3 slice
cd "$GRAFT_REPO_ROOT" || exit 1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/quick_c3.txt
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from omm_b200 import load_product_library, Baker, workloads as W
import parity_cases as PC
lib = load_product_library()
for n in (20000, 100000):
    wl = W.config3(num_tris=n, tex_size=4096, level=6)
    with Baker(lib) as b:
        inp, tex = W.make_input(b, wl)
        for it in range(2):
            t = time.time(); r = b.bake(inp); dt = time.time() - t
            tm = r.timings
            print(f"C3 slice n={n} it={it}: wall {dt*1e3:.1f} ms  h2d {tm.h2dMs:.2f} setup {tm.setupMs:.2f} classify {tm.classifyMs:.2f} post {tm.postMs:.2f} d2h {tm.d2hMs:.2f} total {tm.totalDeviceMs:.2f} ms; utris {tm.microTriangles} items {tm.workItems} arr {tm.arrayDataBytes} descs {tm.descCount} launches {tm.kernelLaunches}")
            print(f"   classify rate {tm.microTriangles/tm.classifyMs/1e6:.1f} Gutri/s... wall rate {tm.microTriangles/dt/1e9:.2f} Gutri/s")
        tex.destroy()
PY
