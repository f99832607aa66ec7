import sys, time, ctypes as C
sys.path.insert(0, '.')
import numpy as np
from omm_b200 import load_product_library, Baker, capi, workloads as W
lib = load_product_library()
for name, wl in (("C5small", W.config5(num_tris=20000, tex_size=1024, distinct=512, flat_tris=2000, max_level=9)), ("C3-200k", W.config3(num_tris=200000))):
    with Baker(lib) as b:
        inp, tex = W.make_input(b, wl)
        for it in range(5):
            t = time.time(); r = b.bake(inp); dt = time.time() - t
            tm = r.timings
            print(f"{name} it{it}: wall {dt*1e3:.1f} ms hostBake {tm.hostBakeMs:.2f} device {tm.totalDeviceMs:.2f} (setup {tm.setupMs:.2f} classify {tm.classifyMs:.2f} post {tm.postMs:.2f}) stage {tm.hostStageMs:.2f} dl {tm.hostDownloadMs:.2f}")
        tex.destroy()
