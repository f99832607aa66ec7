#!/usr/bin/env python
"""Randomized pin of the plain-C oracle port against the SDK build (both CPU libraries, no GPU): random meshes / textures / sampler and
bake settings (tests/campaign.py::random_bake) through ommCpuBake of oracle/liboracle_port.so and oracle/_ref/libomm-lib.so, results
compared byte for byte.
usage: python scripts/oracle_campaign.py [seed=1] [seconds=300]   (needs /root/reference-built oracle/_ref, i.e. the build container)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import campaign  # noqa: E402
from omm_b200 import Baker, capi  # noqa: E402
from omm_b200 import workloads as W  # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 300.0
ref = capi.OmmLib(os.path.join(ROOT, "oracle", "_ref", "libomm-lib.so"))
port = capi.OmmLib(os.path.join(ROOT, "oracle", "liboracle_port.so"))
rng = np.random.default_rng(seed)
t0, runs = time.time(), 0
while time.time() - t0 < budget:
    wl, kw = campaign.random_bake(rng)
    if os.environ.get("CAMPAIGN_VERBOSE"):
        print(runs, kw, flush=True)
    res = []
    for lib in (ref, port):
        with Baker(lib) as b:
            inp, tex = W.make_input(b, wl)
            try:
                res.append(b.bake(inp))
            except Exception as e:  # both must fail alike
                res.append(repr(e))
            tex.destroy()
    same = (res[0] == res[1]) if isinstance(res[0], str) or isinstance(res[1], str) else res[0].diff(res[1]) == []
    if not same:
        print("MISMATCH", seed, runs, kw, res[0] if isinstance(res[0], str) else res[0].diff(res[1]))
        sys.exit(1)
    runs += 1
print(f"oracle campaign ok: seed {seed}, {runs} random bakes identical in {time.time() - t0:.0f} s")
