#!/usr/bin/env python
"""Randomized pin of the plain-C oracle port against the SDK build (both CPU libraries, no GPU): random meshes / textures / sampler and
bake settings through ommCpuBake of oracle/liboracle_port.so and oracle/_ref/libomm-lib.so, results compared byte for byte.
usage: python scripts/oracle_campaign.py [seed=1] [seconds=300]   (needs /root/reference-built oracle/_ref, i.e. the build container)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from omm_b200 import Baker, capi  # noqa: E402
from omm_b200 import workloads as W  # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 300.0
ref = capi.OmmLib(os.path.join(ROOT, "oracle", "_ref", "libomm-lib.so"))
port = capi.OmmLib(os.path.join(ROOT, "oracle", "liboracle_port.so"))
rng = np.random.default_rng(seed)
t0, runs = time.time(), 0
while time.time() - t0 < budget:
    kw = dict(
        tex_size=(int(rng.choice([8, 64, 100, 128, 256])),) * 2, tri_texels=float(10 ** rng.uniform(0.3, 1.7)),
        uv_lo=float(rng.choice([0.0, -0.5, -1.5])), tex_kind=str(rng.choice(["noise", "circle", "blocky"])), unorm8=bool(rng.random() < 0.5),
        mips=int(rng.choice([1, 1, 2, 4])), index_dtype=[np.uint32, np.uint16][int(rng.integers(2))], degenerate_frac=float(rng.choice([0.0, 0.0, 0.2])),
        nan_frac=float(rng.choice([0.0, 0.0, 0.05])), reuse_frac=float(rng.choice([0.0, 0.3])),
        addressing_mode=int(rng.integers(5)), filter=int(rng.choice([capi.FILTER_LINEAR, capi.FILTER_LINEAR, capi.FILTER_NEAREST])),
        alpha_cutoff=float(rng.choice([0.5, 0.3, 0.7])), border_alpha=float(rng.random()), format=int(rng.choice([capi.FORMAT_4_STATE, capi.FORMAT_2_STATE])),
        unknown_state_promotion=int(rng.integers(3)), max_subdivision_level=int(rng.integers(0, 6)),
        dynamic_subdivision_scale=float(rng.choice([0.0, 0.0, 1.5, 3.0])), rejection_threshold=float(rng.choice([0.0, 0.0, 0.3])),
    )
    kw["uv_hi"] = kw["uv_lo"] + float(rng.choice([1.0, 2.5]))
    if kw["addressing_mode"] == capi.ADDR_BORDER:
        # the SDK reads out of bounds (and can crash) when a footprint leaves the texture under Border addressing:
        # keep those meshes inside
        kw["uv_lo"], kw["uv_hi"] = 0.3, 0.7
        kw["tri_texels"] = min(kw["tri_texels"], kw["tex_size"][0] / 8.0)
        kw["mips"] = 1          # a 1 x 1 mip is left by every footprint
    if kw["mips"] == 1 and rng.random() < 0.4:
        kw["tex_alpha_cutoff"] = kw["alpha_cutoff"] if rng.random() < 0.7 else 0.4
    if os.environ.get("CAMPAIGN_VERBOSE"):
        print(runs, kw, flush=True)
    wl = W.random_mesh(int(rng.integers(1 << 30)), int(rng.integers(20, 200)), **kw)
    res = []
    for lib in (ref, port):
        if os.environ.get("CAMPAIGN_VERBOSE"):
            print("  ->", "ref" if lib is ref else "port", flush=True)
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
