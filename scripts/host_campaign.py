#!/usr/bin/env python
"""Randomized exactness campaign for the hierarchical classifier, no GPU needed: the host build of omm_hier.cuh (tests/hier_host) runs
the whole descent and compares every micro-triangle with the plain reference walk, over random textures (FP32 / UNORM8, pow2 / npot /
tiny / non-square), address modes, cutoffs (incl. exact texel values), border alphas, promotions, formats, levels 0-9, mip chains,
SAT on / off, triangle shapes and UV ranges (incl. far from the origin and across mirror axes / period boundaries).  The generator lives
in tests/campaign.py; a bounded slice of this campaign runs in pytest (tests/test_hier_host.py::test_random_campaign_slice).
usage: python scripts/host_campaign.py [seed=1] [seconds=600]      A mismatch is saved to /tmp/host_campaign_fail_<seed>.npz."""
import ctypes
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import campaign  # noqa: E402
import test_hier_host as T  # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 600.0
subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "hier_host")])
lib = ctypes.CDLL(T.LIB)
rng = np.random.default_rng(seed)
fixed = T.textures(rng)
t0, total, runs = time.time(), 0, 0
while time.time() - t0 < budget:
    try:
        st, _ = campaign.host_campaign_step(T, lib, rng, fixed)
    except AssertionError as e:
        case = e.campaign_case
        arrays = case.pop("arrays")
        path = f"/tmp/host_campaign_fail_{seed}.npz"
        np.savez(path, **arrays, **{k: v for k, v in case.items() if k not in ("tex", "dtype")})
        print("MISMATCH", dict(seed=seed, run=runs, **case), e, "->", path)
        sys.exit(1)
    total += st.microTriangles
    runs += 1
print(f"campaign ok: seed {seed}, {runs} runs, {total} micro-triangles in {time.time() - t0:.0f} s")
