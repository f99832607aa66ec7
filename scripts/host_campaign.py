#!/usr/bin/env python
"""Randomized exactness campaign for the hierarchical classifier, no GPU needed: the host build of omm_hier.cuh (tests/hier_host) runs
the whole descent and compares every micro-triangle with the plain reference walk, over random textures (FP32 / UNORM8, pow2 / npot /
tiny / non-square), address modes, cutoffs (incl. exact texel values), border alphas, promotions, formats, levels 0-9, mip chains,
SAT on / off, triangle shapes and UV ranges (incl. far from the origin and across mirror axes / period boundaries).
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
import test_hier_host as T  # noqa: E402
from omm_b200 import capi  # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 600.0
subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "hier_host")])
lib = ctypes.CDLL(T.LIB)
rng = np.random.default_rng(seed)
fixed = T.textures(rng)
names = list(fixed.keys())


def random_texture():
    w = int(rng.choice([1, 2, 3, 4, 8, 17, 64, 96, 128, 200, 256]))
    h = int(rng.choice([1, 2, 4, 8, 31, 64, 128, 256])) if rng.random() < 0.4 else w
    cell = int(rng.choice([1, 2, 4, 8, 16]))
    base = rng.random((h // cell + 2, w // cell + 2))
    yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    fy, fx = (yy % cell) / cell, (xx % cell) / cell
    a = base[yy // cell, xx // cell] * (1 - fx) * (1 - fy) + base[yy // cell, xx // cell + 1] * fx * (1 - fy) + \
        base[yy // cell + 1, xx // cell] * (1 - fx) * fy + base[yy // cell + 1, xx // cell + 1] * fx * fy
    kind = rng.integers(4)
    if kind == 0:
        a = (a > 0.5).astype(np.float64)                       # binary
    elif kind == 1:
        a = np.round(a * 4) / 4                                 # few levels: many planar / constant cells
    tex = a.astype(np.float32)
    return (np.round(tex * 255)).astype(np.uint8) if rng.random() < 0.5 else tex


t0, total, runs = time.time(), 0, 0
while time.time() - t0 < budget:
    tx = fixed[names[rng.integers(len(names))]] if rng.random() < 0.4 else random_texture()
    n = int(rng.integers(20, 120))
    size = float(10 ** rng.uniform(-0.5, 1.8))
    lo = float(rng.choice([0.0, 0.0, -0.5, -2.0, 50.0, -300.0, 2000.0]))
    hi = lo + float(rng.choice([1.0, 2.0, 0.05]))
    kind = rng.integers(4)
    uv = T.tris(rng, n, size, max(tx.shape), lo, hi, axis_aligned=(kind == 1), skinny=(kind == 2))
    lv = rng.integers(0, 8, n) if rng.random() < 0.7 else np.full(n, int(rng.integers(0, 10 if size > 20 else 7)))
    if lv.max() > 7:
        uv, lv = uv[:6], lv[:6]
    addr = int(rng.choice([capi.ADDR_WRAP, capi.ADDR_MIRROR, capi.ADDR_CLAMP, capi.ADDR_BORDER, capi.ADDR_MIRROR_ONCE]))
    promo = int(rng.choice([capi.PROMOTE_FORCE_OPAQUE, capi.PROMOTE_FORCE_TRANSPARENT, capi.PROMOTE_NEAREST]))
    fmt = int(rng.choice([capi.FORMAT_4_STATE, capi.FORMAT_2_STATE]))
    texel = float(tx.flat[rng.integers(tx.size)]) * (1.0 / 255.0 if tx.dtype == np.uint8 else 1.0)
    cutoff = float(rng.choice([0.5, 0.3, 0.0, 1.0, 0.5000001, texel, np.nextafter(np.float32(texel), np.float32(2)), float(rng.random())]))
    mips = int(rng.choice([1, 1, 1, 2, 3, 5])) if min(tx.shape) >= 32 else 1
    use_sat = bool(rng.random() < 0.35) and mips == 1
    border = float(rng.choice([0.0, 1.0, cutoff, float(rng.random())]))
    gt, le = (capi.STATE_O, capi.STATE_T) if rng.random() < 0.7 else (int(rng.choice([capi.STATE_T, capi.STATE_UO])), int(rng.choice([capi.STATE_O, capi.STATE_UT])))
    if fmt == capi.FORMAT_2_STATE:
        gt, le = (capi.STATE_O, capi.STATE_T) if rng.random() < 0.5 else (capi.STATE_T, capi.STATE_O)
    try:
        st = T.check(lib, tx, uv, lv, addr=addr, cutoff=cutoff, promotion=promo, fmt=fmt, gt=gt, le=le, border=border, use_sat=use_sat, mips=mips)
    except AssertionError as e:
        path = f"/tmp/host_campaign_fail_{seed}.npz"
        np.savez(path, tx=tx, uv=uv, lv=lv, addr=addr, cutoff=cutoff, promo=promo, fmt=fmt, gt=gt, le=le, border=border, use_sat=use_sat, mips=mips)
        print("MISMATCH", dict(seed=seed, run=runs, tex=tx.shape, dtype=str(tx.dtype), addr=addr, promo=promo, fmt=fmt, cutoff=cutoff, mips=mips, use_sat=use_sat,
                               lo=lo, hi=hi, size=size, gt=gt, le=le, border=border), e, "->", path)
        sys.exit(1)
    total += st.microTriangles
    runs += 1
print(f"campaign ok: seed {seed}, {runs} runs, {total} micro-triangles in {time.time() - t0:.0f} s")
