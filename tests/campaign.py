"""Random-configuration generators shared by the randomized parity campaigns (TEST INFRASTRUCTURE):

  random_bake(rng)          one random ommCpuBake configuration (mesh, texture, sampler and bake settings) as a Workload -- used by the GPU
                            campaign (tests/test_gpu_campaign.py: product vs oracle/_ref) and by scripts/oracle_campaign.py (port vs SDK build)
  host_campaign_step(...)   one random configuration of the host build of omm_hier.cuh against the plain reference walk -- used by
                            tests/test_hier_host.py::test_random_campaign_slice and scripts/host_campaign.py

The reference's analogous pin is running itself on varied inputs (support/tests/test_omm_bake_cpu.cpp:323-344 serialize round trip on every bake)."""
import numpy as np

from omm_b200 import capi
from omm_b200 import workloads as W


def random_bake(rng, big_levels: bool = False, border_outside: bool = False):
    """Returns (workload, description dict).  `big_levels` adds a few work items of level 9-12 (CPU cost: ~0.5 us per micro-triangle).
    Border addressing is kept inside the texture unless `border_outside`: the SDK itself reads out of bounds -- and can crash -- when a
    footprint leaves a texture or a 1 x 1 mip under Border."""
    n = int(rng.integers(20, 200))
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
    if kw["addressing_mode"] == capi.ADDR_BORDER and not border_outside:
        kw["uv_lo"], kw["uv_hi"] = 0.3, 0.7
        kw["tri_texels"] = min(kw["tri_texels"], kw["tex_size"][0] / 8.0)
        kw["mips"] = 1          # a 1 x 1 mip is left by every footprint
    if kw["mips"] == 1 and rng.random() < 0.4:
        kw["tex_alpha_cutoff"] = kw["alpha_cutoff"] if rng.random() < 0.7 else 0.4
    if kw["format"] == capi.FORMAT_2_STATE:
        pass  # default states O / T are 2-state compatible
    elif rng.random() < 0.25:
        kw["alpha_cutoff_gt"], kw["alpha_cutoff_le"] = int(rng.choice([capi.STATE_T, capi.STATE_UO, capi.STATE_O])), int(rng.choice([capi.STATE_O, capi.STATE_UT, capi.STATE_T]))
    flags = 0
    if rng.random() < 0.15:
        flags |= capi.BAKE_DISABLE_SPECIAL_INDICES
    if rng.random() < 0.1:
        flags |= capi.BAKE_DISABLE_DUPLICATE_DETECTION
    if rng.random() < 0.1:
        flags |= capi.BAKE_FORCE_32BIT_INDICES
    if flags:
        kw["bake_flags"] = flags
    if rng.random() < 0.35:
        # per-triangle levels (0 .. 8, a 13 = "use the global level" now and then), bounded total work
        top = int(rng.integers(2, 9))
        lv = rng.integers(0, top + 1, n).astype(np.uint8)
        lv[rng.random(n) < 0.05] = 13
        kw["max_subdivision_level"] = int(rng.integers(0, 5))
        eff = lambda: np.where(lv == 13, kw["max_subdivision_level"], lv).astype(np.int64)
        for _ in range(n):
            if int((4 ** eff()).sum()) <= 600_000:
                break
            lv[np.argmax(eff())] = 2
        kw["subdivision_levels"] = lv
    if big_levels:
        # a few items at levels 9-12: hierarchical chunks, XXH64 of megabyte blocks, 4096+ initial regions per item
        n = int(rng.integers(3, 12))
        lv = rng.integers(0, 7, n).astype(np.uint8)
        lv[0] = int(rng.choice([9, 10, 10, 11, 12]))
        if rng.random() < 0.5:
            lv[1] = 9
        kw["subdivision_levels"] = lv
        kw["tri_texels"] = float(rng.choice([6.0, 40.0, 300.0]))
        kw["tex_size"] = (int(rng.choice([128, 256, 1024])),) * 2
        kw["max_subdivision_level"] = 12
        kw["degenerate_frac"] = kw["nan_frac"] = kw["reuse_frac"] = 0.0
        kw["index_dtype"] = np.uint32
    seed = int(rng.integers(1 << 30))
    wl = W.random_mesh(seed, n, **dict(kw))
    kw["seed"], kw["n"] = seed, n
    if "subdivision_levels" in kw:
        kw["subdivision_levels"] = kw["subdivision_levels"].tolist()
    kw["index_dtype"] = np.dtype(kw["index_dtype"]).name
    return wl, kw


def random_texture(rng):
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


def host_campaign_step(T, lib, rng, fixed):
    """One random configuration through tests/test_hier_host.check (raises AssertionError on a mismatch); returns (Stats, description)."""
    names = list(fixed.keys())
    tx = fixed[names[rng.integers(len(names))]] if rng.random() < 0.4 else random_texture(rng)
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
    what = dict(tex=tx.shape, dtype=str(tx.dtype), addr=addr, promo=promo, fmt=fmt, cutoff=cutoff, mips=mips, use_sat=use_sat, lo=lo, hi=hi, size=size, gt=gt, le=le,
                border=border, arrays=dict(tx=tx, uv=uv, lv=lv))
    try:
        st = T.check(lib, tx, uv, lv, addr=addr, cutoff=cutoff, promotion=promo, fmt=fmt, gt=gt, le=le, border=border, use_sat=use_sat, mips=mips)
    except AssertionError as e:
        e.campaign_case = what
        raise
    return st, what
