"""Host-side fuzz of the exact shortcuts of the hierarchical classifier (omm_b200/csrc/omm_hier.cuh).

tests/hier_host/hier_host_check.cpp compiles the very header the CUDA kernels use for the HOST (g++, -ffp-contract=off),
runs the descent of HierTestInitial / HierTestList / HierLeaves serially -- whole-cell bitmap (F), region tests (A)-(C),(G),
leaf walk with the edge filter (D),(E) and the queued edge tests -- and compares every micro-triangle with the plain
reference walk (ClassifyMicroTriangle), which the GPU parity suite pins to the SDK build byte for byte.  No GPU needed; this
is test infrastructure, the product has no CPU path.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from omm_b200 import capi
from omm_b200 import workloads as W

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "hier_host", "libhier_host_check.so")


class Stats(ctypes.Structure):
    _fields_ = [("microTriangles", ctypes.c_uint64), ("mismatches", ctypes.c_uint64), ("tests", ctypes.c_uint64 * 4), ("passes", ctypes.c_uint64 * 4),
                ("fullEvals", ctypes.c_uint64), ("firstBadItem", ctypes.c_uint64), ("firstBadIndex", ctypes.c_uint64), ("firstBadGot", ctypes.c_int32),
                ("firstBadWant", ctypes.c_int32)]


@pytest.fixture(scope="module")
def lib():
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "hier_host")])
    return ctypes.CDLL(LIB)


def check(lib, tex, uvs, levels, addr=capi.ADDR_WRAP, cutoff=0.5, promotion=capi.PROMOTE_FORCE_OPAQUE, fmt=capi.FORMAT_4_STATE, gt=capi.STATE_O,
          le=capi.STATE_T, border=0.0, use_sat=False, mips=1):
    tex = np.ascontiguousarray(tex)
    uvs = np.ascontiguousarray(uvs, dtype=np.float32)
    levels = np.ascontiguousarray(levels, dtype=np.uint8)
    st = Stats()
    h, w = tex.shape
    lib.hier_host_check(tex.ctypes.data_as(ctypes.c_void_p), int(tex.dtype == np.float32), w, h, addr, ctypes.c_float(border), ctypes.c_float(cutoff), gt, le,
                        fmt, promotion, uvs.ctypes.data_as(ctypes.c_void_p), levels.ctypes.data_as(ctypes.c_void_p), len(levels), ctypes.byref(st), int(use_sat), int(mips))
    assert st.mismatches == 0, (f"{st.mismatches} of {st.microTriangles} micro-triangles differ from the reference walk; first: item {st.firstBadItem} "
                                f"index {st.firstBadIndex} got {st.firstBadGot} want {st.firstBadWant}")
    return st


def tris(rng, n, size_texels, texsize, lo=0.0, hi=1.0, axis_aligned=False, skinny=False):
    c = lo + (hi - lo) * rng.random((n, 1, 2))
    r = size_texels / texsize
    if axis_aligned:  # two edges parallel to the texel grid: the vertical-edge and steep-slope branches of the edge test
        base = np.array([[0, 0], [1, 0], [0, 1]], dtype=np.float64)[None] * r * (0.3 + 0.7 * rng.random((n, 1, 1)))
        flip = rng.integers(0, 4, (n, 1, 1))
        base = np.where(flip == 1, base * [-1, 1], base)
        base = np.where(flip == 2, base * [1, -1], base)
        base = np.where(flip == 3, base[:, :, ::-1], base)
        uv = c + base
    else:
        ang = 2 * np.pi * (rng.random((n, 3)) / 3 + np.arange(3)[None] / 3)
        rad = r * (0.3 + 0.7 * rng.random((n, 3)))
        if skinny:
            rad[:, 2] *= 0.02
        uv = c + np.stack([rad * np.cos(ang), rad * np.sin(ang)], axis=2)
    return uv.astype(np.float32).reshape(n, 6)


def textures(rng):
    yy, xx = np.meshgrid(np.arange(256), np.arange(256), indexing="ij")
    return {
        "noise": W.noise_texture(1024),
        "noise8": W.noise_texture(512, as_unorm8=True),
        "smooth": (0.5 + 0.4 * np.sin(xx * 0.11) * np.cos(yy * 0.07) + 0.05 * np.sin(xx * 0.9 + yy * 0.7)).astype(np.float32),
        "ramp": ((xx + yy) / 510.0).astype(np.float32),
        "checker": (((xx >> 3) + (yy >> 3)) & 1).astype(np.float32),
        "nearcut": (0.5 + 1e-6 * rng.standard_normal((128, 128))).astype(np.float32),
        "npot": W.noise_texture(256)[:200, :173].copy(),
    }


def test_c3_slice_level6(lib):
    wl = W.config3(num_tris=400, tex_size=4096, level=6)
    st = check(lib, wl.mips[0], wl.texcoords.reshape(-1, 6), np.full(400, 6))
    # the point of the hierarchy: almost everything is decided by region tests, only a few per cent reach the leaf walk
    assert st.fullEvals < 0.08 * st.microTriangles
    assert st.passes[0] > 0.8 * st.tests[0]


@pytest.mark.parametrize("addr", [capi.ADDR_WRAP, capi.ADDR_MIRROR, capi.ADDR_CLAMP, capi.ADDR_BORDER, capi.ADDR_MIRROR_ONCE])
def test_address_modes(lib, addr):
    rng = np.random.default_rng(100 + addr)
    t = textures(rng)
    check(lib, t["noise"], tris(rng, 120, 6, 1024, -0.5, 1.5), np.full(120, 6), addr=addr, border=0.4)
    check(lib, t["npot"], tris(rng, 120, 8, 200, -1.0, 2.0), np.full(120, 5), addr=addr, border=0.6)


def test_shapes_and_slopes(lib):
    rng = np.random.default_rng(7)
    t = textures(rng)
    check(lib, t["noise"], tris(rng, 150, 6, 1024, axis_aligned=True), np.full(150, 6))
    check(lib, t["ramp"], tris(rng, 150, 10, 256, axis_aligned=True), np.full(150, 6))
    check(lib, t["checker"], tris(rng, 150, 9, 256, axis_aligned=True), np.full(150, 5))
    check(lib, t["noise"], tris(rng, 150, 12, 1024, skinny=True), np.full(150, 6))
    check(lib, t["smooth"], tris(rng, 150, 6, 256, -0.02, 0.02, axis_aligned=True), np.full(150, 6))   # around the UV origin: tiny ulps
    check(lib, t["noise"], tris(rng, 150, 6, 1024, 100, 101), np.full(150, 6))                            # far from it: coarse ulps
    check(lib, t["noise"], tris(rng, 150, 0.7, 1024), np.full(150, 6))                                    # micro-triangles << rounding slack


def test_margins_and_formats(lib):
    rng = np.random.default_rng(8)
    t = textures(rng)
    check(lib, t["nearcut"], tris(rng, 100, 5, 128), np.full(100, 5))                # every texel within 1e-6 of the cutoff: nothing may be skipped
    check(lib, t["smooth"], tris(rng, 150, 10, 256, -0.2, 1.2), np.full(150, 6))
    check(lib, t["noise8"], tris(rng, 150, 8, 512), np.full(150, 5), promotion=capi.PROMOTE_NEAREST)
    check(lib, t["noise8"], tris(rng, 150, 8, 512), np.full(150, 5), promotion=capi.PROMOTE_FORCE_TRANSPARENT, fmt=capi.FORMAT_2_STATE)
    check(lib, t["noise"], tris(rng, 150, 8, 1024), np.full(150, 5), gt=capi.STATE_T, le=capi.STATE_UO, cutoff=0.3)


def test_levels(lib):
    rng = np.random.default_rng(9)
    t = textures(rng)
    check(lib, t["noise"], tris(rng, 400, 10, 1024), rng.integers(0, 7, 400))       # levels 0..6: items smaller than one initial region
    check(lib, t["noise"], tris(rng, 6, 40, 1024), np.full(6, 9))                   # more than 64 initial regions per item
    check(lib, t["smooth"], tris(rng, 100, 60, 256), np.full(100, 3))               # micro-triangles of many texels: footprints too large to shortcut


def test_constant_areas_and_large_micro_triangles(lib):
    """(H): triangles of hundreds of texels at low levels over fully opaque / transparent areas with islands of detail."""
    rng = np.random.default_rng(10)
    tex = np.zeros((512, 512), dtype=np.float32)
    tex[:, 256:] = 1.0
    tex[100:140, 60:120] = W.noise_texture(64)[:40, :60]          # detail inside the transparent half
    tex[300:330, 300:360] = 0.5 + 1e-6                              # constant, but too close to the cutoff to be trusted
    st = check(lib, tex, tris(rng, 200, 120, 512, 0.1, 0.9), rng.integers(0, 5, 200))
    assert st.passes[0] + st.passes[1] + st.passes[2] > 0
    check(lib, tex, tris(rng, 100, 150, 512, -0.3, 1.3), rng.integers(0, 4, 100), addr=capi.ADDR_CLAMP)
    check(lib, (tex * 255).astype(np.uint8), tris(rng, 100, 100, 512, 0.1, 0.9), rng.integers(1, 5, 100))


def test_sat_pass_with_the_bake_cutoff(lib):
    """Textures created with an alpha cutoff equal to the bake's: the reference's SAT pass runs before the fine classification and
    the hierarchical path must agree with it (leaves apply it first; region proofs are compatible with any decisive SAT answer)."""
    rng = np.random.default_rng(11)
    t = textures(rng)
    check(lib, t["noise"], tris(rng, 200, 6, 1024, -0.3, 1.3), np.full(200, 6), use_sat=True)
    check(lib, t["checker"], tris(rng, 150, 9, 256, axis_aligned=True), np.full(150, 5), use_sat=True, addr=capi.ADDR_CLAMP)
    check(lib, t["noise8"], tris(rng, 150, 20, 512), rng.integers(0, 6, 150), use_sat=True, gt=capi.STATE_UO, le=capi.STATE_T)
    check(lib, t["smooth"], tris(rng, 150, 10, 256, -0.2, 1.2), np.full(150, 6), use_sat=True, promotion=capi.PROMOTE_NEAREST)


def test_sat_pass_where_the_address_mapping_folds(lib):
    """(S): the SDK's SAT rectangle spans the address-mapped texels of a micro-triangle's two bounding-box corners; where the mapping
    folds inside the range (Mirror / MirrorOnce across a mirror axis) it is not the set of texels the micro-triangle uses, and the
    region proofs must leave such work items to the exact walk.  First the case the randomized campaign found (MirrorOnce, UVs
    straddling u = 0: the proofs said "opaque", the folded rectangle held one transparent texel column), then every address mode
    with triangles across the axes and period boundaries."""
    tex = W.noise_texture(512, as_unorm8=True)
    uv = np.array([[0.00641006, 1.0755265, -0.0430925, 1.0589049, 0.00442714, 1.0566009]], dtype=np.float32)
    check(lib, tex, uv, np.array([4]), addr=capi.ADDR_MIRROR_ONCE, cutoff=0.5000001, promotion=capi.PROMOTE_NEAREST, fmt=capi.FORMAT_2_STATE, use_sat=True)
    rng = np.random.default_rng(13)
    t = textures(rng)
    for addr in (capi.ADDR_WRAP, capi.ADDR_MIRROR, capi.ADDR_CLAMP, capi.ADDR_BORDER, capi.ADDR_MIRROR_ONCE):
        for lo, hi in ((-0.06, 0.06), (0.94, 1.06), (-1.05, -0.95), (1.95, 2.05)):
            check(lib, t["noise8"], tris(rng, 60, 10, 512, lo, hi), rng.integers(2, 6, 60), addr=addr, use_sat=True, border=0.3)
            check(lib, t["npot"], tris(rng, 60, 7, 200, lo, hi), rng.integers(2, 6, 60), addr=addr, use_sat=True, promotion=capi.PROMOTE_NEAREST)
    small = W.noise_texture(256)[:8, :8].copy()
    check(lib, small, tris(rng, 150, 6, 8, -1.0, 2.0), rng.integers(0, 3, 150), addr=capi.ADDR_WRAP, use_sat=True)   # micro-triangles wider than the texture


def test_mip_chains(lib):
    """Several mips: a region must pass on every mip with the same side; leaves walk the mips like the reference (stop at Unknown)."""
    rng = np.random.default_rng(12)
    t = textures(rng)
    check(lib, t["noise"], tris(rng, 150, 8, 1024, -0.2, 1.2), np.full(150, 5), mips=4)
    check(lib, t["smooth"], tris(rng, 150, 12, 256), np.full(150, 6), mips=3, promotion=capi.PROMOTE_NEAREST)
    check(lib, t["noise8"], tris(rng, 150, 10, 512), rng.integers(0, 6, 150), mips=5, addr=capi.ADDR_MIRROR, le=capi.STATE_UO, gt=capi.STATE_T)
    check(lib, t["npot"], tris(rng, 120, 9, 200, -0.5, 1.5), np.full(120, 5), mips=3, addr=capi.ADDR_CLAMP)


def _sdk_states(ref_lib, tex, uv, lv, addr, cutoff, promotion, fmt, gt, le, border, use_sat):
    """Per-triangle micro-triangle states of the SDK build for the same inputs (every triangle keeps its own block)."""
    from omm_b200 import Baker
    n = len(lv)
    wl = W.Workload(name="plain-walk pin", mips=[np.ascontiguousarray(tex)], indices=np.arange(3 * n, dtype=np.uint32),
                    texcoords=np.ascontiguousarray(uv, dtype=np.float32).reshape(-1, 2), tex_alpha_cutoff=cutoff if use_sat else -1.0,
                    subdivision_levels=np.ascontiguousarray(lv, dtype=np.uint8),
                    desc=dict(addressing_mode=addr, filter=capi.FILTER_LINEAR, border_alpha=border, alpha_cutoff=cutoff, alpha_cutoff_gt=gt, alpha_cutoff_le=le,
                              format=fmt, unknown_state_promotion=promotion, max_subdivision_level=12, dynamic_subdivision_scale=0.0,
                              bake_flags=capi.BAKE_DISABLE_SPECIAL_INDICES | capi.BAKE_DISABLE_DUPLICATE_DETECTION))
    with Baker(ref_lib) as b:
        inp, t = W.make_input(b, wl)
        res = b.bake(inp)
        t.destroy()
    out = []
    for tri in range(n):
        d = res.desc_array[int(res.index_buffer[tri])]
        level, off = int(d["subdivisionLevel"]), int(d["offset"])
        m = 1 << (2 * level)
        i = np.arange(m)
        if fmt == capi.FORMAT_4_STATE:
            out.append((res.array_data[off + (i >> 2)] >> ((i & 3) * 2)) & 3)
        else:
            out.append((res.array_data[off + (i >> 3)] >> (i & 7)) & 1)
    return np.concatenate(out).astype(np.uint8)


def test_plain_walk_equals_the_sdk_build(lib):
    """Closes the loop on the CPU: the reference walk of omm_device_math.cuh (host build; the very code the flat kernels and the leaves run,
    and the yardstick of every test above) against the SDK build, micro-triangle by micro-triangle, on random configurations.  On the GPU
    box the same link is the byte-level parity suite."""
    ref_path = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libomm-lib.so")
    if not os.path.exists(ref_path):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    ref = capi.OmmLib(ref_path)
    rng = np.random.default_rng(77)
    t = textures(rng)
    lib.hier_host_set_state_sink.argtypes = [ctypes.c_void_p]
    for k in range(40):
        tex = t[["noise", "noise8", "smooth", "ramp", "checker", "npot", "nearcut"][k % 7]]
        n = 40
        lv = rng.integers(0, 6, n)
        addr = int(rng.choice([capi.ADDR_WRAP, capi.ADDR_MIRROR, capi.ADDR_CLAMP, capi.ADDR_MIRROR_ONCE]))  # Border: the SDK reads out of bounds
        lo = float(rng.choice([0.0, -0.5, -1.2]))
        uv = tris(rng, n, float(rng.choice([3, 8, 25])), max(tex.shape), lo, lo + float(rng.choice([1.0, 2.0])), axis_aligned=bool(k % 3 == 1), skinny=bool(k % 5 == 2))
        fmt = int(rng.choice([capi.FORMAT_4_STATE, capi.FORMAT_2_STATE]))
        promo = int(rng.integers(3))
        cutoff = float(rng.choice([0.5, 0.3, 0.5000001]))
        use_sat = bool(rng.random() < 0.4)
        sink = np.zeros(int((4 ** lv).sum()), dtype=np.uint8)
        lib.hier_host_set_state_sink(sink.ctypes.data_as(ctypes.c_void_p))
        try:
            check(lib, tex, uv, lv, addr=addr, cutoff=cutoff, promotion=promo, fmt=fmt, use_sat=use_sat)
        finally:
            lib.hier_host_set_state_sink(None)
        want = _sdk_states(ref, tex, uv, lv, addr, cutoff, promo, fmt, capi.STATE_O, capi.STATE_T, 0.0, use_sat)
        bad = np.nonzero(sink != want)[0]
        assert bad.size == 0, f"run {k}: {bad.size} of {want.size} micro-triangles differ from the SDK build (first at {bad[:5]}, got {sink[bad[:5]]}, SDK {want[bad[:5]]})"


def test_random_campaign_slice(lib):
    """A bounded, seeded slice of scripts/host_campaign.py inside the CPU suite, so that a regression in one of the exactness bounds of
    omm_hier.cuh (the constants of (A)-(G), (S)) is caught by `pytest -m "not gpu"` and not only by a hand-run campaign.  ~25 s; the
    generator is tests/campaign.py::host_campaign_step (random textures, address modes, cutoffs on texel values, levels 0-9, mips, SAT)."""
    import time

    import campaign
    rng = np.random.default_rng(20261017)
    fixed = textures(rng)
    t0, runs, total = time.time(), 0, 0
    while runs < 400 and time.time() - t0 < 25.0:
        st, _ = campaign.host_campaign_step(__import__("test_hier_host"), lib, rng, fixed)
        total += st.microTriangles
        runs += 1
    assert runs >= 20 and total > 1_000_000, (runs, total)
