"""Synthetic bake inputs for the BASELINE.json configs, generated with numpy only.

Every generator is a pure function of its arguments (splitmix64 streams with fixed seeds), so the oracle, the
SDK build and the CUDA library are always fed byte-identical host buffers.

    config1()                  C1: one quad, 256x256 checkerboard, level 3, 2-state
    config2(...)               C2: 10k-triangle leaf-card mesh with 25 % UV reuse, 1024x1024 value noise, level 4, 4-state
    config3(num_tris, ...)     C3: jittered 708x708-cell triangle grid, 4096x4096 2-octave value noise, level 6, 4-state
    config5(...)               C5: mixed per-triangle levels + heavy UV / block reuse
The scaled-down variants used by the parity tests keep the texel-per-triangle ratio of the full config.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import capi

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    """Vectorised splitmix64 finaliser of (seed + index * golden) -- public-domain algorithm by S. Vigna."""
    with np.errstate(over="ignore"):
        z = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def _stream(seed: int, n: int, salt: int = 0) -> np.ndarray:
    with np.errstate(over="ignore"):
        idx = np.arange(n, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
        return splitmix64(idx + np.uint64(seed) + np.uint64(salt) * np.uint64(0xD1B54A32D192ED03))


def _unit(seed: int, n: int, salt: int = 0) -> np.ndarray:
    """float64 uniform in [0,1) from the top 53 bits."""
    return (_stream(seed, n, salt) >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


def value_noise_u8(size: int, cell: int, seed: int) -> np.ndarray:
    """8-bit integer value noise, periodic over `size`, lattice spacing `cell` texels, integer bilinear weights."""
    assert size % cell == 0
    n = size // cell
    with np.errstate(over="ignore"):
        ix, iy = np.meshgrid(np.arange(n, dtype=np.uint64), np.arange(n, dtype=np.uint64))
        lat = (splitmix64(ix * np.uint64(0x9E3779B97F4A7C15) + iy * np.uint64(0xC2B2AE3D27D4EB4F) + np.uint64(seed)) & np.uint64(0xFF)).astype(np.int64)
    x = np.arange(size)
    cx, fx = x // cell, x % cell
    cx1 = (cx + 1) % n
    v00 = lat[np.ix_(cx, cx)]
    v10 = lat[np.ix_(cx, cx1)]   # rows index y, cols index x
    v01 = lat[np.ix_(cx1, cx)]
    v11 = lat[np.ix_(cx1, cx1)]
    wx = fx[None, :]
    wy = fx[:, None]
    num = (cell - wx) * (cell - wy) * v00 + wx * (cell - wy) * v10 + (cell - wx) * wy * v01 + wx * wy * v11
    return (num // (cell * cell)).astype(np.uint8)


def noise_texture(size: int, cells=(8, 4), weights=(3, 1), seed: int = 0x0A11, as_unorm8: bool = False) -> np.ndarray:
    """2-octave integer value noise; FP32 texels are u8 * (1/255) in float32 arithmetic."""
    acc = np.zeros((size, size), dtype=np.int64)
    for k, (c, w) in enumerate(zip(cells, weights)):
        acc += w * value_noise_u8(size, c, seed + 0x1000 * k).astype(np.int64)
    u8 = (acc // sum(weights)).astype(np.uint8)
    if as_unorm8:
        return u8
    return (u8.astype(np.float32) * np.float32(1.0 / 255.0)).astype(np.float32)


@dataclass
class Workload:
    """Host buffers + desc fields of one bake.  `mips` are 2-D arrays (float32 or uint8)."""
    name: str
    mips: list
    indices: np.ndarray
    texcoords: np.ndarray
    texcoord_format: int = capi.UV32_FLOAT
    tex_alpha_cutoff: float = -1.0
    tex_flags: int = capi.TEXFLAG_NONE
    desc: dict = field(default_factory=dict)   # keyword overrides for baker.BakeInput
    subdivision_levels: Optional[np.ndarray] = None
    formats: Optional[np.ndarray] = None

    @property
    def num_triangles(self) -> int:
        return self.indices.size // 3

    def micro_triangles(self) -> int:
        """Upper bound: 4^level per triangle (before UV pre-dedup)."""
        if self.subdivision_levels is not None:
            lv = np.minimum(self.subdivision_levels.astype(np.int64), 12)
            return int((4 ** lv).sum())
        return self.num_triangles * 4 ** int(self.desc.get("max_subdivision_level", 8))


def config1() -> Workload:
    """C1: indices {0,1,2,3,1,2}, UVs of test_omm_bake_cpu.cpp:594-595, 256^2 FP32 checkerboard."""
    j, i = np.meshgrid(np.arange(256), np.arange(256), indexing="ij")
    tex = (((i >> 5) + (j >> 5)) & 1).astype(np.float32)
    return Workload(
        name="C1 quad/256^2 checker/L3/2-state",
        mips=[tex],
        indices=np.array([0, 1, 2, 3, 1, 2], dtype=np.uint32),
        texcoords=np.array([[0, 0], [0, 1], [1, 0], [1, 1]], dtype=np.float32),
        desc=dict(addressing_mode=capi.ADDR_CLAMP, filter=capi.FILTER_LINEAR, alpha_cutoff=0.5, format=capi.FORMAT_2_STATE,
                  unknown_state_promotion=capi.PROMOTE_FORCE_OPAQUE, max_subdivision_level=3, dynamic_subdivision_scale=0.0),
    )


def config2(num_quads: int = 5000, tex_size: int = 1024, level: int = 4, reuse: float = 0.25, seed: int = 0xC2,
            unorm8: bool = False) -> Workload:
    """C2: leaf cards -- randomly placed/rotated quads (2 triangles each, indexed, 16-bit-index friendly when small);
    `reuse` of the quads reference the UVs of an earlier quad (instancing => UV pre-dedup)."""
    cx, cy = _unit(seed, num_quads, 1), _unit(seed, num_quads, 2)
    half = (6.0 + 18.0 * _unit(seed, num_quads, 3)) / tex_size          # 6..24 texel half-extent
    ang = 2 * np.pi * _unit(seed, num_quads, 4)
    ca, sa = np.cos(ang), np.sin(ang)
    corners = np.array([[-1, -1], [1, -1], [-1, 1], [1, 1]], dtype=np.float64)
    uv = np.empty((num_quads, 4, 2), dtype=np.float64)
    for k in range(4):
        dx, dy = corners[k, 0] * half, corners[k, 1] * half * 0.6
        uv[:, k, 0] = cx + ca * dx - sa * dy
        uv[:, k, 1] = cy + sa * dx + ca * dy
    uv = uv.astype(np.float32)
    src = np.arange(num_quads)
    pick = _unit(seed, num_quads, 5) < reuse
    earlier = (_unit(seed, num_quads, 6) * np.maximum(src, 1)).astype(np.int64)
    src = np.where(pick & (src > 0), earlier, src)
    for _ in range(32):                              # resolve chains so reused quads copy an original
        src = src[src]
    uv = uv[src]
    base = (np.arange(num_quads, dtype=np.uint32) * 4)[:, None]
    idx = (base + np.array([0, 1, 2, 3, 2, 1], dtype=np.uint32)[None, :]).reshape(-1)
    tex = noise_texture(tex_size, cells=(32, 8), weights=(3, 1), seed=0xA1FA, as_unorm8=unorm8)
    return Workload(
        name=f"C2 {2 * num_quads} tris/{tex_size}^2 noise/L{level}/4-state",
        mips=[tex], indices=idx, texcoords=uv.reshape(-1, 2),
        desc=dict(addressing_mode=capi.ADDR_WRAP, filter=capi.FILTER_LINEAR, alpha_cutoff=0.5, format=capi.FORMAT_4_STATE,
                  unknown_state_promotion=capi.PROMOTE_FORCE_OPAQUE, max_subdivision_level=level, dynamic_subdivision_scale=0.0),
    )


def config3(num_tris: int = 1_000_000, tex_size: int = 4096, level: int = 6, grid: Optional[int] = None, seed: int = 0xB200,
            cells=(8, 4), first_tri: int = 0, promotion: int = capi.PROMOTE_FORCE_OPAQUE, texture: Optional[np.ndarray] = None,
            tex_alpha_cutoff: float = -1.0) -> Workload:
    """C3 (headline): unindexed triangles on a jittered cell grid (708 cells across a 4096-texel texture, i.e. one cell =
    5.79 texels), two triangles per cell, every vertex jittered +-0.25 cell so no two triangles share UVs.
    `first_tri`/`num_tris` select a slice of the same global triangle sequence (used for the bounded CPU-baseline sample)."""
    if grid is None:
        grid = max(2, int(round(708 * tex_size / 4096)))
    t = np.arange(first_tri, first_tri + num_tris, dtype=np.int64)
    cell = t // 2
    cxi, cyi = cell % grid, (cell // grid) % grid
    upper = (t % 2).astype(np.int64)
    # corner offsets (in cells) of the two triangles of a cell: lower = (0,0)(1,0)(0,1), upper = (1,1)(0,1)(1,0)
    ox = np.stack([upper, 1 - upper, upper], axis=1).astype(np.float64)
    oy = np.stack([upper, upper, 1 - upper], axis=1).astype(np.float64)
    vid = (t[:, None] * 3 + np.arange(3)[None, :]).astype(np.uint64)
    with np.errstate(over="ignore"):
        jx = (splitmix64(vid * np.uint64(2) + np.uint64(seed)) >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))
        jy = (splitmix64(vid * np.uint64(2) + np.uint64(1) + np.uint64(seed)) >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))
    # shrink the triangle slightly towards its cell so jitter keeps it inside [0,1)
    u = (cxi[:, None] + 0.25 + 0.5 * ox + (jx - 0.5) * 0.5) / grid
    v = (cyi[:, None] + 0.25 + 0.5 * oy + (jy - 0.5) * 0.5) / grid
    uv = np.stack([u, v], axis=2).astype(np.float32).reshape(-1, 2)
    idx = np.arange(3 * num_tris, dtype=np.uint32)
    tex = texture if texture is not None else noise_texture(tex_size, cells=cells, weights=(3, 1), seed=0x0A11)
    return Workload(
        name=f"C3 {num_tris} tris/{tex_size}^2 noise{cells}/L{level}/4-state",
        mips=[tex], indices=idx, texcoords=uv, tex_alpha_cutoff=tex_alpha_cutoff,
        desc=dict(addressing_mode=capi.ADDR_WRAP, filter=capi.FILTER_LINEAR, alpha_cutoff=0.5, format=capi.FORMAT_4_STATE,
                  unknown_state_promotion=promotion, max_subdivision_level=level, dynamic_subdivision_scale=0.0),
    )


def config5(num_tris: int = 1_000_000, tex_size: int = 4096, distinct: int = 4096, flat_tris: int = 65536, max_level: int = 12,
            seed: int = 0xC5) -> Workload:
    """C5: `num_tris` triangles drawn (Zipf-like) from `distinct` UV triangles, plus `flat_tris` triangles lying in
    constant-alpha regions (identical block content => XXH64 dedup stress); per-triangle levels with P(l) ~ 4^-l."""
    base = config3(distinct, tex_size=tex_size, level=0, seed=seed)
    buv = base.texcoords.reshape(distinct, 3, 2)
    r = _unit(seed, num_tris, 1)
    pick = np.minimum((distinct * r ** 3).astype(np.int64), distinct - 1)       # heavy head
    uv = buv[pick].copy()
    # per-DISTINCT-triangle level so that reused UVs share level (=> UV pre-dedup hits)
    lr = _unit(seed, distinct, 2)
    lv_d = np.minimum(np.floor(-np.log(np.maximum(lr, 1e-12)) / np.log(4.0) * 1.6).astype(np.int64), max_level)
    levels = lv_d[pick].astype(np.uint8)
    # flat region: paint a constant block into the texture and put small distinct triangles there
    tex = base.mips[0].copy()
    q = tex_size // 4
    tex[:q, :q] = np.float32(1.0)
    tex[:q, q:2 * q] = np.float32(0.0)
    nf = min(flat_tris, num_tris)
    fu = _unit(seed, nf * 3, 3).reshape(nf, 3)
    fv = _unit(seed, nf * 3, 4).reshape(nf, 3)
    side = (_unit(seed, nf, 5) < 0.5)
    x0 = np.where(side, 0.02, 0.27)[:, None]
    fuv = np.stack([x0 + 0.2 * fu, 0.02 + 0.2 * fv], axis=2).astype(np.float32)
    uv[:nf] = fuv
    levels[:nf] = np.minimum(3, max_level)
    idx = np.arange(3 * num_tris, dtype=np.uint32)
    return Workload(
        name=f"C5 {num_tris} tris mixed L0-{max_level}, {distinct} distinct + {nf} flat",
        mips=[tex], indices=idx, texcoords=uv.reshape(-1, 2), subdivision_levels=levels,
        desc=dict(addressing_mode=capi.ADDR_WRAP, filter=capi.FILTER_LINEAR, alpha_cutoff=0.5, format=capi.FORMAT_4_STATE,
                  unknown_state_promotion=capi.PROMOTE_FORCE_OPAQUE, max_subdivision_level=max_level, dynamic_subdivision_scale=0.0),
    )


def make_input(baker, wl: Workload, **overrides):
    """Create the texture on `baker` and return (BakeInput, texture)."""
    from .baker import BakeInput
    tex = baker.create_texture(wl.mips, alpha_cutoff=wl.tex_alpha_cutoff, flags=wl.tex_flags)
    kw = dict(wl.desc)
    kw.update(overrides)
    inp = BakeInput(texture=tex, indices=wl.indices, texcoords=wl.texcoords, texcoord_format=wl.texcoord_format,
                    subdivision_levels=wl.subdivision_levels, formats=wl.formats, **kw)
    return inp, tex


def random_mesh(seed: int, num_tris: int, tex_size=(256, 256), tri_texels: float = 12.0, uv_lo: float = 0.0, uv_hi: float = 1.0,
                tex_kind: str = "noise", unorm8: bool = False, mips: int = 1, index_dtype=np.uint32, shared_vertices: bool = True,
                degenerate_frac: float = 0.0, nan_frac: float = 0.0, reuse_frac: float = 0.0, **desc) -> Workload:
    """General-purpose parity workload: random triangles of about `tri_texels` texels across, centred uniformly in
    [uv_lo, uv_hi]^2 (values outside [0,1] exercise the address modes), over a noise / circle / blocky texture.
    `degenerate_frac` of the triangles are collapsed to lines or points, `nan_frac` get a NaN/Inf coordinate and
    `reuse_frac` repeat the UVs of an earlier triangle."""
    w, h = tex_size
    n = num_tris
    cx = uv_lo + (uv_hi - uv_lo) * _unit(seed, n, 1)
    cy = uv_lo + (uv_hi - uv_lo) * _unit(seed, n, 2)
    rad = tri_texels / max(w, h) * (0.3 + 0.7 * _unit(seed, n, 3))
    uv = np.empty((n, 3, 2), dtype=np.float64)
    for k in range(3):
        ang = 2 * np.pi * (_unit(seed, n, 4 + k) / 3.0 + k / 3.0)
        uv[:, k, 0] = cx + rad * np.cos(ang)
        uv[:, k, 1] = cy + rad * np.sin(ang)
    uv = uv.astype(np.float32)
    sel = _unit(seed, n, 8)
    kind = _unit(seed, n, 9)
    deg = sel < degenerate_frac
    line = deg & (kind < 0.6)
    point = deg & (kind >= 0.6)
    # line: third vertex on the segment p0-p1 (exact midpoint of equal endpoints keeps area exactly 0 for axis-aligned lines)
    uv[line, 2] = uv[line, 0]
    uv[point, 1] = uv[point, 0]
    uv[point, 2] = uv[point, 0]
    bad = (sel >= degenerate_frac) & (sel < degenerate_frac + nan_frac)
    uv[bad & (kind < 0.5), 1, 0] = np.float32(np.nan)
    uv[bad & (kind >= 0.5), 2, 1] = np.float32(np.inf)
    if reuse_frac > 0:
        pick = _unit(seed, n, 10) < reuse_frac
        src = np.arange(n)
        earlier = (_unit(seed, n, 11) * np.maximum(src, 1)).astype(np.int64)
        src = np.where(pick & (src > 0), earlier, src)
        for _ in range(32):
            src = src[src]
        uv = uv[src]
    if shared_vertices and np.dtype(index_dtype) != np.dtype(np.uint32):
        maxv = {np.dtype(np.uint8): 255, np.dtype(np.uint16): 65535}[np.dtype(index_dtype)]
        assert 3 * n <= maxv + 1
    idx = np.arange(3 * n).astype(index_dtype)
    size = max(w, h)
    if tex_kind == "noise":
        base = noise_texture(1 << int(np.ceil(np.log2(size))), cells=(8, 4), weights=(3, 1), seed=seed ^ 0x7E57, as_unorm8=True)[:h, :w]
    elif tex_kind == "blocky":
        yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
        base = ((((xx // 7) + (yy // 5)) % 3 == 0) * 255).astype(np.uint8)
    elif tex_kind == "circle":
        yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
        r = np.sqrt(((xx + 0.5) / w - 0.5) ** 2 + ((yy + 0.5) / h - 0.5) ** 2)
        base = np.clip((r - 0.3) * 4 * 255, 0, 255).astype(np.uint8)
    else:
        raise ValueError(tex_kind)
    chain = [base]
    for _ in range(1, mips):
        p = chain[-1]
        hh, ww = max(1, p.shape[0] // 2), max(1, p.shape[1] // 2)
        q = p[:hh * 2, :ww * 2].astype(np.uint16) if p.shape[0] >= 2 and p.shape[1] >= 2 else None
        if q is None:
            chain.append(p[:hh, :ww].copy())
        else:
            chain.append(((q[0::2, 0::2] + q[1::2, 0::2] + q[0::2, 1::2] + q[1::2, 1::2]) // 4).astype(np.uint8))
    if not unorm8:
        chain = [(m.astype(np.float32) * np.float32(1.0 / 255.0)).astype(np.float32) for m in chain]
    d = dict(addressing_mode=capi.ADDR_WRAP, filter=capi.FILTER_LINEAR, alpha_cutoff=0.5, format=capi.FORMAT_4_STATE,
             unknown_state_promotion=capi.PROMOTE_FORCE_OPAQUE, max_subdivision_level=4, dynamic_subdivision_scale=0.0)
    tex_cut = desc.pop("tex_alpha_cutoff", -1.0)
    tex_flags = desc.pop("tex_flags", capi.TEXFLAG_NONE)
    levels = desc.pop("subdivision_levels", None)
    formats = desc.pop("formats", None)
    d.update(desc)
    return Workload(name=f"random_mesh(seed={seed},n={n},{w}x{h},{tex_kind})", mips=chain, indices=idx, texcoords=uv.reshape(-1, 2),
                    tex_alpha_cutoff=tex_cut, tex_flags=tex_flags, desc=d, subdivision_levels=levels, formats=formats)


def pack_unorm16(uv: np.ndarray) -> np.ndarray:
    """glm::packUnorm2x16 of float UVs (round(clamp(v,0,1) * 65535))."""
    q = np.round(np.clip(uv.astype(np.float32), 0.0, 1.0) * np.float32(65535.0)).astype(np.uint32)
    return (q[:, 0] | (q[:, 1] << 16)).astype(np.uint32)


def pack_half(uv: np.ndarray) -> np.ndarray:
    hbits = uv.astype(np.float16).view(np.uint16).astype(np.uint32)
    return (hbits[:, 0] | (hbits[:, 1] << 16)).astype(np.uint32)
