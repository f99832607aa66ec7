"""BASELINE config 3 at FULL size (1 M triangles, 4096^2 alpha, level 6 = 4.1e9 micro-triangles) on the GPU, tied to the oracle by a
size-independent property: classification is per work item, so the block (or special index) a triangle ends up with in the full
bake must be byte-identical with what the SDK's CPU baker produces for the same triangle in a bake of only the first few thousand
triangles of the same sequence (layout differs -- dedup survivors, sort order, offsets -- content per triangle does not)."""
import os

import numpy as np
import pytest

from omm_b200 import Baker, capi, load_product_library
from omm_b200 import workloads as W

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "libomm-lib.so")
PORT = os.path.join(ROOT, "oracle", "liboracle_port.so")


def _bake(lib, wl, **kw):
    with Baker(lib) as b:
        inp, tex = W.make_input(b, wl, **kw)
        try:
            return b.bake(inp)
        finally:
            tex.destroy()


def _blocks(res, tris):
    """per triangle: ('special', idx) or ('block', level, format, bytes)"""
    out = []
    idx = res.index_buffer.astype(np.int64)
    for t in tris:
        i = int(idx[t])
        if i < 0:
            out.append(("special", i))
            continue
        d = res.desc_array[i]
        lvl, fmt, off = int(d["subdivisionLevel"]), int(d["format"]), int(d["offset"])
        n = 1 << (2 * lvl)
        nbytes = max(1, (n * fmt) // 8)
        out.append(("block", lvl, fmt, res.array_data[off:off + nbytes].tobytes()))
    return out


def test_config3_full_size_matches_the_oracle_per_triangle():
    lib = load_product_library()
    oracle_path = REF if os.path.exists(REF) else PORT
    sample = 6000 if oracle_path == REF else 400
    full = _bake(lib, W.config3())
    assert full.index_buffer.size == 1_000_000
    # structural invariants of the result (ref: bake_cpu_impl.cpp:1756-1920)
    descs = full.desc_array
    sizes = np.maximum(1, ((1 << (2 * descs["subdivisionLevel"].astype(np.int64))) * descs["format"].astype(np.int64)) // 8)
    assert np.array_equal(descs["offset"].astype(np.int64), np.concatenate([[0], np.cumsum(sizes)[:-1]]))
    assert int(sizes.sum()) == full.array_data.size
    idx = full.index_buffer.astype(np.int64)
    assert idx.min() >= -4 and idx.max() == descs.size - 1
    assert int(full.desc_histogram["count"].sum()) == descs.size
    assert int(full.index_histogram["count"].sum()) == int((idx >= 0).sum())
    # determinism: a second bake is byte-identical
    assert not full.diff(_bake(lib, W.config3()))
    # per-triangle content against the CPU oracle on the first `sample` triangles of the same sequence
    oracle = _bake(capi.OmmLib(oracle_path), W.config3(num_tris=sample), bake_flags=capi.BAKE_ENABLE_INTERNAL_THREADS)
    tris = range(sample)
    assert _blocks(full, tris) == _blocks(oracle, tris)


def test_config5_full_size_matches_the_oracle_per_triangle():
    """BASELINE config 5 (1 M triangles drawn from 4096 distinct UV triangles + 65 k large triangles over constant areas, per-triangle
    levels 0..12): the full bake on the GPU against the CPU oracle on a subset of its triangles (those of level <= 7, so that the
    CPU finishes in seconds)."""
    import copy
    lib = load_product_library()
    oracle_path = REF if os.path.exists(REF) else PORT
    wl = W.config5()
    full = _bake(lib, wl)
    assert full.index_buffer.size == wl.num_triangles
    rng = np.random.default_rng(5)
    lv = wl.subdivision_levels
    cand = np.nonzero((lv <= 7) & (np.arange(lv.size) >= 65536))[0]
    flat = np.nonzero(np.arange(lv.size) < 65536)[0]
    pick = np.sort(np.concatenate([rng.choice(cand, 1500 if oracle_path == REF else 150, replace=False), rng.choice(flat, 4, replace=False)]))
    sub = copy.copy(wl)
    uv = wl.texcoords.reshape(-1, 3, 2)
    sub.texcoords = np.ascontiguousarray(uv[pick].reshape(-1, 2))
    sub.indices = np.arange(3 * pick.size, dtype=np.uint32)
    sub.subdivision_levels = np.ascontiguousarray(lv[pick])
    oracle = _bake(capi.OmmLib(oracle_path), sub, bake_flags=capi.BAKE_ENABLE_INTERNAL_THREADS)
    assert _blocks(full, pick) == _blocks(oracle, range(pick.size))
