"""Unit known-answer tests of the restated helpers against golden vectors produced from the reference checkout by
tests/golden/make_golden.py (address modes: support/tests/test_texture.cpp; XXH64: external/xxHash; std::hash<float>:
libstdc++; Morton: src/util/bit_tricks.h) and against the SDK build's result digests."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

import parity_cases as PC

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


def test_texcoord_address_modes(port_lib):
    f = port_lib.dll.omm_oracle_texcoord
    f.restype, f.argtypes = C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]
    kats = _load("texcoord_kat.json")
    assert len(kats) >= 180
    for mode, x, y, w, h, ex, ey in kats:
        pow2 = int((w & (w - 1)) == 0 and (h & (h - 1)) == 0)
        assert f(mode, pow2, x, w) == ex, (mode, x, w)
        assert f(mode, pow2, y, h) == ey, (mode, y, h)


def _xorshift_bytes(n):
    s, out = 88172645463325252, bytearray()
    for _ in range(n):
        s ^= (s << 13) & 0xFFFFFFFFFFFFFFFF
        s ^= s >> 7
        s ^= (s << 17) & 0xFFFFFFFFFFFFFFFF
        out.append((s >> 32) & 0xFF)
    return bytes(out)


def test_xxh64_matches_vendored_xxhash(port_lib):
    f = port_lib.dll.omm_oracle_xxh64
    f.restype, f.argtypes = C.c_uint64, [C.c_char_p, C.c_size_t, C.c_uint64]
    g = _load("xxh64.json")
    buf = _xorshift_bytes(5000)
    for length, seed, want in g["plain"]:
        assert f(buf, length, seed) == int(want), (length, seed)
    for lvl, want in g["states"]:
        n = 1 << (2 * lvl)
        st = bytes((3 if (b % 3) == 2 else (b % 3)) for b in buf[:n])
        assert f(st, n, 42) == int(want), lvl


def test_std_hash_float(port_lib):
    f = port_lib.dll.omm_oracle_std_hash_float
    f.restype, f.argtypes = C.c_uint64, [C.c_float]
    for bits, want in _load("std_hash_float.json"):
        v = np.array([bits], dtype=np.uint32).view(np.float32)[0]
        assert f(float(v)) == int(want), hex(bits)


def test_morton(port_lib):
    f = port_lib.dll.omm_oracle_morton
    f.restype, f.argtypes = C.c_uint32, [C.c_uint32, C.c_uint32]
    for x, y, want in _load("morton.json"):
        assert f(x, y) == want


def test_bird_curve_is_a_bijection_with_shared_lattice_vertices(port_lib):
    """index2bary: every level-3 micro-triangle is distinct, has area 4^-3/2 in barycentric space and consecutive indices
    touch (the curve is contiguous)."""
    f = port_lib.dll.omm_oracle_index2bary
    f.restype, f.argtypes = None, [C.c_uint32, C.c_uint32, C.POINTER(C.c_float)]
    out = (C.c_float * 6)()
    seen, prev = set(), None
    for i in range(64):
        f(i, 3, out)
        tri = tuple(round(v * 8) for v in out)
        verts = frozenset([tri[0:2], tri[2:4], tri[4:6]])
        assert len(verts) == 3 and verts not in seen
        seen.add(verts)
        if prev is not None:
            assert len(verts & prev) >= 1, i
        prev = verts
    assert len(seen) == 64


def _digests(res):
    d = {k: hashlib.sha256(getattr(res, k).tobytes()).hexdigest() for k in ("array_data", "desc_array", "desc_histogram", "index_buffer", "index_histogram")}
    d.update(index_format=int(res.index_format), array_bytes=int(res.array_data.size), descs=int(res.desc_array.size))
    return d


GOLDEN_BAKES = _load("bake_digests.json")


def _case(name):
    return PC.sdk_only_cases()[name[len("sdk_only:"):]] if name.startswith("sdk_only:") else PC.cases()[name]


@pytest.mark.parametrize("name", sorted(n for n in GOLDEN_BAKES if not n.startswith("sdk_only:")))
def test_port_matches_golden_sdk_digests(name, port_lib):
    mk, ov = PC.cases()[name]
    assert _digests(PC.run_bake(port_lib, mk(), **ov)) == GOLDEN_BAKES[name]


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GOLDEN_BAKES))
def test_product_matches_golden_sdk_digests(name, product_lib):
    mk, ov = _case(name)
    assert _digests(PC.run_bake(product_lib, mk(), **ov)) == GOLDEN_BAKES[name]
