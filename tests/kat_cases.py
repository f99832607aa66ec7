"""Known-answer tests of the SDK, transcribed from support/tests/test_omm_bake_cpu.cpp (the line of each expected-count
block is cited).  The SDK pins STATE COUNTS (ommDebugStats), identical for all six of its suite configurations
(Default / TextureDisableZOrder / Force32BitIndices / TextureAsUNORM8* / AlphaCutoff / Serialize).

Harness defaults of the SDK tests (test_omm_bake_cpu.cpp:42-61, 165-205): Clamp addressing, Linear filter, Nearest
unknown-state promotion, 4-state, dynamicSubdivisionScale 0, quad indices {0,1,2,3,1,2} with UVs (0,0)(0,1)(1,0)(1,1).
Textures are generated with the C library's sinf so that texel bits equal the SDK test's std::sin(float)."""
import ctypes
import ctypes.util

import numpy as np

from omm_b200 import capi
from omm_b200.workloads import Workload

_libm = ctypes.CDLL(ctypes.util.find_library("m"))
_libm.sinf.restype = ctypes.c_float
_libm.sinf.argtypes = [ctypes.c_float]

QUAD_IDX = np.array([0, 1, 2, 3, 1, 2], dtype=np.uint32)
QUAD_UV = np.array([[0, 0], [0, 1], [1, 0], [1, 1]], dtype=np.float32)


def _tex(w, h, fn, dtype=np.float32):
    t = np.empty((h, w), dtype=dtype)
    for j in range(h):
        for i in range(w):
            t[j, i] = fn(i, j, w, h)
    return t


def _const(v):
    return lambda w, h: np.full((h, w), np.float32(v), dtype=np.float32)


def standard_circle(w, h):
    """ref: test_omm_bake_cpu.cpp:64-76"""
    i, j = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
    u = (i / np.float32(w)).astype(np.float32) - np.float32(0.5)
    v = (j / np.float32(w)).astype(np.float32) - np.float32(0.5)
    length = np.sqrt((u * u + v * v).astype(np.float32)).astype(np.float32)
    t = np.where(length < np.float32(0.4), np.float32(0.0), np.float32(1.0)).astype(np.float32)
    t[0, 0] = np.float32(0.6)
    return t


def sine_fp32(w, h):
    """ref: test_omm_bake_cpu.cpp:1025-1032"""
    row = np.array([np.float32(1.0) - np.float32(_libm.sinf(np.float32(np.float32(i) / np.float32(w)) * np.float32(15))) for i in range(w)], dtype=np.float32)
    t = np.tile(row, (h, 1))
    t[0, 0] = np.float32(0.6)
    return t


def sine_unorm8(w, h):
    """ref: test_omm_bake_cpu.cpp:1005-1010"""
    def val(i):
        uv = np.float32(np.float32(i) / np.float32(w))
        v = np.float32(0.5) - np.float32(0.5) * np.float32(_libm.sinf(np.float32(uv * np.float32(15))))
        return np.uint8(int(np.float32(v * np.float32(255.0))))
    row = np.array([val(i) for i in range(w)], dtype=np.uint8)
    return np.tile(row, (h, 1))


def diag8(on, off):
    def f(w, h):
        i, j = np.meshgrid(np.arange(w), np.arange(h))
        return np.where((i % 8) != (j % 8), np.float32(on), np.float32(off)).astype(np.float32)
    return f


def corner(w, h):
    t = np.full((h, w), np.float32(0.4), dtype=np.float32)
    t[0, 0] = np.float32(0.6)
    return t


def uniform4(w, h):
    """ref: test_omm_bake_cpu.cpp:1399-1412"""
    vals = [0.9, 0.1, 0.1, 0.7]
    return _tex(w, h, lambda i, j, w_, h_: np.float32(1.0) - np.float32(vals[(i % 2) + 2 * (j % 2)]))


def _wl(name, tex, level, idx=QUAD_IDX, uv=QUAD_UV, **desc):
    d = dict(addressing_mode=capi.ADDR_CLAMP, filter=capi.FILTER_LINEAR, alpha_cutoff=0.5, format=capi.FORMAT_4_STATE,
             unknown_state_promotion=capi.PROMOTE_NEAREST, max_subdivision_level=level, dynamic_subdivision_scale=0.0,
             bake_flags=capi.BAKE_ENABLE_INTERNAL_THREADS)
    d.update(desc)
    return Workload(name=name, mips=[tex], indices=idx, texcoords=uv, desc=d)


# name -> (workload factory, expected stats dict [unlisted fields are 0], reference line)
def kats():
    k = {}
    for lvl in range(5):
        k[f"AllOpaque{lvl}"] = (lambda lvl=lvl: _wl("AllOpaque", _const(0.6)(1024, 1024), lvl), dict(totalFullyOpaque=2), "791-849")
    for lvl in range(1, 5):
        k[f"AllTransparent{lvl}"] = (lambda lvl=lvl: _wl("AllTransparent", _const(0.4)(1024, 1024), lvl), dict(totalFullyTransparent=2), "851-897")
    k["AllUnknownTransparent"] = (lambda: _wl("AUT", diag8(0.0, 1.0)(1024, 1024), 1), dict(totalFullyUnknownTransparent=2), "899-911")
    k["AllUnknownOpaque"] = (lambda: _wl("AUO", diag8(1.0, 0.0)(1024, 1024), 1), dict(totalFullyUnknownOpaque=2), "913-925")
    k["AllTransparentOpaqueCorner4"] = (lambda: _wl("corner", corner(1024, 1024), 4),
                                        dict(totalTransparent=255, totalUnknownTransparent=1, totalFullyTransparent=1), "927-943")
    k["Circle"] = (lambda: _wl("Circle", standard_circle(1024, 1024), 4),
                   dict(totalOpaque=204, totalTransparent=219, totalUnknownTransparent=39, totalUnknownOpaque=50), "958-971")
    k["CircleOC2"] = (lambda: _wl("CircleOC2", standard_circle(1024, 1024), 4, format=capi.FORMAT_2_STATE),
                      dict(totalOpaque=254, totalTransparent=258), "988-999")
    k["SineUNORM8"] = (lambda: _wl("SineUNORM8", sine_unorm8(1024, 1024), 4),
                       dict(totalOpaque=128, totalTransparent=256, totalUnknownTransparent=48, totalUnknownOpaque=80), "1001-1018")
    k["Sine"] = (lambda: _wl("Sine", sine_fp32(1024, 1024), 4),
                 dict(totalOpaque=224, totalTransparent=128, totalUnknownTransparent=96, totalUnknownOpaque=64), "1020-1039")
    k["SineOC2"] = (lambda: _wl("SineOC2", sine_fp32(1024, 1024), 4, format=capi.FORMAT_2_STATE),
                    dict(totalOpaque=288, totalTransparent=224), "1041-1058")
    k["Uniform"] = (lambda: _wl("Uniform", uniform4(4, 4), 6, idx=np.array([0, 1, 2, 1, 2, 3], dtype=np.uint32),
                                uv=np.array([[0, 0], [0, 1], [1, 1], [1, 0]], dtype=np.float32)),
                    dict(totalOpaque=5132, totalTransparent=2393, totalUnknownTransparent=357, totalUnknownOpaque=310), "1389-1420")
    return k


STAT_FIELDS = ["totalOpaque", "totalTransparent", "totalUnknownTransparent", "totalUnknownOpaque", "totalFullyOpaque", "totalFullyTransparent",
               "totalFullyUnknownOpaque", "totalFullyUnknownTransparent"]


def collect_stats(res):
    """Python restatement of CollectStats (ref: libraries/omm-lib/src/debug_impl.cpp:512-641) over a BakeResult."""
    s = dict.fromkeys(STAT_FIELDS, 0)
    idx = res.index_buffer.astype(np.int64)
    s["totalFullyTransparent"] = int((idx == -1).sum())
    s["totalFullyOpaque"] = int((idx == -2).sum())
    s["totalFullyUnknownTransparent"] = int((idx == -3).sum())
    s["totalFullyUnknownOpaque"] = int((idx == -4).sum())
    refs = np.bincount(idx[idx >= 0], minlength=res.desc_array.size)
    for di, nref in enumerate(refs):
        if nref == 0:
            continue
        d = res.desc_array[di]
        n = 1 << (2 * int(d["subdivisionLevel"]))
        data = res.array_data[int(d["offset"]):]
        if int(d["format"]) == capi.FORMAT_2_STATE:
            bits = np.unpackbits(data[: max(1, n // 8)], bitorder="little")[:n]
            cnt = [int((bits == 0).sum()), int((bits == 1).sum()), 0, 0]
        else:
            nb = max(1, n // 4)
            b = data[:nb].astype(np.uint8)
            st = np.stack([(b >> (2 * k)) & 3 for k in range(4)], axis=1).reshape(-1)[:n]
            cnt = [int((st == v).sum()) for v in range(4)]
        s["totalTransparent"] += int(nref) * cnt[0]
        s["totalOpaque"] += int(nref) * cnt[1]
        s["totalUnknownTransparent"] += int(nref) * cnt[2]
        s["totalUnknownOpaque"] += int(nref) * cnt[3]
    return s
