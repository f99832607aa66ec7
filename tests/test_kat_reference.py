"""Replays the SDK's own known-answer counts (support/tests/test_omm_bake_cpu.cpp) against the oracle port on CPU, and
-- in the gpu-marked half -- against libomm-b200.so.  This pins the oracle independently of the SDK build."""
import ctypes as C

import pytest

import kat_cases as K
import parity_cases as PC
from omm_b200 import capi

KATS = K.kats()


def _expected(d):
    e = dict.fromkeys(K.STAT_FIELDS, 0)
    e.update(d)
    return e


@pytest.mark.parametrize("name", sorted(KATS))
def test_port_reproduces_sdk_known_answers(name, port_lib):
    mk, want, line = KATS[name]
    res = PC.run_bake(port_lib, mk())
    assert K.collect_stats(res) == _expected(want), f"{name} (test_omm_bake_cpu.cpp:{line})"


@pytest.mark.parametrize("name", ["Circle", "Sine", "Uniform"])
def test_sdk_build_reproduces_its_known_answers(name, ref_lib):
    mk, want, line = KATS[name]
    assert K.collect_stats(PC.run_bake(ref_lib, mk())) == _expected(want)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(KATS))
def test_product_reproduces_sdk_known_answers(name, product_lib):
    mk, want, line = KATS[name]
    res = PC.run_bake(product_lib, mk())
    assert K.collect_stats(res) == _expected(want), f"{name} (test_omm_bake_cpu.cpp:{line})"


@pytest.mark.gpu
def test_product_circle_merge_similar_kat(product_lib):
    """ref: test_omm_bake_cpu.cpp:973-986 (Circle with EnableNearDuplicateDetection)."""
    mk, _, _ = KATS["Circle"]
    wl = mk()
    wl.desc["bake_flags"] = capi.BAKE_ENABLE_INTERNAL_THREADS | capi.BAKE_ENABLE_NEAR_DUPLICATE_DETECTION
    res = PC.run_bake(product_lib, wl)
    assert K.collect_stats(res) == _expected(dict(totalOpaque=200, totalTransparent=216, totalUnknownTransparent=42, totalUnknownOpaque=54))


@pytest.mark.gpu
def test_product_debug_stats_entry_point(product_lib):
    """ommDebugGetStats of the product agrees with the Python restatement."""
    from omm_b200 import Baker
    from omm_b200 import workloads as W
    mk, want, _ = KATS["Circle"]
    wl = mk()
    with Baker(product_lib) as b:
        inp, tex = W.make_input(b, wl)
        desc = inp.to_desc()
        rc, h = b.bake_raw(desc)
        assert rc == capi.SUCCESS
        pdesc = C.POINTER(capi.CpuBakeResultDesc)()
        assert product_lib.dll.ommCpuGetBakeResultDesc(h, C.byref(pdesc)) == capi.SUCCESS
        st = capi.DebugStats()
        assert product_lib.dll.ommDebugGetStats(b.handle, pdesc, C.byref(st)) == capi.SUCCESS
        got = {f: int(getattr(st, f)) for f in K.STAT_FIELDS}
        product_lib.dll.ommCpuDestroyBakeResult(h)
        tex.destroy()
    assert got == _expected(want)
