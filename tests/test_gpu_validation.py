"""Validation behaviour that needs a texture (hence a GPU): return codes and message texts of ommCpuBake for invalid
descs, compared verbatim with the SDK build when it travelled (ref: bake_cpu_impl.cpp:235-290, 652-657, 682-713;
support/tests/test_omm_log.cpp:146-208)."""
import ctypes as C
import os

import numpy as np
import pytest

from omm_b200 import Baker, capi
from omm_b200 import workloads as W

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _attempt(lib, mutate, tex_cutoff=-1.0, flags=0):
    msgs = []
    wl = W.random_mesh(5, 50, tex_size=(64, 64), tex_alpha_cutoff=tex_cutoff)
    with Baker(lib, on_message=lambda sev, m: msgs.append((sev, m))) as b:
        inp, tex = W.make_input(b, wl, bake_flags=flags)
        d = inp.to_desc()
        mutate(d)
        h = C.c_void_p()
        rc = lib.dll.ommCpuBake(b.handle, C.byref(d), C.byref(h))
        if rc == capi.SUCCESS:
            lib.dll.ommCpuDestroyBakeResult(h)
        else:
            assert not h.value
        tex.destroy()
    return rc, msgs


def _set(field, value):
    def f(d):
        setattr(d, field, value)
    return f


def _sampler(addr=None, filt=None):
    def f(d):
        if addr is not None:
            d.runtimeSamplerDesc.addressingMode = addr
        if filt is not None:
            d.runtimeSamplerDesc.filter = filt
    return f


CASES = {
    "alphaMode": (_set("alphaMode", capi.ALPHA_MAX), "[Invalid Argument] - alphaMode is not set"),
    "texCoordFormat": (_set("texCoordFormat", capi.UV_MAX), "[Invalid Argument] - texCoordFormat is not set"),
    "texCoords": (_set("texCoords", None), "[Invalid Argument] - texCoords is not set"),
    "indexFormat": (_set("indexFormat", capi.INDEX_MAX), "[Invalid Argument] - indexFormat is not set"),
    "indexBuffer": (_set("indexBuffer", None), "[Invalid Argument] - indexBuffer is not set"),
    "indexCount": (_set("indexCount", 0), "[Invalid Argument] - indexCount is not set"),
    "maxLevel": (_set("maxSubdivisionLevel", 13), "[Invalid Argument] - maxSubdivisionLevel (13) is greater than maximum supported (12)"),
    "gt_state_2state": (lambda d: (setattr(d, "format", capi.FORMAT_2_STATE), setattr(d, "alphaCutoffGreater", capi.STATE_UO)),
                        "[Invalid Argument] - alphaCutoffGreater=UnknownOpaque is not compatible with OC1_2_State"),
    "le_state_2state": (lambda d: (setattr(d, "format", capi.FORMAT_2_STATE), setattr(d, "alphaCutoffLessEqual", capi.STATE_UT)),
                        "[Invalid Argument] - alphaCutoffLessEqual=UnknownTransparent is not compatible with OC1_2_State"),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_invalid_desc_messages(name, product_lib, checker_lib):
    mutate, text = CASES[name]
    rc, msgs = _attempt(product_lib, mutate)
    assert rc == capi.INVALID_ARGUMENT
    assert msgs == [(capi.SEVERITY_FATAL, text)]
    if "libomm-lib" in checker_lib.path:
        assert _attempt(checker_lib, mutate) == (rc, msgs)


def test_unset_sampler_fields_fail_silently_like_the_sdk(product_lib, checker_lib):
    """The SDK dispatches on (addressing mode, filter) before validating (ref: bake_cpu_impl.cpp:297-303): FAILURE, no message."""
    for mutate in (_sampler(addr=capi.ADDR_MAX), _sampler(filt=capi.FILTER_MAX)):
        assert _attempt(product_lib, mutate) == (capi.FAILURE, [])
        if "libomm-lib" in checker_lib.path:
            assert _attempt(checker_lib, mutate) == (capi.FAILURE, [])


def test_texture_cutoff_mismatch_message(product_lib, checker_lib):
    rc, msgs = _attempt(product_lib, _set("alphaCutoff", 0.25), tex_cutoff=0.5)
    assert rc == capi.INVALID_ARGUMENT
    assert msgs == [(capi.SEVERITY_FATAL, "[Invalid Argument] - Texture object alpha cutoff threshold (0.500000) is different from alpha cutoff threshold in bake input (0.250000)")]
    if "libomm-lib" in checker_lib.path:
        assert _attempt(checker_lib, _set("alphaCutoff", 0.25), tex_cutoff=0.5) == (rc, msgs)


def test_workload_limit_and_validation_messages(product_lib, checker_lib):
    # ref: test_omm_bake_cpu.cpp:2022-2032 (WORKLOAD_TOO_BIG) and bake_cpu_impl.cpp:652-657 / 700-709 (info + perf warning)
    rc, msgs = _attempt(product_lib, _set("maxWorkloadSize", 10))
    assert rc == capi.WORKLOAD_TOO_BIG and msgs == []
    wl = W.random_mesh(6, 200, tex_size=(64, 64), nan_frac=0.2, uv_lo=-40.0, uv_hi=40.0, tri_texels=64 * 300.0)
    out = []
    for lib in [product_lib] + ([checker_lib] if "libomm-lib" in checker_lib.path else []):
        msgs = []
        with Baker(lib, on_message=lambda sev, m: msgs.append((sev, m))) as b:
            inp, tex = W.make_input(b, wl, bake_flags=capi.BAKE_ENABLE_VALIDATION | capi.BAKE_INT_DISABLE_FINE, max_subdivision_level=1)
            res = b.bake(inp)
            tex.destroy()
        out.append((msgs, res))
    msgs = out[0][0]
    assert any(sev == capi.SEVERITY_INFO and "unclassifiable triangles" in m for sev, m in msgs)
    assert any(sev == capi.SEVERITY_PERF_WARNING and "This is unusually large" in m for sev, m in msgs)
    if len(out) == 2:
        assert out[0][0] == out[1][0], "validation messages differ from the SDK build"
        assert out[0][1].diff(out[1][1]) == []


def test_resident_bake_and_device_result(product_lib):
    """ommB200StageInputs / BakeResident / GetDeviceResultDesc / DownloadResult give the same bytes as ommCpuBake."""
    wl = W.config3(num_tris=500, tex_size=256, level=5)
    with Baker(product_lib) as b:
        inp, tex = W.make_input(b, wl)
        want = b.bake(inp)
        d = inp.to_desc()
        staged = C.c_void_p()
        assert product_lib.dll.ommB200StageInputs(b.handle, C.byref(d), C.byref(staged)) == capi.SUCCESS
        h = C.c_void_p()
        assert product_lib.dll.ommB200BakeResident(b.handle, staged, None, C.byref(h)) == capi.SUCCESS
        dev = capi.B200DeviceResultDesc()
        assert product_lib.dll.ommB200GetDeviceResultDesc(h, C.byref(dev)) == capi.SUCCESS
        assert dev.arrayDataSize == want.array_data.size and dev.descArrayCount == want.desc_array.size and dev.arrayData
        assert product_lib.dll.ommB200DownloadResult(h) == capi.SUCCESS
        p = C.POINTER(capi.CpuBakeResultDesc)()
        assert product_lib.dll.ommCpuGetBakeResultDesc(h, C.byref(p)) == capi.SUCCESS
        from omm_b200.baker import _copy_result
        got = _copy_result(p.contents)
        product_lib.dll.ommCpuDestroyBakeResult(h)
        product_lib.dll.ommB200DestroyStagedInputs(staged)
        tex.destroy()
    assert got.diff(want) == []


def test_sliced_download_equals_plain_download(product_lib):
    """ommCpuBake sends arrayData over PCIe slice by slice while it is packed (results >= 8 MiB into the page-locked pool); the bytes
    must equal those of the resident bake followed by ommB200DownloadResult, which packs in one launch and copies afterwards.  With
    a user allocator (pageable destination) the sliced path is not taken: same bytes again."""
    from omm_b200.baker import _copy_result
    wl = W.config3(num_tris=60000)
    with Baker(product_lib) as b:
        inp, tex = W.make_input(b, wl)
        sliced = b.bake(inp)
        assert sliced.array_data.size >= 8 << 20 and sliced.desc_array.size >= 512, "workload too small to take the sliced path"
        d = inp.to_desc()
        staged = C.c_void_p()
        assert product_lib.dll.ommB200StageInputs(b.handle, C.byref(d), C.byref(staged)) == capi.SUCCESS
        h = C.c_void_p()
        assert product_lib.dll.ommB200BakeResident(b.handle, staged, None, C.byref(h)) == capi.SUCCESS
        assert product_lib.dll.ommB200DownloadResult(h) == capi.SUCCESS
        p = C.POINTER(capi.CpuBakeResultDesc)()
        assert product_lib.dll.ommCpuGetBakeResultDesc(h, C.byref(p)) == capi.SUCCESS
        plain = _copy_result(p.contents)
        product_lib.dll.ommCpuDestroyBakeResult(h)
        product_lib.dll.ommB200DestroyStagedInputs(staged)
        again = b.bake(inp)  # the pool block of the first result is reused
        tex.destroy()
    assert sliced.diff(plain) == []
    assert again.diff(plain) == []


def test_user_allocator_is_honoured(product_lib):
    """All host memory visible through the ABI comes from ommMemoryAllocatorInterface (ref: omm.h:214-232, bake.cpp:123-124)."""
    live, total = {}, [0]
    libc = C.CDLL(None)
    libc.aligned_alloc.restype, libc.aligned_alloc.argtypes = C.c_void_p, [C.c_size_t, C.c_size_t]
    libc.free.argtypes = [C.c_void_p]

    def alloc(user, size, align):
        align = max(int(align), 16)
        p = libc.aligned_alloc(align, (int(size) + align - 1) // align * align)
        live[p] = size
        total[0] += 1
        return p

    def free(user, p):
        if p:
            assert p in live, "free of a pointer the allocator never returned"
            del live[p]
            libc.free(p)

    a_cb, f_cb = capi.ALLOCATE_FN(alloc), capi.FREE_FN(free)
    r_cb = capi.REALLOCATE_FN(lambda user, p, size, align: None)
    desc = capi.BakerCreationDesc()
    desc.type = capi.BAKER_CPU
    desc.memoryAllocatorInterface.allocate, desc.memoryAllocatorInterface.reallocate, desc.memoryAllocatorInterface.free = a_cb, r_cb, f_cb
    hb = C.c_void_p()
    assert product_lib.dll.ommCreateBaker(C.byref(desc), C.byref(hb)) == capi.SUCCESS
    b = Baker.__new__(Baker)
    b.lib, b.handle, b._cb, b.messages = product_lib, hb.value, None, []
    wl = W.config3(num_tris=300, tex_size=256, level=5)
    inp, tex = W.make_input(b, wl)
    res = b.bake(inp)
    assert res.array_data.size > 0 and total[0] >= 4
    tex.destroy()
    b.destroy()
    assert live == {}, f"leaked {len(live)} host allocations"


def test_formats_the_sdk_has_undefined_behaviour_on_are_refused(product_lib):
    """desc.format outside {OC1_2_State, OC1_4_State}: the SDK only asserts (bake_cpu_impl.cpp:328-333, compiled out in its release build) and then
    indexes its histograms with format - 1.  Per-triangle formats that differ from desc.format: its arrays are sized from desc.format alone
    (:1763-1771).  Both are refused before any device work (round-1 advisor finding: the second used to fail only after the whole classification)."""
    rc, msgs = _attempt(product_lib, _set("format", 0))
    assert rc == capi.INVALID_ARGUMENT and any("format is not set" in m for _, m in msgs), (rc, msgs)
    rc, msgs = _attempt(product_lib, _set("format", 3))
    assert rc == capi.INVALID_ARGUMENT

    keep = []

    def mixed(d):
        n = d.indexCount // 3
        f = np.full(n, capi.FORMAT_4_STATE, dtype=np.int32)
        f[n // 2] = capi.FORMAT_2_STATE
        keep.append(f)
        d.formats = f.ctypes.data
    rc, msgs = _attempt(product_lib, mixed)
    assert rc == capi.FAILURE and any("per-triangle formats" in m for _, m in msgs), (rc, msgs)
