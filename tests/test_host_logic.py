"""Host-side behaviour of the C-ABI shim that needs no GPU: argument checks, return codes and message texts of the
entry points (ref: libraries/omm-lib/src/bake.cpp:44-135, 410-479; bake_cpu_impl.cpp:97-103, 235-257), compared with the
SDK build where it is available."""
import ctypes as C
import os

import pytest

from omm_b200 import Baker, capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    return capi.OmmLib(capi.PRODUCT_LIB)


def _libs(lib, request):
    out = [lib]
    ref = os.path.join(ROOT, "oracle", "_ref", "libomm-lib.so")
    if os.path.exists(ref):
        out.append(capi.OmmLib(ref))
    return out


def test_create_baker_argument_checks(lib, request):
    for l in _libs(lib, request):
        h = C.c_void_p()
        assert l.dll.ommCreateBaker(None, C.byref(h)) == capi.INVALID_ARGUMENT
        d = capi.BakerCreationDesc()
        d.type = capi.BAKER_MAX
        assert l.dll.ommCreateBaker(C.byref(d), C.byref(h)) == capi.INVALID_ARGUMENT
        assert l.dll.ommDestroyBaker(None) == capi.INVALID_ARGUMENT


def test_gpu_baker_type_is_not_offered(lib):
    d = capi.BakerCreationDesc()
    d.type = capi.BAKER_GPU
    h = C.c_void_p()
    assert lib.dll.ommCreateBaker(C.byref(d), C.byref(h)) == capi.NOT_IMPLEMENTED


def _bake_messages(l, desc_mutator):
    msgs = []
    with Baker(l, on_message=lambda sev, m: msgs.append((sev, m))) as b:
        d = capi.bake_input_desc_default()
        desc_mutator(d)
        h = C.c_void_p()
        rc = l.dll.ommCpuBake(b.handle, C.byref(d), C.byref(h))
        assert not h.value, "outBakeResult must stay untouched on failure (ref: test_omm_bake_cpu.cpp:783-789)"
    return rc, msgs


def test_null_desc_and_null_texture_messages(lib, request):
    results = []
    for l in _libs(lib, request):
        with Baker(l, on_message=lambda sev, m: None) as b:
            h = C.c_void_p()
            assert l.dll.ommCpuBake(None, None, C.byref(h)) == capi.INVALID_ARGUMENT
        msgs = []
        with Baker(l, on_message=lambda sev, m: msgs.append((sev, m))) as b:
            h = C.c_void_p()
            assert l.dll.ommCpuBake(b.handle, None, C.byref(h)) == capi.INVALID_ARGUMENT
        rc, m2 = _bake_messages(l, lambda d: None)   # default desc: no texture
        assert rc == capi.INVALID_ARGUMENT
        results.append((msgs, m2))
    assert results[0][0] == [(capi.SEVERITY_FATAL, "input desc was not set")]
    assert results[0][1] == [(capi.SEVERITY_FATAL, "[Invalid Argument] - ommCpuBakeInputDesc has no texture set")]
    for r in results[1:]:
        assert r == results[0], "messages differ from the SDK build"


def test_texture_argument_checks(lib, request):
    for l in _libs(lib, request):
        msgs = []
        with Baker(l, on_message=lambda sev, m: msgs.append(m)) as b:
            h = C.c_void_p()
            assert l.dll.ommCpuCreateTexture(None, None, C.byref(h)) == capi.INVALID_ARGUMENT
            assert l.dll.ommCpuCreateTexture(b.handle, None, C.byref(h)) == capi.INVALID_ARGUMENT
            td = capi.CpuTextureDesc()
            td.format, td.mipCount = capi.TEX_FP32, 0
            assert l.dll.ommCpuCreateTexture(b.handle, C.byref(td), C.byref(h)) == capi.INVALID_ARGUMENT
            td.format, td.mipCount = capi.TEX_MAX, 1
            mip = (capi.CpuTextureMipDesc * 1)()
            td.mips = mip
            assert l.dll.ommCpuCreateTexture(b.handle, C.byref(td), C.byref(h)) == capi.INVALID_ARGUMENT
            td.format = capi.TEX_FP32
            assert l.dll.ommCpuCreateTexture(b.handle, C.byref(td), C.byref(h)) == capi.INVALID_ARGUMENT   # no data
            buf = (C.c_float * 4)()
            mip[0].textureData = C.cast(buf, C.c_void_p)
            mip[0].width, mip[0].height = 0, 2
            assert l.dll.ommCpuCreateTexture(b.handle, C.byref(td), C.byref(h)) == capi.INVALID_ARGUMENT
            mip[0].width, mip[0].height = 65537, 2
            assert l.dll.ommCpuCreateTexture(b.handle, C.byref(td), C.byref(h)) == capi.INVALID_ARGUMENT
            assert l.dll.ommCpuDestroyTexture(b.handle, None) == capi.INVALID_ARGUMENT
            assert l.dll.ommCpuGetTextureDesc(None, None) == capi.INVALID_ARGUMENT
        assert msgs == ["texture desc was not set", "[Invalid Arg] - mipCount must be non-zero", "[Invalid Arg] - format is not set",
                        "[Invalid Arg] - mips.textureData is not set", "[Invalid Arg] - mips.width must be non-zero",
                        "[Invalid Arg] - mips.width must be less than kMaxDim.x (65536)"], l.path


def test_result_entry_points_reject_null(lib, request):
    for l in _libs(lib, request):
        assert l.dll.ommCpuDestroyBakeResult(None) == capi.INVALID_ARGUMENT
        p = C.POINTER(capi.CpuBakeResultDesc)()
        assert l.dll.ommCpuGetBakeResultDesc(None, C.byref(p)) == capi.INVALID_ARGUMENT


def test_product_fails_loudly_without_a_gpu(lib):
    """No CPU fallback: with no CUDA device, creating a texture must fail with a message (never silently succeed)."""
    if lib.dll.ommB200GetDeviceCount() > 0:
        pytest.skip("a CUDA device is visible")
    import numpy as np
    msgs = []
    with Baker(lib, on_message=lambda sev, m: msgs.append(m)) as b:
        with pytest.raises(Exception):
            b.create_texture([np.zeros((4, 4), dtype=np.float32)])
    assert any("no CUDA device" in m for m in msgs)


def test_b200_extension_argument_checks(lib):
    """Host-only behaviour of the round-2 extension entry points (include/omm_b200.h): result mode of a sharded baker, host pool trim."""
    with Baker(lib) as b:
        assert lib.dll.ommB200SetShardedResultMode(b.handle, capi.SHARDED_RESULT_ON_RANK0) == capi.SUCCESS
        assert lib.dll.ommB200SetShardedResultMode(b.handle, capi.SHARDED_RESULT_REPLICATED) == capi.SUCCESS
        assert lib.dll.ommB200SetShardedResultMode(b.handle, 7) == capi.INVALID_ARGUMENT
    assert lib.dll.ommB200SetShardedResultMode(None, capi.SHARDED_RESULT_ON_RANK0) == capi.INVALID_ARGUMENT
    assert lib.dll.ommB200TrimHostPool(0) == 0   # nothing is cached without a bake
