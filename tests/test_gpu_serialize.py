"""ommCpuSerialize / ommCpuDeserialize against the SDK build: byte-identical blobs where the SDK's blob is
deterministic, cross-library round trips (also LZ4-compressed SDK blobs) everywhere else, and bakes of deserialized inputs."""
import ctypes as C
import os

import numpy as np
import pytest

from omm_b200 import Baker, capi, load_product_library
from omm_b200 import workloads as W

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "libomm-lib.so")


@pytest.fixture(scope="module")
def libs():
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/libomm-lib.so not built")
    return load_product_library(), capi.OmmLib(REF)


def _blob_and_result(lib, wl, flags=capi.SERIALIZE_NONE, with_result=True):
    with Baker(lib) as b:
        inp, tex = W.make_input(b, wl)
        desc = inp.to_desc()
        res = b.bake(inp)
        results = []
        keep = None
        if with_result:
            # serialize the bake result too: rebuild a CpuBakeResultDesc over the copied arrays
            r = capi.CpuBakeResultDesc()
            keep = (np.ascontiguousarray(res.array_data), np.ascontiguousarray(res.desc_array), np.ascontiguousarray(res.desc_histogram),
                    np.ascontiguousarray(res.index_buffer), np.ascontiguousarray(res.index_histogram))
            r.arrayData, r.arrayDataSize = keep[0].ctypes.data, keep[0].size
            r.descArray, r.descArrayCount = C.cast(keep[1].ctypes.data, C.POINTER(capi.CpuOpacityMicromapDesc)), keep[1].size
            r.descArrayHistogram, r.descArrayHistogramCount = C.cast(keep[2].ctypes.data, C.POINTER(capi.CpuOpacityMicromapUsageCount)), keep[2].size
            r.indexBuffer, r.indexCount, r.indexFormat = keep[3].ctypes.data, keep[3].size, res.index_format
            r.indexHistogram, r.indexHistogramCount = C.cast(keep[4].ctypes.data, C.POINTER(capi.CpuOpacityMicromapUsageCount)), keep[4].size
            results = [r]
        blob = b.serialize([desc], results, flags)
        tex.destroy()
    return blob, res


def _bake_blob(lib, blob):
    with Baker(lib) as b:
        rc, h, pd = b.deserialize_raw(blob)
        assert rc == capi.SUCCESS, rc
        try:
            d = pd.contents
            assert d.numInputDescs == 1
            res = b.bake_desc(d.inputDescs[0])
            stored = None
            if d.numResultDescs:
                from omm_b200.baker import _copy_result
                stored = _copy_result(d.resultDescs[0])
            return res, stored
        finally:
            lib.dll.ommCpuDestroyDeserializedResult(h)


CASES = {
    # square power-of-two FP32 texture: no padding anywhere in the SDK's texture dump -> its blob is deterministic
    "pow2_fp32": lambda: W.config3(num_tris=200, tex_size=256, level=4),
    "pow2_fp32_sat": lambda: W.config3(num_tris=200, tex_size=128, level=4, tex_alpha_cutoff=0.5),
    "c1_checker": lambda: W.config1(),
    "npot_unorm8_mips": lambda: W.random_mesh(201, 150, tex_size=(100, 75), unorm8=True, mips=3, addressing_mode=capi.ADDR_MIRROR),
    "linear_tiling_levels": lambda: W.random_mesh(202, 150, tex_flags=capi.TEXFLAG_DISABLE_ZORDER, subdivision_levels=(np.arange(150) % 5).astype(np.uint8),
                                                  index_dtype=np.uint16),
}


@pytest.mark.parametrize("name", ["pow2_fp32", "pow2_fp32_sat", "c1_checker"])
def test_blob_is_byte_identical_with_the_sdk(libs, name):
    ours, sdk = libs
    blob_o, _ = _blob_and_result(ours, CASES[name]())
    blob_s, _ = _blob_and_result(sdk, CASES[name]())
    assert blob_o == blob_s


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("compress", [False, True])
def test_cross_library_round_trip(libs, name, compress):
    ours, sdk = libs
    wl = CASES[name]()
    flags = capi.SERIALIZE_COMPRESS if compress else capi.SERIALIZE_NONE
    blob_s, direct_s = _blob_and_result(sdk, wl, flags)
    blob_o, direct_o = _blob_and_result(ours, wl, flags)
    assert not direct_o.diff(direct_s)
    # SDK blob (LZ4 block when compressed) -> this library
    res, stored = _bake_blob(ours, blob_s)
    assert not res.diff(direct_s)
    assert stored is not None and not stored.diff(direct_s)
    # this library's blob -> SDK
    res, stored = _bake_blob(sdk, blob_o)
    assert not res.diff(direct_s)
    assert stored is not None and not stored.diff(direct_s)
    # and back into itself
    res, _ = _bake_blob(ours, blob_o)
    assert not res.diff(direct_s)


def test_corrupted_blob_is_rejected(libs):
    ours, _ = libs
    blob, _ = _blob_and_result(ours, CASES["c1_checker"](), with_result=False)
    bad = bytearray(blob)
    bad[len(bad) // 2] ^= 0x40
    msgs = []
    with Baker(ours, on_message=lambda sev, m: msgs.append(m)) as b:
        rc, h, _ = b.deserialize_raw(bytes(bad))
        assert rc == capi.INVALID_ARGUMENT and h is None
        assert any("corrupted" in m for m in msgs)
        rc, h, _ = b.deserialize_raw(blob[:20])
        assert rc != capi.SUCCESS


def _golden_cases():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "serialized_inputs.json")) as f:
        return json.load(f)["cases"]


@pytest.mark.parametrize("case", _golden_cases(), ids=lambda c: c["name"])
def test_golden_blobs_of_older_sdk_versions(case):
    """The SDK's own DeserializeInput_* tests: blobs written by SDK 1.4 - 1.7 (older header / texture layouts, one LZ4-compressed)
    must deserialize and bake to the state totals those tests expect (committed fixture, tests/golden/make_golden.py)."""
    import kat_cases as K
    lib = load_product_library()
    res, _ = _bake_blob(lib, bytes.fromhex(case["blob_hex"]))
    got = K.collect_stats(res)
    for k, v in case["expect"].items():
        assert got[k] == v, (case["name"], case["line"], k, got)
