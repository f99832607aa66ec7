"""CPU pin of tests/golden/serialized_inputs.json: the UNMODIFIED SDK build deserializes the golden blobs of its own
DeserializeInput_* tests and bakes them to the state totals those tests expect (the GPU counterpart is tests/test_gpu_serialize.py)."""
import json
import os

import pytest

import kat_cases as K
from omm_b200 import Baker, capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "libomm-lib.so")
with open(os.path.join(ROOT, "tests", "golden", "serialized_inputs.json")) as f:
    CASES = json.load(f)["cases"]


@pytest.mark.parametrize("case", CASES, ids=lambda c: c["name"])
def test_sdk_build_bakes_its_golden_blobs(case):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/libomm-lib.so not built (only possible where /root/reference exists)")
    lib = capi.OmmLib(REF)
    with Baker(lib) as b:
        rc, h, pd = b.deserialize_raw(bytes.fromhex(case["blob_hex"]))
        assert rc == capi.SUCCESS
        try:
            assert pd.contents.numInputDescs == 1
            got = K.collect_stats(b.bake_desc(pd.contents.inputDescs[0]))
        finally:
            lib.dll.ommCpuDestroyDeserializedResult(h)
    for k, v in case["expect"].items():
        assert got[k] == v, (case["name"], case["line"], k, got)
