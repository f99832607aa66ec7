"""GPU parity proper: libomm-b200.so (CUDA, through the C ABI) against the strongest CPU checker available (the SDK
build when it travelled with the repo, else the port) on the same host buffers.  Bit-exact: arrayData, descArray,
descArrayHistogram, indexBuffer (+format), indexHistogram."""
import pytest

import parity_cases as PC

CASES = PC.cases()
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(CASES))
def test_product_matches_checker(name, product_lib, checker_lib):
    mk, ov = CASES[name]
    wl = mk()
    want = PC.run_bake(checker_lib, wl, **ov)
    got = PC.run_bake(product_lib, wl, **ov)
    assert got.diff(want) == [], f"{name}: CUDA result differs from {checker_lib.path}"


@pytest.mark.parametrize("name", ["uv16_unorm", "uv16_float", "uv32_stride20"])
def test_product_matches_checker_uv_formats(name, product_lib, checker_lib):
    wl = PC.uv_format_cases()[name]
    want = PC.run_bake(checker_lib, wl)
    got = PC.run_bake(product_lib, wl)
    assert got.diff(want) == []


def test_product_matches_port_too(product_lib, port_lib):
    """The port travels everywhere; keep one direct comparison against it."""
    from omm_b200 import workloads as W
    wl = W.config3(num_tris=400, tex_size=256, level=5)
    assert PC.run_bake(product_lib, wl).diff(PC.run_bake(port_lib, wl)) == []


SDK_ONLY = PC.sdk_only_cases()


@pytest.mark.parametrize("name", sorted(SDK_ONLY))
def test_optional_passes_match_sdk(name, product_lib, ref_lib):
    """Near-duplicate merge (LSH / brute force) and the Compress budget pass against the SDK build itself."""
    mk, ov = SDK_ONLY[name]
    wl = mk()
    want = PC.run_bake(ref_lib, wl, **ov)
    got = PC.run_bake(product_lib, wl, **ov)
    assert got.diff(want) == [], name


def test_streamed_download_matches_checker(product_lib, checker_lib):
    """OMM_B200_STREAMED_DOWNLOAD=1: classifier chunks are merged, packed and sent to the host while later chunks are classified (possible because
    the work items are in output order).  Both outcomes are covered: a bake whose blocks do not repeat across chunks (streamed to the end) and
    one where a later chunk holds the SDK's survivor of an already-sent block (detected, merged again the ordinary way)."""
    import os
    from omm_b200 import capi
    from omm_b200 import workloads as W
    os.environ["OMM_B200_STREAMED_DOWNLOAD"] = "1"
    os.environ["OMM_B200_STREAM_CHUNK_REGIONS"] = "8192"   # many chunks at test size
    try:
        for wl in (W.config3(num_tris=8000, tex_size=1024, level=6), W.config3(num_tris=4096, tex_size=512, level=6, cells=(64, 32)),
                   W.config5(num_tris=9000, tex_size=512, distinct=2500, flat_tris=1500, max_level=7),
                   # both unknown states in play (Nearest promotion): a fully-unknown special item and an item mixing them share a digest
                   W.config3(num_tris=6000, tex_size=256, level=4, promotion=capi.PROMOTE_NEAREST)):
            want = PC.run_bake(checker_lib, wl)
            got = PC.run_bake(product_lib, wl)
            assert got.diff(want) == [], wl.name
    finally:
        del os.environ["OMM_B200_STREAMED_DOWNLOAD"]
        del os.environ["OMM_B200_STREAM_CHUNK_REGIONS"]


def test_chunk_lanes_match_checker(product_lib, checker_lib):
    """Classifier chunks in flight side by side on several streams (OMM_B200_CHUNK_LANES; two by default, but only bakes of more than one
    32 M-region chunk have anything to put side by side): forced here with small chunks on mid-size bakes, for 1 to 4 lanes, incl. a bake
    whose big blocks keep the per-item post pass behind the classification, mixed levels and both unknown states."""
    import os
    from omm_b200 import capi
    from omm_b200 import workloads as W
    wls = (W.config3(num_tris=8000, tex_size=1024, level=6), W.config3(num_tris=40, tex_size=512, level=9),
           W.config5(num_tris=9000, tex_size=512, distinct=2500, flat_tris=1500, max_level=8),
           W.config3(num_tris=20000, tex_size=512, level=5, promotion=capi.PROMOTE_NEAREST))
    want = [PC.run_bake(checker_lib, wl) for wl in wls]
    os.environ["OMM_B200_CHUNK_REGIONS"] = "65536"
    try:
        for lanes in (1, 2, 3, 4):
            os.environ["OMM_B200_CHUNK_LANES"] = str(lanes)
            for wl, w in zip(wls, want):
                got = PC.run_bake(product_lib, wl)
                assert got.diff(w) == [], (wl.name, lanes)
                if lanes > 1 and wl is wls[0]:
                    assert got.timings.kernelLaunches > 60, "the bake was expected to run in several chunks"
    finally:
        del os.environ["OMM_B200_CHUNK_REGIONS"]
        os.environ.pop("OMM_B200_CHUNK_LANES", None)
