"""Pins the oracle: the plain-C port (oracle/omm_oracle.c) must be byte-identical with the unmodified SDK build
(oracle/_ref/libomm-lib.so) on every parity workload -- the five result arrays the SDK's own serialize round-trip test
compares (ref: support/tests/test_omm_bake_cpu.cpp:323-344).  Skipped where the SDK build is absent."""
import pytest

import parity_cases as PC

CASES = PC.cases()


@pytest.mark.parametrize("name", sorted(CASES))
def test_port_matches_sdk(name, ref_lib, port_lib):
    mk, ov = CASES[name]
    wl = mk()
    a = PC.run_bake(ref_lib, wl, **ov)
    b = PC.run_bake(port_lib, wl, **ov)
    assert a.diff(b) == [], f"{name}: port differs from the SDK"


@pytest.mark.parametrize("name", ["uv16_unorm", "uv16_float", "uv32_stride20"])
def test_port_matches_sdk_uv_formats(name, ref_lib, port_lib):
    wl = PC.uv_format_cases()[name]
    a = PC.run_bake(ref_lib, wl)
    b = PC.run_bake(port_lib, wl)
    assert a.diff(b) == []
