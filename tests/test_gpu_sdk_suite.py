"""The SDK's OWN gtest suite (support/tests of the reference, CPU part: 700+ tests -- known-answer bakes through ommDebugGetStats,
serialization round trips of every bake, log texts, golden blobs) linked against libomm-b200.so.

oracle/Makefile target `reftests` compiles the unmodified test sources where they lie under /root/reference (build container only)
into oracle/_ref/tests_b200; the binary travels to the GPU box like the other built artefacts.  Excluded: the tests of the
D3D12 / Vulkan command-list baker, which is out of scope (GpuTest.*, Baker.CreateDestroyGPU, Baker.StaticDataGPU)."""
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "tests_b200")
OUT_OF_SCOPE = "GpuTest.*:Baker.CreateDestroyGPU:Baker.StaticDataGPU"


def test_sdk_gtest_suite_passes_against_this_library():
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/tests_b200 not built (make -C oracle reftests, needs /root/reference)")
    p = subprocess.run([BIN, f"--gtest_filter=-{OUT_OF_SCOPE}", "--gtest_brief=1"], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                       timeout=1500)
    tail = "\n".join(p.stdout.splitlines()[-40:])
    m = re.search(r"\[\s+PASSED\s+\]\s+(\d+) tests", p.stdout)
    assert p.returncode == 0 and m, tail
    assert int(m.group(1)) >= 700, tail
