"""The drop-in boundary: libomm-b200.so loads without a GPU, exports every symbol include/omm_b200.h declares, and the
header's struct layouts equal the ctypes mirror (and the SDK's own omm.h where the reference checkout is present)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

from omm_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "omm_b200.h")
SDK_HEADER_DIR = "/root/reference/libraries/omm-lib/include"

STRUCTS = ["ommLibraryDesc", "ommSamplerDesc", "ommMemoryAllocatorInterface", "ommMessageInterface", "ommBakerCreationDesc", "ommCpuTextureMipDesc",
           "ommCpuTextureDesc", "ommCpuBakeInputDesc", "ommCpuOpacityMicromapDesc", "ommCpuOpacityMicromapUsageCount", "ommCpuBakeResultDesc", "ommDebugStats"]
CTYPES = {"ommLibraryDesc": capi.LibraryDesc, "ommSamplerDesc": capi.SamplerDesc, "ommMemoryAllocatorInterface": capi.MemoryAllocatorInterface,
          "ommMessageInterface": capi.MessageInterface, "ommBakerCreationDesc": capi.BakerCreationDesc, "ommCpuTextureMipDesc": capi.CpuTextureMipDesc,
          "ommCpuTextureDesc": capi.CpuTextureDesc, "ommCpuBakeInputDesc": capi.CpuBakeInputDesc, "ommCpuOpacityMicromapDesc": capi.CpuOpacityMicromapDesc,
          "ommCpuOpacityMicromapUsageCount": capi.CpuOpacityMicromapUsageCount, "ommCpuBakeResultDesc": capi.CpuBakeResultDesc, "ommDebugStats": capi.DebugStats}
FIELDS = ["bakeFlags", "texture", "runtimeSamplerDesc", "alphaMode", "texCoordFormat", "texCoords", "texCoordStrideInBytes", "indexFormat", "indexBuffer",
          "indexCount", "dynamicSubdivisionScale", "rejectionThreshold", "alphaCutoff", "nearDuplicateDeduplicationFactor", "alphaCutoffLessEqual",
          "alphaCutoffGreater", "format", "formats", "unknownStatePromotion", "unresolvedTriState", "maxSubdivisionLevel", "maxArrayDataSize",
          "subdivisionLevels", "maxWorkloadSize"]


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"OMM_API\s+[\w\s\*]+?\s+\**(omm\w+)\s*\(", src)))


def test_header_declares_the_bake_path():
    syms = declared_symbols()
    for s in capi.CORE_SYMBOLS + capi.SERIALIZE_SYMBOLS + capi.B200_SYMBOLS:
        assert s in syms, s


def test_library_loads_without_gpu_and_exports_everything():
    if not os.path.exists(capi.PRODUCT_LIB):
        pytest.fail(f"{capi.PRODUCT_LIB} missing -- run __graft_entry__.build()")
    lib = capi.OmmLib(capi.PRODUCT_LIB)
    for s in declared_symbols():
        assert lib.exported(s), f"{s} declared in include/omm_b200.h but not exported"
    d = lib.dll.ommGetLibraryDesc()
    assert (d.versionMajor, d.versionMinor, d.versionBuild) == (1, 9, 0)  # ref: support/tests/test_basic.cpp:19-26


def _layout_program(header_dir, header, cxx=False):
    lines = ["#include <stdio.h>", "#include <stddef.h>", f'#include "{header}"', "int main(void){"]
    for s in STRUCTS:
        lines.append(f'printf("{s} %zu\\n", sizeof({s}));')
    for f in FIELDS:
        lines.append(f'printf("ommCpuBakeInputDesc.{f} %zu\\n", offsetof(ommCpuBakeInputDesc, {f}));')
    for e in ["ommResult_WORKLOAD_TOO_BIG", "ommOpacityState_UnknownOpaque", "ommSpecialIndex_FullyUnknownOpaque", "ommFormat_OC1_4_State",
              "ommIndexFormat_UINT_8", "ommTextureAddressMode_MirrorOnce", "ommTextureFilterMode_Linear", "ommCpuTextureFormat_FP32",
              "ommCpuBakeFlags_Allow8BitIndices", "ommUnknownStatePromotion_ForceTransparent", "ommTexCoordFormat_UV32_FLOAT", "ommBakerType_CPU"]:
        lines.append(f'printf("{e} %d\\n", (int){e});')
    lines += ["return 0;}"]
    with tempfile.TemporaryDirectory() as td:
        src, exe = os.path.join(td, "l.cpp" if cxx else "l.c"), os.path.join(td, "l")
        open(src, "w").write("\n".join(lines))
        cc = ["/usr/bin/g++", "-std=gnu++17", "-Wno-invalid-offsetof"] if cxx else ["/usr/bin/gcc", "-std=c11"]
        subprocess.check_call(cc + ["-I", header_dir, "-o", exe, src])
        out = subprocess.check_output([exe]).decode()
    return dict(ln.rsplit(" ", 1) for ln in out.strip().splitlines())


def test_header_layout_matches_ctypes_mirror():
    lay = _layout_program(os.path.join(ROOT, "include"), "omm_b200.h")
    for s in STRUCTS:
        assert int(lay[s]) == C.sizeof(CTYPES[s]), s
    assert int(lay["ommCpuBakeInputDesc"]) == 136
    for f in FIELDS:
        assert int(lay[f"ommCpuBakeInputDesc.{f}"]) == getattr(capi.CpuBakeInputDesc, f).offset, f


@pytest.mark.skipif(not os.path.exists(os.path.join(SDK_HEADER_DIR, "omm.h")), reason="reference checkout not present")
def test_header_layout_matches_the_sdk_header():
    ours = _layout_program(os.path.join(ROOT, "include"), "omm_b200.h", cxx=True)   # the SDK header is only valid C++
    sdk = _layout_program(SDK_HEADER_DIR, "omm.h", cxx=True)
    assert ours == sdk
