"""ctypes mirror of include/omm_b200.h (== the CPU-bake subset of the SDK's omm.h, ref:
/root/reference/libraries/omm-lib/include/omm.h).

The same binding drives three shared libraries that export this ABI:
  * omm_b200/lib/libomm-b200.so -- the product (CUDA, sm_100a),
  * oracle/liboracle_port.so    -- the plain-C restatement (test infrastructure),
  * oracle/_ref/libomm-lib.so   -- the unmodified SDK build (test infrastructure).
Only `tests/`, `bench.py`'s baseline legs and `__graft_entry__.smoke()` ever load the last two.
"""
from __future__ import annotations

import ctypes as C
import os

# ---- enums (ref: omm.h:78-192, 282-334) -------------------------------------------------------
SUCCESS, FAILURE, INVALID_ARGUMENT, INSUFFICIENT_SCRATCH_MEMORY, NOT_IMPLEMENTED, WORKLOAD_TOO_BIG = range(6)
RESULT_NAMES = ["SUCCESS", "FAILURE", "INVALID_ARGUMENT", "INSUFFICIENT_SCRATCH_MEMORY", "NOT_IMPLEMENTED", "WORKLOAD_TOO_BIG"]

SEVERITY_INFO, SEVERITY_PERF_WARNING, SEVERITY_ERROR, SEVERITY_FATAL = range(4)

STATE_T, STATE_O, STATE_UT, STATE_UO = range(4)
SPECIAL_FT, SPECIAL_FO, SPECIAL_FUT, SPECIAL_FUO = -1, -2, -3, -4
FORMAT_INVALID, FORMAT_2_STATE, FORMAT_4_STATE = 0, 1, 2
PROMOTE_NEAREST, PROMOTE_FORCE_OPAQUE, PROMOTE_FORCE_TRANSPARENT = range(3)
BAKER_GPU, BAKER_CPU, BAKER_MAX = range(3)
UV16_UNORM, UV16_FLOAT, UV32_FLOAT, UV_MAX = range(4)
INDEX_UINT16, INDEX_UINT32, INDEX_UINT8, INDEX_MAX = range(4)
ADDR_WRAP, ADDR_MIRROR, ADDR_CLAMP, ADDR_BORDER, ADDR_MIRROR_ONCE, ADDR_MAX = range(6)
FILTER_NEAREST, FILTER_LINEAR, FILTER_MAX = range(3)
ALPHA_TEST, ALPHA_BLEND, ALPHA_MAX = range(3)
TEX_UNORM8, TEX_FP32, TEX_MAX = range(3)
TEXFLAG_NONE, TEXFLAG_DISABLE_ZORDER = 0, 1

BAKE_NONE = 0
BAKE_ENABLE_INTERNAL_THREADS = 1 << 0
BAKE_DISABLE_SPECIAL_INDICES = 1 << 1
BAKE_FORCE_32BIT_INDICES = 1 << 2
BAKE_DISABLE_DUPLICATE_DETECTION = 1 << 3
BAKE_ENABLE_NEAR_DUPLICATE_DETECTION = 1 << 4
BAKE_ENABLE_VALIDATION = 1 << 5
BAKE_ALLOW_8BIT_INDICES = 1 << 6
# undocumented internal bits (ref: bake_cpu_impl.cpp:44-48)
BAKE_INT_AABB_TESTING = 1 << 7
BAKE_INT_DISABLE_LEVEL_LINE = 1 << 8
BAKE_INT_DISABLE_FINE = 1 << 9
BAKE_INT_NEAR_DUP_BRUTE_FORCE = 1 << 10
BAKE_INT_EDGE_HEURISTIC = 1 << 11

INDEX_FORMAT_BYTES = {INDEX_UINT16: 2, INDEX_UINT32: 4, INDEX_UINT8: 1}


# ---- structs ---------------------------------------------------------------------------------
class LibraryDesc(C.Structure):
    _fields_ = [("versionMajor", C.c_uint8), ("versionMinor", C.c_uint8), ("versionBuild", C.c_uint8)]


class SamplerDesc(C.Structure):
    _fields_ = [("addressingMode", C.c_int), ("filter", C.c_int), ("borderAlpha", C.c_float)]


ALLOCATE_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t)
REALLOCATE_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t)
FREE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p)
MESSAGE_FN = C.CFUNCTYPE(None, C.c_int, C.c_char_p, C.c_void_p)


class MemoryAllocatorInterface(C.Structure):
    _fields_ = [("allocate", ALLOCATE_FN), ("reallocate", REALLOCATE_FN), ("free", FREE_FN), ("userArg", C.c_void_p)]


class MessageInterface(C.Structure):
    _fields_ = [("messageCallback", MESSAGE_FN), ("userArg", C.c_void_p)]


class BakerCreationDesc(C.Structure):
    _fields_ = [("type", C.c_int), ("memoryAllocatorInterface", MemoryAllocatorInterface), ("messageInterface", MessageInterface)]


class CpuTextureMipDesc(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("rowPitch", C.c_uint32), ("textureData", C.c_void_p)]


class CpuTextureDesc(C.Structure):
    _fields_ = [("format", C.c_int), ("flags", C.c_int), ("mips", C.POINTER(CpuTextureMipDesc)), ("mipCount", C.c_uint32),
                ("alphaCutoff", C.c_float)]


class CpuBakeInputDesc(C.Structure):
    _fields_ = [
        ("bakeFlags", C.c_int),
        ("texture", C.c_void_p),
        ("runtimeSamplerDesc", SamplerDesc),
        ("alphaMode", C.c_int),
        ("texCoordFormat", C.c_int),
        ("texCoords", C.c_void_p),
        ("texCoordStrideInBytes", C.c_uint32),
        ("indexFormat", C.c_int),
        ("indexBuffer", C.c_void_p),
        ("indexCount", C.c_uint32),
        ("dynamicSubdivisionScale", C.c_float),
        ("rejectionThreshold", C.c_float),
        ("alphaCutoff", C.c_float),
        ("nearDuplicateDeduplicationFactor", C.c_float),
        ("alphaCutoffLessEqual", C.c_int),
        ("alphaCutoffGreater", C.c_int),
        ("format", C.c_int),
        ("formats", C.c_void_p),
        ("unknownStatePromotion", C.c_int),
        ("unresolvedTriState", C.c_int),
        ("maxSubdivisionLevel", C.c_uint8),
        ("maxArrayDataSize", C.c_uint32),
        ("subdivisionLevels", C.c_void_p),
        ("maxWorkloadSize", C.c_uint64),
    ]


assert C.sizeof(CpuBakeInputDesc) == 136  # ref: serialize_impl.cpp:86


class CpuOpacityMicromapDesc(C.Structure):
    _fields_ = [("offset", C.c_uint32), ("subdivisionLevel", C.c_uint16), ("format", C.c_uint16)]


class CpuOpacityMicromapUsageCount(C.Structure):
    _fields_ = [("count", C.c_uint32), ("subdivisionLevel", C.c_uint16), ("format", C.c_uint16)]


class CpuBakeResultDesc(C.Structure):
    _fields_ = [
        ("arrayData", C.c_void_p),
        ("arrayDataSize", C.c_uint32),
        ("descArray", C.POINTER(CpuOpacityMicromapDesc)),
        ("descArrayCount", C.c_uint32),
        ("descArrayHistogram", C.POINTER(CpuOpacityMicromapUsageCount)),
        ("descArrayHistogramCount", C.c_uint32),
        ("indexBuffer", C.c_void_p),
        ("indexCount", C.c_uint32),
        ("indexFormat", C.c_int),
        ("indexHistogram", C.POINTER(CpuOpacityMicromapUsageCount)),
        ("indexHistogramCount", C.c_uint32),
    ]


class DebugStats(C.Structure):
    _fields_ = [
        ("totalOpaque", C.c_uint64),
        ("totalTransparent", C.c_uint64),
        ("totalUnknownTransparent", C.c_uint64),
        ("totalUnknownOpaque", C.c_uint64),
        ("totalFullyOpaque", C.c_uint32),
        ("totalFullyTransparent", C.c_uint32),
        ("totalFullyUnknownOpaque", C.c_uint32),
        ("totalFullyUnknownTransparent", C.c_uint32),
        ("knownAreaMetric", C.c_float),
    ]


class B200BakeTimings(C.Structure):
    _fields_ = [
        ("h2dMs", C.c_float), ("setupMs", C.c_float), ("classifyMs", C.c_float), ("postMs", C.c_float), ("d2hMs", C.c_float),
        ("totalDeviceMs", C.c_float),
        ("microTriangles", C.c_uint64), ("workItems", C.c_uint32), ("kernelLaunches", C.c_uint32),
        ("h2dBytes", C.c_uint64), ("d2hBytes", C.c_uint64), ("arrayDataBytes", C.c_uint64),
        ("descCount", C.c_uint32), ("reserved", C.c_uint32),
        ("hostStageMs", C.c_float), ("hostBakeMs", C.c_float), ("hostDownloadMs", C.c_float), ("hostTotalMs", C.c_float),
        ("itemPostMs", C.c_float), ("gatherMs", C.c_float),
    ]


class B200DeviceResultDesc(C.Structure):
    _fields_ = [("arrayData", C.c_void_p), ("descArray", C.c_void_p), ("indexBuffer", C.c_void_p), ("arrayDataSize", C.c_uint32),
                ("descArrayCount", C.c_uint32), ("indexCount", C.c_uint32), ("indexFormat", C.c_int)]


class CpuBlobDesc(C.Structure):  # ref: omm.h:532-536
    _fields_ = [("data", C.c_void_p), ("size", C.c_uint64)]


class CpuDeserializedDesc(C.Structure):  # ref: omm.h:546-555
    _fields_ = [("flags", C.c_int), ("numInputDescs", C.c_int), ("inputDescs", C.POINTER(CpuBakeInputDesc)), ("numResultDescs", C.c_int),
                ("resultDescs", C.POINTER(CpuBakeResultDesc))]


SERIALIZE_NONE, SERIALIZE_COMPRESS = 0, 1


def bake_input_desc_default() -> CpuBakeInputDesc:
    """ref: omm.h:462-490 (ommCpuBakeInputDescDefault)."""
    d = CpuBakeInputDesc()
    d.bakeFlags = BAKE_NONE
    d.texture = None
    d.runtimeSamplerDesc = SamplerDesc(ADDR_MAX, FILTER_MAX, 0.0)
    d.alphaMode = ALPHA_MAX
    d.texCoordFormat = UV_MAX
    d.texCoords = None
    d.texCoordStrideInBytes = 0
    d.indexFormat = INDEX_MAX
    d.indexBuffer = None
    d.indexCount = 0
    d.dynamicSubdivisionScale = 2.0
    d.rejectionThreshold = 0.0
    d.alphaCutoff = 0.5
    d.nearDuplicateDeduplicationFactor = 0.15
    d.alphaCutoffLessEqual = STATE_T
    d.alphaCutoffGreater = STATE_O
    d.format = FORMAT_4_STATE
    d.formats = None
    d.unknownStatePromotion = PROMOTE_FORCE_OPAQUE
    d.unresolvedTriState = SPECIAL_FUO
    d.maxSubdivisionLevel = 8
    d.maxArrayDataSize = 0xFFFFFFFF
    d.subdivisionLevels = None
    d.maxWorkloadSize = 0xFFFFFFFFFFFFFFFF
    return d


REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# OMM_B200_LIB overrides the library path (used to A/B kernel build variants; the default is the in-tree build)
PRODUCT_LIB = os.environ.get("OMM_B200_LIB") or os.path.join(REPO_ROOT, "omm_b200", "lib", "libomm-b200.so")

# the ABI every library must export (include/omm_b200.h, first half)
CORE_SYMBOLS = [
    "ommGetLibraryDesc", "ommCreateBaker", "ommDestroyBaker", "ommCpuCreateTexture", "ommCpuGetTextureDesc",
    "ommCpuDestroyTexture", "ommCpuBake", "ommCpuDestroyBakeResult", "ommCpuGetBakeResultDesc",
]
# SURVEY 8f row N2 (the product and the SDK build export them; the plain-C port of the bake path does not)
SERIALIZE_SYMBOLS = [
    "ommCpuSerialize", "ommCpuGetSerializedResultDesc", "ommCpuDestroySerializedResult", "ommCpuDeserialize",
    "ommCpuGetDeserializedDesc", "ommCpuDestroyDeserializedResult",
]
# product-only symbols (include/omm_b200.h, second half + ommDebugGetStats)
B200_SYMBOLS = [
    "ommDebugGetStats", "ommB200SetDevice", "ommB200GetDeviceCount", "ommB200GetLastBakeTimings", "ommB200StageInputs",
    "ommB200DestroyStagedInputs", "ommB200BakeResident", "ommB200GetDeviceResultDesc", "ommB200DownloadResult",
    "ommB200InitSharding", "ommB200GetNcclUniqueId", "ommB200ComputeShardBounds", "ommB200ShardsPerRank", "ommB200ShardOwner",
    "ommB200TrimHostPool", "ommB200SetShardedResultMode",
]
SHARDED_RESULT_REPLICATED, SHARDED_RESULT_ON_RANK0 = 0, 1


class OmmLib:
    """A loaded library exporting the omm C ABI."""

    def __init__(self, path: str):
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} is missing -- build it first (python -c 'import __graft_entry__ as g; g.build()')")
        self.path = path
        self.dll = C.CDLL(path, mode=C.RTLD_LOCAL)
        d = self.dll
        d.ommGetLibraryDesc.restype = LibraryDesc
        d.ommGetLibraryDesc.argtypes = []
        d.ommCreateBaker.restype = C.c_int
        d.ommCreateBaker.argtypes = [C.POINTER(BakerCreationDesc), C.POINTER(C.c_void_p)]
        d.ommDestroyBaker.restype = C.c_int
        d.ommDestroyBaker.argtypes = [C.c_void_p]
        d.ommCpuCreateTexture.restype = C.c_int
        d.ommCpuCreateTexture.argtypes = [C.c_void_p, C.POINTER(CpuTextureDesc), C.POINTER(C.c_void_p)]
        d.ommCpuGetTextureDesc.restype = C.c_int
        d.ommCpuGetTextureDesc.argtypes = [C.c_void_p, C.POINTER(CpuTextureDesc)]
        d.ommCpuDestroyTexture.restype = C.c_int
        d.ommCpuDestroyTexture.argtypes = [C.c_void_p, C.c_void_p]
        d.ommCpuBake.restype = C.c_int
        d.ommCpuBake.argtypes = [C.c_void_p, C.POINTER(CpuBakeInputDesc), C.POINTER(C.c_void_p)]
        d.ommCpuDestroyBakeResult.restype = C.c_int
        d.ommCpuDestroyBakeResult.argtypes = [C.c_void_p]
        d.ommCpuGetBakeResultDesc.restype = C.c_int
        d.ommCpuGetBakeResultDesc.argtypes = [C.c_void_p, C.POINTER(C.POINTER(CpuBakeResultDesc))]
        # serialization: omm.h passes the two descs by C++ reference = by pointer at the ABI level
        self.has_serialize = hasattr(d, "ommCpuSerialize")
        if self.has_serialize:
            self._bind_serialize(d)
        self.has_debug_stats = hasattr(d, "ommDebugGetStats")
        if self.has_debug_stats:
            d.ommDebugGetStats.restype = C.c_int
            d.ommDebugGetStats.argtypes = [C.c_void_p, C.POINTER(CpuBakeResultDesc), C.POINTER(DebugStats)]
        self._bind_b200(d)

    @staticmethod
    def _bind_serialize(d):
        d.ommCpuSerialize.restype = C.c_int
        d.ommCpuSerialize.argtypes = [C.c_void_p, C.POINTER(CpuDeserializedDesc), C.POINTER(C.c_void_p)]
        d.ommCpuGetSerializedResultDesc.restype = C.c_int
        d.ommCpuGetSerializedResultDesc.argtypes = [C.c_void_p, C.POINTER(C.POINTER(CpuBlobDesc))]
        d.ommCpuDestroySerializedResult.restype = C.c_int
        d.ommCpuDestroySerializedResult.argtypes = [C.c_void_p]
        d.ommCpuDeserialize.restype = C.c_int
        d.ommCpuDeserialize.argtypes = [C.c_void_p, C.POINTER(CpuBlobDesc), C.POINTER(C.c_void_p)]
        d.ommCpuGetDeserializedDesc.restype = C.c_int
        d.ommCpuGetDeserializedDesc.argtypes = [C.c_void_p, C.POINTER(C.POINTER(CpuDeserializedDesc))]
        d.ommCpuDestroyDeserializedResult.restype = C.c_int
        d.ommCpuDestroyDeserializedResult.argtypes = [C.c_void_p]

    def _bind_b200(self, d):
        self.is_b200 = hasattr(d, "ommB200BakeResident")
        if self.is_b200:
            d.ommB200SetDevice.restype = C.c_int
            d.ommB200SetDevice.argtypes = [C.c_int]
            d.ommB200GetDeviceCount.restype = C.c_int
            d.ommB200GetDeviceCount.argtypes = []
            d.ommB200GetLastBakeTimings.restype = C.c_int
            d.ommB200GetLastBakeTimings.argtypes = [C.c_void_p, C.POINTER(B200BakeTimings)]
            d.ommB200StageInputs.restype = C.c_int
            d.ommB200StageInputs.argtypes = [C.c_void_p, C.POINTER(CpuBakeInputDesc), C.POINTER(C.c_void_p)]
            d.ommB200DestroyStagedInputs.restype = C.c_int
            d.ommB200DestroyStagedInputs.argtypes = [C.c_void_p]
            d.ommB200BakeResident.restype = C.c_int
            d.ommB200BakeResident.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
            d.ommB200GetDeviceResultDesc.restype = C.c_int
            d.ommB200GetDeviceResultDesc.argtypes = [C.c_void_p, C.POINTER(B200DeviceResultDesc)]
            d.ommB200DownloadResult.restype = C.c_int
            d.ommB200DownloadResult.argtypes = [C.c_void_p]
            d.ommB200InitSharding.restype = C.c_int
            d.ommB200InitSharding.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
            d.ommB200GetNcclUniqueId.restype = C.c_int
            d.ommB200GetNcclUniqueId.argtypes = [C.c_void_p, C.c_size_t]
            d.ommB200ComputeShardBounds.restype = C.c_int
            d.ommB200ComputeShardBounds.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_void_p]
            d.ommB200ShardsPerRank.restype = C.c_int
            d.ommB200ShardsPerRank.argtypes = [C.c_int]
            d.ommB200ShardOwner.restype = C.c_int
            d.ommB200ShardOwner.argtypes = [C.c_int, C.c_int]
            d.ommB200SetShardedResultMode.restype = C.c_int
            d.ommB200SetShardedResultMode.argtypes = [C.c_void_p, C.c_int]
            d.ommB200TrimHostPool.restype = C.c_size_t
            d.ommB200TrimHostPool.argtypes = [C.c_size_t]

    def exported(self, name: str) -> bool:
        return hasattr(self.dll, name)


_product: OmmLib | None = None


def load_product_library() -> OmmLib:
    """Load libomm-b200.so.  There is no CPU fallback: a missing library is an error."""
    global _product
    if _product is None:
        _product = OmmLib(PRODUCT_LIB)
        if not _product.is_b200:
            raise RuntimeError(f"{PRODUCT_LIB} does not export the ommB200* entry points")
    return _product
