"""Host-side mirror of the SDK's C++ wrapper for the CPU bake path (ref: libraries/omm-lib/include/omm.hpp:973-1088,
`omm::CreateBaker`, `omm::Cpu::CreateTexture`, `omm::Cpu::Bake`, `omm::Cpu::GetBakeResultDesc`), written over ctypes so
the parity tests read like the SDK's own tests (support/tests/test_omm_bake_cpu.cpp:165-205).

Everything here goes through the C ABI of whichever library it is given; no arithmetic of the bake path lives in Python.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Callable, Optional, Sequence

import numpy as np

from . import capi
from .capi import OmmLib

DESC_DTYPE = np.dtype([("offset", "<u4"), ("subdivisionLevel", "<u2"), ("format", "<u2")])
USAGE_DTYPE = np.dtype([("count", "<u4"), ("subdivisionLevel", "<u2"), ("format", "<u2")])
_INDEX_NP = {capi.INDEX_UINT8: np.int8, capi.INDEX_UINT16: np.int16, capi.INDEX_UINT32: np.int32}
_INDEX_IN_NP = {capi.INDEX_UINT8: np.uint8, capi.INDEX_UINT16: np.uint16, capi.INDEX_UINT32: np.uint32}


class OmmError(RuntimeError):
    def __init__(self, what: str, result: int):
        super().__init__(f"{what} failed: ommResult_{capi.RESULT_NAMES[result] if 0 <= result < 6 else result}")
        self.result = result


@dataclass
class BakeResult:
    """Host copy of an ommCpuBakeResultDesc (ref: omm.h:512-530)."""
    array_data: np.ndarray          # uint8[arrayDataSize]
    desc_array: np.ndarray          # DESC_DTYPE[descArrayCount]
    desc_histogram: np.ndarray      # USAGE_DTYPE[...]
    index_buffer: np.ndarray        # int8/int16/int32[indexCount]
    index_format: int
    index_histogram: np.ndarray     # USAGE_DTYPE[...]
    timings: Optional[capi.B200BakeTimings] = None

    def same_bytes(self, other: "BakeResult") -> bool:
        return not self.diff(other)

    def diff(self, other: "BakeResult") -> list:
        """The five comparisons of the SDK's serialize round-trip test (ref: test_omm_bake_cpu.cpp:323-344)."""
        out = []
        if self.index_format != other.index_format:
            out.append(f"indexFormat {self.index_format} != {other.index_format}")
        for name in ("array_data", "desc_array", "desc_histogram", "index_buffer", "index_histogram"):
            a, b = getattr(self, name), getattr(other, name)
            if a.shape != b.shape or a.dtype != b.dtype:
                out.append(f"{name}: shape/dtype {a.shape}{a.dtype} != {b.shape}{b.dtype}")
            elif a.tobytes() != b.tobytes():
                av, bv = a.view(np.uint8).ravel(), b.view(np.uint8).ravel()
                nz = np.nonzero(av != bv)[0]
                out.append(f"{name}: {nz.size} differing bytes, first at {int(nz[0])}")
        return out


def result_sha256(res: BakeResult) -> str:
    """One digest over everything ommCpuBakeResultDesc exposes: the five arrays in the order of the SDK's serialize round-trip comparison
    (test_omm_bake_cpu.cpp:323-344) plus the index format."""
    import hashlib
    h = hashlib.sha256()
    for k in ("array_data", "desc_array", "desc_histogram", "index_buffer", "index_histogram"):
        h.update(np.ascontiguousarray(getattr(res, k)).tobytes())
    h.update(bytes([res.index_format]))
    return h.hexdigest()


class Texture:
    def __init__(self, baker: "Baker", handle: int, keepalive):
        self.baker, self.handle, self._keepalive = baker, handle, keepalive

    def destroy(self):
        if self.handle:
            rc = self.baker.lib.dll.ommCpuDestroyTexture(self.baker.handle, self.handle)
            self.handle = None
            if rc != capi.SUCCESS:
                raise OmmError("ommCpuDestroyTexture", rc)


@dataclass
class BakeInput:
    """Python view of ommCpuBakeInputDesc (ref: omm.h:384-460); defaults are ommCpuBakeInputDescDefault()."""
    texture: Texture
    indices: np.ndarray                      # uint8 / uint16 / uint32, 3 per triangle
    texcoords: np.ndarray                    # float32 (n,2) for UV32_FLOAT, uint32 (n,) packed pairs for the 16-bit formats
    texcoord_format: int = capi.UV32_FLOAT
    texcoord_stride: int = 0
    addressing_mode: int = capi.ADDR_CLAMP
    filter: int = capi.FILTER_LINEAR
    border_alpha: float = 0.0
    alpha_mode: int = capi.ALPHA_TEST
    alpha_cutoff: float = 0.5
    alpha_cutoff_le: int = capi.STATE_T
    alpha_cutoff_gt: int = capi.STATE_O
    format: int = capi.FORMAT_4_STATE
    formats: Optional[np.ndarray] = None     # int32 per triangle
    unknown_state_promotion: int = capi.PROMOTE_FORCE_OPAQUE
    unresolved_tri_state: int = capi.SPECIAL_FUO
    max_subdivision_level: int = 8
    subdivision_levels: Optional[np.ndarray] = None  # uint8 per triangle
    dynamic_subdivision_scale: float = 2.0
    rejection_threshold: float = 0.0
    near_duplicate_factor: float = 0.15
    max_array_data_size: int = 0xFFFFFFFF
    max_workload_size: int = 0xFFFFFFFFFFFFFFFF
    bake_flags: int = capi.BAKE_NONE
    _keep: list = field(default_factory=list, repr=False)

    def to_desc(self) -> capi.CpuBakeInputDesc:
        d = capi.bake_input_desc_default()
        idx = np.ascontiguousarray(self.indices)
        index_format = {np.dtype(np.uint8): capi.INDEX_UINT8, np.dtype(np.uint16): capi.INDEX_UINT16,
                        np.dtype(np.uint32): capi.INDEX_UINT32}[idx.dtype]
        uv = np.ascontiguousarray(self.texcoords)
        self._keep = [idx, uv]
        d.bakeFlags = self.bake_flags
        d.texture = self.texture.handle if self.texture is not None else None
        d.runtimeSamplerDesc = capi.SamplerDesc(self.addressing_mode, self.filter, self.border_alpha)
        d.alphaMode = self.alpha_mode
        d.texCoordFormat = self.texcoord_format
        d.texCoords = uv.ctypes.data
        d.texCoordStrideInBytes = self.texcoord_stride
        d.indexFormat = index_format
        d.indexBuffer = idx.ctypes.data
        d.indexCount = idx.size
        d.dynamicSubdivisionScale = self.dynamic_subdivision_scale
        d.rejectionThreshold = self.rejection_threshold
        d.alphaCutoff = self.alpha_cutoff
        d.nearDuplicateDeduplicationFactor = self.near_duplicate_factor
        d.alphaCutoffLessEqual = self.alpha_cutoff_le
        d.alphaCutoffGreater = self.alpha_cutoff_gt
        d.format = self.format
        if self.formats is not None:
            f = np.ascontiguousarray(self.formats, dtype=np.int32)
            self._keep.append(f)
            d.formats = f.ctypes.data
        d.unknownStatePromotion = self.unknown_state_promotion
        d.unresolvedTriState = self.unresolved_tri_state
        d.maxSubdivisionLevel = self.max_subdivision_level
        d.maxArrayDataSize = self.max_array_data_size
        if self.subdivision_levels is not None:
            s = np.ascontiguousarray(self.subdivision_levels, dtype=np.uint8)
            self._keep.append(s)
            d.subdivisionLevels = s.ctypes.data
        d.maxWorkloadSize = self.max_workload_size
        return d


def _copy_result(desc: capi.CpuBakeResultDesc) -> BakeResult:
    def arr(ptr, count, dtype):
        if not ptr or count == 0:
            return np.zeros(0, dtype=dtype)
        nbytes = count * np.dtype(dtype).itemsize
        addr = C.cast(ptr, C.c_void_p).value
        buf = (C.c_uint8 * nbytes).from_address(addr)
        return np.frombuffer(bytes(buf), dtype=dtype).copy()

    ifmt = desc.indexFormat
    return BakeResult(
        array_data=arr(desc.arrayData, desc.arrayDataSize, np.uint8),
        desc_array=arr(desc.descArray, desc.descArrayCount, DESC_DTYPE),
        desc_histogram=arr(desc.descArrayHistogram, desc.descArrayHistogramCount, USAGE_DTYPE),
        index_buffer=arr(desc.indexBuffer, desc.indexCount, _INDEX_NP[ifmt]),
        index_format=ifmt,
        index_histogram=arr(desc.indexHistogram, desc.indexHistogramCount, USAGE_DTYPE),
    )


class Baker:
    """ommBaker of type CPU on a given library (ref: omm.hpp `omm::CreateBaker`)."""

    def __init__(self, lib: OmmLib, on_message: Optional[Callable[[int, str], None]] = None, baker_type: int = capi.BAKER_CPU):
        self.lib = lib
        self.messages: list[tuple[int, str]] = []
        desc = capi.BakerCreationDesc()
        desc.type = baker_type
        self._cb = None
        if on_message is not None:
            def _cb(sev, msg, _user):
                on_message(int(sev), msg.decode("utf-8", "replace"))
            self._cb = capi.MESSAGE_FN(_cb)
            desc.messageInterface.messageCallback = self._cb
        h = C.c_void_p()
        rc = lib.dll.ommCreateBaker(C.byref(desc), C.byref(h))
        if rc != capi.SUCCESS:
            raise OmmError("ommCreateBaker", rc)
        self.handle = h.value

    def destroy(self):
        if self.handle:
            self.lib.dll.ommDestroyBaker(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.destroy()

    # -- textures ------------------------------------------------------------------------------
    def create_texture(self, mips: Sequence[np.ndarray], alpha_cutoff: float = -1.0, flags: int = capi.TEXFLAG_NONE,
                       row_pitch: Optional[Sequence[int]] = None, widths: Optional[Sequence[int]] = None) -> Texture:
        """mips: 2-D arrays (rows = y), float32 -> FP32, uint8 -> UNORM8 (ref: omm.h:339-382).
        `widths` / `row_pitch` override the per-mip width / rowPitch fields (to describe padded rows)."""
        mips = [np.ascontiguousarray(m) for m in mips]
        fmt = capi.TEX_FP32 if mips[0].dtype == np.float32 else capi.TEX_UNORM8
        if fmt == capi.TEX_UNORM8:
            assert mips[0].dtype == np.uint8
        marr = (capi.CpuTextureMipDesc * len(mips))()
        for i, m in enumerate(mips):
            marr[i].width = m.shape[1] if widths is None else widths[i]
            marr[i].height = m.shape[0]
            marr[i].rowPitch = 0 if row_pitch is None else row_pitch[i]
            marr[i].textureData = m.ctypes.data
        td = capi.CpuTextureDesc()
        td.format, td.flags, td.mips, td.mipCount, td.alphaCutoff = fmt, flags, marr, len(mips), alpha_cutoff
        h = C.c_void_p()
        rc = self.lib.dll.ommCpuCreateTexture(self.handle, C.byref(td), C.byref(h))
        if rc != capi.SUCCESS:
            raise OmmError("ommCpuCreateTexture", rc)
        return Texture(self, h.value, (mips, marr))

    # -- bake ------------------------------------------------------------------------------------
    def bake_raw(self, desc: capi.CpuBakeInputDesc) -> tuple[int, Optional[int]]:
        h = C.c_void_p()
        rc = self.lib.dll.ommCpuBake(self.handle, C.byref(desc), C.byref(h))
        return rc, h.value

    # -- serialization (ommCpuSerialize / ommCpuDeserialize) -----------------------------------------
    def serialize(self, input_descs=(), result_descs=(), flags: int = capi.SERIALIZE_NONE) -> bytes:
        """Blob of the given capi.CpuBakeInputDesc / capi.CpuBakeResultDesc structures."""
        d = capi.CpuDeserializedDesc()
        ins = (capi.CpuBakeInputDesc * max(1, len(input_descs)))(*input_descs)
        outs = (capi.CpuBakeResultDesc * max(1, len(result_descs)))(*result_descs)
        d.flags, d.numInputDescs, d.inputDescs = flags, len(input_descs), ins
        d.numResultDescs, d.resultDescs = len(result_descs), outs
        h = C.c_void_p()
        rc = self.lib.dll.ommCpuSerialize(self.handle, C.byref(d), C.byref(h))
        if rc != capi.SUCCESS:
            raise OmmError("ommCpuSerialize", rc)
        try:
            pb = C.POINTER(capi.CpuBlobDesc)()
            rc = self.lib.dll.ommCpuGetSerializedResultDesc(h, C.byref(pb))
            if rc != capi.SUCCESS:
                raise OmmError("ommCpuGetSerializedResultDesc", rc)
            return C.string_at(pb.contents.data, pb.contents.size)
        finally:
            self.lib.dll.ommCpuDestroySerializedResult(h)

    def deserialize_raw(self, blob: bytes):
        """(rc, handle, POINTER(CpuDeserializedDesc)); the caller destroys the handle with ommCpuDestroyDeserializedResult."""
        buf = C.create_string_buffer(blob, len(blob))
        bd = capi.CpuBlobDesc(C.cast(buf, C.c_void_p).value, len(blob))
        h = C.c_void_p()
        rc = self.lib.dll.ommCpuDeserialize(self.handle, C.byref(bd), C.byref(h))
        if rc != capi.SUCCESS:
            return rc, None, None
        pd = C.POINTER(capi.CpuDeserializedDesc)()
        rc = self.lib.dll.ommCpuGetDeserializedDesc(h, C.byref(pd))
        return rc, h, pd

    def bake_desc(self, desc: capi.CpuBakeInputDesc) -> BakeResult:
        rc, h = self.bake_raw(desc)
        if rc != capi.SUCCESS:
            raise OmmError("ommCpuBake", rc)
        try:
            pdesc = C.POINTER(capi.CpuBakeResultDesc)()
            rc = self.lib.dll.ommCpuGetBakeResultDesc(h, C.byref(pdesc))
            if rc != capi.SUCCESS:
                raise OmmError("ommCpuGetBakeResultDesc", rc)
            return _copy_result(pdesc.contents)
        finally:
            self.lib.dll.ommCpuDestroyBakeResult(h)

    def bake(self, inp: BakeInput) -> BakeResult:
        desc = inp.to_desc()
        rc, h = self.bake_raw(desc)
        if rc != capi.SUCCESS:
            raise OmmError("ommCpuBake", rc)
        try:
            pdesc = C.POINTER(capi.CpuBakeResultDesc)()
            rc = self.lib.dll.ommCpuGetBakeResultDesc(h, C.byref(pdesc))
            if rc != capi.SUCCESS:
                raise OmmError("ommCpuGetBakeResultDesc", rc)
            res = _copy_result(pdesc.contents)
            if self.lib.is_b200:
                t = capi.B200BakeTimings()
                if self.lib.dll.ommB200GetLastBakeTimings(self.handle, C.byref(t)) == capi.SUCCESS:
                    res.timings = t
            return res
        finally:
            self.lib.dll.ommCpuDestroyBakeResult(h)
