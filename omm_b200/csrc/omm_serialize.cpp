// omm_serialize.cpp -- ommCpuSerialize / ommCpuDeserialize: the SDK's blob format for bake inputs
// and results, so that captured inputs (the SDK's viewer blobs, its leaves.bin example, the v1.4-v1.7 golden blobs of its tests)
// can drive this baker and its results can be read back by SDK tools.
//
// Format (ref: libraries/omm-lib/src/serialize_impl.cpp:79-221, serialize_impl.h:43-66, texture_impl.h:232-345), little endian,
// fields written back to back without padding:
//   header   u64 XXH64(seed 42) of everything after it | i32 major, minor, patch | i32 inputDescVersion (5) | i32 flags |
//            i32 decompressedSize (version >= 2; non-zero = the rest is one LZ4 block)
//   body     i32 numInputDescs, input descs, i32 numResultDescs, result descs
//   input    bakeFlags | TEXTURE | addressingMode, filter, borderAlpha, alphaMode | texCoordFormat, u64 texCoordBytes, texCoords,
//            texCoordStride | indexFormat, indexCount, indices | dynamicSubdivisionScale, rejectionThreshold, alphaCutoff,
//            alphaCutoffLessEqual, alphaCutoffGreater, format | u64 numFormats, formats | unknownStatePromotion,
//            unresolvedTriState (v >= 2), maxSubdivisionLevel, maxArrayDataSize (v >= 4) | u64 numLevels, levels | maxWorkloadSize
//   texture  i32 numMips, per mip {i32 w, h; f32 rcpW, rcpH; u64 dataOffset, numElements, dataOffsetSAT} | tilingMode,
//            flags + alphaCutoff (v >= 3), format | u64 dataSize, data | u64 satSize, sat
//            -- the SDK dumps its internal texture memory: Z-order tiled (padded to nextPow2(max(w,h))^2 elements per mip) unless
//            DisableZOrder, each mip aligned to 64 bytes, plus its summed-area tables.  The SDK leaves the padding uninitialised; it
//            is zero here, so blobs are byte-identical with the SDK's exactly when there is no padding (square power-of-two mips of
//            at least 64 bytes).
//   result   u32 count + bytes for arrayData, descArray, descArrayHistogram | indexFormat | u32 count + indices | indexHistogram
//
// ommCpuSerializeFlags_Compress: the SDK compresses the body with LZ4_compress_default (vendored lz4 1.10.0).  Reproducing that
// compressor's exact output is out of scope; this library writes the body uncompressed (decompressedSize = 0, a valid blob the SDK
// reads), and READS compressed blobs with its own LZ4 block decoder (format: lz4_Block_format.md).
#include <algorithm>
#include <cstring>
#include <vector>

#include "omm_internal.h"
#include "omm_xxh64.h"

namespace ommb200 {


namespace {

constexpr int kSerializeVersion = 5;  // ref: serialize_impl.h:54-56
constexpr int kLibMajor = 1, kLibMinor = 9, kLibPatch = 0;

struct Writer {
    std::vector<uint8_t> bytes;
    template <class T>
    void put(const T& v) {
        const uint8_t* p = reinterpret_cast<const uint8_t*>(&v);
        bytes.insert(bytes.end(), p, p + sizeof(T));
    }
    void raw(const void* p, size_t n) {
        if (n) bytes.insert(bytes.end(), (const uint8_t*)p, (const uint8_t*)p + n);
    }
    void zeros(size_t n) { bytes.insert(bytes.end(), n, (uint8_t)0); }
};

struct Reader {
    const uint8_t* p;
    const uint8_t* end;
    bool ok = true;
    template <class T>
    T get() {
        T v{};
        if ((size_t)(end - p) < sizeof(T)) {
            ok = false;
            p = end;
            return v;
        }
        memcpy(&v, p, sizeof(T));
        p += sizeof(T);
        return v;
    }
    const uint8_t* take(size_t n) {
        if ((size_t)(end - p) < n) {
            ok = false;
            p = end;
            return nullptr;
        }
        const uint8_t* q = p;
        p += n;
        return q;
    }
};

uint32_t NextPow2(uint32_t v) {  // ref: util/bit_tricks.h:25-34
    v += (v == 0);
    v--;
    v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16;
    return ++v;
}
uint32_t Part1By1(uint32_t x) {
    x &= 0xFFFFu;
    x = (x | (x << 8)) & 0x00FF00FFu; x = (x | (x << 4)) & 0x0F0F0F0Fu; x = (x | (x << 2)) & 0x33333333u; x = (x | (x << 1)) & 0x55555555u;
    return x;
}
uint32_t Morton(uint32_t x, uint32_t y) { return Part1By1(x) | (Part1By1(y) << 1); }  // ref: util/bit_tricks.h:40-64, x in the even bits
size_t Align64(size_t v) { return (v + 63) & ~(size_t)63; }

uint32_t MaxIndexOf(const ommCpuBakeInputDesc& d) {  // ref: serialize_impl.cpp:62-77
    uint32_t m = 0;
    const size_t n = (size_t)(d.indexCount / 3u) * 3u;
    for (size_t i = 0; i < n; ++i) {
        const uint32_t v = d.indexFormat == ommIndexFormat_UINT_8 ? ((const uint8_t*)d.indexBuffer)[i]
                         : d.indexFormat == ommIndexFormat_UINT_16 ? ((const uint16_t*)d.indexBuffer)[i] : ((const uint32_t*)d.indexBuffer)[i];
        m = v > m ? v : m;
    }
    return m;
}
size_t TexCoordSize(ommTexCoordFormat f) { return f == ommTexCoordFormat_UV32_FLOAT ? 8 : 4; }
size_t IndexSize(ommIndexFormat f) { return f == ommIndexFormat_UINT_8 ? 1 : (f == ommIndexFormat_UINT_16 ? 2 : 4); }

// ---- texture ----------------------------------------------------------------------------------------------------------------
void WriteTexture(Writer& w, const TextureObject& t) {  // ref: texture_impl.h:232-270, texture_impl.cpp:77-224
    const bool linear = ((uint32_t)t.flags & (uint32_t)ommCpuTextureFlags_DisableZOrder) != 0;
    const size_t spp = t.format == ommCpuTextureFormat_FP32 ? 4 : 1;
    const bool sat = t.alphaCutoff >= 0.f;
    const int numMips = (int)t.mipCount;
    struct MipLayout { uint64_t dataOffset, numElements, satOffset; };
    std::vector<MipLayout> lay(numMips);
    size_t dataSize = 0, satSize = 0;
    for (int i = 0; i < numMips; ++i) {
        const DevMip& m = t.dev.mips[i];
        lay[i].dataOffset = dataSize;
        lay[i].satOffset = satSize;
        const uint64_t side = NextPow2((uint32_t)std::max(m.w, m.h));
        lay[i].numElements = linear ? (uint64_t)m.w * m.h : side * side;
        dataSize = Align64(dataSize + spp * lay[i].numElements);
        if (sat) satSize = Align64(satSize + 4 * lay[i].numElements);
    }
    w.put<int32_t>(numMips);
    for (int i = 0; i < numMips; ++i) {
        const DevMip& m = t.dev.mips[i];
        w.put<int32_t>(m.w); w.put<int32_t>(m.h);
        w.put<float>(m.rcpw); w.put<float>(m.rcph);
        w.put<uint64_t>(lay[i].dataOffset); w.put<uint64_t>(lay[i].numElements); w.put<uint64_t>(lay[i].satOffset);
    }
    w.put<int32_t>(linear ? 0 : 1);  // TilingMode: Linear = 0, MortonZ = 1 (ref: texture_impl.h:26-30)
    w.put<int32_t>((int32_t)t.flags);
    w.put<float>(t.alphaCutoff);
    w.put<int32_t>((int32_t)t.format);
    w.put<uint64_t>(dataSize);
    {
        const size_t base = w.bytes.size();
        w.zeros(dataSize);
        for (int i = 0; i < numMips; ++i) {
            const DevMip& m = t.dev.mips[i];
            const uint8_t* src = (const uint8_t*)t.hostTexels + m.texelOffset * spp;
            uint8_t* dst = w.bytes.data() + base + lay[i].dataOffset;
            if (linear) memcpy(dst, src, spp * (size_t)m.w * m.h);
            else
                for (int y = 0; y < m.h; ++y)
                    for (int x = 0; x < m.w; ++x) memcpy(dst + (size_t)Morton((uint32_t)x, (uint32_t)y) * spp, src + ((size_t)y * m.w + x) * spp, spp);
        }
    }
    w.put<uint64_t>(satSize);
    if (sat) {
        const size_t base = w.bytes.size();
        w.zeros(satSize);
        for (int i = 0; i < numMips; ++i) {
            const DevMip& m = t.dev.mips[i];
            const uint8_t* src = (const uint8_t*)t.hostTexels + m.texelOffset * spp;
            std::vector<uint32_t> s((size_t)m.w * m.h);
            for (int y = 0; y < m.h; ++y)
                for (int x = 0; x < m.w; ++x) {
                    const size_t k = (size_t)y * m.w + x;
                    const float a = spp == 4 ? ((const float*)src)[k] : (float)src[k] * (1.f / 255.f);
                    s[k] = (a > t.alphaCutoff ? 1u : 0u) + (x ? s[k - 1] : 0u);
                }
            for (int y = 1; y < m.h; ++y)
                for (int x = 0; x < m.w; ++x) s[(size_t)y * m.w + x] += s[(size_t)(y - 1) * m.w + x];
            memcpy(w.bytes.data() + base + lay[i].satOffset, s.data(), 4 * s.size());  // row-major w x h at the start of the mip's SAT range
        }
    }
}

// Fills a TextureObject (host side) from the blob; the caller uploads it.
bool ReadTexture(Reader& r, int version, const HostAllocator& alloc, TextureObject* t) {
    const int numMips = r.get<int32_t>();
    if (!r.ok || numMips <= 0 || numMips > kMaxMips) return false;
    struct MipLayout { uint64_t dataOffset, numElements, satOffset; };
    std::vector<MipLayout> lay(numMips);
    size_t totalTexels = 0;
    for (int i = 0; i < numMips; ++i) {
        DevMip& m = t->dev.mips[i];
        m.w = r.get<int32_t>(); m.h = r.get<int32_t>();
        r.get<float>(); r.get<float>();  // rcpSize is recomputed
        lay[i].dataOffset = r.get<uint64_t>(); lay[i].numElements = r.get<uint64_t>(); lay[i].satOffset = r.get<uint64_t>();
        if (!r.ok || m.w <= 0 || m.h <= 0 || m.w > 65536 || m.h > 65536) return false;
        int lw = 0, lh = 0;
        for (uint32_t v = (uint32_t)m.w; (v & 1u) == 0; v >>= 1) lw++;
        for (uint32_t v = (uint32_t)m.h; (v & 1u) == 0; v >>= 1) lh++;
        m.log2w = lw; m.log2h = lh;
        m.isPow2 = (m.w & (m.w - 1)) == 0 && (m.h & (m.h - 1)) == 0;
        m.rcpw = 1.f / (float)m.w; m.rcph = 1.f / (float)m.h;
        m.texelOffset = totalTexels; m.satOffset = totalTexels;
        totalTexels += (size_t)m.w * m.h;
    }
    const int tiling = r.get<int32_t>();
    if (version >= 3) {
        t->flags = (ommCpuTextureFlags)r.get<int32_t>();
        t->alphaCutoff = r.get<float>();
    } else {  // ref: texture_impl.h:311-323
        t->flags = tiling == 1 ? ommCpuTextureFlags_None : ommCpuTextureFlags_DisableZOrder;
        t->alphaCutoff = -1.f;
    }
    t->format = (ommCpuTextureFormat)r.get<int32_t>();
    if (!r.ok || (tiling != 0 && tiling != 1) || (t->format != ommCpuTextureFormat_FP32 && t->format != ommCpuTextureFormat_UNORM8)) return false;
    const size_t spp = t->format == ommCpuTextureFormat_FP32 ? 4 : 1;
    const uint64_t dataSize = r.get<uint64_t>();
    const uint8_t* data = r.take((size_t)dataSize);
    const uint64_t satSize = r.get<uint64_t>();
    const uint8_t* satData = r.take((size_t)satSize);
    if (!r.ok) return false;
    (void)satData;  // the summed-area table is rebuilt on the device from the texels
    t->mipCount = (uint32_t)numMips;
    t->dev.mipCount = numMips;
    // every mip must lie inside the texel payload the blob really carries -- checked BEFORE anything is sized from the header fields
    // (a forged header could otherwise ask for 17 x 16 GiB)
    for (int i = 0; i < numMips; ++i) {
        const DevMip& m = t->dev.mips[i];
        const uint64_t need = tiling == 0 ? (uint64_t)m.w * m.h : (uint64_t)Morton((uint32_t)m.w - 1, (uint32_t)m.h - 1) + 1;
        if (lay[i].dataOffset > dataSize || need * spp > dataSize - lay[i].dataOffset) return false;
    }
    t->hostBytes = totalTexels * spp;
    t->hostTexels = alloc.alloc(t->hostBytes, 64);
    if (!t->hostTexels) return false;
    for (int i = 0; i < numMips; ++i) {
        const DevMip& m = t->dev.mips[i];
        uint8_t* dst = (uint8_t*)t->hostTexels + m.texelOffset * spp;
        const uint8_t* src = data + lay[i].dataOffset;
        if (tiling == 0) memcpy(dst, src, spp * (size_t)m.w * m.h);
        else
            for (int y = 0; y < m.h; ++y)
                for (int x = 0; x < m.w; ++x) memcpy(dst + ((size_t)y * m.w + x) * spp, src + (size_t)Morton((uint32_t)x, (uint32_t)y) * spp, spp);
    }
    t->hasSerializedSat = satSize != 0;
    return true;
}

// ---- descs -------------------------------------------------------------------------------------------------------------------
void WriteInput(Writer& w, const ommCpuBakeInputDesc& d) {  // ref: serialize_impl.cpp:79-159
    w.put<int32_t>((int32_t)d.bakeFlags);
    WriteTexture(w, *HandlePtr<TextureObject>(d.texture));
    w.put<int32_t>((int32_t)d.runtimeSamplerDesc.addressingMode);
    w.put<int32_t>((int32_t)d.runtimeSamplerDesc.filter);
    w.put<float>(d.runtimeSamplerDesc.borderAlpha);
    w.put<int32_t>((int32_t)d.alphaMode);
    w.put<int32_t>((int32_t)d.texCoordFormat);
    const uint64_t texCoordBytes = TexCoordSize(d.texCoordFormat) * ((uint64_t)MaxIndexOf(d) + 1);
    w.put<uint64_t>(texCoordBytes);
    w.raw(d.texCoords, (size_t)texCoordBytes);
    w.put<uint32_t>(d.texCoordStrideInBytes);
    w.put<int32_t>((int32_t)d.indexFormat);
    w.put<uint32_t>(d.indexCount);
    w.raw(d.indexBuffer, (size_t)d.indexCount * IndexSize(d.indexFormat));
    w.put<float>(d.dynamicSubdivisionScale);
    w.put<float>(d.rejectionThreshold);
    w.put<float>(d.alphaCutoff);
    w.put<int32_t>((int32_t)d.alphaCutoffLessEqual);
    w.put<int32_t>((int32_t)d.alphaCutoffGreater);
    w.put<int32_t>((int32_t)d.format);
    const uint64_t numFormats = d.formats ? d.indexCount : 0;
    w.put<uint64_t>(numFormats);
    w.raw(d.formats, (size_t)numFormats * 4);
    w.put<int32_t>((int32_t)d.unknownStatePromotion);
    w.put<int32_t>((int32_t)d.unresolvedTriState);
    w.put<uint8_t>(d.maxSubdivisionLevel);
    w.put<uint32_t>(d.maxArrayDataSize);
    const uint64_t numLevels = d.subdivisionLevels ? d.indexCount : 0;
    w.put<uint64_t>(numLevels);
    w.raw(d.subdivisionLevels, (size_t)numLevels);
    w.put<uint64_t>(d.maxWorkloadSize);
}

template <class T>
void WriteArray(Writer& w, const T* data, uint32_t count) {  // ref: serialize_impl.cpp:23-29
    w.put<uint32_t>(count);
    if (count) w.raw(data, sizeof(T) * (size_t)count);
}
void WriteResult(Writer& w, const ommCpuBakeResultDesc& r) {  // ref: serialize_impl.cpp:161-187
    WriteArray<uint8_t>(w, (const uint8_t*)r.arrayData, r.arrayDataSize);
    WriteArray<ommCpuOpacityMicromapDesc>(w, r.descArray, r.descArrayCount);
    WriteArray<ommCpuOpacityMicromapUsageCount>(w, r.descArrayHistogram, r.descArrayHistogramCount);
    w.put<int32_t>((int32_t)r.indexFormat);
    w.put<uint32_t>(r.indexCount);
    if (r.indexCount) w.raw(r.indexBuffer, IndexSize(r.indexFormat) * (size_t)r.indexCount);
    WriteArray<ommCpuOpacityMicromapUsageCount>(w, r.indexHistogram, r.indexHistogramCount);
}

// LZ4 block decoder (lz4_Block_format.md): sequences of [token][literal length+][literals][offset u16][match length+]
bool Lz4Decode(const uint8_t* src, size_t srcSize, uint8_t* dst, size_t dstSize) {
    const uint8_t* ip = src;
    const uint8_t* const iend = src + srcSize;
    uint8_t* op = dst;
    uint8_t* const oend = dst + dstSize;
    while (ip < iend) {
        const unsigned token = *ip++;
        size_t lit = token >> 4;
        if (lit == 15) {
            unsigned b;
            do {
                if (ip >= iend) return false;
                b = *ip++;
                lit += b;
            } while (b == 255);
        }
        if ((size_t)(iend - ip) < lit || (size_t)(oend - op) < lit) return false;
        memcpy(op, ip, lit);
        op += lit;
        ip += lit;
        if (ip >= iend) break;  // the last sequence has no match part
        if (iend - ip < 2) return false;
        const size_t offset = (size_t)ip[0] | ((size_t)ip[1] << 8);
        ip += 2;
        if (offset == 0 || offset > (size_t)(op - dst)) return false;
        size_t len = (token & 15u);
        if (len == 15) {
            unsigned b;
            do {
                if (ip >= iend) return false;
                b = *ip++;
                len += b;
            } while (b == 255);
        }
        len += 4;
        if ((size_t)(oend - op) < len) return false;
        const uint8_t* match = op - offset;
        for (size_t i = 0; i < len; ++i) op[i] = match[i];  // byte-wise: matches may overlap their own output
        op += len;
    }
    return op == oend;
}

}  // namespace

// ---- objects ----------------------------------------------------------------------------------------------------------------
struct SerializedResultObject {
    HostAllocator alloc;
    ommCpuBlobDesc desc{nullptr, 0};
};
struct DeserializedResultObject {
    HostAllocator alloc;
    Logger log;
    ommCpuDeserializedDesc desc{};
    std::vector<ommCpuBakeInputDesc> inputs;
    std::vector<ommCpuBakeResultDesc> results;
    std::vector<void*> owned;             // buffers allocated with `alloc`
    std::vector<TextureObject*> textures;
};

ommResult SerializeImpl(BakerObject* b, const ommCpuDeserializedDesc& d, SerializedResultObject** out) {
    if (d.numInputDescs < 0 || d.numResultDescs < 0 || (d.numInputDescs && !d.inputDescs) || (d.numResultDescs && !d.resultDescs)) return ommResult_INVALID_ARGUMENT;
    for (int i = 0; i < d.numInputDescs; ++i)
        if (d.inputDescs[i].texture == 0) return b->log.InvalidArg("[Invalid Argument] - ommCpuBakeInputDesc has no texture set");
    Writer w;
    w.put<uint64_t>(0);  // digest, patched below
    w.put<int32_t>(kLibMajor); w.put<int32_t>(kLibMinor); w.put<int32_t>(kLibPatch);
    w.put<int32_t>(kSerializeVersion);
    w.put<int32_t>((int32_t)d.flags);
    w.put<int32_t>(0);   // decompressedSize: the body is never compressed here (see the header of this file)
    w.put<int32_t>(d.numInputDescs);
    for (int i = 0; i < d.numInputDescs; ++i) WriteInput(w, d.inputDescs[i]);
    w.put<int32_t>(d.numResultDescs);
    for (int i = 0; i < d.numResultDescs; ++i) WriteResult(w, d.resultDescs[i]);
    const uint64_t digest = HostXxh64(w.bytes.data() + 8, w.bytes.size() - 8, 42);
    memcpy(w.bytes.data(), &digest, 8);
    SerializedResultObject* r = AllocObject<SerializedResultObject>(b->alloc);
    if (!r) return ommResult_FAILURE;
    r->alloc = b->alloc;
    r->desc.size = w.bytes.size();
    r->desc.data = b->alloc.alloc(w.bytes.size(), 16);
    if (!r->desc.data) {
        FreeObject(b->alloc, r);
        return ommResult_FAILURE;
    }
    memcpy(r->desc.data, w.bytes.data(), w.bytes.size());
    *out = r;
    return ommResult_SUCCESS;
}
const ommCpuBlobDesc* SerializedDesc(const SerializedResultObject* r) { return &r->desc; }
void DestroySerialized(SerializedResultObject* r) {
    const HostAllocator alloc = r->alloc;
    alloc.release(r->desc.data);
    FreeObject(alloc, r);
}

void DestroyDeserialized(DeserializedResultObject* r) {
    for (TextureObject* t : r->textures) {
        DestroyTextureDevice(t);
        r->alloc.release(t->hostTexels);
        FreeObject(r->alloc, t);
    }
    for (void* p : r->owned) r->alloc.release(p);
    const HostAllocator alloc = r->alloc;
    FreeObject(alloc, r);
}

static const void* ReadOwned(Reader& rd, DeserializedResultObject* r, size_t bytes) {
    if (bytes == 0) return nullptr;
    const uint8_t* src = rd.take(bytes);
    if (!src) return nullptr;
    void* p = r->alloc.alloc(bytes, 16);
    if (!p) {
        rd.ok = false;
        return nullptr;
    }
    memcpy(p, src, bytes);
    r->owned.push_back(p);
    return p;
}

static ommCpuBakeInputDesc DefaultInputDesc() {  // ref: omm.h:462-490
    ommCpuBakeInputDesc v;
    memset(&v, 0, sizeof(v));
    v.bakeFlags = ommCpuBakeFlags_None;
    v.runtimeSamplerDesc.addressingMode = ommTextureAddressMode_MAX_NUM;
    v.runtimeSamplerDesc.filter = ommTextureFilterMode_MAX_NUM;
    v.runtimeSamplerDesc.borderAlpha = 0;
    v.alphaMode = ommAlphaMode_MAX_NUM;
    v.texCoordFormat = ommTexCoordFormat_MAX_NUM;
    v.indexFormat = ommIndexFormat_MAX_NUM;
    v.dynamicSubdivisionScale = 2;
    v.rejectionThreshold = 0;
    v.alphaCutoff = 0.5f;
    v.nearDuplicateDeduplicationFactor = 0.15f;  // not part of the blob: the SDK's default survives a round trip
    v.alphaCutoffLessEqual = ommOpacityState_Transparent;
    v.alphaCutoffGreater = ommOpacityState_Opaque;
    v.format = ommFormat_OC1_4_State;
    v.unknownStatePromotion = ommUnknownStatePromotion_ForceOpaque;
    v.unresolvedTriState = ommSpecialIndex_FullyUnknownOpaque;
    v.maxSubdivisionLevel = 8;
    v.maxArrayDataSize = 0xFFFFFFFFu;
    v.maxWorkloadSize = 0xFFFFFFFFFFFFFFFFull;
    return v;
}

static ommResult DeserializeChecked(BakerObject* b, const ommCpuBlobDesc& blob, DeserializedResultObject*& r, DeserializedResultObject** out);
// The header digest is an integrity check, not an authentication: the contents are validated field by field, and an allocation
// failure while following them (std::bad_alloc from the containers) ends as FAILURE instead of crossing the C ABI.
ommResult DeserializeImpl(BakerObject* b, const ommCpuBlobDesc& blob, DeserializedResultObject** out) {
    DeserializedResultObject* r = nullptr;
    try {
        return DeserializeChecked(b, blob, r, out);
    } catch (...) {
        if (r) DestroyDeserialized(r);
        return ommResult_FAILURE;
    }
}
static ommResult DeserializeChecked(BakerObject* b, const ommCpuBlobDesc& blob, DeserializedResultObject*& r, DeserializedResultObject** out) {
    const Logger& log = b->log;
    if (blob.data == nullptr) return log.InvalidArg("data must be non-null");
    if (blob.size == 0) return log.InvalidArg("size must be non-zero");
    if (blob.size < 8 + 5 * 4) return ommResult_FAILURE;
    const uint8_t* bytes = (const uint8_t*)blob.data;
    const uint64_t digest = HostXxh64(bytes + 8, (size_t)blob.size - 8, 42);
    Reader hdr{bytes, bytes + blob.size};
    const uint64_t stored = hdr.get<uint64_t>();
    if (stored != digest) {
        log.Logf(ommMessageSeverity_Fatal, "The serialized blob appears corrupted, computed digest != header value %llu, %llu", (unsigned long long)digest,
                 (unsigned long long)stored);
        return ommResult_INVALID_ARGUMENT;
    }
    const int major = hdr.get<int32_t>(), minor = hdr.get<int32_t>(), patch = hdr.get<int32_t>(), version = hdr.get<int32_t>(), flags = hdr.get<int32_t>();
    int decompressedSize = 0;
    if (version >= 2) decompressedSize = hdr.get<int32_t>();
    if (!hdr.ok) return ommResult_FAILURE;
    if (version > kSerializeVersion) {
        log.Logf(ommMessageSeverity_Fatal, "The serialized blob appears to be generated from an incompatible version of the SDK (%d.%d.%d:%d)", major, minor, patch,
                 version);
        return ommResult_INVALID_ARGUMENT;
    }
    if (version < 1) return ommResult_FAILURE;
    std::vector<uint8_t> inflated;
    Reader rd{hdr.p, bytes + blob.size};
    if (decompressedSize != 0) {
        if (decompressedSize < 0) return ommResult_FAILURE;
        // an LZ4 block cannot expand by more than 255x (one length byte per 255 output bytes)
        if ((uint64_t)decompressedSize > 255ull * (uint64_t)(bytes + blob.size - hdr.p) + 64ull) return ommResult_FAILURE;
        inflated.resize((size_t)decompressedSize);
        if (!Lz4Decode(hdr.p, (size_t)(bytes + blob.size - hdr.p), inflated.data(), inflated.size())) return ommResult_FAILURE;
        rd = Reader{inflated.data(), inflated.data() + inflated.size()};
    }
    r = AllocObject<DeserializedResultObject>(b->alloc);
    if (!r) return ommResult_FAILURE;
    r->alloc = b->alloc;
    r->log = b->log;
    ommResult rc = ommResult_SUCCESS;
    const int numInputs = rd.get<int32_t>();
    if (!rd.ok || numInputs < 0) rc = ommResult_FAILURE;
    for (int i = 0; rc == ommResult_SUCCESS && i < numInputs; ++i) {
        ommCpuBakeInputDesc d = DefaultInputDesc();
        d.bakeFlags = (ommCpuBakeFlags)rd.get<int32_t>();
        TextureObject* t = AllocObject<TextureObject>(b->alloc);
        if (!t) { rc = ommResult_FAILURE; break; }
        t->alloc = b->alloc;
        t->device = b->device;
        r->textures.push_back(t);
        if (!ReadTexture(rd, version, b->alloc, t)) { rc = ommResult_FAILURE; break; }
        d.texture = MakeHandle<ommCpuTexture>(t, HandleTag::Texture);
        d.runtimeSamplerDesc.addressingMode = (ommTextureAddressMode)rd.get<int32_t>();
        d.runtimeSamplerDesc.filter = (ommTextureFilterMode)rd.get<int32_t>();
        d.runtimeSamplerDesc.borderAlpha = rd.get<float>();
        d.alphaMode = (ommAlphaMode)rd.get<int32_t>();
        d.texCoordFormat = (ommTexCoordFormat)rd.get<int32_t>();
        const uint64_t texCoordBytes = rd.get<uint64_t>();
        d.texCoords = ReadOwned(rd, r, (size_t)texCoordBytes);
        d.texCoordStrideInBytes = rd.get<uint32_t>();
        d.indexFormat = (ommIndexFormat)rd.get<int32_t>();
        d.indexCount = rd.get<uint32_t>();
        d.indexBuffer = ReadOwned(rd, r, (size_t)d.indexCount * IndexSize(d.indexFormat));
        d.dynamicSubdivisionScale = rd.get<float>();
        d.rejectionThreshold = rd.get<float>();
        d.alphaCutoff = rd.get<float>();
        d.alphaCutoffLessEqual = (ommOpacityState)rd.get<int32_t>();
        d.alphaCutoffGreater = (ommOpacityState)rd.get<int32_t>();
        d.format = (ommFormat)rd.get<int32_t>();
        const uint64_t numFormats = rd.get<uint64_t>();
        d.formats = (const ommFormat*)ReadOwned(rd, r, (size_t)numFormats * 4);
        d.unknownStatePromotion = (ommUnknownStatePromotion)rd.get<int32_t>();
        if (version >= 2) d.unresolvedTriState = (ommSpecialIndex)rd.get<int32_t>();
        d.maxSubdivisionLevel = rd.get<uint8_t>();
        if (version >= 4) d.maxArrayDataSize = rd.get<uint32_t>();
        const uint64_t numLevels = rd.get<uint64_t>();
        d.subdivisionLevels = (const uint8_t*)ReadOwned(rd, r, (size_t)numLevels);
        d.maxWorkloadSize = rd.get<uint64_t>();
        if (!rd.ok) { rc = ommResult_FAILURE; break; }
        // A bake of this desc reads texCoords[index], formats[t] and subdivisionLevels[t] for every triangle: the arrays the blob
        // carries must cover what its own index buffer references.
        {
            const uint32_t isz = IndexSize(d.indexFormat);
            if ((uint32_t)d.indexFormat >= (uint32_t)ommIndexFormat_MAX_NUM || (uint32_t)d.texCoordFormat >= (uint32_t)ommTexCoordFormat_MAX_NUM) { rc = ommResult_FAILURE; break; }
            uint64_t maxIndex = 0;
            const uint8_t* ib = (const uint8_t*)d.indexBuffer;
            for (uint32_t k = 0; k < d.indexCount; ++k) {
                uint32_t v = 0;
                memcpy(&v, ib + (size_t)k * isz, isz);
                maxIndex = v > maxIndex ? v : maxIndex;
            }
            const uint64_t uvSize = d.texCoordFormat == ommTexCoordFormat_UV32_FLOAT ? 8 : 4;
            const uint64_t stride = d.texCoordStrideInBytes ? d.texCoordStrideInBytes : uvSize;
            const uint64_t tris = d.indexCount / 3;
            if (d.indexCount && texCoordBytes < maxIndex * stride + uvSize) { rc = ommResult_FAILURE; break; }
            if (numFormats && numFormats < tris) { rc = ommResult_FAILURE; break; }
            if (numLevels && numLevels < tris) { rc = ommResult_FAILURE; break; }
        }
        // ref: serialize_impl.cpp:463-468 -- blobs older than v3 did not store the texture's alpha cutoff although they stored its SAT
        if (t->hasSerializedSat && version < 3) t->alphaCutoff = d.alphaCutoff;
        rc = UploadTexture(t, log);
        if (rc != ommResult_SUCCESS) {
            // UploadTexture released the device side; the host side is released with the object below
            break;
        }
        r->inputs.push_back(d);
    }
    if (rc == ommResult_SUCCESS) {
        const int numResults = rd.get<int32_t>();
        if (!rd.ok || numResults < 0) rc = ommResult_FAILURE;
        for (int i = 0; rc == ommResult_SUCCESS && i < numResults; ++i) {
            ommCpuBakeResultDesc o;
            memset(&o, 0, sizeof(o));
            o.arrayDataSize = rd.get<uint32_t>();
            o.arrayData = ReadOwned(rd, r, o.arrayDataSize);
            o.descArrayCount = rd.get<uint32_t>();
            o.descArray = (const ommCpuOpacityMicromapDesc*)ReadOwned(rd, r, sizeof(ommCpuOpacityMicromapDesc) * (size_t)o.descArrayCount);
            o.descArrayHistogramCount = rd.get<uint32_t>();
            o.descArrayHistogram = (const ommCpuOpacityMicromapUsageCount*)ReadOwned(rd, r, sizeof(ommCpuOpacityMicromapUsageCount) * (size_t)o.descArrayHistogramCount);
            o.indexFormat = (ommIndexFormat)rd.get<int32_t>();
            o.indexCount = rd.get<uint32_t>();
            o.indexBuffer = ReadOwned(rd, r, IndexSize(o.indexFormat) * (size_t)o.indexCount);
            o.indexHistogramCount = rd.get<uint32_t>();
            o.indexHistogram = (const ommCpuOpacityMicromapUsageCount*)ReadOwned(rd, r, sizeof(ommCpuOpacityMicromapUsageCount) * (size_t)o.indexHistogramCount);
            if (!rd.ok) { rc = ommResult_FAILURE; break; }
            r->results.push_back(o);
        }
    }
    if (rc != ommResult_SUCCESS) {
        DestroyDeserialized(r);
        r = nullptr;
        return rc;
    }
    r->desc.flags = (ommCpuSerializeFlags)flags;
    r->desc.numInputDescs = (int)r->inputs.size();
    r->desc.inputDescs = r->inputs.empty() ? nullptr : r->inputs.data();
    r->desc.numResultDescs = (int)r->results.size();
    r->desc.resultDescs = r->results.empty() ? nullptr : r->results.data();
    *out = r;
    return ommResult_SUCCESS;
}
const ommCpuDeserializedDesc* DeserializedDesc(const DeserializedResultObject* r) { return &r->desc; }

}  // namespace ommb200
