// omm_api.cpp -- the C ABI of libomm-b200.so (include/omm_b200.h): argument checks, handles, allocator and message
// plumbing with the SDK's observable behaviour (return codes, message texts, ownership), and the glue that turns one
// ommCpuBake call into stage -> device pipeline -> download.
//
// Reference behaviour restated here: libraries/omm-lib/src/bake.cpp:36-135, 410-479 (entry points),
// bake_cpu_impl.cpp:97-119, 235-290 (validation + messages), texture_impl.cpp:43-224 (texture validation and copy),
// std_allocator.h:45-117 (default allocator), omm_handle.h:17-53 (handle tags), debug_impl.cpp:512-641 (stats).
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <map>

#include "omm_internal.h"

using namespace ommb200;

namespace ommb200 {

// ---- default allocator: aligned malloc with the original pointer stored in front (ref: std_allocator.h:45-94) ----
// Two words sit in front of every block: the malloc pointer and the requested size (the latter makes reallocate a real one).
static void* DefaultAllocate(void*, size_t size, size_t alignment) {
    if (alignment < sizeof(void*)) alignment = sizeof(void*);
    const size_t header = 2 * sizeof(void*);
    uint8_t* raw = (uint8_t*)std::malloc(size + header + alignment - 1);
    if (!raw) return nullptr;
    uint8_t* aligned = (uint8_t*)(((uintptr_t)(raw + header) + alignment - 1) & ~(uintptr_t)(alignment - 1));
    ((void**)aligned)[-1] = raw;
    ((size_t*)aligned)[-2] = size;
    return aligned;
}
static void DefaultFree(void*, void* memory) {
    if (memory) std::free(((void**)memory)[-1]);
}
static void* DefaultReallocate(void* user, void* memory, size_t size, size_t alignment) {  // ref: std_allocator.h:96-117 (realloc semantics)
    if (!memory) return DefaultAllocate(user, size, alignment);
    if (size == 0) {
        DefaultFree(user, memory);
        return nullptr;
    }
    void* fresh = DefaultAllocate(user, size, alignment);
    if (!fresh) return nullptr;  // the old block stays valid, like realloc
    const size_t old = ((size_t*)memory)[-2];
    memcpy(fresh, memory, old < size ? old : size);
    DefaultFree(user, memory);
    return fresh;
}
void SetDefaultAllocatorIfUnset(ommMemoryAllocatorInterface& iface) {
    if (iface.allocate != nullptr) return;
    iface.allocate = DefaultAllocate;
    iface.reallocate = DefaultReallocate;
    iface.free = DefaultFree;
}

static thread_local int g_requestedDevice = -1;

static const char* OpacityStateName(ommOpacityState s) {  // ref: util/util.h:41-55
    switch (s) {
    case ommOpacityState_Transparent: return "Transparent";
    case ommOpacityState_Opaque: return "Opaque";
    case ommOpacityState_UnknownTransparent: return "UnknownTransparent";
    case ommOpacityState_UnknownOpaque: return "UnknownOpaque";
    default: return "Unknown";
    }
}
static const char* FormatName(ommFormat f) {  // ref: util/util.h:57-68
    switch (f) {
    case ommFormat_OC1_2_State: return "OC1_2_State";
    case ommFormat_OC1_4_State: return "OC1_4_State";
    default: return "Unknown";
    }
}
static bool StateCompatible(ommOpacityState s, ommFormat f) {  // ref: util/util.h:27-34
    if (f == ommFormat_OC1_2_State) return s == ommOpacityState_Opaque || s == ommOpacityState_Transparent;
    return true;
}

// ref: bake_cpu_impl.cpp:235-290 -- same order, same texts
static ommResult ValidateBakeDesc(const Logger& log, const ommCpuBakeInputDesc& d) {
    const uint32_t flags = (uint32_t)d.bakeFlags;
    const bool nearDup = (flags & ommCpuBakeFlags_EnableNearDuplicateDetection) != 0, nearDupBrute = (flags & (1u << 10)) != 0;
    if (d.texture == 0) return log.InvalidArg("[Invalid Argument] - texture is not set");
    if (HandleTagOf(d.texture) != HandleTag::Texture) return log.InvalidArg("[Invalid Argument] - desc.texture is of incorrect type");
    if (d.alphaMode == ommAlphaMode_MAX_NUM) return log.InvalidArg("[Invalid Argument] - alphaMode is not set");
    if (d.runtimeSamplerDesc.addressingMode == ommTextureAddressMode_MAX_NUM)
        return log.InvalidArg("[Invalid Argument] - runtimeSamplerDesc.addressingMode is not set");
    if (d.runtimeSamplerDesc.filter == ommTextureFilterMode_MAX_NUM) return log.InvalidArg("[Invalid Argument] - runtimeSamplerDesc.filter is not set");
    if (d.texCoordFormat == ommTexCoordFormat_MAX_NUM) return log.InvalidArg("[Invalid Argument] - texCoordFormat is not set");
    if (d.texCoords == nullptr) return log.InvalidArg("[Invalid Argument] - texCoords is not set");
    if (d.indexFormat == ommIndexFormat_MAX_NUM) return log.InvalidArg("[Invalid Argument] - indexFormat is not set");
    if (d.indexBuffer == nullptr) return log.InvalidArg("[Invalid Argument] - indexBuffer is not set");
    if (d.indexCount == 0) return log.InvalidArg("[Invalid Argument] - indexCount is not set");
    if (d.maxSubdivisionLevel > 12) {
        log.Logf(ommMessageSeverity_Fatal, "[Invalid Argument] - maxSubdivisionLevel (%d) is greater than maximum supported (%d)", d.maxSubdivisionLevel, 12);
        return ommResult_INVALID_ARGUMENT;
    }
    if ((nearDup || nearDupBrute) && (flags & ommCpuBakeFlags_DisableDuplicateDetection))
        return log.InvalidArg(
            "[Invalid Argument] - EnableNearDuplicateDetection or EnableNearDuplicateDetectionBruteForce is used together with DisableDuplicateDetection");
    if ((flags & ommCpuBakeFlags_EnableValidation) && !log.HasLogger())
        return log.InvalidArg("[Invalid Argument] - EnableValidation is set but no message callback was provided");
    const TextureObject* tex = HandlePtr<TextureObject>(d.texture);
    if (tex->HasAlphaCutoff() && tex->alphaCutoff != d.alphaCutoff) {
        log.Logf(ommMessageSeverity_Fatal,
                 "[Invalid Argument] - Texture object alpha cutoff threshold (%.6f) is different from alpha cutoff threshold in bake input (%.6f)",
                 tex->alphaCutoff, d.alphaCutoff);
        return ommResult_INVALID_ARGUMENT;
    }
    if (!StateCompatible(d.alphaCutoffGreater, d.format)) {
        log.Logf(ommMessageSeverity_Fatal, "[Invalid Argument] - alphaCutoffGreater=%s is not compatible with %s", OpacityStateName(d.alphaCutoffGreater),
                 FormatName(d.format));
        return ommResult_INVALID_ARGUMENT;
    }
    if (!StateCompatible(d.alphaCutoffLessEqual, d.format)) {
        log.Logf(ommMessageSeverity_Fatal, "[Invalid Argument] - alphaCutoffLessEqual=%s is not compatible with %s", OpacityStateName(d.alphaCutoffLessEqual),
                 FormatName(d.format));
        return ommResult_INVALID_ARGUMENT;
    }
    return ommResult_SUCCESS;
}

// Results outlive nothing in the SDK's contract except their own handle, so a result may be asked for its host copy (or destroyed) after
// its baker is gone: live bakers are looked up here, never dereferenced blindly.
static std::mutex g_liveBakersMu;
static std::vector<BakerObject*> g_liveBakers;

static void DestroyResult(BakeResultObject* r) {
    if (!r) return;
    DestroyResultDevice(r);
    const HostAllocator alloc = r->alloc;
    if (r->sharedWindowId >= 0) {
        // the array lives in a shared window of the baker's sharding (unmapped with the sharding if the baker went first)
        std::lock_guard<std::mutex> live(g_liveBakersMu);
        if (std::find(g_liveBakers.begin(), g_liveBakers.end(), r->baker) != g_liveBakers.end()) ReleaseSharedWindow(r->baker, r->sharedWindowId);
    } else if (r->arrayDataFromPinnedPool) PinnedPoolRelease(r->hostArrayData);
    else alloc.release(r->hostArrayData);
    if (r->descFromPinnedPool) PinnedPoolRelease(r->hostDescArray);
    else alloc.release(r->hostDescArray);
    if (r->indexFromPinnedPool) PinnedPoolRelease(r->hostIndexBuffer);
    else alloc.release(r->hostIndexBuffer);
    FreeObject(alloc, r);
}

}  // namespace ommb200

// ======================================================================================================================
// SDK entry points
// ======================================================================================================================
OMM_API ommLibraryDesc ommGetLibraryDesc(void) {
    ommLibraryDesc d = {OMM_VERSION_MAJOR, OMM_VERSION_MINOR, OMM_VERSION_BUILD};
    return d;
}

// the timings of a deferred download of a result whose baker is gone are simply not recorded
static void RecordDeferredDownload(BakerObject* baker, float d2hMs, uint64_t d2hBytes, float hostMs) {
    std::lock_guard<std::mutex> live(g_liveBakersMu);
    if (std::find(g_liveBakers.begin(), g_liveBakers.end(), baker) == g_liveBakers.end()) return;
    std::lock_guard<std::mutex> g(baker->mu);
    baker->last.d2hMs = d2hMs;
    baker->last.d2hBytes = d2hBytes;
    if (hostMs > 0.f) {
        baker->last.hostDownloadMs = hostMs;
        baker->last.hostTotalMs += hostMs;
    }
}

OMM_API ommResult ommCreateBaker(const ommBakerCreationDesc* desc, ommBaker* outBaker) {  // ref: bake.cpp:410-455
    if (desc == nullptr) return ommResult_INVALID_ARGUMENT;
    if (desc->type == ommBakerType_GPU) return ommResult_NOT_IMPLEMENTED;  // the SDK's D3D12/VK command-list baker is out of scope
    if (desc->type != ommBakerType_CPU) return ommResult_INVALID_ARGUMENT;
    HostAllocator alloc;
    alloc.iface = desc->memoryAllocatorInterface;
    const bool defaultAllocator = alloc.iface.allocate == nullptr;
    SetDefaultAllocatorIfUnset(alloc.iface);
    BakerObject* b = AllocObject<BakerObject>(alloc);
    if (!b) return ommResult_FAILURE;
    b->alloc = alloc;
    b->usesDefaultAllocator = defaultAllocator;
    b->log.sink = desc->messageInterface;
    b->device = g_requestedDevice >= 0 ? g_requestedDevice : CurrentDeviceOr(0);
    {
        std::lock_guard<std::mutex> live(g_liveBakersMu);
        g_liveBakers.push_back(b);
    }
    *outBaker = MakeHandle<ommBaker>(b, HandleTag::CpuBaker);
    return ommResult_SUCCESS;
}

OMM_API ommResult ommDestroyBaker(ommBaker baker) {  // ref: bake.cpp:457-479
    if (baker == 0) return ommResult_INVALID_ARGUMENT;
    if (HandleTagOf(baker) != HandleTag::CpuBaker) return ommResult_FAILURE;
    BakerObject* b = HandlePtr<BakerObject>(baker);
    bool lastBaker = false;
    {
        std::lock_guard<std::mutex> live(g_liveBakersMu);
        g_liveBakers.erase(std::remove(g_liveBakers.begin(), g_liveBakers.end(), b), g_liveBakers.end());
        lastBaker = g_liveBakers.empty();
    }
    DestroySharding(b);
    const HostAllocator alloc = b->alloc;
    FreeObject(alloc, b);
    if (lastBaker) PinnedPoolTrim(0);  // page-locked memory is a system resource: nothing stays cached once the last baker is gone
    return ommResult_SUCCESS;
}

OMM_API ommResult ommCpuCreateTexture(ommBaker baker, const ommCpuTextureDesc* desc, ommCpuTexture* outTexture) {  // ref: bake.cpp:44-69
    if (baker == 0) return ommResult_INVALID_ARGUMENT;
    BakerObject* b = HandlePtr<BakerObject>(baker);
    if (desc == 0) return b->log.InvalidArg("texture desc was not set");
    if (HandleTagOf(baker) != HandleTag::CpuBaker) return b->log.InvalidArg("Baker was not created as the right type");
    const Logger& log = b->log;
    // ref: texture_impl.cpp:43-64
    if (desc->mipCount == 0) return log.InvalidArg("[Invalid Arg] - mipCount must be non-zero");
    if (desc->format == ommCpuTextureFormat_MAX_NUM) return log.InvalidArg("[Invalid Arg] - format is not set");
    for (uint32_t i = 0; i < desc->mipCount; ++i) {
        if (!desc->mips[i].textureData) return log.InvalidArg("[Invalid Arg] - mips.textureData is not set");
        if (desc->mips[i].width == 0) return log.InvalidArg("[Invalid Arg] - mips.width must be non-zero");
        if (desc->mips[i].height == 0) return log.InvalidArg("[Invalid Arg] - mips.height must be non-zero");
        if (desc->mips[i].width > 65536) return log.InvalidArg("[Invalid Arg] - mips.width must be less than kMaxDim.x (65536)");
        if (desc->mips[i].height > 65536) return log.InvalidArg("[Invalid Arg] - mips.height must be less than kMaxDim.y (65536)");
    }
    if (desc->mipCount > (uint32_t)kMaxMips) return log.InvalidArg("[Invalid Arg] - more mips than a 65536-texel texture can have");

    TextureObject* t = AllocObject<TextureObject>(b->alloc);
    if (!t) return ommResult_FAILURE;
    t->alloc = b->alloc;
    t->format = desc->format;
    t->flags = desc->flags;
    t->alphaCutoff = desc->alphaCutoff;
    t->mipCount = desc->mipCount;
    t->device = b->device;
    const size_t spp = desc->format == ommCpuTextureFormat_FP32 ? 4 : 1;
    const bool linear = ((uint32_t)desc->flags & (uint32_t)ommCpuTextureFlags_DisableZOrder) != 0;
    size_t totalTexels = 0;
    for (uint32_t i = 0; i < desc->mipCount; ++i) {
        DevMip& m = t->dev.mips[i];
        m.w = (int)desc->mips[i].width;
        m.h = (int)desc->mips[i].height;
        int lw = 0, lh = 0;
        for (uint32_t v = (uint32_t)m.w; (v & 1u) == 0; v >>= 1) lw++;  // ctz; sizes are non-zero
        for (uint32_t v = (uint32_t)m.h; (v & 1u) == 0; v >>= 1) lh++;
        m.log2w = lw;
        m.log2h = lh;
        m.isPow2 = (m.w & (m.w - 1)) == 0 && (m.h & (m.h - 1)) == 0;
        m.rcpw = 1.f / (float)m.w;
        m.rcph = 1.f / (float)m.h;
        m.texelOffset = totalTexels;
        m.satOffset = totalTexels;
        totalTexels += (size_t)m.w * m.h;
    }
    t->dev.mipCount = (int)desc->mipCount;
    t->hostBytes = totalTexels * spp;
    t->hostTexels = b->alloc.alloc(t->hostBytes, 64);
    if (!t->hostTexels) {
        FreeObject(b->alloc, t);
        return ommResult_FAILURE;
    }
    for (uint32_t i = 0; i < desc->mipCount; ++i) {
        const ommCpuTextureMipDesc& s = desc->mips[i];
        const DevMip& m = t->dev.mips[i];
        // ref: texture_impl.cpp:137-184 -- the linear layout reads rowPitch in BYTES, the (default) Z-order layout in TEXELS
        const size_t rowBytes = linear ? (s.rowPitch == 0 ? spp * s.width : (size_t)s.rowPitch) : spp * (s.rowPitch == 0 ? (size_t)s.width : (size_t)s.rowPitch);
        uint8_t* dst = (uint8_t*)t->hostTexels + m.texelOffset * spp;
        const uint8_t* src = (const uint8_t*)s.textureData;
        for (int y = 0; y < m.h; ++y) memcpy(dst + spp * (size_t)m.w * y, src + rowBytes * (size_t)y, spp * (size_t)m.w);
    }
    const ommResult rc = UploadTexture(t, log);
    if (rc != ommResult_SUCCESS) {
        b->alloc.release(t->hostTexels);
        FreeObject(b->alloc, t);
        return rc;
    }
    *outTexture = MakeHandle<ommCpuTexture>(t, HandleTag::Texture);
    return ommResult_SUCCESS;
}

OMM_API ommResult ommCpuGetTextureDesc(ommCpuTexture texture, ommCpuTextureDesc* outDesc) {  // ref: bake.cpp:71-82, texture_impl.cpp:280-325
    if (texture == 0) return ommResult_INVALID_ARGUMENT;
    TextureObject* t = HandlePtr<TextureObject>(texture);
    if (t == nullptr || outDesc == nullptr) return ommResult_INVALID_ARGUMENT;
    outDesc->format = t->format;
    outDesc->flags = t->flags;
    outDesc->alphaCutoff = t->alphaCutoff;
    outDesc->mipCount = t->mipCount;
    if (outDesc->mips == nullptr) return ommResult_SUCCESS;
    const size_t spp = t->format == ommCpuTextureFormat_FP32 ? 4 : 1;
    for (uint32_t i = 0; i < t->mipCount; ++i) {
        ommCpuTextureMipDesc& m = const_cast<ommCpuTextureMipDesc&>(outDesc->mips[i]);
        m.width = (uint32_t)t->dev.mips[i].w;
        m.height = (uint32_t)t->dev.mips[i].h;
        m.rowPitch = (uint32_t)t->dev.mips[i].w;
        if (m.textureData != nullptr)
            memcpy(const_cast<void*>(m.textureData), (const uint8_t*)t->hostTexels + t->dev.mips[i].texelOffset * spp, spp * (size_t)m.width * m.height);
    }
    return ommResult_SUCCESS;
}

OMM_API ommResult ommCpuDestroyTexture(ommBaker baker, ommCpuTexture texture) {  // ref: bake.cpp:84-101
    if (texture == 0) return ommResult_INVALID_ARGUMENT;
    BakerObject* b = HandlePtr<BakerObject>(baker);
    if (HandleTagOf(baker) != HandleTag::CpuBaker) return b ? b->log.InvalidArg("Baker was not created as the right type") : ommResult_INVALID_ARGUMENT;
    TextureObject* t = HandlePtr<TextureObject>(texture);
    DestroyTextureDevice(t);
    const HostAllocator alloc = t->alloc;
    alloc.release(t->hostTexels);
    FreeObject(alloc, t);
    return ommResult_SUCCESS;
}

// ---- serialization (ref: bake.cpp:137-252) ---------------------------------------------------------------------------------------
OMM_API ommResult ommCpuSerialize(ommBaker baker, const ommCpuDeserializedDesc* desc, ommCpuSerializedResult* outResult) {
    if (baker == 0 || outResult == nullptr) return ommResult_INVALID_ARGUMENT;
    BakerObject* b = HandlePtr<BakerObject>(baker);
    if (HandleTagOf(baker) != HandleTag::CpuBaker) return b->log.InvalidArg("Baker was not created as the right type");
    if (desc == nullptr) return ommResult_INVALID_ARGUMENT;
    SerializedResultObject* r = nullptr;
    const ommResult rc = SerializeImpl(b, *desc, &r);
    *outResult = rc == ommResult_SUCCESS ? MakeHandle<ommCpuSerializedResult>(r, HandleTag::SerializeResult) : (ommCpuSerializedResult) nullptr;
    return rc;
}
OMM_API ommResult ommCpuGetSerializedResultDesc(ommCpuSerializedResult result, const ommCpuBlobDesc** desc) {
    if (result == 0 || desc == nullptr) return ommResult_INVALID_ARGUMENT;
    *desc = SerializedDesc(HandlePtr<SerializedResultObject>(result));
    return ommResult_SUCCESS;
}
OMM_API ommResult ommCpuDestroySerializedResult(ommCpuSerializedResult result) {
    if (result == 0 || HandleTagOf(result) != HandleTag::SerializeResult) return ommResult_INVALID_ARGUMENT;
    DestroySerialized(HandlePtr<SerializedResultObject>(result));
    return ommResult_SUCCESS;
}
OMM_API ommResult ommCpuDeserialize(ommBaker baker, const ommCpuBlobDesc* desc, ommCpuDeserializedResult* outResult) {
    if (baker == 0) return ommResult_INVALID_ARGUMENT;
    BakerObject* b = HandlePtr<BakerObject>(baker);
    if (HandleTagOf(baker) != HandleTag::CpuBaker) return b->log.InvalidArg("Baker was not created as the right type");
    if (desc == nullptr || outResult == nullptr) return ommResult_INVALID_ARGUMENT;
    DeserializedResultObject* r = nullptr;
    const ommResult rc = DeserializeImpl(b, *desc, &r);
    *outResult = rc == ommResult_SUCCESS ? MakeHandle<ommCpuDeserializedResult>(r, HandleTag::DeserializeResult) : (ommCpuDeserializedResult) nullptr;
    return rc;
}
OMM_API ommResult ommCpuGetDeserializedDesc(ommCpuDeserializedResult result, const ommCpuDeserializedDesc** desc) {
    if (result == 0 || HandleTagOf(result) != HandleTag::DeserializeResult || desc == nullptr) return ommResult_INVALID_ARGUMENT;
    *desc = DeserializedDesc(HandlePtr<DeserializedResultObject>(result));
    return ommResult_SUCCESS;
}
OMM_API ommResult ommCpuDestroyDeserializedResult(ommCpuDeserializedResult result) {
    if (result == 0 || HandleTagOf(result) != HandleTag::DeserializeResult) return ommResult_INVALID_ARGUMENT;
    DestroyDeserialized(HandlePtr<DeserializedResultObject>(result));
    return ommResult_SUCCESS;
}

// ---- entry points of the SDK that are outside this library's scope: exported so that programs written against
// omm.h -- in particular the SDK's own test binary, oracle/Makefile target `reftests` -- link; every one reports NOT_IMPLEMENTED.
// The D3D12 / Vulkan command-list baker (ommGpu*), image dumps and ommDebugGetStats2 (stats from a result handle).
OMM_API ommResult ommGpuGetStaticResourceData(int, uint8_t*, size_t*) { return ommResult_NOT_IMPLEMENTED; }
OMM_API ommResult ommGpuCreatePipeline(ommBaker, const void*, void**) { return ommResult_NOT_IMPLEMENTED; }
OMM_API ommResult ommGpuDestroyPipeline(ommBaker, void*) { return ommResult_NOT_IMPLEMENTED; }
OMM_API ommResult ommGpuGetPipelineDesc(void*, const void**) { return ommResult_NOT_IMPLEMENTED; }
OMM_API ommResult ommGpuGetPreDispatchInfo(void*, const void*, void*) { return ommResult_NOT_IMPLEMENTED; }
OMM_API ommResult ommGpuDispatch(void*, const void*, const void**) { return ommResult_NOT_IMPLEMENTED; }
OMM_API ommResult ommDebugSaveAsImages(ommBaker, const ommCpuBakeInputDesc*, const ommCpuBakeResultDesc*, const void*) { return ommResult_NOT_IMPLEMENTED; }
OMM_API ommResult ommDebugSaveBinaryToDisk(ommBaker, const ommCpuBlobDesc*, const char*) { return ommResult_NOT_IMPLEMENTED; }

static ommResult CheckBakeArgs(ommBaker baker, const ommCpuBakeInputDesc* d, BakerObject** outBaker) {
    if (baker == 0) return ommResult_INVALID_ARGUMENT;
    BakerObject* b = HandlePtr<BakerObject>(baker);
    if (d == 0) return b->log.InvalidArg("input desc was not set");  // ref: bake.cpp:110-113
    if (HandleTagOf(baker) != HandleTag::CpuBaker) return b->log.InvalidArg("Baker was not created as the right type");
    if (d->texture == 0) return b->log.InvalidArg("[Invalid Argument] - ommCpuBakeInputDesc has no texture set");  // ref: bake_cpu_impl.cpp:97-103
    // ref: bake_cpu_impl.cpp:297-303 -- the SDK looks its kernel up by (format, tiling, addressing mode, filter, pow2) BEFORE it
    // validates the desc, so an out-of-range addressing mode or filter yields FAILURE without a message.
    if ((uint32_t)d->runtimeSamplerDesc.addressingMode >= (uint32_t)ommTextureAddressMode_MAX_NUM ||
        (uint32_t)d->runtimeSamplerDesc.filter >= (uint32_t)ommTextureFilterMode_MAX_NUM)
        return ommResult_FAILURE;
    const ommResult v = ValidateBakeDesc(b->log, *d);
    if (v != ommResult_SUCCESS) return v;
    // The SDK only asserts on the format (bake_cpu_impl.cpp:328-333, 365: compiled out in its release build, then it indexes its
    // histograms with format - 1); here anything but the two OC1 formats is refused before any work is done.
    if (d->format != ommFormat_OC1_2_State && d->format != ommFormat_OC1_4_State) return b->log.InvalidArg("[Invalid Argument] - format is not set");
    const uint32_t flags = (uint32_t)d->bakeFlags;
    if ((flags & (1u << 7)) && !(flags & (1u << 8)))  // ref: bake_cpu_impl.cpp:718-719
        return b->log.InvalidArg("[Invalid Arg] - EnableAABBTesting can't be used without also setting DisableLevelLineIntersection");
    *outBaker = b;
    return ommResult_SUCCESS;
}

// download: make the host copy before returning; early: the caller will want the host copy (ommCpuBake), so the device pipeline may
// start sending the array while it is still packing (and, sharded, every rank sends its own shards to the root's host memory)
static ommResult RunBake(BakerObject* b, const StagedInputs& staged, void* stream, bool download, bool early, float stageMs, ommCpuBakeResult* out) {
    BakeResultObject* r = AllocObject<BakeResultObject>(b->alloc);
    if (!r) return ommResult_FAILURE;
    r->alloc = b->alloc;
    r->usesDefaultAllocator = b->usesDefaultAllocator;
    r->log = b->log;
    r->baker = b;
    ommB200BakeTimings tm{};
    tm.h2dMs = staged.h2dMs;
    tm.h2dBytes = staged.h2dBytes;
    tm.hostStageMs = stageMs;
    const auto t0 = std::chrono::steady_clock::now();
    HostTrace::Mark("result object");
    ommResult rc = BakeOnDevice(b, staged, stream, r, &tm, early);
    HostTrace::Mark("BakeOnDevice returned");
    const auto t1 = std::chrono::steady_clock::now();
    if (rc == ommResult_SUCCESS && download) rc = DownloadResult(r, &tm.d2hMs, &tm.d2hBytes);
    const auto t2 = std::chrono::steady_clock::now();
    tm.hostBakeMs = std::chrono::duration<float, std::milli>(t1 - t0).count();
    tm.hostDownloadMs = std::chrono::duration<float, std::milli>(t2 - t1).count();
    tm.hostTotalMs = stageMs + tm.hostBakeMs + tm.hostDownloadMs;
    if (rc != ommResult_SUCCESS) {
        DestroyResult(r);
        return rc;
    }
    {
        std::lock_guard<std::mutex> g(b->mu);
        b->last = tm;
        b->haveTimings = true;
    }
    *out = (ommCpuBakeResult)r;  // raw pointer, like the SDK (ref: bake_cpu_impl.cpp:113)
    return ommResult_SUCCESS;
}

OMM_API ommResult ommCpuBake(ommBaker baker, const ommCpuBakeInputDesc* d, ommCpuBakeResult* outBakeResult) {  // ref: bake.cpp:103-116
    BakerObject* b = nullptr;
    HostTrace::Mark("ommCpuBake entry");
    const ommResult v = CheckBakeArgs(baker, d, &b);
    if (v != ommResult_SUCCESS) return v;
    HostTrace::Mark("args checked");
    StagedInputs staged;
    const auto t0 = std::chrono::steady_clock::now();
    ommResult rc = StageInputs(b, *d, &staged);
    if (rc != ommResult_SUCCESS) return rc;
    const float stageMs = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    HostTrace::Mark("inputs staged");
    // Sharded bakes leave the complete result in every rank's HBM; the host copy is made by ommCpuGetBakeResultDesc on the ranks
    // that ask for it (N simultaneous downloads through one host were measured at a quarter of the single-download speed).
    rc = RunBake(b, staged, nullptr, b->shard.world <= 1, true, stageMs, outBakeResult);
    HostTrace::Mark("bake + download");
    DestroyStagedDevice(&staged);
    HostTrace::Mark("staged inputs freed");
    HostTrace::Dump();
    return rc;
}

OMM_API ommResult ommCpuDestroyBakeResult(ommCpuBakeResult bakeResult) {  // ref: bake.cpp:118-127
    if (bakeResult == 0) return ommResult_INVALID_ARGUMENT;
    DestroyResult((BakeResultObject*)bakeResult);
    return ommResult_SUCCESS;
}

OMM_API ommResult ommCpuGetBakeResultDesc(ommCpuBakeResult bakeResult, const ommCpuBakeResultDesc** desc) {  // ref: bake.cpp:129-135
    if (bakeResult == 0) return ommResult_INVALID_ARGUMENT;
    BakeResultObject* r = (BakeResultObject*)bakeResult;
    if (desc == nullptr) return r->log.InvalidArg("[Invalid Arg] - No BakeResultDesc provided");  // ref: bake_cpu_impl.h:113-120
    if (!r->arrayOnThisRank) return r->log.InvalidArg("[omm-b200] the complete result of this sharded bake lives on rank 0 (ommB200ShardedResultMode_OnRank0)");
    if (!r->downloaded) {
        float d2hMs = 0.f;
        uint64_t d2hBytes = 0;
        const auto t0 = std::chrono::steady_clock::now();
        const ommResult rc = DownloadResult(r, &d2hMs, &d2hBytes);
        if (rc != ommResult_SUCCESS) return rc;
        const float hostMs = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (r->baker) RecordDeferredDownload(r->baker, d2hMs, d2hBytes, hostMs);  // it belongs to the timings of the bake that produced it
    }
    *desc = &r->desc;
    return ommResult_SUCCESS;
}

// ref: debug_impl.cpp:512-641 (CollectStats; area is not part of the public result, so knownAreaMetric is 0 as in ommDebugGetStats)
OMM_API ommResult ommDebugGetStats(ommBaker baker, const ommCpuBakeResultDesc* res, ommDebugStats* out) {
    if (baker == 0 || res == nullptr || out == nullptr) return ommResult_INVALID_ARGUMENT;
    ommDebugStats s{};
    std::map<uint32_t, uint32_t> refs;
    for (uint32_t i = 0; i < res->indexCount; ++i) {
        int32_t idx;
        if (res->indexFormat == ommIndexFormat_UINT_8) idx = ((const int8_t*)res->indexBuffer)[i];
        else if (res->indexFormat == ommIndexFormat_UINT_16) idx = ((const int16_t*)res->indexBuffer)[i];
        else idx = ((const int32_t*)res->indexBuffer)[i];
        if (idx == ommSpecialIndex_FullyTransparent) s.totalFullyTransparent++;
        else if (idx == ommSpecialIndex_FullyOpaque) s.totalFullyOpaque++;
        else if (idx == ommSpecialIndex_FullyUnknownTransparent) s.totalFullyUnknownTransparent++;
        else if (idx == ommSpecialIndex_FullyUnknownOpaque) s.totalFullyUnknownOpaque++;
        else refs[(uint32_t)idx]++;
    }
    for (const auto& kv : refs) {
        if (kv.first >= res->descArrayCount) return ommResult_FAILURE;
        const ommCpuOpacityMicromapDesc& d = res->descArray[kv.first];
        const uint8_t* data = (const uint8_t*)res->arrayData + d.offset;
        const uint32_t n = 1u << (d.subdivisionLevel << 1);
        const uint32_t is2 = d.format == ommFormat_OC1_2_State ? 1 : 0;
        uint64_t cnt[4] = {0, 0, 0, 0};
        for (uint32_t u = 0; u < n; ++u) {
            const uint8_t v = data[u >> (2 + is2)];
            const uint32_t st = is2 ? ((v >> (u & 7)) & 1u) : ((v >> ((u << 1) & 7)) & 3u);
            cnt[st]++;
        }
        s.totalTransparent += kv.second * cnt[0];
        s.totalOpaque += kv.second * cnt[1];
        s.totalUnknownTransparent += kv.second * cnt[2];
        s.totalUnknownOpaque += kv.second * cnt[3];
    }
    *out = s;
    return ommResult_SUCCESS;
}

// ======================================================================================================================
// B200 extension
// ======================================================================================================================
OMM_API ommResult ommB200SetDevice(int cudaDevice) {
    if (cudaDevice < 0 || cudaDevice >= DeviceCount()) return ommResult_INVALID_ARGUMENT;
    g_requestedDevice = cudaDevice;
    return ommResult_SUCCESS;
}
OMM_API int ommB200GetDeviceCount(void) { return DeviceCount(); }
OMM_API size_t ommB200TrimHostPool(size_t keepBytes) { return PinnedPoolTrim(keepBytes); }

OMM_API ommResult ommB200GetLastBakeTimings(ommBaker baker, ommB200BakeTimings* out) {
    if (baker == 0 || out == nullptr || HandleTagOf(baker) != HandleTag::CpuBaker) return ommResult_INVALID_ARGUMENT;
    BakerObject* b = HandlePtr<BakerObject>(baker);
    std::lock_guard<std::mutex> g(b->mu);
    if (!b->haveTimings) return ommResult_FAILURE;
    *out = b->last;
    return ommResult_SUCCESS;
}

OMM_API ommResult ommB200StageInputs(ommBaker baker, const ommCpuBakeInputDesc* desc, ommB200StagedInputs* outStaged) {
    BakerObject* b = nullptr;
    const ommResult v = CheckBakeArgs(baker, desc, &b);
    if (v != ommResult_SUCCESS) return v;
    if (outStaged == nullptr) return ommResult_INVALID_ARGUMENT;
    StagedInputs* s = AllocObject<StagedInputs>(b->alloc);
    if (!s) return ommResult_FAILURE;
    const ommResult rc = StageInputs(b, *desc, s);
    if (rc != ommResult_SUCCESS) {
        FreeObject(b->alloc, s);
        return rc;
    }
    *outStaged = (ommB200StagedInputs)s;
    return ommResult_SUCCESS;
}
OMM_API ommResult ommB200DestroyStagedInputs(ommB200StagedInputs staged) {
    if (staged == 0) return ommResult_INVALID_ARGUMENT;
    StagedInputs* s = (StagedInputs*)staged;
    DestroyStagedDevice(s);
    const HostAllocator alloc = s->baker->alloc;
    FreeObject(alloc, s);
    return ommResult_SUCCESS;
}
OMM_API ommResult ommB200BakeResident(ommBaker baker, ommB200StagedInputs staged, void* cudaStream, ommCpuBakeResult* outBakeResult) {
    if (baker == 0 || staged == 0 || outBakeResult == nullptr || HandleTagOf(baker) != HandleTag::CpuBaker) return ommResult_INVALID_ARGUMENT;
    BakerObject* b = HandlePtr<BakerObject>(baker);
    StagedInputs* s = (StagedInputs*)staged;
    if (s->baker != b) return b->log.InvalidArg("[omm-b200] staged inputs belong to a different baker");
    HostTrace::Mark("ommB200BakeResident entry");
    const ommResult rc = RunBake(b, *s, cudaStream, false, false, 0.f, outBakeResult);
    HostTrace::Mark("bake");
    HostTrace::Dump();
    return rc;
}
OMM_API ommResult ommB200GetDeviceResultDesc(ommCpuBakeResult bakeResult, ommB200DeviceResultDesc* out) {
    if (bakeResult == 0 || out == nullptr) return ommResult_INVALID_ARGUMENT;
    const BakeResultObject* r = (const BakeResultObject*)bakeResult;
    if (!r->arrayOnThisRank || !r->deviceArrayComplete)
        return r->log.InvalidArg(r->arrayOnThisRank ? "[omm-b200] this result was assembled in host memory (sharded ommCpuBake, rank-0 mode); use ommB200BakeResident for a device-resident array"
                                                    : "[omm-b200] the complete result of this sharded bake lives on rank 0 (ommB200ShardedResultMode_OnRank0)");
    out->arrayData = r->devArrayData;
    out->descArray = r->devDescArray;
    out->indexBuffer = r->devIndexBuffer;
    out->arrayDataSize = r->descCount ? r->arrayDataSize : 0;
    out->descArrayCount = r->descCount;
    out->indexCount = r->indexCount;
    out->indexFormat = r->indexFormat;
    return ommResult_SUCCESS;
}
OMM_API ommResult ommB200DownloadResult(ommCpuBakeResult bakeResult) {
    if (bakeResult == 0) return ommResult_INVALID_ARGUMENT;
    BakeResultObject* r = (BakeResultObject*)bakeResult;
    if (!r->arrayOnThisRank) return r->log.InvalidArg("[omm-b200] the complete result of this sharded bake lives on rank 0 (ommB200ShardedResultMode_OnRank0)");
    float ms = 0.f;
    uint64_t bytes = 0;
    const bool was = r->downloaded;
    const ommResult rc = DownloadResult(r, &ms, &bytes);
    if (rc == ommResult_SUCCESS && !was && r->baker) RecordDeferredDownload(r->baker, ms, bytes, 0.f);
    return rc;
}
OMM_API ommResult ommB200InitSharding(ommBaker baker, int rank, int worldSize, const void* ncclUniqueIdBytes, size_t idSize) {
    if (baker == 0 || HandleTagOf(baker) != HandleTag::CpuBaker) return ommResult_INVALID_ARGUMENT;
    return InitSharding(HandlePtr<BakerObject>(baker), rank, worldSize, ncclUniqueIdBytes, idSize);
}
OMM_API ommResult ommB200GetNcclUniqueId(void* outBytes, size_t idSize) { return GetNcclUniqueId(outBytes, idSize); }
OMM_API ommResult ommB200SetShardedResultMode(ommBaker baker, ommB200ShardedResultMode mode) {
    if (baker == 0 || HandleTagOf(baker) != HandleTag::CpuBaker) return ommResult_INVALID_ARGUMENT;
    if (mode != ommB200ShardedResultMode_Replicated && mode != ommB200ShardedResultMode_OnRank0) return ommResult_INVALID_ARGUMENT;
    HandlePtr<BakerObject>(baker)->shard.resultMode = (int)mode;
    return ommResult_SUCCESS;
}
OMM_API int ommB200ShardsPerRank(int worldSize) { return ShardsPerRankOf(worldSize); }
OMM_API int ommB200ShardOwner(int shard, int worldSize) { return ShardOwnerOf(shard, worldSize); }
OMM_API ommResult ommB200ComputeShardBounds(const uint64_t* unitPrefix, uint32_t entries, int worldSize, uint32_t* outFirstItem) {
    return ComputeShardBounds((const unsigned long long*)unitPrefix, entries, worldSize, outFirstItem);
}
