// omm_internal.h -- types shared by the C-ABI shim (omm_api.cpp) and the device pipeline (omm_bake.cu).
#pragma once

#define OMMB200_BUILDING_LIBRARY 1
#include "../../include/omm_b200.h"

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <ctime>
#include <utility>
#include <mutex>
#include <new>
#include <vector>

namespace ommb200 {

// ---- host allocator plumbing (ref: libraries/omm-lib/src/std_allocator.h:45-117) -----------------------------
// Every host allocation whose lifetime is visible through the ABI goes through the caller's
// ommMemoryAllocatorInterface; when the caller gives none, an aligned malloc is used.
struct HostAllocator {
    ommMemoryAllocatorInterface iface{};
    void* alloc(size_t size, size_t alignment = 64) const { return iface.allocate(iface.userArg, size ? size : 1, alignment); }
    void release(void* p) const {
        if (p) iface.free(iface.userArg, p);
    }
};
void SetDefaultAllocatorIfUnset(ommMemoryAllocatorInterface& iface);

template <class T, class... Args>
T* AllocObject(const HostAllocator& a, Args&&... args) {
    void* mem = a.alloc(sizeof(T), alignof(T) < 16 ? 16 : alignof(T));
    if (!mem) return nullptr;
    return new (mem) T(static_cast<Args&&>(args)...);
}
template <class T>
void FreeObject(const HostAllocator& a, T* obj) {
    if (!obj) return;
    obj->~T();
    a.release(obj);
}

// ---- logger (ref: libraries/omm-lib/src/log.h:33-140) ---------------------------------------------------------
struct Logger {
    ommMessageInterface sink{};
    bool HasLogger() const { return sink.messageCallback != nullptr; }
    void Log(ommMessageSeverity sev, const char* msg) const {
        if (sink.messageCallback) sink.messageCallback(sev, msg, sink.userArg);
    }
    void Logf(ommMessageSeverity sev, const char* fmt, ...) const {
        if (!sink.messageCallback) return;
        char buf[256];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        sink.messageCallback(sev, buf, sink.userArg);
    }
    ommResult InvalidArg(const char* msg) const {
        Log(ommMessageSeverity_Fatal, msg);
        return ommResult_INVALID_ARGUMENT;
    }
};

// ---- host trace (OMM_B200_TRACE=1): wall-clock marks of one ommCpuBake call, printed to stderr when the call returns ----
struct HostTrace {
    static bool Enabled() { static const bool on = getenv("OMM_B200_TRACE") != nullptr; return on; }
    static std::vector<std::pair<const char*, double>>& Marks() { thread_local std::vector<std::pair<const char*, double>> m; return m; }
    static double Now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
    static void Mark(const char* label) { if (Enabled()) Marks().push_back({label, Now()}); }
    static void Dump() {
        if (!Enabled()) return;
        auto& m = Marks();
        for (size_t i = 1; i < m.size(); ++i) fprintf(stderr, "[omm-b200 trace] %-28s +%8.3f ms  (at %8.3f)\n", m[i].first, m[i].second - m[i - 1].second, m[i].second - m[0].second);
        m.clear();
    }
};

// ---- handle tags (ref: libraries/omm-lib/src/omm_handle.h:17-53): low 3 bits of the pointer ------------------
enum class HandleTag : uintptr_t { Reserved = 0, GpuBaker = 1, Pipeline = 2, CpuBaker = 3, Texture = 4, BakeResult = 5, SerializeResult = 6, DeserializeResult = 7 };
template <class H, class T>
H MakeHandle(T* p, HandleTag tag) { return (H)((uintptr_t)p | (uintptr_t)tag); }
template <class T, class H>
T* HandlePtr(H h) { return (T*)((uintptr_t)h & ~(uintptr_t)7); }
template <class H>
HandleTag HandleTagOf(H h) { return (HandleTag)((uintptr_t)h & 7); }

// ---- device-side view of the alpha texture --------------------------------------------------------------------
constexpr int kMaxMips = 17;  // 65536 = 2^16 is the largest legal dimension (ref: texture_impl.h:153)
struct DevMip {
    int w, h;
    int log2w, log2h;  // ctz(size) (ref: util/bit_tricks.h:66-78)
    int isPow2;
    float rcpw, rcph;           // 1.f / size (ref: texture_impl.cpp:102)
    unsigned long long texelOffset;  // element offset of this mip in the texel buffer
    unsigned long long satOffset;    // element offset of this mip in the SAT buffer
};
struct DevTexture {
    const void* texels;       // row-major, float or uint8_t
    const uint32_t* sat;      // inclusive summed-area table of (alpha > cutoff), or nullptr
    const uint32_t* flatSat;  // mip 0: inclusive summed-area table over the (w-1) x (h-1) interior cells of "not a constant, clearly one-sided
                              // cell" for the bake's alpha cutoff (omm_hier.cuh (H)), or nullptr
    const uint32_t* strongPlus;   // mip 0: summed-area tables of "not an item-independent whole-cell pass above / below the cutoff"
    const uint32_t* strongMinus;  // (omm_hier.cuh (I)), same layout as flatSat, or nullptr
    int isFp32;
    int mipCount;
    DevMip mips[kMaxMips];
};

struct CellTables {
    float cutoff = 0.f;
    uint32_t* flatSat = nullptr;
    uint32_t* strongPlus = nullptr;
    uint32_t* strongMinus = nullptr;
    int refs = 0;          // bakes between AcquireCellTables and ReleaseCellTables
    uint64_t lastUse = 0;  // TextureObject::cellTableClock at the last acquire (eviction order of idle sets)
};

struct TextureObject {
    HostAllocator alloc;
    ommCpuTextureFormat format = ommCpuTextureFormat_MAX_NUM;
    ommCpuTextureFlags flags = ommCpuTextureFlags_None;
    float alphaCutoff = -1.f;
    uint32_t mipCount = 0;
    int device = 0;
    void* hostTexels = nullptr;  // tight row-major copy of all mips (for ommCpuGetTextureDesc)
    size_t hostBytes = 0;
    void* devTexels = nullptr;
    uint32_t* devSat = nullptr;
    // constant-cell tables of the hierarchical classifier: one immutable set per alpha cutoff the texture has been baked with, built on
    // first use, shared by concurrent bakes (reference-counted) and never rewritten once published (see AcquireCellTables, omm_bake.cu)
    std::mutex flatMu;
    std::vector<struct CellTables*> cellTables;
    uint64_t cellTableClock = 0;
    bool hasSerializedSat = false;  // deserialized textures: the blob carried a summed-area table (ref: texture_impl.h:105-108)
    DevTexture dev{};
    bool HasAlphaCutoff() const { return alphaCutoff >= 0.f; }
};

// ---- result ------------------------------------------------------------------------------------------------------
struct BakeResultObject {
    HostAllocator alloc;
    Logger log;
    ommCpuBakeResultDesc desc{};
    // host arrays (allocated with `alloc`)
    void* hostArrayData = nullptr;
    void* hostDescArray = nullptr;
    void* hostIndexBuffer = nullptr;
    ommCpuOpacityMicromapUsageCount hostArrayHist[26];
    ommCpuOpacityMicromapUsageCount hostIndexHist[26];
    // device-resident copies (freed with the result)
    int device = 0;
    void* devArrayData = nullptr;
    void* devDescArray = nullptr;
    void* devIndexBuffer = nullptr;
    uint32_t arrayDataSize = 0, descCount = 0, indexCount = 0;
    ommIndexFormat indexFormat = ommIndexFormat_UINT_32;
    bool downloaded = false;
    bool arrayDataDownloaded = false;  // hostArrayData was filled slice by slice while the array was packed (ommCpuBake on one GPU)
    float earlyD2hMs = 0.f;
    bool arrayDataFromPinnedPool = false;  // hostArrayData came from the library's page-locked pool (default allocator only)
    bool descFromPinnedPool = false, indexFromPinnedPool = false;  // the same for hostDescArray / hostIndexBuffer (1 MiB and more)
    int sharedWindowId = -1;               // hostArrayData is a SharedHostWindow of the baker's sharding (root rank of a sharded ommCpuBake)
    bool arrayOnThisRank = true;           // false: sharded bake in rank-0 mode, seen from another rank (descriptors and index buffer only)
    bool deviceArrayComplete = true;       // false: rank 0 of a sharded ommCpuBake in rank-0 mode -- the array was assembled in host memory only
    bool usesDefaultAllocator = false;
    struct BakerObject* baker = nullptr;
};

// ---- staged (HBM-resident) inputs -------------------------------------------------------------------------------
struct StagedInputs {
    struct BakerObject* baker = nullptr;
    ommCpuBakeInputDesc desc{};  // copy; texCoords/indexBuffer/... pointers are NOT used after staging
    int device = 0;
    void* devIndices = nullptr;
    void* devTexCoords = nullptr;
    uint8_t* devLevels = nullptr;
    int32_t* devFormats = nullptr;
    uint32_t triangleCount = 0;
    uint32_t texCoordStride = 0;
    size_t texCoordBytes = 0;
    uint64_t h2dBytes = 0;
    float h2dMs = 0.f;
};

// Sharded bakes on one box: a page-locked host buffer that EVERY rank can write with its own copy engine over its own PCIe link
// (POSIX shared memory, registered with CUDA in each process), so that the host copy of a result is assembled in parallel instead of
// being pulled through one GPU's link.  Created by the root rank on demand, cached for the life of the sharding.
struct SharedHostWindow {
    int id = -1;
    void* ptr = nullptr;
    size_t capacity = 0;
    bool inUse = false;     // root only: a live result owns it
    bool unlinked = false;  // root only: the name is gone from /dev/shm (every rank has mapped it)
};
struct ShmControl;  // lives in shared memory (omm_bake.cu)
struct ShardState {
    int rank = 0, world = 1;
    void* ncclComm = nullptr;  // ncclComm_t
    ShmControl* ctl = nullptr;
    unsigned long long idHash = 0;  // names of the shared-memory objects derive from the ncclUniqueId
    unsigned long long bakeSeq = 0; // sharded ommCpuBake calls so far (the same number on every rank: the call is collective)
    int resultMode = 0;             // ommB200ShardedResultMode: 0 = every rank gets the complete array, 1 = rank 0 only
    std::vector<SharedHostWindow> windows;
};

struct BakerObject {
    HostAllocator alloc;
    bool usesDefaultAllocator = false;
    Logger log;
    int device = 0;
    ShardState shard;
    std::mutex mu;
    ommB200BakeTimings last{};
    bool haveTimings = false;
};

// a17 / a18 (near-duplicate merge, size-budget compression) are requested by this desc (omm_post_passes.cuh)
bool HostPassesNeeded(const ommCpuBakeInputDesc& desc);

// implemented in omm_serialize.cpp
struct SerializedResultObject;
struct DeserializedResultObject;
ommResult SerializeImpl(BakerObject* baker, const ommCpuDeserializedDesc& desc, SerializedResultObject** out);
const ommCpuBlobDesc* SerializedDesc(const SerializedResultObject* r);
void DestroySerialized(SerializedResultObject* r);
ommResult DeserializeImpl(BakerObject* baker, const ommCpuBlobDesc& blob, DeserializedResultObject** out);
const ommCpuDeserializedDesc* DeserializedDesc(const DeserializedResultObject* r);
void DestroyDeserialized(DeserializedResultObject* r);

// implemented in omm_bake.cu
ommResult UploadTexture(TextureObject* tex, const Logger& log);
void DestroyTextureDevice(TextureObject* tex);
ommResult StageInputs(BakerObject* baker, const ommCpuBakeInputDesc& desc, StagedInputs* out);
void DestroyStagedDevice(StagedInputs* s);
// Runs the whole device pipeline.  On success fills res (device buffers; host histograms) and timings.
ommResult BakeOnDevice(BakerObject* baker, const StagedInputs& in, void* userStream, BakeResultObject* res, ommB200BakeTimings* t, bool earlyDownload);
ommResult DownloadResult(BakeResultObject* res, float* d2hMs, uint64_t* d2hBytes);
void DestroyResultDevice(BakeResultObject* res);
int DeviceCount();
ommResult InitSharding(BakerObject* baker, int rank, int world, const void* id, size_t idSize);
ommResult GetNcclUniqueId(void* out, size_t size);
void DestroySharding(BakerObject* baker);
void ReleaseSharedWindow(BakerObject* baker, int windowId);
ommResult ComputeShardBounds(const unsigned long long* unitStart, uint32_t entries, int world, uint32_t* outFirstItem);
int ShardsPerRankOf(int world);
int ShardOwnerOf(int shard, int world);
int CurrentDeviceOr(int fallback);
// Page-locked host blocks recycled across bakes (used for big result arrays when the caller left the allocator to us).
void* PinnedPoolAcquire(size_t bytes);
void PinnedPoolRelease(void* p);
// Frees cached (not in use) blocks until at most keepBytes stay cached; returns the bytes still cached.
size_t PinnedPoolTrim(size_t keepBytes);

}  // namespace ommb200
