// omm_bake.cu -- device pipeline of libomm-b200.so: everything ommCpuBake does between "inputs are in HBM" and
// "result arrays are in HBM" (SURVEY.md section 8a, rows a3-a21), as CUDA kernels for sm_100a.
//
// Pipeline (one stream -- plus a second one for every other classifier chunk of a single-GPU bake, see HierChunkLanes -- and two host
// read-back points of a few counters each; DESIGN.md section 4 has the table):
//   K1  SetupTriangles         fetch indices/UVs, pick subdivision level (a3)
//   K2  UvTableInsert/Resolve  "first triangle wins" UV pre-dedup via a CAS hash table + atomicMin (a3)
//   K3  BuildItems             compact unique triangles into work items (the SDK's first-seen order)
//   K3b ItemKeysAll, radix sort, PermuteItems   the work items in OUTPUT order (the SDK's spatial sort, a20, moved in front)
//   K4  Hier* kernels          hierarchical classification of the micro-triangles (a5-a14)        <-- 90 % of a bake
//       ClassifyKernel(Q)      the flat kernels: Nearest filter, foreign SAT cutoff, internal flags
//   K5  ItemPostKernel, ItemPostBigPipelined   special-index detection + XXH64 of the 3-state bytes (a15, a16)
//   K6  DigestInsert/Resolve   "lowest first-seen item wins" exact dedup (a16)
//   P   omm_post_passes.cuh    optional: near-duplicate merge, size-budget compression (a17, a18)
//   K7  EmitInfo, prefix sums, TriangleFinalItems   histograms, descriptor slots, byte offsets (a19, a21)
//   K8  WriteDescs, PackItems, WriteIndexBuffer      serialization (a21)
// Sharded over several GPUs: K1-K3b replicated, K4-K5 on a rank's shards, one all-gather of digests + special indices, K6-K7
// replicated, K8 on the rank's shards, byte ranges exchanged (BakeOnDevice, "the exchange").
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>  // types only: the library is bound at run time (see NcclApi) so single-GPU use needs no libnccl

#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/statvfs.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cmath>

#include "omm_device_math.cuh"
#include "omm_hier.cuh"
#include "omm_xxh64.h"

namespace ommb200 {

#define CUDA_TRY(expr)                                                                                                  \
    do {                                                                                                                \
        cudaError_t _e = (expr);                                                                                        \
        if (_e != cudaSuccess) {                                                                                        \
            log.Logf(ommMessageSeverity_Fatal, "[omm-b200] CUDA error %s at %s:%d (%s)", cudaGetErrorName(_e), __FILE__, __LINE__, #expr); \
            rc = ommResult_FAILURE;                                                                                     \
            goto cleanup;                                                                                               \
        }                                                                                                               \
    } while (0)

constexpr int kMaxLevel = 12;
constexpr uint32_t kNoItem = 0xFFFFFFFFu;

// NVTX ranges around the stages of a bake (visible in Nsight Systems / Compute timelines).  libnvToolsExt is bound at run time like
// NCCL: when it is not installed the ranges are no-ops.
struct NvtxApi {
    int (*push)(const char*) = nullptr;
    int (*pop)() = nullptr;
};
static NvtxApi& Nvtx() {
    static NvtxApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        if (getenv("OMM_B200_NO_NVTX")) return;
        const char* names[] = {"libnvToolsExt.so.1", "libnvToolsExt.so"};
        for (const char* n : names)
            if (void* lib = dlopen(n, RTLD_NOW | RTLD_LOCAL)) {
                api.push = (int (*)(const char*))dlsym(lib, "nvtxRangePushA");
                api.pop = (int (*)())dlsym(lib, "nvtxRangePop");
                if (api.push && api.pop) break;
                api.push = nullptr;
                api.pop = nullptr;
            }
    });
    return api;
}
struct NvtxRange {
    bool on;
    explicit NvtxRange(const char* name) : on(Nvtx().push != nullptr) {
        if (on) Nvtx().push(name);
    }
    ~NvtxRange() {
        if (on) Nvtx().pop();
    }
};
// Events and non-blocking streams are recycled across bakes (per process and device): creating and destroying the dozen a bake uses cost
// ~0.35 ms of host time per call -- 2 % of a config-3 bake on one GPU, 10 % of its eighth on eight.
struct CudaObjectPool {
    std::mutex mu;
    std::vector<cudaEvent_t> timing[64], plain[64];
    std::vector<cudaStream_t> streams[64];
};
static CudaObjectPool& ObjectPool() {
    static CudaObjectPool* pool = new CudaObjectPool();  // never destroyed: CUDA may be gone by the time static destructors run
    return *pool;
}
static int PoolDevice() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess) cudaGetLastError();
    return d >= 0 && d < 64 ? d : 0;
}
static cudaError_t PoolEventCreate(cudaEvent_t* e, bool timing = true) {
    CudaObjectPool& p = ObjectPool();
    const int d = PoolDevice();
    {
        std::lock_guard<std::mutex> g(p.mu);
        std::vector<cudaEvent_t>& v = timing ? p.timing[d] : p.plain[d];
        if (!v.empty()) {
            *e = v.back();
            v.pop_back();
            return cudaSuccess;
        }
    }
    return timing ? cudaEventCreate(e) : cudaEventCreateWithFlags(e, cudaEventDisableTiming);
}
static void PoolEventRelease(cudaEvent_t e, bool timing = true) {
    if (!e) return;
    CudaObjectPool& p = ObjectPool();
    const int d = PoolDevice();
    std::lock_guard<std::mutex> g(p.mu);
    (timing ? p.timing[d] : p.plain[d]).push_back(e);
}
static cudaError_t PoolStreamCreate(cudaStream_t* s) {
    CudaObjectPool& p = ObjectPool();
    const int d = PoolDevice();
    {
        std::lock_guard<std::mutex> g(p.mu);
        if (!p.streams[d].empty()) {
            *s = p.streams[d].back();
            p.streams[d].pop_back();
            return cudaSuccess;
        }
    }
    return cudaStreamCreateWithFlags(s, cudaStreamNonBlocking);
}
// (the caller has synchronised the stream: nothing of the finished bake is pending on it)
static void PoolStreamRelease(cudaStream_t s) {
    if (!s) return;
    CudaObjectPool& p = ObjectPool();
    const int d = PoolDevice();
    std::lock_guard<std::mutex> g(p.mu);
    p.streams[d].push_back(s);
}


// ---------------------------------------------------------------------------------------------------------------------
// Work item record (one per unique UV triangle).
// ---------------------------------------------------------------------------------------------------------------------
struct ItemRec {
    float2 p0, p1, p2;
    uint32_t tri;        // triangle that created the item (first occurrence)
    uint8_t level;
    uint8_t format;      // ommFormat of the item
    uint8_t degenerate;  // base UV triangle is degenerate (ref: util/geometry.h:44-47)
    uint8_t hashLevel;   // level the item was classified at: the exact-dedup digest always covers 4^hashLevel bytes, also after Compress
};

struct SetupArgs {
    const void* indices;
    const void* texCoords;
    const uint8_t* levels;    // optional per-triangle levels
    const int32_t* formats;   // optional per-triangle formats
    uint32_t triCount;
    uint32_t texCoordStride;
    int indexFormat;
    int texCoordFormat;
    int globalFormat;
    int maxLevel;
    float dynScale;
    int edgeHeuristic;
    int disableLevelLine;
    uint32_t texW, texH;
};

// ---------------------------------------------------------------------------------------------------------------------
// hashes
// ---------------------------------------------------------------------------------------------------------------------
// libstdc++ std::hash<float>: murmur-style _Hash_bytes(&v, 4, 0xc70f6907); +-0.0f hash to 0.
__device__ __forceinline__ uint64_t StdHashFloat(float v) {
    if (v == 0.0f) return 0;
    const uint64_t mul = (((uint64_t)0xc6a4a793UL) << 32) + (uint64_t)0x5bd1e995UL;
    uint64_t hash = (uint64_t)0xc70f6907UL ^ (4 * mul);
    hash ^= (uint64_t)__float_as_uint(v);
    hash *= mul;
    hash = (hash ^ (hash >> 47)) * mul;
    hash = hash ^ (hash >> 47);
    return hash;
}
__device__ __forceinline__ void GlmHashCombine(uint64_t& seed, uint64_t hash) {  // ref: external/glm/glm/gtx/hash.inl:6-10
    hash += 0x9e3779b9 + (seed << 6) + (seed >> 2);
    seed ^= hash;
}
__device__ __forceinline__ uint64_t HashFloat2(float2 v) {
    uint64_t seed = 0;
    GlmHashCombine(seed, StdHashFloat(v.x));
    GlmHashCombine(seed, StdHashFloat(v.y));
    return seed;
}
__device__ __forceinline__ void OmmHashCombine(uint64_t& seed, uint64_t h) {  // ref: util/geometry.h:141-146
    seed ^= h + 0x9e3779b9 + (seed << 6) + (seed >> 2);
}

// ---------------------------------------------------------------------------------------------------------------------
// K1: per-triangle set-up (ref: bake_cpu_impl.cpp:579-633, util/geometry.h:148-239)
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 FetchUV(const SetupArgs& a, uint32_t index) {
    const uint8_t* base = (const uint8_t*)a.texCoords + (size_t)a.texCoordStride * index;
    if (a.texCoordFormat == ommTexCoordFormat_UV32_FLOAT) {
        // the stride may be any byte count: read component-wise through byte loads when unaligned
        float x, y;
        if ((((uintptr_t)base) & 3) == 0) {
            x = __ldg((const float*)base);
            y = __ldg((const float*)base + 1);
        } else {
            uint32_t xb = 0, yb = 0;
            for (int i = 0; i < 4; ++i) {
                xb |= (uint32_t)base[i] << (8 * i);
                yb |= (uint32_t)base[4 + i] << (8 * i);
            }
            x = __uint_as_float(xb);
            y = __uint_as_float(yb);
        }
        return make_float2(x, y);
    }
    uint32_t packed = 0;
    for (int i = 0; i < 4; ++i) packed |= (uint32_t)base[i] << (8 * i);
    const uint16_t lo = (uint16_t)(packed & 0xFFFFu), hi = (uint16_t)(packed >> 16);
    if (a.texCoordFormat == ommTexCoordFormat_UV16_UNORM)  // glm::unpackUnorm2x16
        return make_float2((float)lo * 1.5259021896696421759365224689097e-5f, (float)hi * 1.5259021896696421759365224689097e-5f);
    // glm::unpackHalf2x16 -> detail::toFloat32: exact half -> float widening
    return make_float2(__half2float(__ushort_as_half(lo)), __half2float(__ushort_as_half(hi)));
}
__device__ __forceinline__ uint32_t FetchIndex(const SetupArgs& a, size_t i) {
    if (a.indexFormat == ommIndexFormat_UINT_8) return ((const uint8_t*)a.indices)[i];
    if (a.indexFormat == ommIndexFormat_UINT_16) return ((const uint16_t*)a.indices)[i];
    return ((const uint32_t*)a.indices)[i];
}
__device__ __forceinline__ float Area2D(float2 p0, float2 p1, float2 p2) {  // ref: bake_cpu_impl.cpp:464-468
    const float v0x = p2.x - p0.x, v0y = p2.y - p0.y, v1x = p1.x - p0.x, v1y = p1.y - p0.y;
    const float cz = v0x * v1y - v1x * v0y;
    return 0.5f * sqrtf(cz * cz);  // |(0,0,cz)|; the zero components add exact zeros
}
__device__ __forceinline__ uint32_t AreaHeuristic(const SetupArgs& a, float2 p0, float2 p1, float2 p2) {  // ref: :470-509
    const float sx = (float)a.texW, sy = (float)a.texH;
    const float area = Area2D(make_float2(p0.x * sx, p0.y * sy), make_float2(p1.x * sx, p1.y * sy), make_float2(p2.x * sx, p2.y * sy));
    const float target = a.dynScale * a.dynScale;
    const float q = area / target;
    // uint(float) as GCC/x86-64 does it: cvttss2si to int64, keep the low 32 bits
    long long q64 = (q >= -9223372036854775808.f && q < 9223372036854775808.f) ? __float2ll_rz(q) : (long long)0x8000000000000000ull;
    uint32_t v = (uint32_t)q64;
    v--; v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16; v++;
    uint32_t r = (v & 0xAAAAAAAAu) != 0;
    r |= (uint32_t)((v & 0xFFFF0000u) != 0) << 4;
    r |= (uint32_t)((v & 0xFF00FF00u) != 0) << 3;
    r |= (uint32_t)((v & 0xF0F0F0F0u) != 0) << 2;
    r |= (uint32_t)((v & 0xCCCCCCCCu) != 0) << 1;
    const uint32_t lvl = r >> 1;
    return min(lvl, (uint32_t)a.maxLevel);
}
// squared longest edge in texels; the log2 part of the edge heuristic is finished on the host (see FinishEdgeHeuristic)
__device__ __forceinline__ float EdgeHeuristicEMax(const SetupArgs& a, float2 p0, float2 p1, float2 p2) {  // ref: :511-522
    const float sx = (float)a.texW, sy = (float)a.texH;
    const float e0x = sx * (p1.x - p0.x), e0y = sy * (p1.y - p0.y);
    const float e1x = sx * (p2.x - p0.x), e1y = sy * (p2.y - p0.y);
    const float e2x = sx * (p2.x - p1.x), e2y = sy * (p2.y - p1.y);
    const float l0 = e0x * e0x + e0y * e0y, l1 = e1x * e1x + e1y * e1y, l2 = e2x * e2x + e2y * e2y;
    float eMax = l0;
    if (eMax < l1) eMax = l1;
    if (eMax < l2) eMax = l2;
    return eMax;
}

struct HostLevelFix {  // triangles whose level needs libm's log2f (edge heuristic): finished on the host
    uint32_t tri;
    float eMax;
};

__global__ void SetupTriangles(SetupArgs a, float2* __restrict__ triUV, int8_t* __restrict__ triLevel, uint8_t* __restrict__ triFormat,
                               uint8_t* __restrict__ triDegenerate, HostLevelFix* __restrict__ fixList, uint32_t* __restrict__ fixCount) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.triCount) return;
    const uint32_t i0 = FetchIndex(a, (size_t)3 * t), i1 = FetchIndex(a, (size_t)3 * t + 1), i2 = FetchIndex(a, (size_t)3 * t + 2);
    const float2 p0 = FetchUV(a, i0), p1 = FetchUV(a, i1), p2 = FetchUV(a, i2);
    triUV[(size_t)3 * t] = p0;
    triUV[(size_t)3 * t + 1] = p1;
    triUV[(size_t)3 * t + 2] = p2;
    const bool invalid = isnan(p0.x) || isnan(p0.y) || isnan(p1.x) || isnan(p1.y) || isnan(p2.x) || isnan(p2.y) || isinf(p0.x) || isinf(p0.y) ||
                         isinf(p1.x) || isinf(p1.y) || isinf(p2.x) || isinf(p2.y);
    const bool degenerate = TriIsDegenerate(p0, p1, p2);
    int level;  // ref: bake_cpu_impl.cpp:542-560
    if (a.levels && a.levels[t] <= 12) level = a.levels[t];
    else if (a.dynScale > 0.f) {
        if (degenerate || a.edgeHeuristic) {
            level = -2;  // resolved by the host with the same libm as the reference
            if (!invalid) {
                const uint32_t slot = atomicAdd(fixCount, 1u);
                fixList[slot].tri = t;
                fixList[slot].eMax = EdgeHeuristicEMax(a, p0, p1, p2);
            }
        } else
            level = (int)AreaHeuristic(a, p0, p1, p2);
    } else
        level = a.maxLevel;
    // ref: bake_cpu_impl.cpp:562-575, 616 -- NaN/Inf triangles (and degenerate ones when the level-line test is off) are skipped
    if (invalid || (a.disableLevelLine && degenerate)) level = -1;
    triLevel[t] = (int8_t)level;
    triFormat[t] = (uint8_t)((!a.formats || a.formats[t] == ommFormat_INVALID) ? a.globalFormat : a.formats[t]);
    triDegenerate[t] = degenerate ? 1 : 0;
}

__global__ void ApplyLevelFixes(const HostLevelFix* __restrict__ fixList, const int8_t* __restrict__ fixedLevels, uint32_t n, int8_t* __restrict__ triLevel) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) triLevel[fixList[i].tri] = fixedLevels[i];
}

// ---------------------------------------------------------------------------------------------------------------------
// K2: UV pre-dedup.  Key = the SDK's 64-bit hash_combine chain (ref: bake_cpu_impl.cpp:626-633); the table maps
// key -> lowest triangle index having it, which is exactly "first seen wins" of the serial loop (:635-649).
// ---------------------------------------------------------------------------------------------------------------------
constexpr uint64_t kEmptyKey = 0xFFFFFFFFFFFFFFFFull;  // remapped below if a real key equals it

__device__ __forceinline__ uint64_t TableSlot(uint64_t key, uint64_t mask) { return ((key * 0x9E3779B97F4A7C15ull) >> 17) & mask; }

__device__ __forceinline__ void TableInsertMin(uint64_t* keys, uint32_t* vals, uint64_t mask, uint64_t key, uint32_t val) {
    if (key == kEmptyKey) key = 0x7FFFFFFFFFFFFFFFull;  // keep the sentinel free (a 2^-64 aliasing, same class as a hash collision)
    uint64_t slot = TableSlot(key, mask);
    while (true) {
        const uint64_t prev = atomicCAS((unsigned long long*)&keys[slot], (unsigned long long)kEmptyKey, (unsigned long long)key);
        if (prev == kEmptyKey || prev == key) {
            atomicMin(&vals[slot], val);
            return;
        }
        slot = (slot + 1) & mask;
    }
}
__device__ __forceinline__ uint32_t TableFind(const uint64_t* keys, const uint32_t* vals, uint64_t mask, uint64_t key) {
    if (key == kEmptyKey) key = 0x7FFFFFFFFFFFFFFFull;
    uint64_t slot = TableSlot(key, mask);
    while (true) {
        const uint64_t k = keys[slot];
        if (k == key) return vals[slot];
        if (k == kEmptyKey) return kNoItem;
        slot = (slot + 1) & mask;
    }
}

__global__ void FillTable(uint64_t* keys, uint32_t* vals, uint64_t cap) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cap) {
        keys[i] = kEmptyKey;
        vals[i] = 0xFFFFFFFFu;
    }
}

__global__ void UvTableInsert(const float2* __restrict__ triUV, const int8_t* __restrict__ triLevel, const uint8_t* __restrict__ triFormat, uint32_t triCount,
                              uint64_t* __restrict__ triKey, uint64_t* keys, uint32_t* vals, uint64_t mask) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= triCount) return;
    const int level = triLevel[t];
    if (level < 0) return;
    uint64_t seed = 42;
    OmmHashCombine(seed, HashFloat2(triUV[(size_t)3 * t]));
    OmmHashCombine(seed, HashFloat2(triUV[(size_t)3 * t + 1]));
    OmmHashCombine(seed, HashFloat2(triUV[(size_t)3 * t + 2]));
    OmmHashCombine(seed, (uint64_t)(int64_t)level);
    OmmHashCombine(seed, (uint64_t)(int64_t)(int32_t)triFormat[t]);
    triKey[t] = seed;
    TableInsertMin(keys, vals, mask, seed, t);
}

// flag[t] = 1 when triangle t opens a work item
__global__ void UvTableResolve(const int8_t* __restrict__ triLevel, const uint64_t* __restrict__ triKey, uint32_t triCount, const uint64_t* __restrict__ keys,
                               const uint32_t* __restrict__ vals, uint64_t mask, int disableDup, uint32_t* __restrict__ triFirst,
                               uint32_t* __restrict__ isItem) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= triCount) return;
    if (triLevel[t] < 0) {
        triFirst[t] = kNoItem;
        isItem[t] = 0;
        return;
    }
    // ref: bake_cpu_impl.cpp:636 -- with DisableDuplicateDetection every valid triangle becomes its own work item
    const uint32_t first = disableDup ? t : TableFind(keys, vals, mask, triKey[t]);
    triFirst[t] = first;
    isItem[t] = (first == t) ? 1u : 0u;
}

// ---------------------------------------------------------------------------------------------------------------------
// K3: build work items.  State block of an item: 2 bits per micro-triangle, padded to four 32-bit words.
// A "unit" is the work of one warp: 32 consecutive micro-triangles of one item.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void BuildItems(const float2* __restrict__ triUV, const int8_t* __restrict__ triLevel, const uint8_t* __restrict__ triFormat,
                           const uint8_t* __restrict__ triDegenerate, const uint32_t* __restrict__ isItem, const uint32_t* __restrict__ itemScan,
                           uint32_t triCount, ItemRec* __restrict__ itemsW, uint32_t* __restrict__ levelHist, uint32_t* __restrict__ numItemsOut) {
    __shared__ uint32_t sh[16];
    if (threadIdx.x < 16) sh[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = t < triCount && isItem[t];
    if (t + 1 == triCount) *numItemsOut = itemScan[t] + (isItem[t] ? 1u : 0u);
    if (active) atomicAdd(&sh[triLevel[t]], 1u);
    __syncthreads();
    if (threadIdx.x < 16 && sh[threadIdx.x]) atomicAdd(&levelHist[threadIdx.x], sh[threadIdx.x]);
    if (!active) return;
    const uint32_t w = itemScan[t];
    ItemRec it;
    it.p0 = triUV[(size_t)3 * t];
    it.p1 = triUV[(size_t)3 * t + 1];
    it.p2 = triUV[(size_t)3 * t + 2];
    it.tri = t;
    it.level = (uint8_t)triLevel[t];
    it.format = triFormat[t];
    it.degenerate = triDegenerate[t];
    it.hashLevel = it.level;
    itemsW[w] = it;
}

// ---------------------------------------------------------------------------------------------------------------------
// K3b: work items in OUTPUT order.  The SDK serializes the surviving work items sorted by (key, index) descending, key = level and
// Morton code of the quantised UV centroid (ref: bake_cpu_impl.cpp:1707-1754) -- a function of the geometry alone.  So the order is
// fixed BEFORE classification: the first-seen items (index w) are sorted once and every later stage works on positions s of that
// order.  What this buys: the survivors of any contiguous run of positions occupy a contiguous range of descriptors and of
// arrayData bytes, so a classifier chunk (one GPU) or a shard (several GPUs) can be packed -- and sent to the host, or to the other
// GPUs -- as soon as ITS items are classified, and the final "sort" is a prefix sum over survivor flags.  Exact-dedup winners are
// still "lowest first-seen index" (ItemRec::tri is monotone in w).
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ItemSortKey(const ItemRec& it) {
    // ref: bake_cpu_impl.cpp:1734-1747 -- level, then Morton code of the 13-bit quantised, MirrorOnce-folded UV centroid
    const float cx = (it.p0.x + it.p1.x + it.p2.x) / 3.f, cy = (it.p0.y + it.p1.y + it.p2.y) / 3.f;
    const int qx = f2i(8192.f * cx), qy = f2i(8192.f * cy);
    const int mx = clampi(f2i(fabsf((float)qx + 0.5f)), 0, 8191), my = clampi(f2i(fabsf((float)qy + 0.5f)), 0, 8191);
    uint32_t x = (uint32_t)mx, y = (uint32_t)my;
    x = (x | (x << 8)) & 0x00FF00FFu; x = (x | (x << 4)) & 0x0F0F0F0Fu; x = (x | (x << 2)) & 0x33333333u; x = (x | (x << 1)) & 0x55555555u;
    y = (y | (y << 8)) & 0x00FF00FFu; y = (y | (y << 4)) & 0x0F0F0F0Fu; y = (y | (y << 2)) & 0x33333333u; y = (y | (y << 1)) & 0x55555555u;
    return (((uint32_t)it.level << 26) | x | (y << 1)) + 1u;
}
// Keys of all `slots` entries (slots = triangle count; entries beyond the item count get key 0 and sort last).  The SDK sorts
// (key, index) pairs descending (std::greater, :1751): a stable descending radix sort fed in reversed index order yields the same
// order -- equal keys keep the higher index first.
__global__ void ItemKeysAll(const ItemRec* __restrict__ itemsW, const uint32_t* __restrict__ numItemsPtr, uint32_t slots, uint32_t* __restrict__ sortKeys,
                            uint32_t* __restrict__ sortVals) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= slots) return;
    const uint32_t pos = slots - 1 - w;
    sortKeys[pos] = w < *numItemsPtr ? ItemSortKey(itemsW[w]) : 0u;
    sortVals[pos] = w;
}
// items[s] = itemsW[perm[s]]; sizes of the state blocks / warp units / initial regions in the new order (zero beyond the item count)
__global__ void PermuteItems(const ItemRec* __restrict__ itemsW, const uint32_t* __restrict__ perm, const uint32_t* __restrict__ numItemsPtr, uint32_t slots,
                             ItemRec* __restrict__ items, uint32_t* __restrict__ posOfOrig, unsigned long long* __restrict__ itemUnits,
                             unsigned long long* __restrict__ itemWords, unsigned long long* __restrict__ itemNodes) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > slots) return;
    if (s >= *numItemsPtr) {
        itemUnits[s] = itemWords[s] = itemNodes[s] = 0ull;
        return;
    }
    const uint32_t w = perm[s];
    const ItemRec it = itemsW[w];
    items[s] = it;
    posOfOrig[w] = s;
    const unsigned long long n = 1ull << (2 * it.level);
    itemUnits[s] = n >= 32 ? n / 32 : 1;
    itemWords[s] = n >= 64 ? n / 16 : 4;  // blocks start 16-byte aligned so the warp stores and the pack copies can be vectorised
    itemNodes[s] = it.level > 3 ? 1ull << (2 * (it.level - 3)) : 1ull;  // initial regions of the hierarchical classifier (HierTestInitial)
}

__global__ void MapTrianglesToItems(const uint32_t* __restrict__ triFirst, const uint32_t* __restrict__ itemScan, const uint32_t* __restrict__ posOfOrig,
                                    uint32_t triCount, uint32_t* __restrict__ triItem) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= triCount) return;
    const uint32_t first = triFirst[t];
    triItem[t] = first == kNoItem ? kNoItem : posOfOrig[itemScan[first]];
}

// ---------------------------------------------------------------------------------------------------------------------
// K4: classification -- the hot kernel.  One thread per micro-triangle; a warp owns 32 consecutive bird-curve indices
// of one work item (spatially adjacent => the same few texels), packs the 32 two-bit states with two warp OR-reductions
// and stores them as one 8-byte word pair.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t FindItem(const unsigned long long* __restrict__ unitStart, uint32_t numItems, unsigned long long unit) {
    // largest i with unitStart[i] <= unit   (unitStart is the exclusive prefix sum of units per item)
    uint32_t lo = 0, hi = numItems;
    while (hi - lo > 1) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(&unitStart[mid]) <= unit) lo = mid;
        else hi = mid;
    }
    return lo;
}

#ifndef OMM_CLASSIFY_WARPS
#define OMM_CLASSIFY_WARPS 4
#endif
#ifndef OMM_CLASSIFY_MIN_BLOCKS
#define OMM_CLASSIFY_MIN_BLOCKS 10
#endif
constexpr int kClassifyWarps = OMM_CLASSIFY_WARPS;  // warps (= units) per block

template <class Cfg>
__global__ void __launch_bounds__(kClassifyWarps * 32, OMM_CLASSIFY_MIN_BLOCKS) ClassifyKernel(const BakeParams P, const ItemRec* __restrict__ items,
                                                                      const unsigned long long* __restrict__ unitStart,
                                                                      const unsigned long long* __restrict__ wordStart, uint32_t itemBegin, uint32_t itemEnd,
                                                                      unsigned long long unitBegin, unsigned long long unitEnd,
                                                                      uint32_t* __restrict__ stateWords) {
    // The block's first unit is located with one binary search; its warps walk forward from there (a block of 8 units
    // spans at most 8 work items, and exactly one for items of level >= 4).
    __shared__ uint32_t sFirstItem;
    const unsigned long long blockUnit = unitBegin + (unsigned long long)blockIdx.x * kClassifyWarps;
    if (threadIdx.x == 0) sFirstItem = itemBegin + FindItem(unitStart + itemBegin, itemEnd - itemBegin, blockUnit);
    __syncthreads();
    const unsigned long long unit = blockUnit + (threadIdx.x >> 5);
    if (unit >= unitEnd) return;
    const uint32_t lane = threadIdx.x & 31;
    uint32_t w = sFirstItem;
    while (w + 1 < itemEnd && __ldg(&unitStart[w + 1]) <= unit) ++w;
    const ItemRec it = items[w];
    const uint32_t level = it.level;
    const uint32_t n = 1u << (2 * level);
    const uint32_t first = (uint32_t)(unit - __ldg(&unitStart[w])) * 32u;
    const uint32_t idx = first + lane;
    uint32_t state = 0;
    if (idx < n) state = (uint32_t)ClassifyMicroTriangle<Cfg>(P, it.p0, it.p1, it.p2, it.degenerate != 0, idx, level);
    const uint32_t mine = state << (2 * (lane & 15));
    const uint32_t lo = __reduce_or_sync(0xFFFFFFFFu, lane < 16 ? mine : 0u);
    const uint32_t hi = __reduce_or_sync(0xFFFFFFFFu, lane >= 16 ? mine : 0u);
    if (lane == 0) {
        uint32_t* dst = stateWords + __ldg(&wordStart[w]) + (first >> 4);
        if (n >= 32) *reinterpret_cast<uint2*>(dst) = make_uint2(lo, hi);
        else *dst = lo;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// K4, load-balanced variant (Linear filter + level-line test + single mip, i.e. the default configuration).
//
// Most micro-triangles cover one texel cell, some two to four, a few (coarse levels on big textures) thousands.  In the
// kernel above the warp then idles through the extra cells of its unluckiest lane.  Here every lane evaluates only the
// FIRST covered cell of its micro-triangle in place; every further (micro-triangle, cell) pair goes to a per-warp queue in
// shared memory and is evaluated 32 pairs at a time by whichever lanes are free.  A warp processes kBatchUnits units per
// batch so that the queue fills.  Coverage counters are sums over visited cells (bake_kernels_cpu.h:317-323, 348-354,
// 371-372), so the evaluation order is immaterial; the set of visited cells is the reference's (RasterNext).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kBatchUnits = 4;
constexpr int kQueueCap = 64;  // >= 32 left over + 32 pushed per step
struct WarpQueue {
    float2 v0[kQueueCap], v1[kQueueCap], v2[kQueueCap];  // micro-triangle vertices (UV space)
    int cx[kQueueCap], cy[kQueueCap];
    uint32_t slot[kQueueCap];                             // unit-in-batch * 32 + lane
    uint32_t above[kBatchUnits * 32], below[kBatchUnits * 32];  // coverage counters per (unit, lane)
    int8_t fixedState[kBatchUnits * 32];                        // >= 0: decided without fine classification
    uint32_t itemOf[kBatchUnits], firstOf[kBatchUnits];
};

template <class Cfg>
__device__ __forceinline__ void DrainQueue(const BakeParams& P, const DevMip& m, WarpQueue& q, uint32_t head, uint32_t count, uint32_t lane) {
    if (lane < count) {
        const uint32_t e = (head + lane) & (kQueueCap - 1);
        const Tri t = MakeTri(q.v0[e], q.v1[e], q.v2[e]);
        Coverage c{0u, 0u};
        LevelLineCell<Cfg, false>(P, m, t, q.cx[e], q.cy[e], c);
        const uint32_t s = q.slot[e];
        if (c.above) atomicAdd(&q.above[s], c.above);
        if (c.below) atomicAdd(&q.below[s], c.below);
    }
    __syncwarp();
}

#ifndef OMM_CLASSIFYQ_MIN_BLOCKS
#define OMM_CLASSIFYQ_MIN_BLOCKS 8
#endif
template <class Cfg>
__global__ void __launch_bounds__(kClassifyWarps * 32, OMM_CLASSIFYQ_MIN_BLOCKS) ClassifyKernelQ(const BakeParams P, const ItemRec* __restrict__ items,
                                                                       const unsigned long long* __restrict__ unitStart,
                                                                       const unsigned long long* __restrict__ wordStart, uint32_t itemBegin, uint32_t itemEnd,
                                                                       unsigned long long unitBegin, unsigned long long unitEnd,
                                                                       uint32_t* __restrict__ stateWords) {
    __shared__ WarpQueue sQueues[kClassifyWarps];
    __shared__ uint32_t sFirstItem;
    const unsigned long long blockUnit = unitBegin + (unsigned long long)blockIdx.x * (kClassifyWarps * kBatchUnits);
    if (threadIdx.x == 0) sFirstItem = itemBegin + FindItem(unitStart + itemBegin, itemEnd - itemBegin, blockUnit);
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long batchUnit = blockUnit + (unsigned long long)warp * kBatchUnits;
    if (batchUnit >= unitEnd) return;
    WarpQueue& q = sQueues[warp];
    const DevMip& m = P.tex.mips[0];
    const bool earlyOut = P.promotion != ommUnknownStatePromotion_Nearest;
    for (int i = lane; i < kBatchUnits * 32; i += 32) { q.above[i] = 0; q.below[i] = 0; }
    __syncwarp();

    uint32_t w = sFirstItem;
    uint32_t head = 0, count = 0;            // queue state (warp-uniform)
#pragma unroll 1
    for (int u = 0; u < kBatchUnits; ++u) {
        const unsigned long long unit = batchUnit + u;
        if (unit >= unitEnd) {                                           // warp-uniform
            if (lane == 0) q.itemOf[u] = kNoItem;
            continue;
        }
        while (w + 1 < itemEnd && __ldg(&unitStart[w + 1]) <= unit) ++w;
        const ItemRec it = items[w];
        const uint32_t level = it.level, n = 1u << (2 * level);
        const uint32_t first = (uint32_t)(unit - __ldg(&unitStart[w])) * 32u;
        const uint32_t idx = first + lane;
        if (lane == 0) {
            q.itemOf[u] = w;
            q.firstOf[u] = first;
        }
        if (it.degenerate) {                                             // warp-uniform: zero-area UV triangles take the serial path
            q.fixedState[u * 32 + lane] = (int8_t)(idx < n ? ClassifyMicroTriangle<Cfg>(P, it.p0, it.p1, it.p2, true, idx, level) : 0);
            continue;
        }
        bool more = false;
        uint32_t regAbove = 0, regBelow = 0;
        Tri st;
        RasterSetup rs;
        RasterCursor cur;
        int fixedState = idx < n ? -1 : 0;
        if (idx < n) {
            st = MicroTri(it.p0, it.p1, it.p2, idx, level);
            int state = ommOpacityState_UnknownOpaque;
            if (P.useCoarse) {
                const int cs = CoarseState<Cfg>(P, st);
                if (cs >= 0) state = cs;
            }
            if (state != ommOpacityState_UnknownOpaque) fixedState = state;  // ref: bake_cpu_impl.cpp:861-864
            else {
                if (P.cutoff < TexBilinear<Cfg>(P, m, st.p0)) regAbove++;
                else regBelow++;
                rs = MakeRasterSetup(st, m.w, m.h, -0.5f);
                cur = RasterBegin(rs);
                int fx, fy;
                more = RasterNext(rs, cur, fx, fy);
                if (more) {
                    Coverage c{regAbove, regBelow};
                    LevelLineCell<Cfg, false>(P, m, st, fx, fy, c);
                    regAbove = c.above;
                    regBelow = c.below;
                }
            }
        }
        q.fixedState[u * 32 + lane] = (int8_t)fixedState;
        // further covered cells -> queue, one per lane per step; full rounds are drained as soon as they exist
        while (true) {
            int ex = 0, ey = 0;
            if (more && earlyOut && regAbove != 0 && regBelow != 0) more = false;  // state already decided (see ClassifyMicroTriangle)
            if (more) more = RasterNext(rs, cur, ex, ey);
            const uint32_t pushMask = __ballot_sync(0xFFFFFFFFu, more);
            if (pushMask == 0) break;
            if (more) {
                const uint32_t e = (head + count + __popc(pushMask & ((1u << lane) - 1u))) & (kQueueCap - 1);
                q.v0[e] = st.p0; q.v1[e] = st.p1; q.v2[e] = st.p2;
                q.cx[e] = ex; q.cy[e] = ey;
                q.slot[e] = (uint32_t)u * 32u + lane;
            }
            count += __popc(pushMask);
            __syncwarp();
            while (count >= 32) {
                DrainQueue<Cfg>(P, m, q, head, 32, lane);
                head = (head + 32) & (kQueueCap - 1);
                count -= 32;
            }
        }
        if (regAbove) atomicAdd(&q.above[u * 32 + lane], regAbove);
        if (regBelow) atomicAdd(&q.below[u * 32 + lane], regBelow);
    }
    if (count) DrainQueue<Cfg>(P, m, q, head, count, lane);
    __syncwarp();
#pragma unroll 1
    for (int u = 0; u < kBatchUnits; ++u) {
        const uint32_t item = q.itemOf[u];
        if (item == kNoItem) continue;
        const int fixedState = q.fixedState[u * 32 + lane];
        uint32_t state;
        if (fixedState >= 0) state = (uint32_t)fixedState;
        else state = (uint32_t)StateFromCoverage(P, q.above[u * 32 + lane], q.below[u * 32 + lane]);
        const uint32_t mine = state << (2 * (lane & 15));
        const uint32_t lo = __reduce_or_sync(0xFFFFFFFFu, lane < 16 ? mine : 0u);
        const uint32_t hi = __reduce_or_sync(0xFFFFFFFFu, lane >= 16 ? mine : 0u);
        if (lane == 0) {
            const uint32_t n = 1u << (2 * items[item].level);
            uint32_t* dst = stateWords + __ldg(&wordStart[item]) + (q.firstOf[u] >> 4);
            if (n >= 32) *reinterpret_cast<uint2*>(dst) = make_uint2(lo, hi);
            else *dst = lo;
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// K4, hierarchical variant (Linear filter + level-line test + single mip + no SAT pass: the default configuration).
//
// Level-synchronous descent of the bird-curve hierarchy with work lists in HBM:
//   HierPrepare       per work item: the constants of the exact shortcuts (HierItem, omm_hier.cuh)
//   HierTestInitial   one thread per INITIAL region (64 micro-triangles; the whole item below level 3): TestRegion proves the
//                     region to be on one side of the cutoff -> its state words are filled; else the region goes to a list
//   HierTestList      one thread per child (16, then 4 micro-triangles) of every listed region: same test, next list
//   HierLeaves        one thread per micro-triangle of every listed 4-region: the reference walk with the exact skips of
//                     LeafCell; four lanes assemble the byte of their 4-region
// Every kernel is small and homogeneous (no divergence between phases, full occupancy); a list entry is (item, region index).
// The lists are sized for the worst case (every test fails) of one CHUNK of initial regions; a bake is a sequence of chunks.
// ---------------------------------------------------------------------------------------------------------------------
// Initial regions per chunk.  The lists are sized for the worst case (every test fails): 38 entries of 8 bytes per initial region.
// The nominal chunk is 32 M regions = 2.1e9 micro-triangles (9.7 GiB of lists per lane, two lanes: 11 % of a B200's HBM, held by the
// stream-ordered pool between bakes): a bake of config-3 size is two chunks, which run side by side (HierChunkLanes); the chunk shrinks
// so that the lists of two lanes take at most an eighth of the free memory on smaller or busier devices.
constexpr unsigned long long kHierChunkRegionsMax = 32ull << 20, kHierChunkRegionsMin = 1ull << 20;
static unsigned long long HierNominalChunkRegions(int device) {
    // cudaMemGetInfo is a slow driver query (milliseconds with a large memory pool): ask once per device
    static std::mutex mu;
    static unsigned long long cached[64] = {};
    if (const char* e = getenv("OMM_B200_CHUNK_REGIONS")) {  // A/B runs: initial regions per chunk (read per bake)
        const unsigned long long v = strtoull(e, nullptr, 10);
        if (v >= (1ull << 16)) return v;
    }
    std::lock_guard<std::mutex> g(mu);
    if (device >= 0 && device < 64 && cached[device]) return cached[device];
    size_t freeB = 0, totalB = 0;
    unsigned long long regions = kHierChunkRegionsMin;
    if (cudaMemGetInfo(&freeB, &totalB) == cudaSuccess) {
        // 38 list entries of 8 bytes per initial region (1 + 4 + 16 failing regions, 1 unresolved, 16 slow-path), two lanes: an eighth of the free memory
        const unsigned long long byMemory = (unsigned long long)(freeB / 8) / (2ull * 38ull * 8ull);
        regions = std::max(kHierChunkRegionsMin, std::min(kHierChunkRegionsMax, byMemory));
    } else
        cudaGetLastError();
    if (device >= 0 && device < 64) cached[device] = regions;
    return regions;
}
// Blocks per SM of the grid-stride classifier kernels (8 are resident).  The tasks of these kernels differ in cost by an order of magnitude
// (a leaf the level line crosses against one its edge filter closes), and a block's share is fixed by its stride: with two waves of blocks
// (16 per SM, rounds 1 and early 2) the last blocks ran long after most SMs had drained.  Many short blocks let the hardware scheduler
// balance: measured at config 3 on one box, classification 12.92 / 12.24 / 11.75 / 11.38 / 11.27 / 11.37 / 12.24 ms at 8 / 16 / 32 / 64 /
// 128 / 256 / 1024 blocks per SM (`scripts/gpu_r2r.sh`).  OMM_B200_LIST_GRID_MULT (all kernels), OMM_B200_INIT_GRID_MULT and
// OMM_B200_LEAF_GRID_MULT override for A/B runs.
static uint32_t GridMultFromEnv(const char* name, uint32_t fallback) {
    const char* e = getenv(name);
    const long n = e ? atol(e) : (long)fallback;
    return (uint32_t)(n < 1 ? 1 : (n > 4096 ? 4096 : n));
}
static uint32_t HierGridBlocksPerSm() {
    const uint32_t v = GridMultFromEnv("OMM_B200_LIST_GRID_MULT", 128);
    return v;
}
static uint32_t HierInitGridBlocksPerSm() {
    const uint32_t v = GridMultFromEnv("OMM_B200_INIT_GRID_MULT", HierGridBlocksPerSm());
    return v;
}
static uint32_t HierLeafGridBlocksPerSm() {
    const uint32_t v = GridMultFromEnv("OMM_B200_LEAF_GRID_MULT", HierGridBlocksPerSm());
    return v;
}
static uint32_t HierSlowGridBlocksPerSm() {  // HierLeavesSlow: zero-area or huge-coordinate items only
    const uint32_t v = GridMultFromEnv("OMM_B200_SLOW_GRID_MULT", HierGridBlocksPerSm());
    return v;
}
// Classifier chunks in flight side by side ("lanes"; OMM_B200_CHUNK_LANES = 1..4, default 2; bakes on one GPU).  The level kernels of ONE
// chunk are a dependent sequence, each with a tail in which the SMs drain, and the per-item post pass (K5) waited for the last of them; chunks
// are independent of each other (whole work items, lists of their own), so the next chunk runs on a second stream, fills those tails, and a
// chunk's post pass follows its leaves on its own lane while the other lane classifies.  Measured at config 3 on one box, one process
// (`scripts/sweep_lanes.py`, `scripts/gpu_r2u.sh`, `gpu_r2v.sh`; every setting reproduces the SDK's digest): 12.53 ms per bake with one lane,
// **12.34** with two (both 32 M-region chunks in flight), 12.33 with three chunks of 22 M on three lanes, 12.30 with four of 16 M on four;
// end to end 18.35 -> 18.20 ms.  What it does NOT buy is L2 residency: chunks small enough for a chunk's state words and lists to stay in the
// 126 MB L2 are slower also when their kernel boundaries overlap (2 lanes: 12.56 ms at 8 M regions per chunk, 12.82 at 4 M; 4 lanes: 12.50 at
// 8 M, 13.1 at 2 M) -- the kernels are issue-bound, the extra launches cost more than the DRAM round trips.
constexpr int kMaxChunkLanes = 4;
static int HierChunkLanes() {  // (read per bake, like the grid switches: one process can compare settings)
    const char* e = getenv("OMM_B200_CHUNK_LANES");
    const long n = e ? atol(e) : 2;
    return (int)(n < 1 ? 1 : (n > kMaxChunkLanes ? kMaxChunkLanes : n));
}
struct HierLists {
    unsigned long long* q[3];        // failing regions of 64, 16 and 4 micro-triangles: (item << 32) | region index within the item
    unsigned long long* unresolved;  // initial regions the whole-cell bitmap did not answer
    unsigned long long* slow;        // 4-regions of items the shortcuts do not cover (HierLeavesSlow)
    unsigned long long* count;       // [0..2] the three lists, [3] unresolved, [5] slow
};

__global__ void HierPrepare(const BakeParams P, const ItemRec* __restrict__ items, uint32_t itemBegin, uint32_t itemEnd, HierItem* __restrict__ hierItems) {
    const uint32_t w = itemBegin + blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= itemEnd) return;
    const ItemRec it = items[w];
    // one set of constants per mip (they depend on the mip's size), laid out [item][mip]
    const int M = P.tex.mipCount;
    for (int k = 0; k < M; ++k) hierItems[(size_t)w * M + k] = MakeHierItemFor(P, P.tex.mips[k], it.p0, it.p1, it.p2, it.level, it.degenerate != 0);
}

__device__ __forceinline__ HierItem LoadHierItem(const HierItem* __restrict__ p) {
    // five 16-byte loads, unpacked field by field: copying through a uint4 view of the struct made ptxas keep it on the stack
    // (72-byte frame, 33 LDL in HierLeaves, profiles/r2_sass_*)
    const uint4* src = reinterpret_cast<const uint4*>(p);
    const uint4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2), d = __ldg(src + 3), e = __ldg(src + 4);
    HierItem hi;
    hi.p0 = make_float2(__uint_as_float(a.x), __uint_as_float(a.y));
    hi.p1 = make_float2(__uint_as_float(a.z), __uint_as_float(a.w));
    hi.p2 = make_float2(__uint_as_float(b.x), __uint_as_float(b.y));
    hi.level = b.z;
    hi.epsRegion = __uint_as_float(b.w);
    hi.epsSingle = __uint_as_float(c.x);
    hi.deltaEdge = __uint_as_float(c.y);
    hi.kmin[0] = __uint_as_float(c.z); hi.kmin[1] = __uint_as_float(c.w); hi.kmin[2] = __uint_as_float(d.x);
    hi.kmax[0] = __uint_as_float(d.y); hi.kmax[1] = __uint_as_float(d.z); hi.kmax[2] = __uint_as_float(d.w);
    hi.pitX = __uint_as_float(e.x);
    hi.pitY = __uint_as_float(e.y);
    hi.ok = (int)e.z;
    hi.pad = 0;
    return hi;
}
static_assert(offsetof(HierItem, level) == 24 && offsetof(HierItem, epsSingle) == 32 && offsetof(HierItem, kmin) == 40 && offsetof(HierItem, kmax) == 52 &&
                  offsetof(HierItem, pitX) == 64 && offsetof(HierItem, ok) == 72,
              "LoadHierItem unpacks this layout");

// warp-aggregated append of (item, idx) for the lanes with `push`
__device__ __forceinline__ void HierAppend(unsigned long long* __restrict__ list, unsigned long long* __restrict__ count, bool push, uint32_t item, uint32_t idx) {
    const uint32_t mask = __ballot_sync(0xFFFFFFFFu, push);
    if (mask == 0) return;
    const uint32_t lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == (uint32_t)(__ffs(mask) - 1)) base = atomicAdd(count, (unsigned long long)__popc(mask));
    base = __shfl_sync(0xFFFFFFFFu, base, __ffs(mask) - 1);
    if (push) list[base + __popc(mask & ((1u << lane) - 1u))] = ((unsigned long long)item << 32) | idx;
}

// state words of region `idx` (size exponent e) of an item whose block starts at `words`
__device__ __forceinline__ void HierFillGlobal(uint32_t* __restrict__ words, uint32_t e, uint32_t idx, uint32_t state) {
    const uint32_t pat = state * 0x55555555u;
    if (e == 3) reinterpret_cast<uint4*>(words)[idx] = make_uint4(pat, pat, pat, pat);
    else if (e == 2) words[idx] = pat;
    else reinterpret_cast<uint8_t*>(words)[idx] = (uint8_t)pat;  // e == 1: four micro-triangles, one byte
}

// Region tests of a whole warp, one region per lane, with the CELLS of all thirty-two footprints spread evenly over the lanes: a
// region has 1 to 16 footprint cells and fails at its first bad one, so evaluating them lane by lane leaves two thirds of the warp
// idle.  Returns this lane's verdict: +1 / -1 = every micro-triangle of the region is on that side, 0 = split it.
template <class Cfg>
__device__ __forceinline__ int WarpTestRegions(const BakeParams& P, const DevMip& m, const HierItem* __restrict__ hierItems, int mip, bool valid, uint32_t w,
                                               const RegionBox& rb) {
    const int M = P.tex.mipCount;
    const uint32_t lane = threadIdx.x & 31;
    const int fw = rb.cx1 - rb.cx0 + 1, fh = rb.cy1 - rb.cy0 + 1;
    int n = valid ? fw * fh : 0;
    int verdict = 0;
    bool decided = !valid;
    if (valid && P.tex.strongPlus != nullptr) {
        const HierItem hi = LoadHierItem(hierItems + (size_t)w * M + mip);
        if (ItemWithinStrongCaps(hi)) {
            verdict = StrongRectSide(P, m, hi, rb, rb.cx0, rb.cy0, rb.cx1, rb.cy1);  // (I)
            if (verdict != 0) {
                decided = true;
                n = 0;
            }
        }
    }
    if (valid && !decided && n > kHierMaxCells) {
        verdict = FlatRectSide<Cfg>(P, m, rb.cx0, rb.cy0, rb.cx1, rb.cy1);  // (H) or split
        decided = true;
        n = 0;
    }
    // inclusive prefix sums of the cell counts
    int incl = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if ((int)lane >= d) incl += o;
    }
    const int excl = incl - n, total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    uint32_t anyPlus = 0, anyMinus = 0, anyFail = 0;
    for (int base = 0; base < total; base += 32) {
        const int task = base + (int)lane;
        const bool active = task < total;
        // owner = the lane whose cell range holds `task`: the number of lanes with incl <= task
        int owner = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            const int v = __shfl_sync(0xFFFFFFFFu, incl, owner + step - 1);
            if (v <= task) owner += step;
        }
        owner = owner > 31 ? 31 : owner;
        RegionBox ob;
        ob.lox = __shfl_sync(0xFFFFFFFFu, rb.lox, owner); ob.loy = __shfl_sync(0xFFFFFFFFu, rb.loy, owner);
        ob.hix = __shfl_sync(0xFFFFFFFFu, rb.hix, owner); ob.hiy = __shfl_sync(0xFFFFFFFFu, rb.hiy, owner);
        ob.r0x = __shfl_sync(0xFFFFFFFFu, rb.r0x, owner); ob.r0y = __shfl_sync(0xFFFFFFFFu, rb.r0y, owner);
        ob.r1x = __shfl_sync(0xFFFFFFFFu, rb.r1x, owner); ob.r1y = __shfl_sync(0xFFFFFFFFu, rb.r1y, owner);
        ob.r2x = __shfl_sync(0xFFFFFFFFu, rb.r2x, owner); ob.r2y = __shfl_sync(0xFFFFFFFFu, rb.r2y, owner);
        ob.eps = __shfl_sync(0xFFFFFFFFu, rb.eps, owner);
        const int ocx0 = __shfl_sync(0xFFFFFFFFu, rb.cx0, owner), ocy0 = __shfl_sync(0xFFFFFFFFu, rb.cy0, owner);
        const int ofw = __shfl_sync(0xFFFFFFFFu, fw, owner), on = __shfl_sync(0xFFFFFFFFu, n, owner), oexcl = __shfl_sync(0xFFFFFFFFu, excl, owner);
        const uint32_t ow = __shfl_sync(0xFFFFFFFFu, w, owner);
        int s = 0;
        if (active) {
            const int local = task - oexcl, ly = local / ofw, lx = local - ly * ofw;
            const HierItem hi = LoadHierItem(hierItems + (size_t)ow * M + mip);
            s = TestRegionCell<Cfg>(P, m, hi, ob, ocx0 + lx, ocy0 + ly, on == 1);
        }
        const uint32_t bp = __ballot_sync(0xFFFFFFFFu, active && s > 0), bm = __ballot_sync(0xFFFFFFFFu, active && s < 0),
                       bf = __ballot_sync(0xFFFFFFFFu, active && s == 0);
        // the tasks of this round that belong to this lane's region
        const int lo = (excl > base ? excl : base) - base, hiE = (incl < base + 32 ? incl : base + 32) - base;
        if (hiE > lo) {
            const uint32_t mine = (hiE - lo == 32 ? 0xFFFFFFFFu : ((1u << (hiE - lo)) - 1u)) << lo;
            anyPlus |= bp & mine;
            anyMinus |= bm & mine;
            anyFail |= bf & mine;
        }
    }
    if (decided) return verdict;
    if (n == 0 || anyFail != 0 || (anyPlus != 0 && anyMinus != 0)) return 0;
    return anyPlus != 0 ? 1 : -1;
}

// The same for every mip of the texture: a region passes when it passes on every mip with the same side (LeafClassifyMips).
template <class Cfg>
__device__ __forceinline__ int WarpTestRegionsAllMips(const BakeParams& P, const HierItem* __restrict__ hierItems, bool valid, uint32_t w, uint32_t idx,
                                                      uint32_t regionLevelBelow) {
    const int M = P.tex.mipCount;
    int verdict = 0;
    for (int k = 0; k < M; ++k) {
        RegionBox rb{};
        bool ok = valid && (k == 0 || verdict != 0);
        if (ok) {
            const HierItem hi = LoadHierItem(hierItems + (size_t)w * M + k);
            ok = MakeRegionBox(P.tex.mips[k], hi, idx, hi.level - regionLevelBelow, rb);  // false: coordinates out of range, the region is split
        }
        const int s = WarpTestRegions<Cfg>(P, P.tex.mips[k], hierItems, k, ok, w, rb);
        if (k == 0) verdict = s;
        else if (s != verdict) verdict = 0;
    }
    return verdict;
}

// One warp per TASK = 64 consecutive initial regions of the chunk (for items of level 6 that is exactly one work item; a level-12
// item is 4096 tasks, sixty-four level-3 items share one).  For every item piece in its task the warp first classifies the cells
// of the piece's footprint as a whole (F): the bitmap answers most region tests without touching the texture again, and a piece
// whose cells are all on one side is finished at once (the common case: most triangles do not meet the level line at all).
constexpr int kHierInitWarps = 4;
constexpr uint32_t kHierTaskRegions = 64;
// resident blocks per SM the list / leaf kernels are compiled for (register budget 64 at 8; A/B builds override)
#ifndef OMM_LIST_MIN_BLOCKS
#define OMM_LIST_MIN_BLOCKS 8
#endif
#ifndef OMM_LEAF_MIN_BLOCKS
#define OMM_LEAF_MIN_BLOCKS 8
#endif

template <class Cfg>
__global__ void __launch_bounds__(kHierInitWarps * 32, 6) HierTestInitial(const BakeParams P, const HierItem* __restrict__ hierItems,
                                                                       const unsigned long long* __restrict__ regionStart,
                                                                       const unsigned long long* __restrict__ wordStart, uint32_t itemBegin, uint32_t itemEnd,
                                                                       HierLists lists, uint32_t* __restrict__ uniformVotes, uint32_t* __restrict__ stateWords) {
    __shared__ uint32_t sPlus[kHierInitWarps][32], sMinus[kHierInitWarps][32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const DevMip& m = P.tex.mips[0];
    const unsigned long long R0 = __ldg(&regionStart[itemBegin]), R1 = __ldg(&regionStart[itemEnd]);
    const unsigned long long numTasks = (R1 - R0 + kHierTaskRegions - 1) / kHierTaskRegions;
    const float itemsPerRegion = (float)(itemEnd - itemBegin) / (float)(R1 - R0);
    // warps stride over the tasks (a grid of resident blocks)
    for (unsigned long long task = (unsigned long long)blockIdx.x * kHierInitWarps + warp; task < numTasks; task += (unsigned long long)gridDim.x * kHierInitWarps) {
        const unsigned long long g0 = R0 + task * kHierTaskRegions, g1 = g0 + kHierTaskRegions < R1 ? g0 + kHierTaskRegions : R1;
        // item of region g0: exact guess when all items of the chunk have the same level, binary search otherwise
        uint32_t w;
        {
            const uint32_t guess = itemBegin + (uint32_t)((float)(g0 - R0) * itemsPerRegion);
            w = guess < itemEnd - 1 ? guess : itemEnd - 1;
            if (!(__ldg(&regionStart[w]) <= g0 && g0 < __ldg(&regionStart[w + 1]))) w = itemBegin + FindItem(regionStart + itemBegin, itemEnd - itemBegin, g0);
        }
        for (; w < itemEnd; ++w) {
            const unsigned long long ws = __ldg(&regionStart[w]);
            if (ws >= g1) break;
            const unsigned long long we = __ldg(&regionStart[w + 1]);
            // piece = regions [a, b) of item w
            const uint32_t a = (uint32_t)((g0 > ws ? g0 : ws) - ws), b = (uint32_t)((g1 < we ? g1 : we) - ws);
            __syncwarp();
            const HierItem hi = LoadHierItem(hierItems + (size_t)w * P.tex.mipCount);
            const uint32_t L = hi.level;
            const uint32_t e = L < 3 ? L : 3;
            uint32_t* words = stateWords + __ldg(&wordStart[w]);
            if (!hi.ok || e == 0) {
                // items the shortcuts do not cover list all their 4-regions (16 per initial region from level 3 on); a level-0 item
                // is one leaf, listed as 4-region 0
                const uint32_t per = L >= 3 ? 16u : (L == 2 ? 4u : 1u);
                const uint32_t n4 = (b - a) * per;
                // level-0 items the shortcuts DO cover go to the ordinary leaf list
                unsigned long long* list = hi.ok ? lists.q[2] : lists.slow;
                unsigned long long* count = hi.ok ? lists.count + 2 : lists.count + 5;
                for (uint32_t base = 0; base < n4; base += 32) HierAppend(list, count, base + lane < n4, w, a * per + base + lane);
                continue;
            }
            // (F) whole-cell bitmap over the footprint of the piece: the whole item, or the aligned node of 64 regions of a bigger item
            ItemCellMap map{0, 0, 0, 0, sPlus[warp], sMinus[warp]};
            RegionBox box;
            bool haveBox = false;
            const bool wholeItem = a == 0 && b == (uint32_t)(we - ws);
            // (the piece-level answers (F), (H), (I) look at mip 0 only: with several mips every region goes to HierTestUnresolved)
            if (P.tex.mipCount == 1) {
                if (L >= 3 && wholeItem) haveBox = MakeItemBox(m, hi, box);
                else if (L > 6 && (a & 63u) == 0 && b - a == 64) haveBox = MakeNodeBox(m, hi, a >> 6, L - 6, box);
            }
            if (haveBox) {
                // (H) the whole piece over a constant area: one table query, whatever its size
                int sFlat = FlatRectSide<Cfg>(P, m, box.cx0, box.cy0, box.cx1, box.cy1);
                if (sFlat == 0 && ItemWithinStrongCaps(hi)) sFlat = StrongRectSide(P, m, hi, box, box.cx0, box.cy0, box.cx1, box.cy1);  // (I)
                if (sFlat != 0) {
                    const uint32_t pat = (uint32_t)(sFlat > 0 ? P.stateGT : P.stateLE) * 0x55555555u;
                    // a WHOLE item proved uniform becomes a special index and its block is never read (ItemPostKernel takes the state
                    // from the votes): skip the write unless special indices are disabled or the host passes want every block
                    if (!(P.skipUniformFill && wholeItem))
                        for (uint32_t i = a + lane; i < b; i += 32) reinterpret_cast<uint4*>(words)[i] = make_uint4(pat, pat, pat, pat);
                    if (lane == 0) atomicAdd(&uniformVotes[2 * (size_t)w + (sFlat > 0 ? 0 : 1)], b - a);
                    continue;
                }
            }
            // (I) answers region queries from the tables when the piece lies inside the texture and within the caps: no bitmap then
            const bool strongOk = ItemWithinStrongCaps(hi) && P.tex.strongPlus != nullptr;
            const bool strongCovers = strongOk && haveBox && box.cx0 >= 0 && box.cy0 >= 0 && box.cx1 <= m.w - 2 && box.cy1 <= m.h - 2 &&
                                      box.hix - box.lox + 1.f + hi.deltaEdge <= kStrongMaxExtent && box.hiy - box.loy + 1.f + hi.deltaEdge <= kStrongMaxExtent;
            if (haveBox && !strongCovers && box.cx1 - box.cx0 < 32 && box.cy1 - box.cy0 < 32) {
                const int fw = box.cx1 - box.cx0 + 1, fh = box.cy1 - box.cy0 + 1;
                sPlus[warp][lane] = 0;
                sMinus[warp][lane] = 0;
                __syncwarp();
                for (int id = (int)lane; id < fw * fh; id += 32) {
                    const int y = id / fw, x = id - y * fw;
                    const int s = WholeCellSide<Cfg>(P, m, hi, box, box.cx0 + x, box.cy0 + y);
                    if (s > 0) atomicOr(&sPlus[warp][y], 1u << x);
                    else if (s < 0) atomicOr(&sMinus[warp][y], 1u << x);
                }
                __syncwarp();
                map.cx0 = box.cx0; map.cy0 = box.cy0; map.fw = fw; map.fh = fh;
                // whole piece on one side?
                const uint32_t rowMask = fw == 32 ? 0xFFFFFFFFu : (1u << fw) - 1u;
                const bool rowPlus = (int)lane >= fh || sPlus[warp][lane] == rowMask, rowMinus = (int)lane >= fh || sMinus[warp][lane] == rowMask;
                const bool allPlus = __all_sync(0xFFFFFFFFu, rowPlus), allMinus = __all_sync(0xFFFFFFFFu, rowMinus);
                if (allPlus || allMinus) {
                    // one 16-byte group per initial region (64 micro-triangles x 2 bits); boxes exist for level >= 3 only
                    const uint32_t pat = (uint32_t)(allPlus ? P.stateGT : P.stateLE) * 0x55555555u;
                    if (!(P.skipUniformFill && wholeItem))
                        for (uint32_t i = a + lane; i < b; i += 32) reinterpret_cast<uint4*>(words)[i] = make_uint4(pat, pat, pat, pat);
                    if (lane == 0) atomicAdd(&uniformVotes[2 * (size_t)w + (allPlus ? 0 : 1)], b - a);
                    continue;
                }
            }
            uint32_t votesUp = 0, votesDown = 0;
            for (uint32_t base = a; base < b; base += 32) {
                const uint32_t idx = base + lane;
                const bool valid = idx < b;
                int s = 0;
                if (valid) {
                    RegionBox rb;
                    if (map.fw != 0 || strongOk) {
                        if (MakeRegionBox(m, hi, idx, L - e, rb)) {
                            if (strongOk) s = StrongRectSide(P, m, hi, rb, rb.cx0, rb.cy0, rb.cx1, rb.cy1);  // (I)
                            if (s == 0) s = LookupCellMap(map, rb);
                        }
                    }
                    if (s != 0) HierFillGlobal(words, e, idx, (uint32_t)(s > 0 ? P.stateGT : P.stateLE));
                }
                votesUp += __popc(__ballot_sync(0xFFFFFFFFu, s > 0));
                votesDown += __popc(__ballot_sync(0xFFFFFFFFu, s < 0));
                // regions the bitmap does not answer get the full test in HierTestUnresolved, where every lane has one
                HierAppend(lists.unresolved, lists.count + 3, valid && s == 0, w, idx);
            }
            if (lane == 0) {
                if (votesUp) atomicAdd(&uniformVotes[2 * (size_t)w], votesUp);
                if (votesDown) atomicAdd(&uniformVotes[2 * (size_t)w + 1], votesDown);
            }
        }
    }
}

// children of the regions in lists.q[src] (size exponent 3 - src); the children have size exponent 2 - src
template <class Cfg>
__global__ void __launch_bounds__(128, OMM_LIST_MIN_BLOCKS) HierTestList(const BakeParams P, const HierItem* __restrict__ hierItems, const unsigned long long* __restrict__ wordStart,
                                                     const unsigned long long* __restrict__ inList, const unsigned long long* __restrict__ inCount,
                                                     unsigned long long* __restrict__ outList, unsigned long long* __restrict__ outCount, int src,
                                                     uint32_t* __restrict__ stateWords) {
    const unsigned long long total = *inCount * 4ull;
    const unsigned long long rounded = (total + 31ull) & ~31ull;
    const uint32_t e = 2u - (uint32_t)src;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < rounded; t += (unsigned long long)gridDim.x * blockDim.x) {
        const bool valid = t < total;
        uint32_t w = 0, idx = 0;
        if (valid) {
            const unsigned long long entry = inList[t >> 2];
            w = (uint32_t)(entry >> 32);
            idx = (uint32_t)entry * 4u + (uint32_t)(t & 3ull);
        }
        const int s = WarpTestRegionsAllMips<Cfg>(P, hierItems, valid, w, idx, e);
        if (s != 0) HierFillGlobal(stateWords + __ldg(&wordStart[w]), e, idx, (uint32_t)(s > 0 ? P.stateGT : P.stateLE));
        HierAppend(outList, outCount, t < total && s == 0, w, idx);
    }
}

// Initial regions the whole-cell bitmap left open: the full region test, one region per thread.
template <class Cfg>
__global__ void __launch_bounds__(128, OMM_LIST_MIN_BLOCKS) HierTestUnresolved(const BakeParams P, const HierItem* __restrict__ hierItems, const unsigned long long* __restrict__ wordStart,
                                                           HierLists lists, uint32_t* __restrict__ uniformVotes, uint32_t* __restrict__ stateWords) {
    const unsigned long long total = lists.count[3];
    const unsigned long long rounded = (total + 31ull) & ~31ull;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < rounded; t += (unsigned long long)gridDim.x * blockDim.x) {
        const bool valid = t < total;
        uint32_t w = 0, idx = 0, e = 3;
        if (valid) {
            const unsigned long long entry = lists.unresolved[t];
            w = (uint32_t)(entry >> 32);
            idx = (uint32_t)entry;
            const uint32_t level = __ldg(&hierItems[(size_t)w * P.tex.mipCount].level);
            e = level < 3 ? level : 3;
        }
        const int s = WarpTestRegionsAllMips<Cfg>(P, hierItems, valid, w, idx, e);
        if (s != 0) {
            HierFillGlobal(stateWords + __ldg(&wordStart[w]), e, idx, (uint32_t)(s > 0 ? P.stateGT : P.stateLE));
            atomicAdd(&uniformVotes[2 * (size_t)w + (s > 0 ? 0 : 1)], 1u);
        }
        const bool fail = t < total && s == 0;
        // e == 3 -> list 0, e == 2 -> list 1, e == 1 -> list 2
        HierAppend(lists.q[0], lists.count + 0, fail && e == 3, w, idx);
        HierAppend(lists.q[1], lists.count + 1, fail && e == 2, w, idx);
        HierAppend(lists.q[2], lists.count + 2, fail && e == 1, w, idx);
    }
}

template <class Cfg>
__global__ void __launch_bounds__(128, OMM_LEAF_MIN_BLOCKS) HierLeaves(const BakeParams P, const ItemRec* __restrict__ items, const HierItem* __restrict__ hierItems,
                                                   const unsigned long long* __restrict__ wordStart, HierLists lists,
                                                   uint32_t* __restrict__ stateWords) {
    const unsigned long long total = lists.count[2] * 4ull;
    const unsigned long long rounded = (total + 31ull) & ~31ull;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < rounded; t += (unsigned long long)gridDim.x * blockDim.x) {
        uint32_t bits = 0, w = 0, region = 0;
        if (t < total) {
            const unsigned long long entry = lists.q[2][t >> 2];
            w = (uint32_t)(entry >> 32);
            region = (uint32_t)entry;
            const uint32_t k = (uint32_t)(t & 3ull), index = region * 4u + k;
            const int M = P.tex.mipCount;
            const HierItem* its = hierItems + (size_t)w * M;
            const HierItem hi = LoadHierItem(its);
            uint32_t st = 0;
            if (index < (1u << (2 * hi.level))) {  // a level-0 item has one micro-triangle in its only "4-region"
                if (hi.ok) {
                    // (a single-micro-triangle TestRegion first was measured slower: the leaf's own edge filter (D) does the same work.  So was, in
                    // round 2, a warp-synchronous walk that deals the three edge tests of every lane needing them out over all 32 lanes each cell
                    // iteration: parity-green, but 14.9 vs 12.5 ms classify on the same box -- an iteration has ~12 requesting lanes, i.e. two dense
                    // rounds against the ~2.3 the short-circuiting per-lane calls take, and the shuffles and spills cost more than that saves.)
                    if (M == 1) st = (uint32_t)LeafClassify<Cfg>(P, P.tex.mips[0], hi, index);
                    else st = (uint32_t)LeafClassifyMips<Cfg>(P, [&](int k) { return k == 0 ? hi : LoadHierItem(its + k); }, index);
                }
            }
            bits = st << (2 * k);
        }
        bits |= __shfl_xor_sync(0xFFFFFFFFu, bits, 1);
        bits |= __shfl_xor_sync(0xFFFFFFFFu, bits, 2);
        if (t < total && (t & 3ull) == 0) reinterpret_cast<uint8_t*>(stateWords + __ldg(&wordStart[w]))[region] = (uint8_t)bits;
    }
}

// Micro-triangles of the items the shortcuts do not cover (zero-area UV triangles, non-finite or huge coordinates): the generic
// reference walk, in a kernel of its own so that HierLeaves stays small.
template <class Cfg>
__global__ void __launch_bounds__(128) HierLeavesSlow(const BakeParams P, const ItemRec* __restrict__ items, const HierItem* __restrict__ hierItems,
                                                       const unsigned long long* __restrict__ wordStart, HierLists lists, uint32_t* __restrict__ stateWords) {
    const unsigned long long total = lists.count[5] * 4ull;
    const unsigned long long rounded = (total + 31ull) & ~31ull;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < rounded; t += (unsigned long long)gridDim.x * blockDim.x) {
        uint32_t bits = 0, w = 0, region = 0;
        if (t < total) {
            const unsigned long long entry = lists.slow[t >> 2];
            w = (uint32_t)(entry >> 32);
            region = (uint32_t)entry;
            const uint32_t k = (uint32_t)(t & 3ull), index = region * 4u + k;
            const ItemRec it = items[w];
            uint32_t st = 0;
            if (index < (1u << (2 * it.level))) st = (uint32_t)ClassifyMicroTriangle<Cfg>(P, it.p0, it.p1, it.p2, it.degenerate != 0, index, it.level);
            bits = st << (2 * k);
        }
        bits |= __shfl_xor_sync(0xFFFFFFFFu, bits, 1);
        bits |= __shfl_xor_sync(0xFFFFFFFFu, bits, 2);
        if (t < total && (t & 3ull) == 0) reinterpret_cast<uint8_t*>(stateWords + __ldg(&wordStart[w]))[region] = (uint8_t)bits;
    }
}

struct HierKernels {
    void (*initial)(const BakeParams, const HierItem*, const unsigned long long*, const unsigned long long*, uint32_t, uint32_t, HierLists, uint32_t*, uint32_t*);
    void (*list)(const BakeParams, const HierItem*, const unsigned long long*, const unsigned long long*, const unsigned long long*, unsigned long long*,
                 unsigned long long*, int, uint32_t*);
    void (*leaves)(const BakeParams, const ItemRec*, const HierItem*, const unsigned long long*, HierLists, uint32_t*);
    void (*unresolved)(const BakeParams, const HierItem*, const unsigned long long*, HierLists, uint32_t*, uint32_t*);
    void (*leavesSlow)(const BakeParams, const ItemRec*, const HierItem*, const unsigned long long*, HierLists, uint32_t*);
};
template <class Cfg>
static HierKernels MakeHierKernels() {
    return HierKernels{HierTestInitial<Cfg>, HierTestList<Cfg>, HierLeaves<Cfg>, HierTestUnresolved<Cfg>, HierLeavesSlow<Cfg>};
}

// OMM_B200_CLASSIFIER=flat|queue selects the older kernels (A/B measurements and parity cross-checks); default = hierarchical.
static int ClassifierOverride() {
    static const int v = [] {
        const char* e = getenv("OMM_B200_CLASSIFIER");
        if (!e) return 0;
        if (!strcmp(e, "flat")) return 1;
        if (!strcmp(e, "queue")) return 2;
        return 0;
    }();
    return v;
}
static bool SelectHierKernels(const BakeParams& P, HierKernels* out) {
    if (ClassifierOverride() != 0) return false;
    if (!(P.filterLinear && !P.disableLevelLine && !P.disableFine)) return false;
    if (P.useCoarse && !P.coarseSameCutoff) return false;  // SAT of another cutoff than the bake's: see LeafClassify
    bool pow2 = true;
    for (int i = 0; i < P.tex.mipCount; ++i) pow2 = pow2 && P.tex.mips[i].isPow2 != 0;
    if (P.tex.isFp32) {
        if (P.addrMode == ommTextureAddressMode_Wrap && pow2) *out = MakeHierKernels<KernelCfg<kAddrWrapPow2, true>>();
        else if (P.addrMode == ommTextureAddressMode_Clamp) *out = MakeHierKernels<KernelCfg<kAddrClamp, true>>();
        else *out = MakeHierKernels<KernelCfg<kAddrGeneric, true>>();
    } else {
        if (P.addrMode == ommTextureAddressMode_Wrap && pow2) *out = MakeHierKernels<KernelCfg<kAddrWrapPow2, false>>();
        else *out = MakeHierKernels<KernelCfg<kAddrGeneric, false>>();
    }
    return true;
}
typedef void (*ClassifyFn)(const BakeParams, const ItemRec*, const unsigned long long*, const unsigned long long*, uint32_t, uint32_t, unsigned long long,
                           unsigned long long, uint32_t*);
static bool UseQueueKernel(const BakeParams& P) {
    if (ClassifierOverride() == 1) return false;
    return P.filterLinear && !P.disableLevelLine && !P.disableFine && P.tex.mipCount == 1;
}
// Picks the compile-time specialisation matching the sampler / texture; every other combination runs the generic kernel.
static ClassifyFn SelectClassifyKernel(const BakeParams& P) {
    const bool allPow2 = [&] {
        for (int i = 0; i < P.tex.mipCount; ++i)
            if (!P.tex.mips[i].isPow2) return false;
        return true;
    }();
    if (UseQueueKernel(P)) {
        if (P.tex.isFp32) {
            if (P.addrMode == ommTextureAddressMode_Wrap && allPow2) return ClassifyKernelQ<KernelCfg<kAddrWrapPow2, true>>;
            if (P.addrMode == ommTextureAddressMode_Clamp) return ClassifyKernelQ<KernelCfg<kAddrClamp, true>>;
            return ClassifyKernelQ<KernelCfg<kAddrGeneric, true>>;
        }
        if (P.addrMode == ommTextureAddressMode_Wrap && allPow2) return ClassifyKernelQ<KernelCfg<kAddrWrapPow2, false>>;
        return ClassifyKernelQ<KernelCfg<kAddrGeneric, false>>;
    }
    if (P.tex.isFp32) {
        if (P.addrMode == ommTextureAddressMode_Wrap && allPow2) return ClassifyKernel<KernelCfg<kAddrWrapPow2, true>>;
        if (P.addrMode == ommTextureAddressMode_Clamp) return ClassifyKernel<KernelCfg<kAddrClamp, true>>;
        return ClassifyKernel<KernelCfg<kAddrGeneric, true>>;
    }
    if (P.addrMode == ommTextureAddressMode_Wrap && allPow2) return ClassifyKernel<KernelCfg<kAddrWrapPow2, false>>;
    return ClassifyKernel<KernelCfg<kAddrGeneric, false>>;
}

// ---------------------------------------------------------------------------------------------------------------------
// K5: per-item post pass, one warp per item: uniform-state / rejection test (ref: bake_cpu_impl.cpp:1432-1472) and
// XXH64(seed 42) of the item's 3-state byte array (UT folded into UO, one byte per micro-triangle; ref: :374-377,
// :1038-1040).  XXH64 follows the published specification (xxHash doc/xxhash_spec.md; SDK pins submodule c961fbe6).
// ---------------------------------------------------------------------------------------------------------------------
constexpr uint64_t XP1 = 0x9E3779B185EBCA87ull, XP2 = 0xC2B2AE3D27D4EB4Full, XP3 = 0x165667B19E3779F9ull, XP4 = 0x85EBCA77C2B2AE63ull,
                   XP5 = 0x27D4EB2F165667C5ull;
__device__ __forceinline__ uint64_t Rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
__device__ __forceinline__ uint64_t XxhRound(uint64_t acc, uint64_t in) { return Rotl64(acc + in * XP2, 31) * XP1; }
__device__ __forceinline__ uint64_t XxhMerge(uint64_t acc, uint64_t v) { return (acc ^ XxhRound(0, v)) * XP1 + XP4; }
__device__ __forceinline__ uint64_t XxhAvalanche(uint64_t h) {
    h ^= h >> 33; h *= XP2; h ^= h >> 29; h *= XP3; h ^= h >> 32;
    return h;
}
// 8 two-bit states (16 bits) -> 8 bytes of 3-state values {0,1,3}
__device__ __forceinline__ uint64_t Expand3State(uint32_t bits16) {
    uint64_t v = (uint64_t)(bits16 & 0xFFu) | ((uint64_t)(bits16 & 0xFF00u) << 24);
    v = (v | (v << 12)) & 0x000F000F000F000Full;
    v = (v | (v << 6)) & 0x0303030303030303ull;
    return v | ((v >> 1) & 0x0101010101010101ull);
}

// XXH64(seed 42) of 4^level bytes that all hold the 3-state value v (0, 1 or 3): the digest of every work item whose
// micro-triangles share one state.  Constants of the algorithm, computed once per process on the host.
struct UniformDigests {
    uint64_t h[3][13];  // [0] Transparent, [1] Opaque, [2] either Unknown state (both hash as 3, ref: bake_cpu_impl.cpp:374-377)
};
static uint64_t HostRotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static uint64_t HostXxhRound(uint64_t acc, uint64_t in) { return HostRotl64(acc + in * XP2, 31) * XP1; }
static uint64_t HostXxhMerge(uint64_t acc, uint64_t v) { return (acc ^ HostXxhRound(0, v)) * XP1 + XP4; }
static uint64_t HostXxh64Constant(uint8_t value, uint64_t n) {
    const uint64_t c8 = 0x0101010101010101ull * value;
    uint64_t h;
    uint64_t left = n;
    if (n >= 32) {
        uint64_t v1 = 42ull + XP1 + XP2, v2 = 42ull + XP2, v3 = 42ull, v4 = 42ull - XP1;
        for (uint64_t i = 0; i < n / 32; ++i) {
            v1 = HostXxhRound(v1, c8); v2 = HostXxhRound(v2, c8); v3 = HostXxhRound(v3, c8); v4 = HostXxhRound(v4, c8);
        }
        h = HostRotl64(v1, 1) + HostRotl64(v2, 7) + HostRotl64(v3, 12) + HostRotl64(v4, 18);
        h = HostXxhMerge(h, v1); h = HostXxhMerge(h, v2); h = HostXxhMerge(h, v3); h = HostXxhMerge(h, v4);
        left = n % 32;
    } else
        h = 42ull + XP5;
    h += n;
    for (; left >= 8; left -= 8) { h ^= HostXxhRound(0, c8); h = HostRotl64(h, 27) * XP1 + XP4; }
    if (left >= 4) { h ^= (c8 & 0xFFFFFFFFull) * XP1; h = HostRotl64(h, 23) * XP2 + XP3; left -= 4; }
    for (; left > 0; --left) { h ^= (uint64_t)value * XP5; h = HostRotl64(h, 11) * XP1; }
    h ^= h >> 33; h *= XP2; h ^= h >> 29; h *= XP3; h ^= h >> 32;
    return h;
}
static const UniformDigests& GetUniformDigests() {
    static UniformDigests table;
    static std::once_flag once;
    std::call_once(once, [] {
        const uint8_t values[3] = {0, 1, 3};
        for (int v = 0; v < 3; ++v)
            for (int l = 0; l <= kMaxLevel; ++l) table.h[v][l] = HostXxh64Constant(values[v], 1ull << (2 * l));
    });
    return table;
}

// Four lanes per work item: lane j carries XXH64 accumulator j and consumes bytes [8j, 8j+8) of every 32-byte stripe (= 8 two-bit
// states = half a state word), so the four accumulators of an item advance in parallel and a warp hashes eight items at once.
constexpr int kItemPostItemsPerBlock = 64;
constexpr uint32_t kBigHashLevel = 9;  // 4^9 bytes = 8192 stripes
__global__ void __launch_bounds__(kItemPostItemsPerBlock * 4) ItemPostKernel(const ItemRec* __restrict__ items, const unsigned long long* __restrict__ wordStart,
                                                      const uint32_t* __restrict__ stateWords, uint32_t itemBegin, uint32_t itemEnd, float rejectionThreshold,
                                                      int disableSpecial, int keepExistingSpecial, const uint32_t* __restrict__ uniformVotes,
                                                      uint32_t stateGT, uint32_t stateLE, const UniformDigests table, uint64_t* __restrict__ digest,
                                                      int32_t* special, uint32_t* __restrict__ bigList, uint32_t* __restrict__ bigCount) {
    // Phase 1, one thread per item of the block's range: items the hierarchical classifier proved uniform (all initial regions on
    // one side, HierTestInitial) get their constant digest without reading anything; the others are compacted into a list.
    __shared__ uint32_t sList[kItemPostItemsPerBlock * 4];
    __shared__ uint32_t sCount;
    if (threadIdx.x == 0) sCount = 0;
    __syncthreads();
    {
        const uint32_t wi = itemBegin + blockIdx.x * (kItemPostItemsPerBlock * 4) + threadIdx.x;
        if (wi < itemEnd) {
            bool done = false;
            if (uniformVotes) {
                const uint32_t lv = items[wi].level;
                const uint32_t nInit = lv > 3 ? 1u << (2 * (lv - 3)) : 1u;
                const uint32_t above = __ldg(&uniformVotes[2 * (size_t)wi]), below = __ldg(&uniformVotes[2 * (size_t)wi + 1]);
                if (above == nInit || below == nInit) {
                    const uint32_t s = above == nInit ? stateGT : stateLE;  // the block itself may not have been written (skipUniformFill)
                    digest[wi] = table.h[s >= 2 ? 2 : s][lv];
                    special[wi] = disableSpecial ? 0 : -(int32_t)s - 1;
                    done = true;
                }
            }
            if (!done) {
                // blocks of 256 KiB and more: a kernel of their own (ItemPostBigPipelined) -- here their one sequential XXH64 chain would
                // hold up the seven items sharing the warp, and runs at half the speed
                if (bigList && items[wi].hashLevel >= kBigHashLevel) bigList[atomicAdd(bigCount, 1u)] = wi;
                else sList[atomicAdd(&sCount, 1u)] = wi;
            }
        }
    }
    __syncthreads();
    // Phase 2, four lanes per listed item
    const uint32_t lane = threadIdx.x & 31, j = lane & 3u, quadBase = lane & ~3u, quadMask = 0xFu << quadBase;
    for (uint32_t slot = threadIdx.x >> 2; slot < sCount; slot += kItemPostItemsPerBlock) {
    const uint32_t w = sList[slot];
    const uint32_t level = items[w].level;
    const uint32_t n = 1u << (2 * level);                       // micro-triangles of the item now (uniformity / rejection test)
    const uint32_t nHash = 1u << (2 * items[w].hashLevel);      // bytes the SDK's digest covers (>= n, differs only after Compress)
    const uint32_t* words = stateWords + wordStart[w];
    const uint32_t word0 = __ldg(words);
    const uint32_t s0 = word0 & 3u;
    const uint32_t fullWords = n >> 4;                           // words entirely inside the first n fields
    const uint32_t headMask = n >= 16 ? 0xFFFFFFFFu : ((1u << (2 * n)) - 1u);  // valid fields of word 0 when n < 16

    uint64_t h;
    bool allEqual;
    uint32_t known;
    if (nHash < 32) {
        // levels 0..2: 1, 4 or 16 micro-triangles in one word
        allEqual = ((word0 ^ (s0 * 0x55555555u)) & headMask) == 0;
        known = __popc(~(word0 >> 1) & 0x55555555u & headMask);
        const uint64_t lo = Expand3State(word0 & 0xFFFFu), hi = Expand3State(word0 >> 16);
        h = 42ull + XP5 + (uint64_t)nHash;
        if (nHash == 16) {
            h ^= XxhRound(0, lo); h = Rotl64(h, 27) * XP1 + XP4;
            h ^= XxhRound(0, hi); h = Rotl64(h, 27) * XP1 + XP4;
        } else if (nHash == 4) {
            h ^= (lo & 0xFFFFFFFFull) * XP1; h = Rotl64(h, 23) * XP2 + XP3;
        } else {
            h ^= (lo & 0xFFull) * XP5; h = Rotl64(h, 11) * XP1;
        }
        h = XxhAvalanche(h);
    } else {
        const uint32_t numWords = nHash >> 4;  // a multiple of 4: the block is read 16 bytes (two stripes) at a time
        const uint32_t pattern = s0 * 0x55555555u;
        uint32_t diff = 0;
        known = 0;
        uint64_t acc = j == 0 ? 42ull + XP1 + XP2 : (j == 1 ? 42ull + XP2 : (j == 2 ? 42ull : 42ull - XP1));
        const uint4* words4 = reinterpret_cast<const uint4*>(words);
        const uint32_t half = j & 1u, pair = j >> 1;
#pragma unroll 4
        for (uint32_t g = 0; g < (numWords >> 2); ++g) {
            const uint4 v = __ldg(words4 + g);
            if (j == 0) {
                // uniformity / known-state statistics over the first n micro-triangles, once per word
                const uint32_t wi = 4 * g;
                const uint32_t m0 = wi < fullWords ? 0xFFFFFFFFu : ((wi == 0 && n < 16) ? headMask : 0u);
                const uint32_t m1 = wi + 1 < fullWords ? 0xFFFFFFFFu : 0u, m2 = wi + 2 < fullWords ? 0xFFFFFFFFu : 0u, m3 = wi + 3 < fullWords ? 0xFFFFFFFFu : 0u;
                diff |= ((v.x ^ pattern) & m0) | ((v.y ^ pattern) & m1) | ((v.z ^ pattern) & m2) | ((v.w ^ pattern) & m3);
                known += __popc(~(v.x >> 1) & 0x55555555u & m0) + __popc(~(v.y >> 1) & 0x55555555u & m1) + __popc(~(v.z >> 1) & 0x55555555u & m2) +
                         __popc(~(v.w >> 1) & 0x55555555u & m3);
            }
            const uint32_t a0 = pair ? v.y : v.x, a1 = pair ? v.w : v.z;  // the word of this lane in stripe 2g and in stripe 2g + 1
            acc = XxhRound(acc, Expand3State(half ? (a0 >> 16) : (a0 & 0xFFFFu)));
            acc = XxhRound(acc, Expand3State(half ? (a1 >> 16) : (a1 & 0xFFFFu)));
        }
        allEqual = diff == 0;  // meaningful on lane j == 0 only, which is the lane that writes
        const uint64_t v1 = __shfl_sync(quadMask, acc, quadBase), v2 = __shfl_sync(quadMask, acc, quadBase + 1), v3 = __shfl_sync(quadMask, acc, quadBase + 2),
                       v4 = __shfl_sync(quadMask, acc, quadBase + 3);
        h = Rotl64(v1, 1) + Rotl64(v2, 7) + Rotl64(v3, 12) + Rotl64(v4, 18);
        h = XxhMerge(h, v1); h = XxhMerge(h, v2); h = XxhMerge(h, v3); h = XxhMerge(h, v4);
        h += (uint64_t)nHash;
        h = XxhAvalanche(h);
    }
    if (j == 0) {
        int common = (int)s0;
        if (!allEqual && rejectionThreshold > 0.f) {
            const float frac = (float)known / (float)n;
            if (frac < rejectionThreshold) {
                allEqual = true;
                common = ommOpacityState_UnknownTransparent;
            }
        }
        digest[w] = h;
        // second promotion pass (ref: bake_cpu_impl.cpp:1439-1440): items that already carry a special index are left alone
        if (!(keepExistingSpecial && special[w] != 0)) special[w] = (allEqual && !disableSpecial) ? (-common - 1) : 0;
    }
    }
}

// The same for ONE big block per warp.  XXH64 is four sequential chains  acc <- rotl(acc + in * P2, 31) * P1  over the 32-byte stripes, and a
// level-12 block has 524 288 stripes: the digest of such a block is a latency chain nothing can shorten (the rotation does not commute with
// the carries of the addition, so there is no parallel-prefix form).  What CAN be removed is everything else from the chain's issue slots: the
// 32 lanes load, expand and pre-multiply eight stripes at a time (lane = stripe * 4 + accumulator), and the chain itself -- kept
// redundantly in every lane for accumulator lane & 3, so there is no divergence -- is a shuffle, a fused multiply-add and two funnel shifts
// per stripe.  Config 5 (one level-12 item): 13.7 ms -> see DESIGN.md section 6.
// One step of that chain, s <- rotl(s, 31) * P1 + x.  kFunnel: both halves of the rotation are ONE funnel shift each and the product is split
// by hand (wide product of the low words with x as its addend, the two cross products added into the high word): six instructions instead
// of the eight the compiler builds from the 64-bit expression (it assembles the low half of the rotation from a multiply, a shift and an
// OR).  Same value for every input (200 M random and edge-pattern inputs on the host; the GPU suite's big-level bakes).  A three-level
// arrangement of the step (rotl(s, 31) * P1 = s * (2^31 P1) + (s >> 33) * P1, everything that depends on the low word alone folded into the
// addend of the wide product) was built as well: ten instructions, and slower -- a single warp is bound by the instructions it can issue,
// not by the depth of the chain (measurements below).
template <bool kFunnel>
__device__ __forceinline__ uint64_t XxhChainStep(uint64_t s, uint64_t x) {
    return kFunnel ? xxh::ChainStep(s, x) : Rotl64(s, 31) * XP1 + x;  // (the funnel form lives in omm_xxh64.h, where the CPU suite pins it)
}
template <bool kFunnel>
__global__ void __launch_bounds__(32) ItemPostBigKernel(const ItemRec* __restrict__ items, const unsigned long long* __restrict__ wordStart,
                                                        const uint32_t* __restrict__ stateWords, const uint32_t* __restrict__ bigList,
                                                        const uint32_t* __restrict__ bigCount, float rejectionThreshold, int disableSpecial, int keepExistingSpecial,
                                                        uint64_t* __restrict__ digest, int32_t* special) {
    if (blockIdx.x >= *bigCount) return;
    const uint32_t w = bigList[blockIdx.x], lane = threadIdx.x;
    const uint32_t level = items[w].level;
    const uint32_t n = 1u << (2 * level);                   // micro-triangles of the item now (uniformity / rejection test)
    const uint32_t nHash = 1u << (2 * items[w].hashLevel);  // bytes the SDK's digest covers (>= n, differs only after Compress)
    const uint32_t* words = stateWords + wordStart[w];
    const uint32_t fullWords = n >> 4, numBatches = nHash >> 8;  // a batch = 8 stripes = 256 bytes = 16 state words
    const uint32_t j = lane & 3u, t = lane >> 2, half = j & 1u, wordInBatch = 2u * t + (j >> 1);
    const uint32_t s0 = __ldg(words) & 3u, pattern = s0 * 0x55555555u;
    uint32_t diff = 0, known = 0;
    auto fetch = [&](uint32_t b) -> uint64_t {
        const uint32_t wi = 16u * b + wordInBatch;
        const uint32_t v = __ldg(words + wi);
        if (half == 0 && wi < fullWords) {  // statistics once per word
            diff |= v ^ pattern;
            known += __popc(~(v >> 1) & 0x55555555u);
        }
        return Expand3State(half ? (v >> 16) : (v & 0xFFFFu)) * XP2;
    };
    uint64_t acc = j == 0 ? 42ull + XP1 + XP2 : (j == 1 ? 42ull + XP2 : (j == 2 ? 42ull : 42ull - XP1));
    uint64_t x = fetch(0);
    for (uint32_t b = 0; b < numBatches; ++b) {
        const uint64_t xNext = b + 1 < numBatches ? fetch(b + 1) : 0ull;
        uint64_t xs[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) xs[q] = __shfl_sync(0xFFFFFFFFu, x, 4 * q + (int)j);
        // s = acc + in * P2 of the pending round; every step is rotl(s, 31) * P1 + (next in * P2): one fused multiply-add on the chain
        uint64_t s = acc + xs[0];
#pragma unroll
        for (int q = 1; q < 8; ++q) s = XxhChainStep<kFunnel>(s, xs[q]);
        acc = XxhChainStep<kFunnel>(s, 0ull);
        x = xNext;
    }
    diff = __reduce_or_sync(0xFFFFFFFFu, diff);
    known = __reduce_add_sync(0xFFFFFFFFu, known);
    const uint64_t v1 = __shfl_sync(0xFFFFFFFFu, acc, 0), v2 = __shfl_sync(0xFFFFFFFFu, acc, 1), v3 = __shfl_sync(0xFFFFFFFFu, acc, 2), v4 = __shfl_sync(0xFFFFFFFFu, acc, 3);
    if (lane != 0) return;
    uint64_t h = Rotl64(v1, 1) + Rotl64(v2, 7) + Rotl64(v3, 12) + Rotl64(v4, 18);
    h = XxhMerge(h, v1); h = XxhMerge(h, v2); h = XxhMerge(h, v3); h = XxhMerge(h, v4);
    h += (uint64_t)nHash;
    h = XxhAvalanche(h);
    bool allEqual = diff == 0;
    int common = (int)s0;
    if (!allEqual && rejectionThreshold > 0.f) {
        const float frac = (float)known / (float)n;
        if (frac < rejectionThreshold) {
            allEqual = true;
            common = ommOpacityState_UnknownTransparent;
        }
    }
    digest[w] = h;
    if (!(keepExistingSpecial && special[w] != 0)) special[w] = (allEqual && !disableSpecial) ? (-common - 1) : 0;
}

// The same with the chain in a warp of its own (the default).  The one-warp kernel above spends 23 cycles per stripe whatever the form of the step: its
// warp also loads, expands, pre-multiplies and shuffles, 11-15 instructions per stripe, and one warp issues about one instruction every
// two cycles -- the chain is bound by issue slots, not by its depth.  So the preparation moves to a second warp (another scheduler of the SM): it
// writes  in * P2  for 64 stripes at a time into one half of a double buffer in shared memory while the chain warp consumes the other half,
// one 8-byte shared load and one six-instruction step per stripe; the two meet at a block barrier per 64 stripes, both running the same
// loop, so the barrier counts cannot differ.  The chain starts from the state whose step with the first stripe gives  acc0 + x0  (P1 is odd:
// the step is invertible), so there is no special case for the first stripe.  Config 5 (one level-12 block = 524 288 stripes), same box, one
// process (`scripts/gpu_r2x.sh`, `gpu_r2z.sh`; every variant reproduces the SDK digest): post pass 6.39 ms with the compiler's step, 6.07 with
// the funnel step, 6.62 with the three-level step (all one warp); 2.14-2.31 ms with the chain warp (2.41 / 2.53 with the three-level steps
// in it), **1.86 ms** (7.0 cycles per stripe) with the prefetching producer below.  Config 5: 7.56 -> 3.07 ms per bake, 8.4 -> 3.9 ms end to end.
// kStripes = stripes per half of the double buffer.  kPrefetch: the producer keeps the state words of the group after next in registers, so
// that a group's loads have a whole group time to arrive -- without it the producer's path per group is L2 latency + its own arithmetic, and
// with the block's data in the far L2 partition that exceeded the chain warp's 512 cycles per 64 stripes (post pass of config 5: 2.1 ms in one
// process, 3.5 ms in another, same kernel).
template <int kStripes, bool kPrefetch>
__global__ void __launch_bounds__(64) ItemPostBigPipelined(const ItemRec* __restrict__ items, const unsigned long long* __restrict__ wordStart,
                                                           const uint32_t* __restrict__ stateWords, const uint32_t* __restrict__ bigList,
                                                           const uint32_t* __restrict__ bigCount, float rejectionThreshold, int disableSpecial,
                                                           int keepExistingSpecial, uint64_t* __restrict__ digest, int32_t* special) {
    constexpr uint32_t kBatches = kStripes / 8;  // a batch = 8 stripes = 16 state words: one word per lane (each word is read by two lanes)
    __shared__ uint2 sx[2][kStripes][4];         // in * P2 per stripe and accumulator
    __shared__ uint32_t sStats[2];
    if (blockIdx.x >= *bigCount) return;  // (the whole block)
    const uint32_t w = bigList[blockIdx.x], lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t level = items[w].level;
    const uint32_t n = 1u << (2 * level);                   // micro-triangles of the item now (uniformity / rejection test)
    const uint32_t nHash = 1u << (2 * items[w].hashLevel);  // bytes the SDK's digest covers (>= n, differs only after Compress)
    const uint32_t* words = stateWords + wordStart[w];
    const uint32_t fullWords = n >> 4;
    const uint32_t numGroups = nHash / (32u * kStripes);  // hashLevel >= 9: at least 64 groups
    const uint32_t j = lane & 3u, t = lane >> 2, half = j & 1u, wordInBatch = 2u * t + (j >> 1);  // producer: lane = (stripe of the batch, accumulator)
    const uint32_t s0 = __ldg(words) & 3u, pattern = s0 * 0x55555555u;
    uint32_t diff = 0, known = 0;
    uint32_t pf[kBatches];  // the producer's words of one group (registers: every loop over it is unrolled)
    auto loadGroup = [&](uint32_t g) {
        const uint32_t firstWord = g * (kStripes * 2u) + wordInBatch;
#pragma unroll
        for (uint32_t b = 0; b < kBatches; ++b) pf[b] = __ldg(words + firstWord + 16u * b);
    };
    auto storeGroup = [&](uint32_t g) {  // group g -> buffer g & 1
        uint2(*buf)[4] = sx[g & 1u];
        const uint32_t firstWord = g * (kStripes * 2u) + wordInBatch;
#pragma unroll
        for (uint32_t b = 0; b < kBatches; ++b) {
            const uint32_t v = pf[b];
            if (half == 0 && firstWord + 16u * b < fullWords) {  // statistics once per word
                diff |= v ^ pattern;
                known += __popc(~(v >> 1) & 0x55555555u);
            }
            const uint64_t x = Expand3State(half ? (v >> 16) : (v & 0xFFFFu)) * XP2;
            buf[8u * b + t][j] = make_uint2((uint32_t)x, (uint32_t)(x >> 32));
        }
    };
    // chain warp: lane & 3 = accumulator (kept redundantly in all lanes: no divergence)
    const uint64_t acc0 = j == 0 ? 42ull + XP1 + XP2 : (j == 1 ? 42ull + XP2 : (j == 2 ? 42ull : 42ull - XP1));
    uint64_t s = xxh::ChainStart(acc0);  // its step with the first stripe gives acc0 + x0
    if (warp == 1) {
        loadGroup(0);
        storeGroup(0);
        if (kPrefetch && numGroups > 1) loadGroup(1);
    }
    __syncthreads();
    for (uint32_t g = 0; g < numGroups; ++g) {
        if (warp == 1) {
            if (g + 1 < numGroups) {
                if (!kPrefetch) loadGroup(g + 1);
                storeGroup(g + 1);
                if (kPrefetch && g + 2 < numGroups) loadGroup(g + 2);
            }
        } else {
            const uint2(*buf)[4] = sx[g & 1u];
#pragma unroll 16
            for (uint32_t q = 0; q < (uint32_t)kStripes; ++q) {
                const uint2 x = buf[q][j];
                s = XxhChainStep<true>(s, ((uint64_t)x.y << 32) | x.x);
            }
        }
        __syncthreads();
    }
    if (warp == 1) {
        diff = __reduce_or_sync(0xFFFFFFFFu, diff);
        known = __reduce_add_sync(0xFFFFFFFFu, known);
        if (lane == 0) {
            sStats[0] = diff;
            sStats[1] = known;
        }
    }
    __syncthreads();
    if (warp != 0) return;
    const uint64_t acc = xxh::ChainEnd(s);
    const uint64_t v1 = __shfl_sync(0xFFFFFFFFu, acc, 0), v2 = __shfl_sync(0xFFFFFFFFu, acc, 1), v3 = __shfl_sync(0xFFFFFFFFu, acc, 2), v4 = __shfl_sync(0xFFFFFFFFu, acc, 3);
    if (lane != 0) return;
    uint64_t h = Rotl64(v1, 1) + Rotl64(v2, 7) + Rotl64(v3, 12) + Rotl64(v4, 18);
    h = XxhMerge(h, v1); h = XxhMerge(h, v2); h = XxhMerge(h, v3); h = XxhMerge(h, v4);
    h += (uint64_t)nHash;
    h = XxhAvalanche(h);
    bool allEqual = sStats[0] == 0;
    int common = (int)s0;
    if (!allEqual && rejectionThreshold > 0.f) {
        const float frac = (float)sStats[1] / (float)n;
        if (frac < rejectionThreshold) {
            allEqual = true;
            common = ommOpacityState_UnknownTransparent;
        }
    }
    digest[w] = h;
    if (!(keepExistingSpecial && special[w] != 0)) special[w] = (allEqual && !disableSpecial) ? (-common - 1) : 0;
}

// ItemPostKernel over [itemBegin, itemEnd) and, when the bake has blocks of level >= kBigHashLevel, ItemPostBigKernel over those of them
struct BigItemList {
    uint32_t* list = nullptr;   // device, capacity entries
    uint32_t* count = nullptr;  // device
    uint32_t capacity = 0;      // work items of level >= kBigHashLevel in the whole bake (0: no big kernel)
};
static cudaError_t LaunchItemPost(cudaStream_t stream, const ItemRec* items, const unsigned long long* wordStart, const uint32_t* stateWords, uint32_t itemBegin,
                                  uint32_t itemEnd, float rejectionThreshold, int disableSpecial, int keepExistingSpecial, const uint32_t* uniformVotes, uint32_t stateGT,
                                  uint32_t stateLE, uint64_t* digest, int32_t* special, const BigItemList& big, uint32_t* launches) {
    if (itemEnd <= itemBegin) return cudaSuccess;
    if (big.capacity) {
        const cudaError_t e = cudaMemsetAsync(big.count, 0, sizeof(uint32_t), stream);
        if (e != cudaSuccess) return e;
    }
    const uint32_t n = itemEnd - itemBegin;
    ItemPostKernel<<<(n + kItemPostItemsPerBlock * 4 - 1) / (kItemPostItemsPerBlock * 4), kItemPostItemsPerBlock * 4, 0, stream>>>(
        items, wordStart, stateWords, itemBegin, itemEnd, rejectionThreshold, disableSpecial, keepExistingSpecial, uniformVotes, stateGT, stateLE, GetUniformDigests(), digest, special,
        big.capacity ? big.list : nullptr, big.count);
    (*launches)++;
    if (big.capacity) {
        // OMM_B200_BIG_HASH (read per bake, A/B runs): the one-warp kernel with the chain step as the compiler builds it (plain) / with funnel
        // shifts (funnel); producer warp + chain warp without prefetch (pipe64); default: producer warp with prefetch + chain warp
        const char* mode = getenv("OMM_B200_BIG_HASH");
        if (mode && !strcmp(mode, "plain"))
            ItemPostBigKernel<false><<<big.capacity, 32, 0, stream>>>(items, wordStart, stateWords, big.list, big.count, rejectionThreshold, disableSpecial, keepExistingSpecial, digest, special);
        else if (mode && !strcmp(mode, "funnel"))
            ItemPostBigKernel<true><<<big.capacity, 32, 0, stream>>>(items, wordStart, stateWords, big.list, big.count, rejectionThreshold, disableSpecial, keepExistingSpecial, digest, special);
        else if (mode && !strcmp(mode, "pipe64"))  // producer without prefetch (the first version of the two-warp kernel)
            ItemPostBigPipelined<64, false><<<big.capacity, 64, 0, stream>>>(items, wordStart, stateWords, big.list, big.count, rejectionThreshold, disableSpecial, keepExistingSpecial, digest, special);
        else  // default: 128 stripes per buffer half, prefetching producer
            ItemPostBigPipelined<128, true><<<big.capacity, 64, 0, stream>>>(items, wordStart, stateWords, big.list, big.count, rejectionThreshold, disableSpecial, keepExistingSpecial, digest, special);
        (*launches)++;
    }
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------------
// K6: exact dedup on digests: the lowest item index with a digest survives (ref: bake_cpu_impl.cpp:1043-1063).
// ---------------------------------------------------------------------------------------------------------------------
// Table value = the item's first triangle: monotone in the SDK's work-item index, which the output order of the items (K3b) is not.
// EVERY item takes part, also those with a special index: the digest is over the 3-state bytes (UnknownTransparent folded into UnknownOpaque),
// so a fully UnknownOpaque item (special index) and an item mixing the two unknown states (none) share a digest, and the SDK lets whichever
// came first absorb the other.  Leaving the special ones out -- tried in round 2 to spare the table the three quarters of the items that are
// uniform -- changed index buffers; the randomized GPU campaign caught it on its 28th bake.
__global__ void DigestInsert(const uint64_t* __restrict__ digest, const ItemRec* __restrict__ items, uint32_t itemBegin, uint32_t itemEnd, uint64_t* keys,
                             uint32_t* vals, uint64_t mask) {
    // Most items of a typical bake are uniform and share a handful of digests: lanes with equal digests elect the lowest
    // first triangle among themselves first, so the table sees one atomic per distinct digest per warp.
    const uint32_t s = itemBegin + blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = s < itemEnd;
    const uint64_t d = valid ? digest[s] : 0ull;
    uint32_t tri = valid ? items[s].tri : 0xFFFFFFFFu;
    const uint32_t active = __ballot_sync(0xFFFFFFFFu, valid);
    if (!valid) return;
    const uint32_t peers = __match_any_sync(active, d);
    if (__reduce_min_sync(peers, tri) == tri) TableInsertMin(keys, vals, mask, d, tri);  // first triangles are unique per item: one lane per group
}
// The same for one classifier chunk of a STREAMED bake (the chunk's survivors are packed and sent to the host while later chunks are still
// being classified, see BakeOnDevice).  A chunk is resolved against the table as it stands after its own insertions, so a survivor of
// an earlier chunk is final only if no later item with its digest has a lower first triangle: when an insertion lowers an entry that
// belonged to a serialized item of an earlier chunk, `conflict` is raised and the bake falls back to the non-streamed merge (rare: equal blocks under
// different UV triangles, the later-sorted one seen first by the SDK).
__global__ void DigestInsertChunk(const uint64_t* __restrict__ digest, const ItemRec* __restrict__ items, const uint32_t* __restrict__ triItem,
                                  const int32_t* __restrict__ special, uint32_t itemBegin, uint32_t itemEnd, uint64_t* keys, uint32_t* vals, uint64_t mask,
                                  uint32_t* __restrict__ conflict) {
    const uint32_t s = itemBegin + blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = s < itemEnd;
    const uint64_t d = valid ? digest[s] : 0ull;
    const uint32_t tri = valid ? items[s].tri : 0xFFFFFFFFu;
    const uint32_t active = __ballot_sync(0xFFFFFFFFu, valid);
    if (!valid) return;
    const uint32_t peers = __match_any_sync(active, d);
    if (__reduce_min_sync(peers, tri) != tri) return;
    uint64_t key = d == kEmptyKey ? 0x7FFFFFFFFFFFFFFFull : d;
    uint64_t slot = TableSlot(key, mask);
    while (true) {
        const uint64_t prev = atomicCAS((unsigned long long*)&keys[slot], (unsigned long long)kEmptyKey, (unsigned long long)key);
        if (prev == kEmptyKey || prev == key) {
            const uint32_t old = atomicMin(&vals[slot], tri);
            // (an earlier item with a special index was never emitted; it only matters when this item does not carry the SAME special index --
            // the digest folds the two unknown states, so a fully-unknown item and one mixing them meet here and the SDK merges them)
            if (old != 0xFFFFFFFFu && old > tri) {
                const uint32_t prev = triItem[old];
                if (prev < itemBegin && !(special[prev] != 0 && special[prev] == special[s]) && atomicOr(conflict, 1u) == 0u) {
                    conflict[1] = prev;  // diagnostics (OMM_B200_TRACE): the first pair that forced the fallback
                    conflict[2] = s;
                }
            }
            return;
        }
        slot = (slot + 1) & mask;
    }
}
// running totals of a streamed bake: descriptors / bytes emitted by the chunks so far (the initial values of the next chunk's scans)
__global__ void AdvanceRunningTotals(const uint32_t* __restrict__ emit, const unsigned long long* __restrict__ blockBytes, const uint32_t* __restrict__ descOfItem,
                                     const unsigned long long* __restrict__ offsetOfItem, uint32_t lastItem, uint32_t* __restrict__ runDesc,
                                     unsigned long long* __restrict__ runBytes, unsigned long long* __restrict__ hostSlot) {
    *runDesc = descOfItem[lastItem] + emit[lastItem];
    const unsigned long long bytes = offsetOfItem[lastItem] + blockBytes[lastItem];
    *runBytes = bytes;
    // the host learns the chunk's end through page-locked memory written from here: a cudaMemcpyAsync on this stream would queue behind
    // the previous chunk's 36 MB transfer on the device-to-host copy engine and stall the classification by that long (measured)
    *hostSlot = bytes;
    __threadfence_system();
}
__global__ void DigestResolve(const uint64_t* __restrict__ digest, const ItemRec* __restrict__ items, const uint32_t* __restrict__ triItem, uint32_t itemBegin,
                              uint32_t itemEnd, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint64_t mask, int disableDup,
                              uint32_t* __restrict__ survivor, int32_t* __restrict__ special) {
    const uint32_t s = itemBegin + blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= itemEnd) return;
    const uint32_t v = disableDup ? s : triItem[TableFind(keys, vals, mask, digest[s])];
    survivor[s] = v;
    if (v != s) special[s] = -1;  // donated its primitives; never serialized (ref: :1059-1060)
}

// ---------------------------------------------------------------------------------------------------------------------
// K7: which items are serialized, where.  The items already are in output order (K3b), so the descriptor slot of an item is the
// number of serialized items before it and its byte offset the sum of their block sizes: two prefix sums.
// ---------------------------------------------------------------------------------------------------------------------
// hist layout: [0..25] array histogram (format-1)*13+level, [26..51] index histogram
// emit[s] / blockBytes[s] for s in [itemBegin, itemEnd]; entry itemEnd is written as zero when `closeRange` (it receives the totals of the scans)
__global__ void EmitInfo(const ItemRec* __restrict__ items, const int32_t* __restrict__ special, uint32_t itemBegin, uint32_t itemEnd, int closeRange,
                         int globalBitCount, uint32_t* __restrict__ hist, uint32_t* __restrict__ emit, unsigned long long* __restrict__ blockBytes) {
    __shared__ uint32_t sh[26];
    if (threadIdx.x < 26) sh[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t s = itemBegin + blockIdx.x * blockDim.x + threadIdx.x;
    if (s < itemEnd) {
        const bool e = special[s] == 0;
        unsigned long long bytes = 0;
        if (e) {
            const ItemRec it = items[s];
            atomicAdd(&sh[(it.format - 1) * 13 + it.level], 1u);
            // ref: bake_cpu_impl.cpp:1819 -- the global format's bit count, at least one byte
            bytes = ((1ull << (2 * it.level)) * (unsigned long long)globalBitCount) >> 3;
            bytes = bytes > 1 ? bytes : 1;
        }
        emit[s] = e ? 1u : 0u;
        blockBytes[s] = bytes;
    } else if (s == itemEnd && closeRange) {
        emit[s] = 0u;
        blockBytes[s] = 0ull;
    }
    __syncthreads();
    if (threadIdx.x < 26 && sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}
// Only after Compress (a18) changed the level of work items -- the level is the leading part of the sort key -- the serialized items
// are sorted again: keys of the items without a special index, fed in reversed FIRST-SEEN order (origOf) for the SDK's tie rule.
__global__ void ResortKeys(const ItemRec* __restrict__ items, const int32_t* __restrict__ special, const uint32_t* __restrict__ origOf, uint32_t numItems,
                           uint32_t* __restrict__ sortKeys, uint32_t* __restrict__ sortVals) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= numItems) return;
    const uint32_t pos = numItems - 1 - origOf[s];
    sortKeys[pos] = special[s] == 0 ? ItemSortKey(items[s]) : 0u;  // special / merged items sort last and are never serialized
    sortVals[pos] = s;
}
__global__ void ResortBlockSizes(const uint32_t* __restrict__ sortedItems, const ItemRec* __restrict__ items, uint32_t numDescs, int globalBitCount,
                                 unsigned long long* __restrict__ blockBytes) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > numDescs) return;
    unsigned long long bytes = 0;
    if (k < numDescs) {
        bytes = ((1ull << (2 * items[sortedItems[k]].level)) * (unsigned long long)globalBitCount) >> 3;
        bytes = bytes > 1 ? bytes : 1;
    }
    blockBytes[k] = bytes;
}
__global__ void ResortScatter(const uint32_t* __restrict__ sortedItems, const unsigned long long* __restrict__ blockOffset, uint32_t numDescs,
                              uint32_t* __restrict__ descOfItem, unsigned long long* __restrict__ offsetOfItem) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= numDescs) return;
    descOfItem[sortedItems[k]] = k;
    offsetOfItem[sortedItems[k]] = blockOffset[k];
}

__global__ void TriangleFinalItems(const uint32_t* __restrict__ triItem, const uint32_t* __restrict__ survivor, const uint32_t* __restrict__ mergeRoot,
                                   const uint32_t* __restrict__ survivor2, const ItemRec* __restrict__ items, const int32_t* __restrict__ special,
                                   uint32_t triCount, uint32_t* __restrict__ triFinal, uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[26];
    if (threadIdx.x < 26) sh[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < triCount) {
        const uint32_t w = triItem[t];
        uint32_t f = kNoItem;
        if (w != kNoItem) {
            f = survivor[w];
            if (survivor2) f = survivor2[mergeRoot[f]];  // near-duplicate merge, then the second exact dedup
            if (special[f] == 0) atomicAdd(&sh[(items[f].format - 1) * 13 + items[f].level], 1u);
        }
        triFinal[t] = f;
    }
    __syncthreads();
    if (threadIdx.x < 26 && sh[threadIdx.x]) atomicAdd(&hist[26 + threadIdx.x], sh[threadIdx.x]);
}

// ---------------------------------------------------------------------------------------------------------------------
// K8: serialization (ref: bake_cpu_impl.cpp:1788-1821, 1856-1902).  One warp per work item of [itemBegin, itemEnd) that is
// serialized: its descriptor, and its block copied / repacked (2-state = even-bit compress) to its final place.  `arrayData2`, when
// given, receives the same bytes: the caller's host copy of the array, page-locked and mapped, written straight from here over PCIe
// (ommCpuBake) so that the "download" of a chunk or shard runs while other items are still being classified.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t CompressEvenBits16(uint32_t x) {  // 16 two-bit fields -> their low bits, 16 bits
    return ExtractEvenBits(x);
}
__global__ void __launch_bounds__(256) PackItems(const ItemRec* __restrict__ items, const int32_t* __restrict__ special,
                                                 const unsigned long long* __restrict__ wordStart, const uint32_t* __restrict__ stateWords,
                                                 const uint32_t* __restrict__ descOfItem, const unsigned long long* __restrict__ offsetOfItem,
                                                 const uint32_t* __restrict__ descBase, const unsigned long long* __restrict__ byteBase, uint32_t itemBegin,
                                                 uint32_t itemEnd, unsigned long long arrayCapacity, uint8_t* __restrict__ arrayData,
                                                 uint8_t* __restrict__ arrayData2) {
    const uint32_t lane = threadIdx.x & 31;
    for (uint32_t w = itemBegin + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < itemEnd; w += gridDim.x * (blockDim.x >> 5)) {
        if (special[w] != 0) continue;
        const ItemRec it = items[w];
        const unsigned long long off = offsetOfItem[w] + (byteBase ? *byteBase : 0ull);
        const uint32_t n = 1u << (2 * it.level);
        const uint32_t* src = stateWords + wordStart[w];
        uint8_t* dst = arrayData + off;
        uint8_t* dst2 = arrayData2 ? arrayData2 + off : nullptr;
        if (it.format == ommFormat_OC1_4_State) {
            if (n >= 16 && (off & 3) == 0) {
                const uint32_t numWords = n >> 4;
                if (numWords >= 4 && (off & 15) == 0 && ((wordStart[w] & 3) == 0)) {
                    const uint4* s4 = reinterpret_cast<const uint4*>(src);
                    uint4* d4 = reinterpret_cast<uint4*>(dst);
                    uint4* e4 = reinterpret_cast<uint4*>(dst2);
                    for (uint32_t i = lane; i < (numWords >> 2); i += 32) {
                        const uint4 v = __ldg(s4 + i);
                        d4[i] = v;
                        if (e4) e4[i] = v;
                    }
                } else {
                    uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);
                    uint32_t* e32 = reinterpret_cast<uint32_t*>(dst2);
                    for (uint32_t i = lane; i < numWords; i += 32) {
                        const uint32_t v = __ldg(src + i);
                        d32[i] = v;
                        if (e32) e32[i] = v;
                    }
                }
            } else {
                const uint32_t numBytes = n >= 4 ? n >> 2 : 1;
                const uint32_t byteMask = n >= 4 ? 0xFFu : ((1u << (2 * n)) - 1u);  // level 0: one 2-bit state in the byte
                for (uint32_t b = lane; b < numBytes; b += 32)
                    if (off + b < arrayCapacity) {
                        const uint8_t v = (uint8_t)((__ldg(src + (b >> 2)) >> ((b & 3) * 8)) & byteMask);
                        dst[b] = v;
                        if (dst2) dst2[b] = v;
                    }
            }
        } else {
            // 2-state: one bit per micro-triangle (the state's low bit)
            const uint32_t numBytes = n >= 8 ? n >> 3 : 1;
            const uint32_t byteMask = n >= 8 ? 0xFFu : ((1u << n) - 1u);  // levels 0/1: 1 or 4 one-bit states in the byte
            for (uint32_t b = lane; b < numBytes; b += 32) {
                const uint32_t bits = CompressEvenBits16(__ldg(src + (b >> 1)));
                if (off + b < arrayCapacity) {
                    const uint8_t v = (uint8_t)((bits >> ((b & 1) * 8)) & byteMask);
                    dst[b] = v;
                    if (dst2) dst2[b] = v;
                }
            }
        }
    }
}

__global__ void WriteIndexBuffer(const uint32_t* __restrict__ triFinal, const int32_t* __restrict__ special, const uint32_t* __restrict__ descOfItem,
                                 uint32_t triCount, int unresolved, int indexBytes, void* __restrict__ out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= triCount) return;
    const uint32_t f = triFinal[t];
    int32_t v = unresolved;
    if (f != kNoItem) v = special[f] != 0 ? special[f] : (int32_t)descOfItem[f];
    if (indexBytes == 4) ((int32_t*)out)[t] = v;
    else if (indexBytes == 2) ((int16_t*)out)[t] = (int16_t)v;
    else ((int8_t*)out)[t] = (int8_t)v;
}

// ---------------------------------------------------------------------------------------------------------------------
// texture upload + summed-area table (ref: texture_impl.cpp:77-224)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void SatRows(const void* __restrict__ texels, int isFp32, unsigned long long texelOffset, int w, int h, float cutoff, uint32_t* __restrict__ sat) {
    // one warp per row: inclusive prefix sum of (alpha > cutoff) along x
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= h) return;
    const int lane = threadIdx.x & 31;
    uint32_t carry = 0;
    for (int base = 0; base < w; base += 32) {
        const int x = base + lane;
        uint32_t v = 0;
        if (x < w) {
            const unsigned long long idx = texelOffset + (unsigned long long)row * w + x;
            const float a = isFp32 ? ((const float*)texels)[idx] : (float)((const uint8_t*)texels)[idx] * (1.f / 255.f);
            v = a > cutoff ? 1u : 0u;
        }
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, v, d);
            if (lane >= d) v += o;
        }
        v += carry;
        if (x < w) sat[(size_t)row * w + x] = v;
        carry = __shfl_sync(0xFFFFFFFFu, v, 31);
    }
}
// (H) row pass of the constant-cell table: inclusive prefix sums of "cell (x, y) is not flat-good" over the (w-1) x (h-1) interior cells
__global__ void FlatSatRows(const void* __restrict__ texels, int isFp32, int w, int h, float cutoff, uint32_t* __restrict__ sat) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= h - 1) return;
    const int lane = threadIdx.x & 31;
    uint32_t carry = 0;
    for (int base = 0; base < w - 1; base += 32) {
        const int x = base + lane;
        uint32_t v = 0;
        if (x < w - 1) {
            const size_t i00 = (size_t)row * w + x;
            float g00, g10, g01, g11;
            if (isFp32) {
                const float* t = (const float*)texels;
                g00 = t[i00]; g10 = t[i00 + 1]; g01 = t[i00 + w]; g11 = t[i00 + w + 1];
            } else {
                const uint8_t* t = (const uint8_t*)texels;
                g00 = (float)t[i00] * (1.f / 255.f); g10 = (float)t[i00 + 1] * (1.f / 255.f); g01 = (float)t[i00 + w] * (1.f / 255.f);
                g11 = (float)t[i00 + w + 1] * (1.f / 255.f);
            }
            v = CellIsFlatGood(g00, g10, g01, g11, cutoff) ? 0u : 1u;
        }
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, v, d);
            if (lane >= d) v += o;
        }
        v += carry;
        if (x < w - 1) sat[(size_t)row * (w - 1) + x] = v;
        carry = __shfl_sync(0xFFFFFFFFu, v, 31);
    }
}
// Column pass of a summed-area table in three parallel steps over segments of kSatSegRows rows: per-segment column totals, their
// running sums, then the in-segment scan with the segment's offset (the single-thread-per-column loop below needs ~1 ms per 1024 rows).
constexpr int kSatSegRows = 64;
__global__ void SatColSegTotals(int w, int h, const uint32_t* __restrict__ sat, uint32_t* __restrict__ segTotals) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, seg = blockIdx.y;
    if (x >= w) return;
    const int y0 = seg * kSatSegRows, y1 = min(h, y0 + kSatSegRows);
    uint32_t acc = 0;
    for (int y = y0; y < y1; ++y) acc += sat[(size_t)y * w + x];
    segTotals[(size_t)seg * w + x] = acc;
}
__global__ void SatColSegScan(int w, int numSegs, uint32_t* __restrict__ segTotals) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= w) return;
    uint32_t acc = 0;
    for (int sgm = 0; sgm < numSegs; ++sgm) {
        const uint32_t v = segTotals[(size_t)sgm * w + x];
        segTotals[(size_t)sgm * w + x] = acc;  // exclusive
        acc += v;
    }
}
__global__ void SatColSegApply(int w, int h, const uint32_t* __restrict__ segTotals, uint32_t* __restrict__ sat) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, seg = blockIdx.y;
    if (x >= w) return;
    const int y0 = seg * kSatSegRows, y1 = min(h, y0 + kSatSegRows);
    uint32_t acc = segTotals[(size_t)seg * w + x];
    for (int y = y0; y < y1; ++y) {
        acc += sat[(size_t)y * w + x];
        sat[(size_t)y * w + x] = acc;
    }
}
static cudaError_t SatColumnPass(int w, int h, uint32_t* sat, cudaStream_t stream) {
    const int numSegs = (h + kSatSegRows - 1) / kSatSegRows;
    uint32_t* segTotals = nullptr;
    cudaError_t e = cudaMallocAsync(&segTotals, sizeof(uint32_t) * (size_t)numSegs * w, stream);
    if (e != cudaSuccess) return e;
    const dim3 grid((w + 127) / 128, numSegs);
    SatColSegTotals<<<grid, 128, 0, stream>>>(w, h, sat, segTotals);
    SatColSegScan<<<(w + 127) / 128, 128, 0, stream>>>(w, numSegs, segTotals);
    SatColSegApply<<<grid, 128, 0, stream>>>(w, h, segTotals, sat);
    cudaFreeAsync(segTotals, stream);
    return cudaGetLastError();
}
// (I) row pass of the two item-independent whole-cell tables (not a cap pass above / below the cutoff), interior cells of mip 0
template <bool kFp32>
__global__ void StrongSatRows(const BakeParams P, uint32_t* __restrict__ satPlus, uint32_t* __restrict__ satMinus) {
    const DevMip& m = P.tex.mips[0];
    const int w = m.w, h = m.h;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= h - 1) return;
    const int lane = threadIdx.x & 31;
    uint32_t carryP = 0, carryM = 0;
    for (int base = 0; base < w - 1; base += 32) {
        const int x = base + lane;
        uint32_t vp = 0, vm = 0;
        if (x < w - 1) {
            const int s = StrongCellSide<KernelCfg<kAddrClamp, kFp32>>(P, m, x, row);
            vp = s > 0 ? 0u : 1u;
            vm = s < 0 ? 0u : 1u;
        }
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t op = __shfl_up_sync(0xFFFFFFFFu, vp, d), om = __shfl_up_sync(0xFFFFFFFFu, vm, d);
            if (lane >= d) { vp += op; vm += om; }
        }
        vp += carryP; vm += carryM;
        if (x < w - 1) {
            satPlus[(size_t)row * (w - 1) + x] = vp;
            satMinus[(size_t)row * (w - 1) + x] = vm;
        }
        carryP = __shfl_sync(0xFFFFFFFFu, vp, 31);
        carryM = __shfl_sync(0xFFFFFFFFu, vm, 31);
    }
}
__global__ void SatCols(int w, int h, uint32_t* __restrict__ sat) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= w) return;
    uint32_t acc = 0;
    for (int y = 0; y < h; ++y) {
        acc += sat[(size_t)y * w + x];
        sat[(size_t)y * w + x] = acc;
    }
}

int DeviceCount() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}
int CurrentDeviceOr(int fallback) {
    int d = fallback;
    if (cudaGetDevice(&d) != cudaSuccess) {
        cudaGetLastError();
        return fallback;
    }
    return d;
}

// Stream-ordered allocations come from the device's default memory pool; keep freed blocks cached between bakes.
static void ConfigurePoolOnce(int device) {
    static std::mutex mu;
    static bool done[64] = {};
    std::lock_guard<std::mutex> g(mu);
    if (device < 0 || device >= 64 || done[device]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long threshold = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
    }
    cudaGetLastError();
    done[device] = true;
}

struct PinnedBlock {
    void* ptr;
    size_t bytes;
    bool inUse;
};
static std::mutex g_pinnedMu;
static std::vector<PinnedBlock> g_pinned;
static size_t g_pinnedTotal = 0;
// cached (not in use) bytes above which a released block is freed at once: OMM_B200_PINNED_CACHE_MB, default 4 GiB (one arrayData of
// the largest legal size); the pool is emptied when the last baker goes away (ommDestroyBaker) or on request (ommB200TrimHostPool)
static size_t PinnedCacheLimit() {
    static const size_t limit = [] {
        const char* e = getenv("OMM_B200_PINNED_CACHE_MB");
        const unsigned long long mb = e ? strtoull(e, nullptr, 10) : 4096ull;
        return (size_t)mb << 20;
    }();
    return limit;
}
void* PinnedPoolAcquire(size_t bytes) {
    std::lock_guard<std::mutex> g(g_pinnedMu);
    // best fit, but never a block more than twice the request (a 1 MiB result must not pin down a multi-GB block)
    int best = -1;
    for (int i = 0; i < (int)g_pinned.size(); ++i)
        if (!g_pinned[i].inUse && g_pinned[i].bytes >= bytes && g_pinned[i].bytes <= 2 * bytes + ((size_t)4 << 20) &&
            (best < 0 || g_pinned[i].bytes < g_pinned[best].bytes))
            best = i;
    if (best >= 0) {
        g_pinned[best].inUse = true;
        return g_pinned[best].ptr;
    }
    const size_t rounded = (bytes + (bytes >> 3) + ((size_t)2 << 20)) & ~(((size_t)2 << 20) - 1);  // 12.5 % slack, 2 MiB granules
    void* p = nullptr;
    if (cudaHostAlloc(&p, rounded, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    g_pinned.push_back({p, rounded, true});
    g_pinnedTotal += rounded;
    return p;
}
static size_t PinnedCachedLocked() {
    size_t cached = 0;
    for (const PinnedBlock& b : g_pinned)
        if (!b.inUse) cached += b.bytes;
    return cached;
}
static size_t PinnedTrimLocked(size_t keepBytes) {
    size_t cached = PinnedCachedLocked();
    // largest blocks go first
    while (cached > keepBytes) {
        int victim = -1;
        for (int i = 0; i < (int)g_pinned.size(); ++i)
            if (!g_pinned[i].inUse && (victim < 0 || g_pinned[i].bytes > g_pinned[victim].bytes)) victim = i;
        if (victim < 0) break;
        cudaFreeHost(g_pinned[victim].ptr);
        cached -= g_pinned[victim].bytes;
        g_pinnedTotal -= g_pinned[victim].bytes;
        g_pinned.erase(g_pinned.begin() + victim);
    }
    cudaGetLastError();
    return cached;
}
void PinnedPoolRelease(void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> g(g_pinnedMu);
    for (size_t i = 0; i < g_pinned.size(); ++i)
        if (g_pinned[i].ptr == p) {
            g_pinned[i].inUse = false;
            PinnedTrimLocked(PinnedCacheLimit());
            return;
        }
}
size_t PinnedPoolTrim(size_t keepBytes) {
    std::lock_guard<std::mutex> g(g_pinnedMu);
    return PinnedTrimLocked(keepBytes);
}

// Host memory of arrayData: the library's page-locked pool under the default allocator (full PCIe speed), the user's allocator otherwise.
static bool AllocHostArrayData(BakeResultObject* res, size_t bytes = 0) {
    if (res->hostArrayData) return true;
    if (bytes == 0) bytes = res->arrayDataSize;
    if (res->usesDefaultAllocator && bytes >= (1u << 20)) {
        res->hostArrayData = PinnedPoolAcquire(bytes);
        res->arrayDataFromPinnedPool = res->hostArrayData != nullptr;
    }
    if (!res->hostArrayData) res->hostArrayData = res->alloc.alloc(bytes, 64);
    return res->hostArrayData != nullptr;
}
// Streamed ommCpuBake (one GPU): classifier chunks per bake = the nominal number times this (more chunks: a shorter tail after the last
// kernel, ~50 us of kernel boundaries each).  OMM_B200_STREAM_DIV overrides (1 = the chunks of a resident bake).
static unsigned StreamChunkDivisor() {
    static const unsigned v = [] {
        const char* e = getenv("OMM_B200_STREAM_DIV");
        const long n = e ? atol(e) : 4;
        return (unsigned)(n < 1 ? 1 : (n > 16 ? 16 : n));
    }();
    return v;
}
// OFF by default (OMM_B200_STREAMED_DOWNLOAD=1 enables it; read per bake so that tests can toggle it).  Measured at config 3: one in 27
// serialized blocks is shared by several UV triangles (almost-uniform blocks whose few odd micro-triangles coincide), half of those pairs
// have their SDK survivor in a later chunk than the copy that was sent first, and every such case shifts all later offsets -- so the
// optimistic merge always falls back there and the call gets 1.7 ms slower (22.0 vs 20.3 ms) instead of 4 ms faster.  It pays for
// inputs whose blocks do not repeat across triangles.
static bool StreamedDownloadEnabled() {
    const char* e = getenv("OMM_B200_STREAMED_DOWNLOAD");
    return e != nullptr && e[0] != '0';
}

static ommResult RequireDevice(const Logger& log, int device) {
    if (DeviceCount() <= 0) {
        log.Log(ommMessageSeverity_Fatal, "[omm-b200] no CUDA device is visible; this library has no CPU fallback");
        return ommResult_FAILURE;
    }
    if (cudaSetDevice(device) != cudaSuccess) {
        cudaGetLastError();
        log.Logf(ommMessageSeverity_Fatal, "[omm-b200] cudaSetDevice(%d) failed", device);
        return ommResult_FAILURE;
    }
    ConfigurePoolOnce(device);
    return ommResult_SUCCESS;
}

ommResult UploadTexture(TextureObject* tex, const Logger& log) {
    ommResult rc = RequireDevice(log, tex->device);
    if (rc != ommResult_SUCCESS) return rc;
    const size_t spp = tex->format == ommCpuTextureFormat_FP32 ? 4 : 1;
    size_t totalTexels = 0;
    for (uint32_t i = 0; i < tex->mipCount; ++i) totalTexels += (size_t)tex->dev.mips[i].w * tex->dev.mips[i].h;
    CUDA_TRY(cudaMalloc(&tex->devTexels, totalTexels * spp));
    CUDA_TRY(cudaMemcpy(tex->devTexels, tex->hostTexels, totalTexels * spp, cudaMemcpyHostToDevice));
    tex->dev.texels = tex->devTexels;
    tex->dev.isFp32 = tex->format == ommCpuTextureFormat_FP32;
    tex->dev.sat = nullptr;
    tex->dev.flatSat = nullptr;
    tex->dev.strongPlus = tex->dev.strongMinus = nullptr;
    if (tex->HasAlphaCutoff()) {  // ref: texture_impl.cpp:91 -- SAT <=> alphaCutoff >= 0
        CUDA_TRY(cudaMalloc(&tex->devSat, totalTexels * sizeof(uint32_t)));
        for (uint32_t i = 0; i < tex->mipCount; ++i) {
            const DevMip& m = tex->dev.mips[i];
            uint32_t* sat = tex->devSat + m.satOffset;
            SatRows<<<(m.h + 7) / 8, 256>>>(tex->devTexels, tex->dev.isFp32, m.texelOffset, m.w, m.h, tex->alphaCutoff, sat);
            SatCols<<<(m.w + 255) / 256, 256>>>(m.w, m.h, sat);
        }
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaDeviceSynchronize());
        tex->dev.sat = tex->devSat;
    }
cleanup:
    if (rc != ommResult_SUCCESS) DestroyTextureDevice(tex);
    return rc;
}
static void FreeCellTables(CellTables* t) {
    if (t->flatSat) cudaFree(t->flatSat);
    if (t->strongPlus) cudaFree(t->strongPlus);
    if (t->strongMinus) cudaFree(t->strongMinus);
    delete t;
}
void DestroyTextureDevice(TextureObject* tex) {
    if (tex->devTexels || tex->devSat || !tex->cellTables.empty()) cudaSetDevice(tex->device);
    if (tex->devTexels) cudaFree(tex->devTexels);
    if (tex->devSat) cudaFree(tex->devSat);
    for (CellTables* t : tex->cellTables) FreeCellTables(t);
    tex->cellTables.clear();
    tex->devTexels = nullptr;
    tex->devSat = nullptr;
}

// (H) The constant-cell table of mip 0 for the bake's cutoff.  A texture keeps one immutable set of tables per cutoff it has been baked
// with: the SDK treats textures as read-only objects shared by concurrent bakes (docs/integration_guide.md:434), so a published set is
// never rewritten -- a bake with another cutoff builds another set.  Sets are reference-counted by the bakes using them; idle sets
// beyond kIdleCellTables are freed, least recently used first (a texture is normally baked with one cutoff).  The build runs on the
// bake's stream and is complete before the set is published, so bakes on other streams may use it at once.  Returns nullptr when the
// texture is too small or memory is short -- the classifier then simply has no O(1) answer for large footprints.
// (I) is OFF by default: measured on B200 at config 3, answering region tests from the two 64 MB tables (four to eight scattered
// 4-byte loads each) is 7 % SLOWER than the per-item bitmap, whose texel gathers are coalesced and shared (13.0 vs 12.2 ms).  The
// code stays for textures / workloads where it may pay; OMM_B200_STRONG_TABLES=1 enables it.
static bool UseStrongTables() {
    static const bool on = getenv("OMM_B200_STRONG_TABLES") != nullptr;
    return on;
}
constexpr int kIdleCellTables = 2;
static CellTables* AcquireCellTables(TextureObject* tex, BakeParams& P, cudaStream_t stream, uint32_t* launches) {
    const DevMip& m = tex->dev.mips[0];
    P.tex.flatSat = P.tex.strongPlus = P.tex.strongMinus = nullptr;
    if (m.w < 2 || m.h < 2) return nullptr;
    const float cutoff = P.cutoff;
    const int numTables = UseStrongTables() ? 3 : 1;
    std::lock_guard<std::mutex> g(tex->flatMu);
    CellTables* set = nullptr;
    for (CellTables* t : tex->cellTables)
        if (FloatAsUint(t->cutoff) == FloatAsUint(cutoff)) set = t;
    if (!set) {
        // evict idle sets first (nobody reads them: refs == 0 means every bake that used them has drained its stream)
        while (true) {
            int idle = 0, victim = -1;
            for (int i = 0; i < (int)tex->cellTables.size(); ++i)
                if (tex->cellTables[i]->refs == 0) {
                    ++idle;
                    if (victim < 0 || tex->cellTables[i]->lastUse < tex->cellTables[victim]->lastUse) victim = i;
                }
            if (idle < kIdleCellTables) break;
            FreeCellTables(tex->cellTables[victim]);
            tex->cellTables.erase(tex->cellTables.begin() + victim);
        }
        set = new (std::nothrow) CellTables();
        if (!set) return nullptr;
        set->cutoff = cutoff;
        const size_t bytes = sizeof(uint32_t) * (size_t)(m.w - 1) * (size_t)(m.h - 1);
        uint32_t** bufs[3] = {&set->flatSat, &set->strongPlus, &set->strongMinus};
        bool ok = true;
        for (int i = 0; i < numTables && ok; ++i) ok = cudaMalloc((void**)bufs[i], bytes) == cudaSuccess;
        if (ok) {
            FlatSatRows<<<(m.h - 1 + 7) / 8, 256, 0, stream>>>(tex->devTexels, tex->dev.isFp32, m.w, m.h, cutoff, set->flatSat);
            if (numTables == 3) {
                if (tex->dev.isFp32) StrongSatRows<true><<<(m.h - 1 + 7) / 8, 256, 0, stream>>>(P, set->strongPlus, set->strongMinus);
                else StrongSatRows<false><<<(m.h - 1 + 7) / 8, 256, 0, stream>>>(P, set->strongPlus, set->strongMinus);
            }
            for (int i = 0; i < numTables && ok; ++i) ok = SatColumnPass(m.w - 1, m.h - 1, *bufs[i], stream) == cudaSuccess;
            *launches += numTables == 3 ? 11 : 4;
            // bakes on other streams may pick the set up as soon as it is in the list
            ok = ok && cudaStreamSynchronize(stream) == cudaSuccess;
        }
        if (!ok) {
            cudaGetLastError();
            FreeCellTables(set);
            return nullptr;
        }
        tex->cellTables.push_back(set);
    }
    set->refs++;
    set->lastUse = ++tex->cellTableClock;
    P.tex.flatSat = set->flatSat;
    P.tex.strongPlus = UseStrongTables() ? set->strongPlus : nullptr;
    P.tex.strongMinus = UseStrongTables() ? set->strongMinus : nullptr;
    return set;
}
static void ReleaseCellTables(TextureObject* tex, CellTables* set) {
    if (!set) return;
    std::lock_guard<std::mutex> g(tex->flatMu);
    set->refs--;
}

// ---------------------------------------------------------------------------------------------------------------------
// NCCL, bound at run time.  The only collective of the path: an all-gather (with per-rank counts, issued as one group of
// broadcasts) of the per-item state blocks, after which dedup / sort / offsets are computed redundantly on every rank.
// ---------------------------------------------------------------------------------------------------------------------
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    bool ok = false;
};
static NcclApi& Nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) return;
        api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
        api.Broadcast = (decltype(api.Broadcast))dlsym(api.lib, "ncclBroadcast");
        api.Send = (decltype(api.Send))dlsym(api.lib, "ncclSend");
        api.Recv = (decltype(api.Recv))dlsym(api.lib, "ncclRecv");
        api.GroupStart = (decltype(api.GroupStart))dlsym(api.lib, "ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))dlsym(api.lib, "ncclGroupEnd");
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Broadcast && api.Send && api.Recv && api.GroupStart && api.GroupEnd;
    });
    return api;
}

// ---------------------------------------------------------------------------------------------------------------------
// staging of the per-bake inputs
// ---------------------------------------------------------------------------------------------------------------------
// Largest vertex index the triangles use: it sizes the upload of the caller's UV buffer (nothing beyond the last referenced vertex may
// be read).  Computed on the device from the index buffer that has just been uploaded (a host scan of 3 M indices is ~1 ms of every
// ommCpuBake; this is one 12 MB pass at HBM speed plus a 4-byte read-back).
__global__ void MaxIndexKernel(const void* __restrict__ indices, int indexFormat, unsigned long long count, uint32_t* __restrict__ result) {
    uint32_t m = 0;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        uint32_t v;
        if (indexFormat == ommIndexFormat_UINT_8) v = ((const uint8_t*)indices)[i];
        else if (indexFormat == ommIndexFormat_UINT_16) v = ((const uint16_t*)indices)[i];
        else v = ((const uint32_t*)indices)[i];
        m = v > m ? v : m;
    }
    m = __reduce_max_sync(0xFFFFFFFFu, m);
    if ((threadIdx.x & 31u) == 0 && m != 0) atomicMax(result, m);
}
static uint32_t TexCoordSize(ommTexCoordFormat f) { return f == ommTexCoordFormat_UV32_FLOAT ? 8u : 4u; }  // ref: util/texture.h:148-160
static uint32_t IndexSize(ommIndexFormat f) { return f == ommIndexFormat_UINT_8 ? 1u : (f == ommIndexFormat_UINT_16 ? 2u : 4u); }

ommResult StageInputs(BakerObject* baker, const ommCpuBakeInputDesc& desc, StagedInputs* out) {
    const Logger& log = baker->log;
    ommResult rc = RequireDevice(log, baker->device);
    if (rc != ommResult_SUCCESS) return rc;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    out->baker = baker;
    out->desc = desc;
    out->device = baker->device;
    out->triangleCount = desc.indexCount / 3u;
    out->texCoordStride = desc.texCoordStrideInBytes == 0 ? TexCoordSize(desc.texCoordFormat) : desc.texCoordStrideInBytes;
    {
        const size_t usedIndices = (size_t)out->triangleCount * 3;
        const size_t indexBytes = usedIndices * IndexSize(desc.indexFormat);
        uint32_t maxIndex = 0;
        CUDA_TRY(PoolEventCreate(&e0));
        CUDA_TRY(PoolEventCreate(&e1));
        CUDA_TRY(cudaEventRecord(e0, 0));
        // Sharded baker: every rank is handed the same inputs (the call is collective), so only rank 0 sends them over PCIe and the
        // others receive them over NVLink -- eight simultaneous uploads of the same 36 MB through one host ran at half speed.
        const int world = baker->shard.world, rank = baker->shard.rank;
        NcclApi& nccl = Nccl();
        const ncclComm_t comm = (ncclComm_t)baker->shard.ncclComm;
        const bool viaRoot = world > 1 && nccl.ok && comm != nullptr;
        auto put = [&](void* dst, const void* src, size_t bytes) -> bool {
            if (bytes == 0) return true;
            if (!viaRoot || rank == 0) {
                if (cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, 0) != cudaSuccess) return false;
                out->h2dBytes += bytes;
            }
            return !viaRoot || nccl.Broadcast(dst, dst, bytes, ncclUint8, 0, comm, 0) == ncclSuccess;
        };
        out->h2dBytes = 0;
        CUDA_TRY(cudaMallocAsync(&out->devIndices, (indexBytes ? indexBytes : 4) + 8, 0));  // + the (aligned) max-index word
        if (!put(out->devIndices, desc.indexBuffer, indexBytes)) { rc = ommResult_FAILURE; goto cleanup; }
        if (usedIndices) {
            uint32_t* devMax = (uint32_t*)((uint8_t*)out->devIndices + ((indexBytes + 3) & ~(size_t)3));
            CUDA_TRY(cudaMemsetAsync(devMax, 0, 4, 0));
            int sms = 0;
            CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, baker->device));
            MaxIndexKernel<<<(uint32_t)std::max(sms, 1) * 4u, 256, 0, 0>>>(out->devIndices, (int)desc.indexFormat, (unsigned long long)usedIndices, devMax);
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaMemcpyAsync(&maxIndex, devMax, 4, cudaMemcpyDeviceToHost, 0));
            CUDA_TRY(cudaStreamSynchronize(0));
        }
        out->texCoordBytes = (size_t)maxIndex * out->texCoordStride + TexCoordSize(desc.texCoordFormat);
        CUDA_TRY(cudaMallocAsync(&out->devTexCoords, out->texCoordBytes + 8, 0));
        if (!put(out->devTexCoords, desc.texCoords, out->texCoordBytes)) { rc = ommResult_FAILURE; goto cleanup; }
        if (desc.subdivisionLevels) {
            CUDA_TRY(cudaMallocAsync((void**)&out->devLevels, out->triangleCount ? out->triangleCount : 1, 0));
            if (!put(out->devLevels, desc.subdivisionLevels, out->triangleCount)) { rc = ommResult_FAILURE; goto cleanup; }
        }
        if (desc.formats) {
            // The SDK sizes its arrays from desc.format alone and then serializes every item (bake_cpu_impl.cpp:1763-1771): a per-triangle
            // format that differs from it overruns there.  Refused here, before any device work.
            for (uint32_t t = 0; t < out->triangleCount; ++t) {
                const int32_t f = (int32_t)desc.formats[t];
                if (f != (int32_t)ommFormat_INVALID && f != (int32_t)desc.format) {
                    log.Log(ommMessageSeverity_Fatal, "[omm-b200] per-triangle formats that differ from desc.format are not supported");
                    rc = ommResult_FAILURE;
                    goto cleanup;
                }
            }
            CUDA_TRY(cudaMallocAsync((void**)&out->devFormats, (size_t)(out->triangleCount ? out->triangleCount : 1) * 4, 0));
            if (!put(out->devFormats, desc.formats, (size_t)out->triangleCount * 4)) { rc = ommResult_FAILURE; goto cleanup; }
        }
        CUDA_TRY(cudaEventRecord(e1, 0));
        CUDA_TRY(cudaEventSynchronize(e1));
        CUDA_TRY(cudaEventElapsedTime(&out->h2dMs, e0, e1));
    }
cleanup:
    PoolEventRelease(e0);
    PoolEventRelease(e1);
    if (rc != ommResult_SUCCESS) DestroyStagedDevice(out);
    return rc;
}
void DestroyStagedDevice(StagedInputs* s) {
    cudaSetDevice(s->device);
    if (s->devIndices) cudaFreeAsync(s->devIndices, 0);
    if (s->devTexCoords) cudaFreeAsync(s->devTexCoords, 0);
    if (s->devLevels) cudaFreeAsync(s->devLevels, 0);
    if (s->devFormats) cudaFreeAsync(s->devFormats, 0);
    s->devIndices = s->devTexCoords = nullptr;
    s->devLevels = nullptr;
    s->devFormats = nullptr;
}

// ref: bake_cpu_impl.cpp:524-527 -- finished on the host so that log2f is the very libm the SDK build calls
static int8_t FinishEdgeHeuristic(float eMax, float dynScale, int maxLevel) {
    const float n = eMax < 1e-6 ? 0 : std::log2(eMax) / 2.f - std::log2(dynScale);
    const int lvl = (int)std::ceil(n);
    return (int8_t)std::clamp<int>(lvl, 0, maxLevel);
}

static uint64_t NextPow2(uint64_t v) {
    uint64_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

// Device scratch of one bake (stream-ordered, freed at the end).  Requests below a quarter of a block are carved out of shared 128 MiB blocks: a
// bake makes about fifty allocations, and fifty cudaMallocAsync / cudaFreeAsync pairs are 0.2 ms of host time -- in the launch-bound set-up
// phase, and after each host read-back while the GPU waits for the next kernels, that time is on the critical path.
struct Scratch {
    cudaStream_t stream;
    std::vector<void*> ptrs;
    uint8_t* block = nullptr;
    size_t blockSize = 0, blockUsed = 0;
    static constexpr size_t kBlock = (size_t)128 << 20, kAlign = 256;
    // OMM_B200_SCRATCH_BLOCKS=0: every request is an allocation of its own, so that compute-sanitizer sees the bounds of every array
    // (scripts/sanitize_cases.py runs that way)
    static bool Shared() {
        static const bool on = [] {
            const char* e = getenv("OMM_B200_SCRATCH_BLOCKS");
            return !(e && e[0] == '0');
        }();
        return on;
    }
    template <class T>
    cudaError_t alloc(T** p, size_t count) {
        const size_t bytes = (((count ? count : 1) * sizeof(T)) + kAlign - 1) & ~(kAlign - 1);
        if (Shared() && bytes <= kBlock / 4) {
            if (blockUsed + bytes > blockSize) {
                void* q = nullptr;
                const cudaError_t e = cudaMallocAsync(&q, kBlock, stream);
                if (e != cudaSuccess) {
                    *p = nullptr;
                    return e;
                }
                ptrs.push_back(q);
                block = (uint8_t*)q;
                blockSize = kBlock;
                blockUsed = 0;
            }
            *p = (T*)(block + blockUsed);
            blockUsed += bytes;
            return cudaSuccess;
        }
        void* q = nullptr;
        const cudaError_t e = cudaMallocAsync(&q, bytes, stream);
        if (e == cudaSuccess) ptrs.push_back(q);
        *p = (T*)q;
        return e;
    }
    void freeAll() {
        for (void* p : ptrs) cudaFreeAsync(p, stream);
        ptrs.clear();
        block = nullptr;
        blockSize = blockUsed = 0;
    }
};

// Sharded bakes: what a work item will cost to classify, for the cut of the shards.  The time of the hierarchical classifier follows the
// level line -- regions it does not touch are proved in bulk, micro-triangles it crosses run the exact edge tests -- so contiguous shards of
// equal micro-triangle count differ by 8-12 % in time.  (Opt-in, see CostBalanceDisabled: folded shards already average that out.)  Proxy: a 5 x 5 lattice of texels over the item's
// bounding box; the share of neighbouring lattice pairs on different sides of the cutoff estimates the density of the level line
// (Cauchy-Crofton), and   cost = units * (8 + 128 * share)   weighs an item the line fills 17 times an untouched one (the ratio of the leaf
// path to the bulk path per micro-triangle at config 3).  Only the balance depends on it, never a result.
template <bool kFp32>
__global__ void ItemCostKernel(const BakeParams P, const ItemRec* __restrict__ items, const uint32_t* __restrict__ numItemsPtr, uint32_t slots,
                               const unsigned long long* __restrict__ itemUnits, unsigned long long* __restrict__ itemCost) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > slots) return;
    if (s >= *numItemsPtr) {
        itemCost[s] = 0ull;
        return;
    }
    const ItemRec it = items[s];
    const DevMip& m = P.tex.mips[0];
    const float W = (float)m.w, H = (float)m.h;
    const float x0 = fminf(fminf(it.p0.x, it.p1.x), it.p2.x) * W, x1 = fmaxf(fmaxf(it.p0.x, it.p1.x), it.p2.x) * W;
    const float y0 = fminf(fminf(it.p0.y, it.p1.y), it.p2.y) * H, y1 = fmaxf(fmaxf(it.p0.y, it.p1.y), it.p2.y) * H;
    uint32_t changes = 0;
    if (x0 > -1e6f && x1 < 1e6f && y0 > -1e6f && y1 < 1e6f) {  // (also false for NaN)
        uint32_t rows[5];
        for (int j = 0; j < 5; ++j) {
            const int ty = (int)floorf(y0 + (y1 - y0) * (0.25f * (float)j));
            const int ay = Addr1Generic(P.addrMode, m.isPow2, ty, m.h, m.log2h);
            uint32_t bits = 0;
            for (int i = 0; i < 5; ++i) {
                const int tx = (int)floorf(x0 + (x1 - x0) * (0.25f * (float)i));
                const int ax = Addr1Generic(P.addrMode, m.isPow2, tx, m.w, m.log2w);
                const float a = TexFetch<KernelCfg<kAddrGeneric, kFp32>>(P, m, ax, ay);
                bits |= (P.cutoff < a ? 1u : 0u) << i;
            }
            rows[j] = bits;
            changes += __popc((bits ^ (bits >> 1)) & 0xFu);
            if (j) changes += __popc(bits ^ rows[j - 1]);
        }
    }
    itemCost[s] = itemUnits[s] * (unsigned long long)(8u + (128u * changes) / 40u);
}

struct ShardBound {
    unsigned long long unit, word, node;
    uint32_t item, pad;
};
// chunk size actually used: the nominal one, or larger when the rank's regions would need more than kHierMaxChunks chunks
__host__ __device__ inline unsigned long long HierChunkRegions(unsigned long long totalRegions, unsigned long long nominal);
constexpr int kHierMaxChunks = 96;  // 96 x 8M initial regions of 64 micro-triangles > the 2^35 micro-triangles a 4 GiB 1-bit array can hold
// First work item of rank r (r = 0..world): items are split where the running unit count crosses r*U/world, so every
// rank owns a contiguous run of whole work items and the state words of a rank are contiguous too.
__host__ __device__ inline unsigned long long HierChunkRegions(unsigned long long totalRegions, unsigned long long nominal) {
    const unsigned long long need = (totalRegions + kHierMaxChunks - 1) / kHierMaxChunks;
    return need > nominal ? need : nominal;
}
__host__ __device__ inline uint32_t ShardFirstItem(const unsigned long long* unitStart, uint32_t entries, int world, int r) {
    const unsigned long long total = unitStart[entries - 1];
    const unsigned long long target = r >= world ? total : (total / (unsigned long long)world) * (unsigned long long)r;
    uint32_t lo = 0, hi = entries - 1;  // smallest i with unitStart[i] >= target
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (unitStart[mid] >= target) hi = mid;
        else lo = mid + 1;
    }
    return lo;
}
// Shards.  The work items are cut into `shards` = world x shardsPerRank contiguous, unit-balanced runs and dealt to the ranks in
// boustrophedon order (0 1 .. N-1, N-1 .. 1 0, ...): the cost of an item follows the level line through it, not its micro-triangle
// count, and varies slowly over a mesh -- one contiguous run per rank left rank 0 waiting 0.7 of 6.4 ms for rank 1 at config 3 on
// two GPUs; folded runs average such trends out without a cost model.  One rank: one shard.
constexpr int kMaxShardsPerRank = 4;
constexpr int kMaxShards = 64;
__host__ __device__ inline int ShardOwner(int shard, int world) {
    const int pass = shard / world, pos = shard % world;
    return (pass & 1) ? world - 1 - pos : pos;
}
// OMM_B200_COST_BALANCE=1 cuts the shards by the cost estimate of ItemCostKernel instead of by micro-triangle count.  OFF by default: measured
// on two B200s at config 3 it does balance one shard per rank (rank 0 waits 0.09 instead of 0.65 ms in the record exchange) -- but two
// count-balanced shards per rank dealt in boustrophedon order already end within the same total (7.87 ms against 7.94-7.96 with the
// estimate, whose kernel and scan cost 0.07 ms on every rank): what a sharded classification loses is the fixed cost per shard
// (2 x 6.45 ms of classification against 12.3 on one GPU), not the balance.
static bool CostBalanceDisabled() {
    static const bool on = getenv("OMM_B200_COST_BALANCE") != nullptr;
    return !on;
}
static int ShardsPerRank(int world) {
    if (world <= 1) return 1;
    // Measured at config 3: two shards per rank win 2.3 % on two GPUs (8.22 -> 8.03 ms: no more waiting, +0.1 ms for the second chunk
    // sequence); on four the second chunk sequence and the doubled launches cost what the balance gains (5.3-5.4 vs 5.4-5.7 ms), so
    // the default folds only on two ranks.  OMM_B200_SHARDS_PER_RANK=1..4 overrides (A/B runs, parity tests of the folded layout).
    static const int requested = [] {
        const char* e = getenv("OMM_B200_SHARDS_PER_RANK");
        const int v = e ? atoi(e) : 0;
        return v < 0 ? 0 : (v > kMaxShardsPerRank ? kMaxShardsPerRank : v);
    }();
    const int r = requested > 0 ? requested : (world == 2 ? 2 : 1);
    return std::max(1, std::min(r, kMaxShards / world));
}
struct OwnedShards {
    int count;
    int shard[kMaxShardsPerRank];
};
__global__ void ShardBounds(const unsigned long long* __restrict__ weightStart /* prefix sums the shards are balanced by */,
                            const unsigned long long* __restrict__ unitStart, const unsigned long long* __restrict__ wordStart,
                            const unsigned long long* __restrict__ nodeStart, uint32_t entries, int world /* number of shards */, OwnedShards owned,
                            unsigned long long chunkRegions, ShardBound* __restrict__ bounds, uint32_t* __restrict__ chunkFirstItem) {
    const int r = threadIdx.x;
    for (int k = 0; k < owned.count; ++k) {
        // chunks of the hierarchical classifier: runs of whole work items of an owned shard holding about `chunkRegions` initial regions each
        const uint32_t ib = ShardFirstItem(weightStart, entries, world, owned.shard[k]), ie = ShardFirstItem(weightStart, entries, world, owned.shard[k] + 1);
        if (r <= kHierMaxChunks) {
            const unsigned long long cr = HierChunkRegions(nodeStart[ie] - nodeStart[ib], chunkRegions);
            const unsigned long long target = nodeStart[ib] + (unsigned long long)r * cr;
            uint32_t lo = ib, hi = ie;  // smallest item in [ib, ie] whose first region is >= target
            while (lo < hi) {
                const uint32_t mid = lo + ((hi - lo) >> 1);
                if (nodeStart[mid] >= target) hi = mid;
                else lo = mid + 1;
            }
            chunkFirstItem[k * (kHierMaxChunks + 1) + r] = lo;
        }
    }
    if (r > world) return;
    const uint32_t lo = ShardFirstItem(weightStart, entries, world, r);
    bounds[r].item = lo;
    bounds[r].unit = unitStart[lo];
    bounds[r].word = wordStart[lo];
    bounds[r].node = nodeStart[lo];
}
// host mirror of the partition (same function), exported for the CPU-side multi-rank tests
ommResult ComputeShardBounds(const unsigned long long* unitStart, uint32_t entries, int world, uint32_t* outFirstItem) {
    if (!unitStart || !outFirstItem || entries == 0 || world < 1) return ommResult_INVALID_ARGUMENT;
    for (int r = 0; r <= world; ++r) outFirstItem[r] = ShardFirstItem(unitStart, entries, world, r);
    return ommResult_SUCCESS;
}

int ShardsPerRankOf(int world) { return world < 1 ? 0 : ShardsPerRank(world); }
int ShardOwnerOf(int shard, int world) { return (world < 1 || shard < 0) ? -1 : ShardOwner(shard, world); }

// workload metric of the SDK (ref: bake_cpu_impl.cpp:662-680): sum over work items of int(aabb.x*texW) * int(aabb.y*texH)
__global__ void WorkloadKernel(const ItemRec* __restrict__ items, uint32_t numItems, float texW, float texH, unsigned long long* __restrict__ total) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long v = 0;
    if (w < numItems) {
        const ItemRec it = items[w];
        const float sx = fminStd(fminStd(it.p0.x, it.p1.x), it.p2.x), sy = fminStd(fminStd(it.p0.y, it.p1.y), it.p2.y);
        const float ex = fmaxStd(fmaxStd(it.p0.x, it.p1.x), it.p2.x), ey = fmaxStd(fmaxStd(it.p0.y, it.p1.y), it.p2.y);
        const int ax = f2i((ex - sx) * texW), ay = f2i((ey - sy) * texH);
        v = (unsigned long long)(long long)(int)((unsigned)ax * (unsigned)ay);
    }
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(total, v);
}

__global__ void SumMicroTriangles(const ItemRec* __restrict__ items, uint32_t itemBegin, uint32_t itemEnd, unsigned long long* __restrict__ total) {
    const uint32_t w = itemBegin + blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long v = w < itemEnd ? (1ull << (2 * items[w].level)) : 0ull;
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(total, v);
}

// primitives per item after the first exact dedup (input of the host passes)
__global__ void CountPrimitives(const uint32_t* __restrict__ triItem, const uint32_t* __restrict__ survivor, uint32_t triCount, uint32_t* __restrict__ primCount) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= triCount) return;
    const uint32_t w = triItem[t];
    if (w != kNoItem) atomicAdd(&primCount[survivor[w]], 1u);
}
__global__ void UpdateItemLevels(ItemRec* __restrict__ items, const uint8_t* __restrict__ levels, uint32_t numItems) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w < numItems) items[w].level = levels[w];
}

__global__ void CountDisabled(const int8_t* __restrict__ triLevel, uint32_t triCount, uint32_t* __restrict__ count) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned m = __ballot_sync(0xFFFFFFFFu, t < triCount && triLevel[t] < 0);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(count, (uint32_t)__popc(m));
}

static const char* SpecialIndexText(int s) {  // ref: log.h:20-31
    switch (s) {
    case ommSpecialIndex_FullyTransparent: return "Fully Transparent";
    case ommSpecialIndex_FullyOpaque: return "Fully Opaque";
    case ommSpecialIndex_FullyUnknownTransparent: return "Fully Unknown Transparent";
    case ommSpecialIndex_FullyUnknownOpaque: return "Fully Unknown Opaque";
    default: return "Unknown State";
    }
}

}  // namespace ommb200
#include "omm_post_passes.cuh"
namespace ommb200 {

bool HostPassesNeeded(const ommCpuBakeInputDesc& desc) {
    const uint32_t flags = (uint32_t)desc.bakeFlags;
    if (flags & ommCpuBakeFlags_DisableDuplicateDetection) return desc.maxArrayDataSize != 0xFFFFFFFFu;  // the merge is off with the dedup (ref: bake_cpu_impl.cpp:1136)
    return (flags & ommCpuBakeFlags_EnableNearDuplicateDetection) != 0 || desc.maxArrayDataSize != 0xFFFFFFFFu;
}

// ---- host side of a sharding on one box: a control block and result windows in POSIX shared memory -------------------------------
struct ShmControl {
    std::atomic<uint32_t> barCount, barGen;  // sense-reversing barrier of the ranks' host threads
    std::atomic<unsigned long long> failSeq; // bakeSeq of the last bake in which some rank could not write its part of the window
    std::atomic<unsigned long long> winSeq;  // bakeSeq for which winId / winCapacity were published by the root
    int winId;
    unsigned long long winCapacity;
};
static_assert(std::atomic<uint32_t>::is_always_lock_free && std::atomic<unsigned long long>::is_always_lock_free, "atomics in shared memory must be address-free");
static void ShmName(char* out, size_t n, unsigned long long idHash, const char* what, int id) { snprintf(out, n, "/ommb200_%016llx_%s%d", idHash, what, id); }
static bool HostBarrier(ShardState& sh, double timeoutSeconds = 120.0) {
    ShmControl* c = sh.ctl;
    if (!c) return false;
    const uint32_t gen = c->barGen.load(std::memory_order_acquire);
    if (c->barCount.fetch_add(1, std::memory_order_acq_rel) + 1 == (uint32_t)sh.world) {
        c->barCount.store(0, std::memory_order_relaxed);
        c->barGen.store(gen + 1, std::memory_order_release);
        return true;
    }
    const double t0 = HostTrace::Now();
    for (uint32_t spins = 0; c->barGen.load(std::memory_order_acquire) == gen; ++spins) {
        if ((spins & 1023u) == 1023u) {
            sched_yield();
            if (HostTrace::Now() - t0 > timeoutSeconds * 1e3) return false;  // a rank died: fail instead of hanging
        }
    }
    return true;
}
static void* MapShm(const char* name, size_t bytes, bool create) {
    const int fd = shm_open(name, create ? (O_CREAT | O_RDWR) : O_RDWR, 0600);
    if (fd < 0) return nullptr;
    if (create && ftruncate(fd, (off_t)bytes) != 0) {
        close(fd);
        return nullptr;
    }
    if (!create) {  // the creator may not have sized it yet
        struct stat st;
        for (int i = 0; i < 20000 && (fstat(fd, &st) != 0 || (size_t)st.st_size < bytes); ++i) usleep(100);
    }
    void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    return p == MAP_FAILED ? nullptr : p;
}
// Spread the pages of a fresh window over the NUMA nodes of the box (before they are touched): the GPUs of an 8-GPU box hang off two
// sockets, and eight copy engines writing memory of ONE node reached 87 GB/s in total.  Best effort: containers may forbid mbind.
static void InterleaveOverNumaNodes(void* ptr, size_t bytes) {
    if (getenv("OMM_B200_NO_NUMA_INTERLEAVE")) return;
    unsigned long mask = 0;
    if (FILE* f = fopen("/sys/devices/system/node/online", "r")) {  // e.g. "0-1" or "0,2-3"
        char buf[128] = {0};
        if (fgets(buf, sizeof(buf), f)) {
            for (char* p = buf; *p;) {
                char* end = nullptr;
                const long a = strtol(p, &end, 10);
                if (end == p) break;
                long b = a;
                p = end;
                if (*p == '-') {
                    b = strtol(p + 1, &end, 10);
                    p = end;
                }
                for (long n = a; n <= b && n < 64; ++n) mask |= 1ul << n;
                if (*p == ',') ++p;
                else break;
            }
        }
        fclose(f);
    }
    if (__builtin_popcountl(mask) < 2) return;
#if defined(SYS_mbind)
    constexpr int kMpolInterleave = 3;
    if (syscall(SYS_mbind, ptr, bytes, kMpolInterleave, &mask, (unsigned long)(8 * sizeof(mask) + 1), 0) != 0 && HostTrace::Enabled())
        fprintf(stderr, "[omm-b200 trace] mbind(MPOL_INTERLEAVE) on the shared window was refused; pages stay on the creating rank's node\n");
#endif
}
// The window the root's host copy of this bake's arrayData lives in (every rank returns its own mapping of the same memory).
static SharedHostWindow* AcquireSharedWindow(ShardState& sh, size_t bytes, const Logger& log) {
    ShmControl* c = sh.ctl;
    if (!c) return nullptr;
    char name[96];
    if (sh.rank == 0) {
        SharedHostWindow* best = nullptr;
        for (SharedHostWindow& w : sh.windows)
            if (!w.inUse && w.capacity >= bytes && (!best || w.capacity < best->capacity)) best = &w;
        if (!best) {
            SharedHostWindow w;
            w.id = (int)sh.windows.size();
            w.capacity = (bytes + (bytes >> 3) + ((size_t)2 << 20)) & ~(((size_t)2 << 20) - 1);
            ShmName(name, sizeof(name), sh.idHash, "w", w.id);
            shm_unlink(name);
            // a tmpfs that cannot back the window would only fail when its pages are touched (SIGBUS): ask first
            struct statvfs vfs;
            const bool roomy = statvfs("/dev/shm", &vfs) != 0 || (unsigned long long)vfs.f_bavail * vfs.f_frsize > (unsigned long long)w.capacity + ((unsigned long long)64 << 20);
            w.ptr = roomy ? MapShm(name, w.capacity, true) : nullptr;
            if (w.ptr) InterleaveOverNumaNodes(w.ptr, w.capacity);
            if (w.ptr && cudaHostRegister(w.ptr, w.capacity, cudaHostRegisterPortable) != cudaSuccess) {
                cudaGetLastError();
                munmap(w.ptr, w.capacity);
                shm_unlink(name);
                w.ptr = nullptr;
            }
            if (!w.ptr) log.Log(ommMessageSeverity_Fatal, "[omm-b200] could not create the shared page-locked result window (shm_open / cudaHostRegister)");
            else {
                sh.windows.push_back(w);
                best = &sh.windows.back();
            }
        }
        // publish (id -1: no window; the ranks then fall back to the root's own download)
        c->winId = best ? best->id : -1;
        c->winCapacity = best ? best->capacity : 0;
        c->winSeq.store(sh.bakeSeq, std::memory_order_release);
        if (best) best->inUse = true;
        return best;
    }
    const double t0 = HostTrace::Now();
    for (uint32_t spins = 0; c->winSeq.load(std::memory_order_acquire) != sh.bakeSeq; ++spins)
        if ((spins & 1023u) == 1023u) {
            sched_yield();
            if (HostTrace::Now() - t0 > 120e3) return nullptr;
        }
    const int id = c->winId;
    const size_t capacity = (size_t)c->winCapacity;
    if (id < 0) return nullptr;
    for (SharedHostWindow& w : sh.windows)
        if (w.id == id) return &w;
    SharedHostWindow w;
    w.id = id;
    w.capacity = capacity;
    ShmName(name, sizeof(name), sh.idHash, "w", id);
    w.ptr = MapShm(name, capacity, false);
    if (w.ptr && cudaHostRegister(w.ptr, capacity, cudaHostRegisterPortable) != cudaSuccess) {
        cudaGetLastError();
        munmap(w.ptr, capacity);
        w.ptr = nullptr;
    }
    if (!w.ptr) {
        log.Log(ommMessageSeverity_Fatal, "[omm-b200] could not map the shared page-locked result window");
        return nullptr;
    }
    sh.windows.push_back(w);
    return &sh.windows.back();
}
void ReleaseSharedWindow(BakerObject* baker, int windowId) {
    std::lock_guard<std::mutex> g(baker->mu);
    for (SharedHostWindow& w : baker->shard.windows)
        if (w.id == windowId) w.inUse = false;
}

__global__ void WriteDescs(const ItemRec* __restrict__ items, const int32_t* __restrict__ special, const uint32_t* __restrict__ descOfItem,
                           const unsigned long long* __restrict__ offsetOfItem, uint32_t itemBegin, uint32_t itemEnd,
                           ommCpuOpacityMicromapDesc* __restrict__ descArray) {
    const uint32_t w = itemBegin + blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= itemEnd || special[w] != 0) return;
    ommCpuOpacityMicromapDesc d;
    d.offset = (uint32_t)offsetOfItem[w];
    d.subdivisionLevel = items[w].level;
    d.format = items[w].format;
    descArray[descOfItem[w]] = d;
}
// byte offset in arrayData at which the blocks of shard r start (r = 0..shards; the last entry is the array size)
__global__ void ShardByteOffsets(const unsigned long long* __restrict__ offsetOfItem, const ShardBound* __restrict__ bounds, int shards,
                                 unsigned long long* __restrict__ out) {
    const int r = threadIdx.x;
    if (r <= shards) out[r] = offsetOfItem[bounds[r].item];
}

struct HostReadback {
    uint32_t counters[32];
    uint32_t hist[52];
    uint32_t stream[4];
    ShardBound bounds[kMaxShards + 1];
    uint32_t chunkFirst[kMaxShardsPerRank][kHierMaxChunks + 1];
    unsigned long long shardOff[kMaxShards + 1];
    unsigned long long totalUnits, totalWords, myMicroTris;
};

ommResult BakeOnDevice(BakerObject* baker, const StagedInputs& in, void* userStream, BakeResultObject* res, ommB200BakeTimings* tm, bool earlyDownload) {
    const Logger& log = baker->log;
    ommResult rc = RequireDevice(log, baker->device);
    if (rc != ommResult_SUCCESS) return rc;
    const ommCpuBakeInputDesc& d = in.desc;
    // the texture's texels are read-only here; only its cache of cell tables (internally locked) is touched through this pointer
    TextureObject* tex = HandlePtr<TextureObject>(d.texture);
    CellTables* cellTables = nullptr;  // released after the final stream synchronisation
    const uint32_t flags = (uint32_t)d.bakeFlags;
    const uint32_t T = in.triangleCount;
    const int disableDup = (flags & ommCpuBakeFlags_DisableDuplicateDetection) != 0;
    const bool validation = (flags & ommCpuBakeFlags_EnableValidation) != 0;
    const bool limitWorkload = d.maxWorkloadSize != 0xFFFFFFFFFFFFFFFFull;
    const int world = baker->shard.world, rank = baker->shard.rank;
    const bool hostPasses = HostPassesNeeded(d);
    const int TPB = 256;
    uint32_t launches = 0;
    // ref: bake_cpu_impl.cpp:1873-1902 -- width of the index buffer
    const bool allow8 = (flags & ommCpuBakeFlags_Allow8BitIndices) != 0, force32 = (flags & ommCpuBakeFlags_Force32BitIndices) != 0;
    const ommIndexFormat ifmt = (allow8 && (int32_t)T <= 127 && !force32) ? ommIndexFormat_UINT_8 : (((int32_t)T <= 32767 && !force32) ? ommIndexFormat_UINT_16 : ommIndexFormat_UINT_32);
    const int indexBytes = ifmt == ommIndexFormat_UINT_8 ? 1 : (ifmt == ommIndexFormat_UINT_16 ? 2 : 4);

    // Everything the host reads back from the device during a bake lives in one page-locked block: the copies are then really asynchronous
    // (a copy into pageable memory holds the host until it has completed), so each of the two read-back points costs one round trip, not
    // one per array.
    HostReadback stackReadback;
    HostReadback* pinnedReadback = (HostReadback*)PinnedPoolAcquire(sizeof(HostReadback));
    HostReadback& hb = pinnedReadback ? *pinnedReadback : stackReadback;
    cudaStream_t stream = (cudaStream_t)userStream;
    bool ownStream = false;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaStream_t copyStream = nullptr;  // early download of the packed array (see K8)
    cudaEvent_t copyEv[2] = {nullptr, nullptr}, sliceEv = nullptr;
    cudaEvent_t gatherEv[3] = {nullptr, nullptr, nullptr};  // OMM_B200_TRACE: phases of the sharded exchange
    cudaStream_t laneStream[kMaxChunkLanes] = {};            // classifier chunks in flight side by side (lane 0 = the bake's stream)
    cudaEvent_t laneEv[kMaxChunkLanes] = {}, lanePrepEv = nullptr;
    Scratch scratch{};
    void* cubTemp = nullptr;
    size_t cubTempBytes = 0;

    // device arrays
    float2* triUV = nullptr;
    int8_t* triLevel = nullptr;
    uint8_t *triFormat = nullptr, *triDegenerate = nullptr;
    HostLevelFix* fixList = nullptr;
    uint32_t* counters = nullptr;  // [0] edge-heuristic fix count, [1] disabled triangles, [2] work items, [8..23] work items per level
    unsigned long long* workloadDev = nullptr;
    uint64_t *triKey = nullptr, *tableKeys = nullptr;
    uint32_t *tableVals = nullptr, *triFirst = nullptr, *isItem = nullptr, *itemScan = nullptr, *triItem = nullptr, *triFinal = nullptr;
    ItemRec *itemsW = nullptr, *items = nullptr;  // first-seen order (the SDK's vmWorkItems order) / output order (K3b)
    uint32_t *posOfOrig = nullptr, *origOf = nullptr;
    unsigned long long *itemUnits = nullptr, *itemWords = nullptr, *unitStart = nullptr, *wordStart = nullptr, *itemNodes = nullptr, *nodeStart = nullptr;
    ShardBound* boundsDev = nullptr;
    uint32_t* chunkFirstDev = nullptr;
    const int numShards = world * ShardsPerRank(world);  // see ShardOwner
    OwnedShards owned{};
    // ommCpuBake on one GPU: the array is packed and written to the caller-visible host memory chunk by chunk WHILE later chunks are
    // classified (possible because the items are in output order, K3b); see "streamed" below
    const bool streamCandidate = earlyDownload && world == 1 && !hostPasses && res->usesDefaultAllocator && StreamedDownloadEnabled() &&
                                 (flags & ommCpuBakeFlags_DisableSpecialIndices) == 0 && T >= 4096;
    unsigned long long hierChunkRegions = HierNominalChunkRegions(baker->device);
    if (streamCandidate) {
        hierChunkRegions = std::max(kHierChunkRegionsMin, hierChunkRegions / StreamChunkDivisor());
        if (const char* e = getenv("OMM_B200_STREAM_CHUNK_REGIONS")) {  // tests: many chunks at small sizes (read per bake)
            const unsigned long long v = strtoull(e, nullptr, 10);
            if (v >= 4096) hierChunkRegions = v;
        }
    }
    bool streaming = false, streamFallback = false;
    bool lanePost = false;  // the per-item post pass ran chunk by chunk on the classifier's lanes
    std::vector<cudaEvent_t> chunkEvs;           // streamed: "chunk c is packed into the device array"
    unsigned long long* chunkEndHost = nullptr;  // streamed: arrayData bytes emitted up to and including chunk c (page-locked, read by the host as chunks complete)
    uint32_t *runDesc = nullptr, *conflictDev = nullptr;
    unsigned long long* runBytes = nullptr;
    unsigned long long worstBytes = 0;
    auto& streamHost = hb.stream;  // [0] descriptors emitted by the chunks, [1] conflict flag of the optimistic dedup, [2], [3] the pair that raised it
    auto& chunkFirst = hb.chunkFirst;
    BigItemList bigItems;  // work items whose blocks get a warp of their own in the post pass
    uint32_t* stateWords = nullptr;
    uint32_t* uniformVotes = nullptr;  // per work item: initial regions proved above / below the cutoff (hierarchical classifier only)
    uint64_t* digest = nullptr;
    int32_t* special = nullptr;
    uint32_t *mergeRoot = nullptr, *survivor2 = nullptr;
    uint32_t *survivor = nullptr, *hist = nullptr, *sortKeysIn = nullptr, *sortValsIn = nullptr, *sortKeysOut = nullptr, *emit = nullptr, *descOfItem = nullptr;
    unsigned long long *blockBytes = nullptr, *offsetOfItem = nullptr, *shardOffDev = nullptr;
    unsigned long long microTris = 0;
    unsigned long long &totalUnits = hb.totalUnits, &totalWords = hb.totalWords, &myMicroTris = hb.myMicroTris;
    auto& shardOff = hb.shardOff;
    uint32_t W = 0;
    auto& countersHost = hb.counters;
    auto& histHost = hb.hist;
    auto& bounds = hb.bounds;
    uint32_t numDescs = 0;
    unsigned long long arrayBytes = 0;
    bool resort = false;  // Compress changed item levels: the serialized items are sorted again (see ResortKeys)
    SharedHostWindow* window = nullptr;  // sharded ommCpuBake: the root's host copy of arrayData, written by every rank
    bool windowProtocol = false;         // this bake takes part in the window hand-shake (every rank decides alike)
    unsigned long long sharedD2hBytes = 0;
    const uint32_t gridT = (T + TPB - 1) / TPB;
    int sms = 1;

    BakeParams P{};
    SetupArgs sa{};
    memset(&hb, 0, sizeof(hb));

    if (world > kMaxShards) return ommResult_INVALID_ARGUMENT;
    for (int v = 0; v < numShards; ++v)
        if (ShardOwner(v, world) == rank) owned.shard[owned.count++] = v;
    if (!stream) {
        if (PoolStreamCreate(&stream) != cudaSuccess) {
            cudaGetLastError();
            return ommResult_FAILURE;
        }
        ownStream = true;
    }
    scratch.stream = stream;
    NvtxRange nvtxBake("omm-b200 bake");
    for (int i = 0; i < 6; ++i) CUDA_TRY(PoolEventCreate(&ev[i]));
    CUDA_TRY(cudaEventRecord(ev[0], stream));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, baker->device));
    HostTrace::Mark("stream + events created");

    // ---- parameters ----
    P.tex = tex->dev;
    P.addrMode = d.runtimeSamplerDesc.addressingMode;
    P.filterLinear = d.runtimeSamplerDesc.filter == ommTextureFilterMode_Linear;
    P.borderAlpha = d.runtimeSamplerDesc.borderAlpha;
    P.cutoff = d.alphaCutoff;
    P.stateGT = d.alphaCutoffGreater;
    P.stateLE = d.alphaCutoffLessEqual;
    P.globalFormat = d.format;
    P.promotion = d.unknownStatePromotion;
    P.pow2Mip0 = tex->dev.mips[0].isPow2;
    P.useCoarse = tex->dev.sat != nullptr && tex->mipCount == 1 && P.filterLinear;
    P.coarseSameCutoff = tex->alphaCutoff == d.alphaCutoff;
    P.disableFine = (flags & (1u << 9)) != 0;
    P.disableLevelLine = (flags & (1u << 8)) != 0;
    P.aabbTesting = (flags & (1u << 7)) != 0;
    P.skipUniformFill = (flags & ommCpuBakeFlags_DisableSpecialIndices) == 0 && !hostPasses;

    sa.indices = in.devIndices;
    sa.texCoords = in.devTexCoords;
    sa.levels = in.devLevels;
    sa.formats = in.devFormats;
    sa.triCount = T;
    sa.texCoordStride = in.texCoordStride;
    sa.indexFormat = d.indexFormat;
    sa.texCoordFormat = d.texCoordFormat;
    sa.globalFormat = d.format;
    sa.maxLevel = d.maxSubdivisionLevel;
    sa.dynScale = d.dynamicSubdivisionScale;
    sa.edgeHeuristic = (flags & (1u << 11)) != 0;
    sa.disableLevelLine = P.disableLevelLine;
    sa.texW = (uint32_t)tex->dev.mips[0].w;
    sa.texH = (uint32_t)tex->dev.mips[0].h;

    // ---- K1: triangles ----
    if (T > 0) {
        NvtxRange nvtxSetup("setup: triangles, UV pre-dedup, work items in output order");
        CUDA_TRY(scratch.alloc(&triUV, (size_t)3 * T));
        CUDA_TRY(scratch.alloc(&triLevel, T));
        CUDA_TRY(scratch.alloc(&triFormat, T));
        CUDA_TRY(scratch.alloc(&triDegenerate, T));
        CUDA_TRY(scratch.alloc(&fixList, T));
        CUDA_TRY(scratch.alloc(&counters, 32));
        CUDA_TRY(scratch.alloc(&workloadDev, 1));
        CUDA_TRY(scratch.alloc(&triKey, T));
        CUDA_TRY(scratch.alloc(&triFirst, T));
        CUDA_TRY(scratch.alloc(&isItem, T));
        CUDA_TRY(scratch.alloc(&itemScan, T));
        CUDA_TRY(scratch.alloc(&triItem, T));
        CUDA_TRY(scratch.alloc(&triFinal, T));
        CUDA_TRY(scratch.alloc(&boundsDev, kMaxShards + 1));
        CUDA_TRY(scratch.alloc(&chunkFirstDev, (size_t)kMaxShardsPerRank * (kHierMaxChunks + 1)));
        CUDA_TRY(cudaMemsetAsync(counters, 0, 32 * sizeof(uint32_t), stream));
        CUDA_TRY(cudaMemsetAsync(workloadDev, 0, sizeof(unsigned long long), stream));
        SetupTriangles<<<gridT, TPB, 0, stream>>>(sa, triUV, triLevel, triFormat, triDegenerate, fixList, counters);
        launches++;
        if (sa.dynScale > 0.f) {
            uint32_t fixCount = 0;
            CUDA_TRY(cudaMemcpyAsync(&fixCount, counters, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaStreamSynchronize(stream));
            if (fixCount) {
                std::vector<HostLevelFix> fixes(fixCount);
                std::vector<int8_t> lv(fixCount);
                CUDA_TRY(cudaMemcpyAsync(fixes.data(), fixList, sizeof(HostLevelFix) * fixCount, cudaMemcpyDeviceToHost, stream));
                CUDA_TRY(cudaStreamSynchronize(stream));
                for (uint32_t i = 0; i < fixCount; ++i) lv[i] = FinishEdgeHeuristic(fixes[i].eMax, sa.dynScale, sa.maxLevel);
                int8_t* lvDev = nullptr;
                CUDA_TRY(scratch.alloc(&lvDev, fixCount));
                CUDA_TRY(cudaMemcpyAsync(lvDev, lv.data(), fixCount, cudaMemcpyHostToDevice, stream));
                ApplyLevelFixes<<<(fixCount + TPB - 1) / TPB, TPB, 0, stream>>>(fixList, lvDev, fixCount, triLevel);
                launches++;
                CUDA_TRY(cudaStreamSynchronize(stream));  // lv (host vector) must outlive the copy
            }
        }
        if (validation) {
            CountDisabled<<<gridT, TPB, 0, stream>>>(triLevel, T, counters + 1);
            launches++;
        }

        // ---- K2: UV pre-dedup ----
        const uint64_t cap = NextPow2((uint64_t)T * 2 + 16);
        if (!disableDup) {
            CUDA_TRY(scratch.alloc(&tableKeys, cap));
            CUDA_TRY(scratch.alloc(&tableVals, cap));
            FillTable<<<(uint32_t)((cap + TPB - 1) / TPB), TPB, 0, stream>>>(tableKeys, tableVals, cap);
            UvTableInsert<<<gridT, TPB, 0, stream>>>(triUV, triLevel, triFormat, T, triKey, tableKeys, tableVals, cap - 1);
            launches += 2;
        }
        UvTableResolve<<<gridT, TPB, 0, stream>>>(triLevel, triKey, T, tableKeys, tableVals, cap - 1, disableDup, triFirst, isItem);
        launches++;

        // ---- K3: items ----
        {
            size_t need = 0, tmp = 0;
            CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, need, isItem, itemScan, (int)T, stream));
            CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int)T + 1, stream));
            need = std::max(need, tmp);
            CUDA_TRY(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr,
                                                                (int)T, 0, 32, stream));
            need = std::max(need, tmp);
            CUDA_TRY(cub::DeviceScan::ExclusiveScan(nullptr, tmp, (unsigned long long*)nullptr, (unsigned long long*)nullptr, cub::Sum(),
                                                     cub::FutureValue<unsigned long long>((unsigned long long*)nullptr), (int)T + 1, stream));
            need = std::max(need, tmp);
            cubTempBytes = need;
            CUDA_TRY(scratch.alloc((uint8_t**)&cubTemp, cubTempBytes));
            tmp = cubTempBytes;
            CUDA_TRY(cub::DeviceScan::ExclusiveSum(cubTemp, tmp, isItem, itemScan, (int)T, stream));
            launches += 2;
        }
        CUDA_TRY(scratch.alloc(&itemsW, T));
        CUDA_TRY(scratch.alloc(&items, T));
        CUDA_TRY(scratch.alloc(&posOfOrig, T));
        CUDA_TRY(scratch.alloc(&origOf, T));
        CUDA_TRY(scratch.alloc(&sortKeysIn, T));
        CUDA_TRY(scratch.alloc(&sortValsIn, T));
        CUDA_TRY(scratch.alloc(&sortKeysOut, T));
        CUDA_TRY(scratch.alloc(&itemUnits, (size_t)T + 1));
        CUDA_TRY(scratch.alloc(&itemWords, (size_t)T + 1));
        CUDA_TRY(scratch.alloc(&unitStart, (size_t)T + 1));
        CUDA_TRY(scratch.alloc(&wordStart, (size_t)T + 1));
        CUDA_TRY(scratch.alloc(&itemNodes, (size_t)T + 1));
        CUDA_TRY(scratch.alloc(&nodeStart, (size_t)T + 1));
        BuildItems<<<gridT, TPB, 0, stream>>>(triUV, triLevel, triFormat, triDegenerate, isItem, itemScan, T, itemsW, counters + 8, counters + 2);
        // ---- K3b: output order first (see ItemSortKey) ----
        ItemKeysAll<<<gridT, TPB, 0, stream>>>(itemsW, counters + 2, T, sortKeysIn, sortValsIn);
        {
            size_t tmp = cubTempBytes;
            // bits [0, 31): the key is (level < 13) << 26 | 26 Morton bits, plus one
            CUDA_TRY(cub::DeviceRadixSort::SortPairsDescending(cubTemp, tmp, sortKeysIn, sortKeysOut, sortValsIn, origOf, (int)T, 0, 31, stream));
        }
        PermuteItems<<<(T + 1 + TPB - 1) / TPB, TPB, 0, stream>>>(itemsW, origOf, counters + 2, T, items, posOfOrig, itemUnits, itemWords, itemNodes);
        MapTrianglesToItems<<<gridT, TPB, 0, stream>>>(triFirst, itemScan, posOfOrig, T, triItem);
        launches += 4 + 8;
        {
            // exclusive scans over T+1 entries: entry [W] (and everything after it) holds the grand total
            size_t tmp = cubTempBytes;
            CUDA_TRY(cub::DeviceScan::ExclusiveSum(cubTemp, tmp, itemUnits, unitStart, (int)T + 1, stream));
            tmp = cubTempBytes;
            CUDA_TRY(cub::DeviceScan::ExclusiveSum(cubTemp, tmp, itemWords, wordStart, (int)T + 1, stream));
            tmp = cubTempBytes;
            CUDA_TRY(cub::DeviceScan::ExclusiveSum(cubTemp, tmp, itemNodes, nodeStart, (int)T + 1, stream));
            launches += 6;
        }
        const unsigned long long* weightStart = unitStart;
        if (world > 1 && !CostBalanceDisabled()) {
            unsigned long long *itemCost = nullptr, *costStart = nullptr;
            CUDA_TRY(scratch.alloc(&itemCost, (size_t)T + 1));
            CUDA_TRY(scratch.alloc(&costStart, (size_t)T + 1));
            if (P.tex.isFp32) ItemCostKernel<true><<<(T + 1 + TPB - 1) / TPB, TPB, 0, stream>>>(P, items, counters + 2, T, itemUnits, itemCost);
            else ItemCostKernel<false><<<(T + 1 + TPB - 1) / TPB, TPB, 0, stream>>>(P, items, counters + 2, T, itemUnits, itemCost);
            size_t tmp = cubTempBytes;
            CUDA_TRY(cub::DeviceScan::ExclusiveSum(cubTemp, tmp, itemCost, costStart, (int)T + 1, stream));
            weightStart = costStart;
            launches += 3;
        }
        ShardBounds<<<1, 128, 0, stream>>>(weightStart, unitStart, wordStart, nodeStart, T + 1, numShards, owned, hierChunkRegions, boundsDev, chunkFirstDev);
        launches++;
        CUDA_TRY(cudaMemcpyAsync(countersHost, counters, sizeof(countersHost), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaMemcpyAsync(bounds, boundsDev, sizeof(ShardBound) * (numShards + 1), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaMemcpyAsync(chunkFirst, chunkFirstDev, sizeof(chunkFirst), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaMemcpyAsync(&totalUnits, unitStart + T, 8, cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaMemcpyAsync(&totalWords, wordStart + T, 8, cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        for (int l = 0; l <= kMaxLevel; ++l) {
            W += countersHost[8 + l];
            microTris += (unsigned long long)countersHost[8 + l] << (2 * l);
        }
        // ref: bake_cpu_impl.cpp:652-657
        if (validation && countersHost[1] != 0)
            log.Logf(ommMessageSeverity_Info, "[Info] - The workload consists of %d unclassifiable triangles, these will be classified as unresolvedTriState = %s.",
                     countersHost[1], SpecialIndexText((int)d.unresolvedTriState));
        // ref: bake_cpu_impl.cpp:682-713
        if ((validation || limitWorkload) && W > 0) {
            unsigned long long workload = 0;
            WorkloadKernel<<<(W + TPB - 1) / TPB, TPB, 0, stream>>>(items, W, (float)tex->dev.mips[0].w, (float)tex->dev.mips[0].h, workloadDev);
            launches++;
            CUDA_TRY(cudaMemcpyAsync(&workload, workloadDev, 8, cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaStreamSynchronize(stream));
            if (limitWorkload && workload > d.maxWorkloadSize) {
                rc = ommResult_WORKLOAD_TOO_BIG;
                goto cleanup;
            }
            if (validation && workload > (1ull << 27))
                log.Logf(ommMessageSeverity_PerfWarning,
                         "[Perf Warning] - The workload consists of %lld work items (number of texels to classify), which corresponds to roughly %lld 1024x1024 "
                         "textures. This is unusually large and may result in long bake times.",
                         (long long)workload, (long long)(workload >> 20));
        }
    }
    CUDA_TRY(cudaEventRecord(ev[1], stream));
    HostTrace::Mark("setup done (host sync 1)");

    // ---- K4: classification of this rank's shard of work items ----
    if (W > 0) {
        NvtxRange nvtxClassify("classify + per-item post pass");
        uint32_t myItems = 0;
        for (int k = 0; k < owned.count; ++k) myItems += bounds[owned.shard[k] + 1].item - bounds[owned.shard[k]].item;
        CUDA_TRY(scratch.alloc(&stateWords, (size_t)totalWords + 4));
        CUDA_TRY(scratch.alloc(&digest, W));
        CUDA_TRY(scratch.alloc(&special, W));
        CUDA_TRY(scratch.alloc(&survivor, W));
        CUDA_TRY(scratch.alloc(&hist, 64));
        CUDA_TRY(scratch.alloc(&emit, (size_t)W + 1));
        CUDA_TRY(scratch.alloc(&descOfItem, (size_t)W + 1));
        CUDA_TRY(scratch.alloc(&blockBytes, (size_t)W + 1));
        CUDA_TRY(scratch.alloc(&offsetOfItem, (size_t)W + 1));
        CUDA_TRY(scratch.alloc(&shardOffDev, kMaxShards + 1));
        CUDA_TRY(cudaMemsetAsync(hist, 0, 64 * sizeof(uint32_t), stream));
        for (uint32_t l = kBigHashLevel; l <= (uint32_t)kMaxLevel; ++l) bigItems.capacity += countersHost[8 + l];
        if (bigItems.capacity) {
            CUDA_TRY(scratch.alloc(&bigItems.list, bigItems.capacity));
            CUDA_TRY(scratch.alloc(&bigItems.count, 1));
        }
        HierKernels hier{};
        const bool useHier = SelectHierKernels(P, &hier);
        const uint64_t tableCap = NextPow2((uint64_t)W * 2 + 16);
        if (streamCandidate && useHier && myItems > 0 && chunkFirst[0][2] < bounds[1].item) {  // at least three chunks
            // ---- streamed: sizes are not known before the last chunk, so the array is laid out in buffers of the worst-case size
            // (every work item serialized); the host one comes from the page-locked pool and is written by PackItems directly ----
            for (int l = 0; l <= kMaxLevel; ++l) {
                const unsigned long long bytes = ((1ull << (2 * l)) * (unsigned long long)d.format) >> 3;
                worstBytes += (unsigned long long)countersHost[8 + l] * (bytes > 1 ? bytes : 1);
            }
            if (worstBytes >= ((unsigned long long)8 << 20) && worstBytes <= ((unsigned long long)3 << 30) && AllocHostArrayData(res, (size_t)worstBytes) &&
                res->arrayDataFromPinnedPool) {
                chunkEndHost = (unsigned long long*)PinnedPoolAcquire(sizeof(unsigned long long) * (kHierMaxChunks + 1));
                if (!chunkEndHost) {
                    rc = ommResult_FAILURE;
                    goto cleanup;
                }
                CUDA_TRY(scratch.alloc(&runDesc, 4));
                CUDA_TRY(scratch.alloc(&runBytes, 2));
                conflictDev = runDesc + 1;
                CUDA_TRY(cudaMemsetAsync(runDesc, 0, 4 * sizeof(uint32_t), stream));
                CUDA_TRY(cudaMemsetAsync(runBytes, 0, 2 * sizeof(unsigned long long), stream));
                CUDA_TRY(cudaMallocAsync(&res->devArrayData, (size_t)worstBytes + 16, stream));
                if (!disableDup) {
                    FillTable<<<(uint32_t)((tableCap + TPB - 1) / TPB), TPB, 0, stream>>>(tableKeys, tableVals, tableCap);
                    launches++;
                }
                streaming = true;
            }
        }
        if (myItems > 0 && useHier) {
            if (P.tex.mipCount == 1) cellTables = AcquireCellTables(tex, P, stream, &launches);  // (H), (I): built on first use per texture and cutoff
            // worst case of a chunk: the nominal number of initial regions plus the rest of its last item (at most 4^9 regions at level 12)
            unsigned long long cap = 0;
            for (int k = 0; k < owned.count; ++k) {
                const unsigned long long regions = bounds[owned.shard[k] + 1].node - bounds[owned.shard[k]].node;
                cap = std::max(cap, std::min<unsigned long long>(HierChunkRegions(regions, hierChunkRegions) + (1ull << 18), regions));
            }
            HierItem* hierItems = nullptr;
            // lanes: chunks in flight side by side, each with lists and a stream of its own (HierChunkLanes); lane 0 is the bake's stream.
            // (A rank of a sharded bake has one chunk per shard at these sizes: nothing to put side by side.)
            int lanes = (streaming || world > 1) ? 1 : HierChunkLanes();
            {
                int chunks = 0;
                for (int k = 0; k < owned.count; ++k)
                    for (int c = 0; c < kHierMaxChunks && chunkFirst[k][c] < bounds[owned.shard[k] + 1].item; ++c) chunks += chunkFirst[k][c + 1] > chunkFirst[k][c];
                lanes = std::max(1, std::min(lanes, chunks));  // a bake of one chunk keeps to the bake's stream
            }
            HierLists laneLists[kMaxChunkLanes] = {};
            CUDA_TRY(scratch.alloc(&hierItems, (size_t)W * (size_t)P.tex.mipCount));
            for (int l = 0; l < lanes; ++l) {
                HierLists& lists = laneLists[l];
                cudaError_t e = scratch.alloc(&lists.q[0], (size_t)cap);
                if (e == cudaSuccess) e = scratch.alloc(&lists.q[1], (size_t)cap * 4);
                if (e == cudaSuccess) e = scratch.alloc(&lists.q[2], (size_t)cap * 16);
                if (e == cudaSuccess) e = scratch.alloc(&lists.unresolved, (size_t)cap);
                if (e == cudaSuccess) e = scratch.alloc(&lists.slow, (size_t)cap * 16);
                if (e == cudaSuccess) e = scratch.alloc(&lists.count, 8);  // [0..2] the lists, [3] unresolved initial regions, [5] slow-path 4-regions
                if (e != cudaSuccess && l > 0) {  // no memory for another set of lists: the bake runs with the lanes it has
                    cudaGetLastError();
                    lanes = l;
                    break;
                }
                CUDA_TRY(e);
            }
            laneStream[0] = stream;
            if (lanes > 1) {
                CUDA_TRY(PoolEventCreate(&lanePrepEv, false));
                for (int l = 1; l < lanes; ++l) {
                    CUDA_TRY(PoolStreamCreate(&laneStream[l]));
                    CUDA_TRY(PoolEventCreate(&laneEv[l], false));
                }
            }
            // with lanes, the per-item post pass of a chunk follows its leaves on the chunk's lane (its state words are still in L2 and the
            // other lanes keep the SMs busy); the big-block digest kernel shares one work list per bake, so bakes with such items post afterwards
            lanePost = lanes > 1 && bigItems.capacity == 0;
            CUDA_TRY(scratch.alloc(&uniformVotes, (size_t)W * 2));
            CUDA_TRY(cudaMemsetAsync(uniformVotes, 0, sizeof(uint32_t) * 2 * (size_t)W, stream));
            const uint32_t listGrid = (uint32_t)std::max(sms, 1) * HierGridBlocksPerSm(), initGrid = (uint32_t)std::max(sms, 1) * HierInitGridBlocksPerSm(),
                           leafGrid = (uint32_t)std::max(sms, 1) * HierLeafGridBlocksPerSm(), slowGrid = (uint32_t)std::max(sms, 1) * HierSlowGridBlocksPerSm();
            for (int k = 0; k < owned.count; ++k) {
                const uint32_t itemBegin = bounds[owned.shard[k]].item, itemEnd = bounds[owned.shard[k] + 1].item;
                if (itemEnd <= itemBegin) continue;
                HierPrepare<<<(itemEnd - itemBegin + TPB - 1) / TPB, TPB, 0, stream>>>(P, items, itemBegin, itemEnd, hierItems);
                launches++;
                if (lanes > 1) {  // the other lanes start behind the shard's constants (and everything else enqueued so far)
                    CUDA_TRY(cudaEventRecord(lanePrepEv, stream));
                    for (int l = 1; l < lanes; ++l) CUDA_TRY(cudaStreamWaitEvent(laneStream[l], lanePrepEv, 0));
                }
                int nextLane = 0;
                for (int c = 0; c < kHierMaxChunks; ++c) {
                    const uint32_t i0 = chunkFirst[k][c], i1 = chunkFirst[k][c + 1];
                    if (i0 >= itemEnd) break;
                    if (i1 <= i0) continue;
                    const HierLists& lists = laneLists[nextLane];
                    const cudaStream_t cs = laneStream[nextLane];
                    nextLane = (nextLane + 1) % lanes;
                    CUDA_TRY(cudaMemsetAsync(lists.count, 0, 8 * sizeof(unsigned long long), cs));
                    hier.initial<<<initGrid, kHierInitWarps * 32, 0, cs>>>(P, hierItems, nodeStart, wordStart, i0, i1, lists, uniformVotes, stateWords);
                    hier.unresolved<<<listGrid, 128, 0, cs>>>(P, hierItems, wordStart, lists, uniformVotes, stateWords);
                    hier.list<<<listGrid, 128, 0, cs>>>(P, hierItems, wordStart, lists.q[0], lists.count + 0, lists.q[1], lists.count + 1, 0, stateWords);
                    hier.list<<<listGrid, 128, 0, cs>>>(P, hierItems, wordStart, lists.q[1], lists.count + 1, lists.q[2], lists.count + 2, 1, stateWords);
                    hier.leaves<<<leafGrid, 128, 0, cs>>>(P, items, hierItems, wordStart, lists, stateWords);
                    hier.leavesSlow<<<slowGrid, 128, 0, cs>>>(P, items, hierItems, wordStart, lists, stateWords);
                    launches += 6;
                    if (lanePost)
                        CUDA_TRY(LaunchItemPost(cs, items, wordStart, stateWords, i0, i1, d.rejectionThreshold, (flags & ommCpuBakeFlags_DisableSpecialIndices) != 0, 0, uniformVotes,
                                                (uint32_t)P.stateGT, (uint32_t)P.stateLE, digest, special, bigItems, &launches));
                    if (streaming) {
                        // the chunk's items are final: special indices + digests, dedup against the chunks so far, descriptor slots and
                        // byte offsets continuing the running totals, then the blocks go to the device array; the host thread forwards each
                        // chunk's byte range to a copy engine as soon as it is packed (below), so the array crosses PCIe while later chunks
                        // are classified.  (PackItems writing the page-locked host array directly was measured first: 52 GB/s, as fast as
                        // the copy engine -- but the SMs it ran on stalled behind the PCIe writes and the whole bake took 5 ms longer.)
                        const uint32_t n = i1 - i0, gridC = (n + TPB - 1) / TPB;
                        CUDA_TRY(LaunchItemPost(stream, items, wordStart, stateWords, i0, i1, d.rejectionThreshold, 0, 0, uniformVotes, (uint32_t)P.stateGT, (uint32_t)P.stateLE, digest,
                                                special, bigItems, &launches));
                        if (!disableDup) DigestInsertChunk<<<gridC, TPB, 0, stream>>>(digest, items, triItem, special, i0, i1, tableKeys, tableVals, tableCap - 1, conflictDev);
                        DigestResolve<<<gridC, TPB, 0, stream>>>(digest, items, triItem, i0, i1, tableKeys, tableVals, tableCap - 1, disableDup, survivor, special);
                        EmitInfo<<<gridC, TPB, 0, stream>>>(items, special, i0, i1, 0, (int)d.format, hist, emit, blockBytes);
                        size_t tmp = cubTempBytes;
                        CUDA_TRY(cub::DeviceScan::ExclusiveScan(cubTemp, tmp, emit + i0, descOfItem + i0, cub::Sum(), cub::FutureValue<uint32_t>(runDesc), (int)n, stream));
                        tmp = cubTempBytes;
                        CUDA_TRY(cub::DeviceScan::ExclusiveScan(cubTemp, tmp, blockBytes + i0, offsetOfItem + i0, cub::Sum(), cub::FutureValue<unsigned long long>(runBytes), (int)n,
                                                                 stream));
                        AdvanceRunningTotals<<<1, 1, 0, stream>>>(emit, blockBytes, descOfItem, offsetOfItem, i1 - 1, runDesc, runBytes, chunkEndHost + chunkEvs.size());
                        PackItems<<<std::min((uint32_t)std::max(sms, 1) * 32u, (n + 7) / 8), 256, 0, stream>>>(items, special, wordStart, stateWords, descOfItem, offsetOfItem, nullptr,
                                                                                                      nullptr, i0, i1, worstBytes, (uint8_t*)res->devArrayData, nullptr);
                        cudaEvent_t e = nullptr;
                        CUDA_TRY(PoolEventCreate(&e, false));
                        chunkEvs.push_back(e);
                        CUDA_TRY(cudaEventRecord(e, stream));
                        launches += 10;
                    }
                }
            }
            for (int l = 1; l < lanes; ++l) {  // the bake's stream continues behind every lane
                CUDA_TRY(cudaEventRecord(laneEv[l], laneStream[l]));
                CUDA_TRY(cudaStreamWaitEvent(stream, laneEv[l], 0));
            }
        } else if (myItems > 0) {
            const ClassifyFn classify = SelectClassifyKernel(P);
            for (int k = 0; k < owned.count; ++k) {
                const uint32_t itemBegin = bounds[owned.shard[k]].item, itemEnd = bounds[owned.shard[k] + 1].item;
                const unsigned long long unitBegin = bounds[owned.shard[k]].unit, unitEnd = bounds[owned.shard[k] + 1].unit;
                if (itemEnd <= itemBegin) continue;
                const unsigned long long unitsPerBlock = (unsigned long long)kClassifyWarps * (UseQueueKernel(P) ? kBatchUnits : 1);
                const unsigned long long blocks = (unitEnd - unitBegin + unitsPerBlock - 1) / unitsPerBlock;
                const unsigned long long kMaxGrid = 0x7FFFFFFFull;
                for (unsigned long long b0 = 0; b0 < blocks; b0 += kMaxGrid) {
                    const unsigned long long nb = std::min(kMaxGrid, blocks - b0);
                    classify<<<(uint32_t)nb, kClassifyWarps * 32, 0, stream>>>(P, items, unitStart, wordStart, itemBegin, itemEnd,
                                                                              unitBegin + b0 * unitsPerBlock, unitEnd, stateWords);
                    launches++;
                }
            }
        }
        myMicroTris = microTris;
        CUDA_TRY(cudaEventRecord(ev[4], stream));  // end of the classification kernels proper
        for (int k = 0; k < owned.count && !streaming && !lanePost; ++k) {
            const uint32_t itemBegin = bounds[owned.shard[k]].item, itemEnd = bounds[owned.shard[k] + 1].item;
            if (itemEnd <= itemBegin) continue;
            // special-index scan + XXH64 of this rank's items (their state words are local already)
            CUDA_TRY(LaunchItemPost(stream, items, wordStart, stateWords, itemBegin, itemEnd, d.rejectionThreshold, (flags & ommCpuBakeFlags_DisableSpecialIndices) != 0, 0, uniformVotes,
                                    (uint32_t)P.stateGT, (uint32_t)P.stateLE, digest, special, bigItems, &launches));
        }
        CUDA_TRY(cudaEventRecord(ev[5], stream));  // end of the per-item post pass
        if (world > 1) {
            // ---- the exchange, part 1: 12 bytes per work item (digest + special index), one group of broadcasts rooted at the shard owners ----
            // exact share of micro-triangles classified on this rank
            CUDA_TRY(cudaMemsetAsync(workloadDev, 0, sizeof(unsigned long long), stream));
            for (int k = 0; k < owned.count; ++k) {
                const uint32_t itemBegin = bounds[owned.shard[k]].item, itemEnd = bounds[owned.shard[k] + 1].item;
                if (itemEnd <= itemBegin) continue;
                SumMicroTriangles<<<(itemEnd - itemBegin + TPB - 1) / TPB, TPB, 0, stream>>>(items, itemBegin, itemEnd, workloadDev);
                launches++;
            }
            // (read back with the histograms: a copy into pageable memory here would hold the host until the classification has finished, and
            // the exchange below would be enqueued -- and start -- that much later)
            NcclApi& nccl = Nccl();
            if (!nccl.ok || !baker->shard.ncclComm) {
                log.Log(ommMessageSeverity_Fatal, "[omm-b200] sharded bake requested but NCCL is not initialised");
                rc = ommResult_FAILURE;
                goto cleanup;
            }
            const ncclComm_t comm = (ncclComm_t)baker->shard.ncclComm;
            bool ncclOk = nccl.GroupStart() == ncclSuccess;
            for (int r = 0; r < numShards && ncclOk; ++r) {  // one broadcast per shard, rooted at its owner
                const int root = ShardOwner(r, world);
                // the optional host passes (a17 / a18) read and rewrite every block on every rank: then the state words travel as well
                const size_t count = (size_t)(bounds[r + 1].word - bounds[r].word);
                uint32_t* seg = stateWords + bounds[r].word;
                if (hostPasses && count) ncclOk = nccl.Broadcast(seg, seg, count, ncclUint32, root, comm, stream) == ncclSuccess;
                const size_t nItems = bounds[r + 1].item - bounds[r].item;
                if (ncclOk && nItems) {
                    ncclOk = nccl.Broadcast(digest + bounds[r].item, digest + bounds[r].item, nItems, ncclUint64, root, comm, stream) == ncclSuccess;
                    ncclOk = ncclOk && nccl.Broadcast(special + bounds[r].item, special + bounds[r].item, nItems, ncclInt32, root, comm, stream) == ncclSuccess;
                }
            }
            ncclOk = (nccl.GroupEnd() == ncclSuccess) && ncclOk;
            if (HostTrace::Enabled()) {
                for (int i = 0; i < 3; ++i) CUDA_TRY(PoolEventCreate(&gatherEv[i]));
                CUDA_TRY(cudaEventRecord(gatherEv[0], stream));
            }
            if (!ncclOk) {
                log.Log(ommMessageSeverity_Fatal, "[omm-b200] NCCL all-gather of the per-item records failed");
                rc = ommResult_FAILURE;
                goto cleanup;
            }
        }
    }
    CUDA_TRY(cudaEventRecord(ev[2], stream));
    HostTrace::Mark("classify + post launched");

    // ---- K6, K7: exact dedup, descriptor slots and byte offsets (replicated on every rank of a sharded bake) ----
    if (W > 0) {
        NvtxRange nvtxMerge("dedup + offsets");
        const uint32_t gridW = (W + TPB - 1) / TPB, gridW1 = (W + 1 + TPB - 1) / TPB;
        const uint64_t cap = NextPow2((uint64_t)W * 2 + 16);
        if (streaming) {
            // Every chunk has been merged already.  Left to do on the device: index histogram, descriptors, index buffer -- launched now
            // so that the GPU never waits for the host; then the host forwards the chunks to the copy engine as they complete.
            TriangleFinalItems<<<gridT, TPB, 0, stream>>>(triItem, survivor, mergeRoot, survivor2, items, special, T, triFinal, hist);
            CUDA_TRY(cudaMallocAsync(&res->devIndexBuffer, (size_t)(T ? T : 1) * 4, stream));
            CUDA_TRY(cudaMallocAsync(&res->devDescArray, (size_t)W * sizeof(ommCpuOpacityMicromapDesc), stream));
            WriteDescs<<<(W + TPB - 1) / TPB, TPB, 0, stream>>>(items, special, descOfItem, offsetOfItem, 0, W, (ommCpuOpacityMicromapDesc*)res->devDescArray);
            WriteIndexBuffer<<<gridT, TPB, 0, stream>>>(triFinal, special, descOfItem, T, (int)d.unresolvedTriState, indexBytes, res->devIndexBuffer);
            launches += 3;
            CUDA_TRY(PoolStreamCreate(&copyStream));
            CUDA_TRY(PoolEventCreate(&copyEv[0]));
            CUDA_TRY(PoolEventCreate(&copyEv[1]));
            {
                unsigned long long sent = 0;
                for (size_t c = 0; c < chunkEvs.size(); ++c) {
                    CUDA_TRY(cudaEventSynchronize(chunkEvs[c]));
                    if (c == 0) CUDA_TRY(cudaEventRecord(copyEv[0], copyStream));
                    const unsigned long long end = chunkEndHost[c];
                    if (end > sent)
                        CUDA_TRY(cudaMemcpyAsync((uint8_t*)res->hostArrayData + sent, (const uint8_t*)res->devArrayData + sent, (size_t)(end - sent), cudaMemcpyDeviceToHost, copyStream));
                    sent = end;
                }
                CUDA_TRY(cudaEventRecord(copyEv[1], copyStream));
            }
            HostTrace::Mark("    streamed: last chunk handed to the copy engine");
            CUDA_TRY(cudaMemcpyAsync(histHost, hist, sizeof(histHost), cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaMemcpyAsync(streamHost, runDesc, sizeof(streamHost), cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaStreamSynchronize(stream));
            HostTrace::Mark("    histograms read (host sync 2, streamed)");
            if (streamHost[1] != 0 && HostTrace::Enabled())
                fprintf(stderr, "[omm-b200 trace] streamed bake falls back: the block of item %u (already sent) equals that of item %u in a later chunk, which the SDK keeps\n",
                        streamHost[2], streamHost[3]);
            if (streamHost[1] != 0) {
                // a later chunk held the true survivor of a digest an earlier chunk had already emitted: merge and pack again, the ordinary way
                streamFallback = true;
                CUDA_TRY(cudaStreamSynchronize(copyStream));
                PoolStreamRelease(copyStream);
                copyStream = nullptr;
                for (int i = 0; i < 2; ++i) {
                    PoolEventRelease(copyEv[i]);
                    copyEv[i] = nullptr;
                }
                CUDA_TRY(cudaMemsetAsync(hist, 0, 64 * sizeof(uint32_t), stream));
                memset(histHost, 0, sizeof(histHost));
            }
        }
        if (!streaming && !disableDup) {
            // the UV table (capacity >= 2T+16 >= 2W+16) is reused for the digests
            FillTable<<<(uint32_t)((cap + TPB - 1) / TPB), TPB, 0, stream>>>(tableKeys, tableVals, cap);
            DigestInsert<<<gridW, TPB, 0, stream>>>(digest, items, 0, W, tableKeys, tableVals, cap - 1);
            launches += 2;
        }
        if (!streaming || streamFallback) {
            DigestResolve<<<gridW, TPB, 0, stream>>>(digest, items, triItem, 0, W, tableKeys, tableVals, cap - 1, disableDup, survivor, special);
            launches++;
        }
        if (hostPasses) {
            // ---- a17 / a18: near-duplicate merge and size-budget compression (omm_post_passes.cuh).  The states stay on the device; the
            // host drives the passes from per-item records (32 B each), hashes, distances and counts.  Both passes walk the work items in
            // the SDK's first-seen order, so the host sees them through posOfOrig. ----
            const bool nearDup = !disableDup && (flags & ommCpuBakeFlags_EnableNearDuplicateDetection) != 0, bruteForce = (flags & (1u << 10)) != 0;
            uint32_t* primCount = nullptr;
            uint8_t* levelsDev = nullptr;
            CUDA_TRY(scratch.alloc(&primCount, W));
            CUDA_TRY(scratch.alloc(&mergeRoot, W));
            CUDA_TRY(scratch.alloc(&survivor2, W));
            CUDA_TRY(scratch.alloc(&levelsDev, W));
            CUDA_TRY(cudaMemsetAsync(primCount, 0, sizeof(uint32_t) * W, stream));
            CountPrimitives<<<gridT, TPB, 0, stream>>>(triItem, survivor, T, primCount);
            launches++;
            std::vector<ItemRec> hItems(W);
            std::vector<int32_t> hSpecial(W);
            std::vector<uint32_t> hPrims(W), hRoot(W), hPos(W);
            std::vector<uint8_t> hLevels(W);
            std::vector<PassItem> pass(W);
            CUDA_TRY(cudaMemcpyAsync(hItems.data(), items, sizeof(ItemRec) * W, cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaMemcpyAsync(hSpecial.data(), special, sizeof(int32_t) * W, cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaMemcpyAsync(hPrims.data(), primCount, sizeof(uint32_t) * W, cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaMemcpyAsync(hPos.data(), posOfOrig, sizeof(uint32_t) * W, cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaStreamSynchronize(stream));
            for (uint32_t w = 0; w < W; ++w) {
                const ItemRec& it = hItems[hPos[w]];
                PassItem& p = pass[w];
                p.pos = hPos[w];
                p.prims = hPrims[p.pos];
                p.special = hSpecial[p.pos];
                p.root = w;
                p.level = p.levelAtStart = it.level;
                p.format = it.format;
                // ref: bake_cpu_impl.cpp:464-468 (area of the UV triangle: half the length of the cross product)
                const float v0x = it.p2.x - it.p0.x, v0y = it.p2.y - it.p0.y, v1x = it.p1.x - it.p0.x, v1y = it.p1.y - it.p0.y;
                const float cz = v0x * v1y - v1x * v0y;
                p.area = 0.5f * std::sqrt(cz * cz);
            }
            PostPasses passes(stream, pass, stateWords, wordStart, totalWords);
            if (nearDup && !(bruteForce ? passes.nearDuplicatesWindowed() : passes.nearDuplicatesLsh(d.nearDuplicateDeduplicationFactor, 3))) {
                log.Log(ommMessageSeverity_Fatal, "[omm-b200] near-duplicate pass failed (CUDA error or out of memory)");
                rc = ommResult_FAILURE;
                goto cleanup;
            }
            // ref: bake_cpu_impl.cpp:1965 -- promotion between the merge and the budget pass: blocks the merge made uniform get their special
            // index now (and are then no candidates for downsampling)
            for (uint32_t w = 0; w < W; ++w) hSpecial[pass[w].pos] = pass[w].special;
            CUDA_TRY(cudaMemcpyAsync(special, hSpecial.data(), sizeof(int32_t) * W, cudaMemcpyHostToDevice, stream));
            CUDA_TRY(LaunchItemPost(stream, items, wordStart, stateWords, 0, W, d.rejectionThreshold, (flags & ommCpuBakeFlags_DisableSpecialIndices) != 0, 1, nullptr, (uint32_t)P.stateGT,
                                    (uint32_t)P.stateLE, digest, special, bigItems, &launches));
            if (d.maxArrayDataSize != 0xFFFFFFFFu) {
                CUDA_TRY(cudaMemcpyAsync(hSpecial.data(), special, sizeof(int32_t) * W, cudaMemcpyDeviceToHost, stream));
                CUDA_TRY(cudaStreamSynchronize(stream));
                for (uint32_t w = 0; w < W; ++w) pass[w].special = hSpecial[pass[w].pos];
                rc = passes.compress(d.maxArrayDataSize);
                if (rc != ommResult_SUCCESS) goto cleanup;
            }
            launches += passes.launches;
            for (uint32_t w = 0; w < W; ++w) {
                uint32_t r = w;
                while (pass[r].root != r) r = pass[r].root;
                const uint32_t s = pass[w].pos;
                hRoot[s] = pass[r].pos;
                hLevels[s] = pass[w].level;
                resort = resort || pass[w].level != pass[w].levelAtStart;
            }
            CUDA_TRY(cudaMemcpyAsync(mergeRoot, hRoot.data(), sizeof(uint32_t) * W, cudaMemcpyHostToDevice, stream));
            CUDA_TRY(cudaMemcpyAsync(levelsDev, hLevels.data(), W, cudaMemcpyHostToDevice, stream));
            UpdateItemLevels<<<gridW, TPB, 0, stream>>>(items, levelsDev, W);
            // ref: bake_cpu_impl.cpp:1969-1971 -- second exact dedup over ALL items with the current states, then the last promotion
            // (computed first here; the dedup overwrites duplicates with -1 exactly as the serial order does)
            CUDA_TRY(LaunchItemPost(stream, items, wordStart, stateWords, 0, W, d.rejectionThreshold, (flags & ommCpuBakeFlags_DisableSpecialIndices) != 0, 1, nullptr, (uint32_t)P.stateGT,
                                    (uint32_t)P.stateLE, digest, special, bigItems, &launches));
            launches++;
            if (!disableDup) {
                FillTable<<<(uint32_t)((cap + TPB - 1) / TPB), TPB, 0, stream>>>(tableKeys, tableVals, cap);
                DigestInsert<<<gridW, TPB, 0, stream>>>(digest, items, 0, W, tableKeys, tableVals, cap - 1);
                launches += 2;
            }
            DigestResolve<<<gridW, TPB, 0, stream>>>(digest, items, triItem, 0, W, tableKeys, tableVals, cap - 1, disableDup, survivor2, special);
            launches++;
            CUDA_TRY(cudaStreamSynchronize(stream));  // the host vectors above must outlive the copies
        }
        if (!streaming || streamFallback) {
            // serialized items, their descriptor slots and byte offsets: prefix sums in output order
            EmitInfo<<<gridW1, TPB, 0, stream>>>(items, special, 0, W, 1, (int)d.format, hist, emit, blockBytes);
            TriangleFinalItems<<<gridT, TPB, 0, stream>>>(triItem, survivor, mergeRoot, survivor2, items, special, T, triFinal, hist);
            {
                size_t tmp = cubTempBytes;
                CUDA_TRY(cub::DeviceScan::ExclusiveSum(cubTemp, tmp, emit, descOfItem, (int)W + 1, stream));
                tmp = cubTempBytes;
                CUDA_TRY(cub::DeviceScan::ExclusiveSum(cubTemp, tmp, blockBytes, offsetOfItem, (int)W + 1, stream));
            }
            ShardByteOffsets<<<1, 96, 0, stream>>>(offsetOfItem, boundsDev, numShards, shardOffDev);
            launches += 7;
            CUDA_TRY(cudaMemcpyAsync(histHost, hist, sizeof(histHost), cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaMemcpyAsync(shardOff, shardOffDev, sizeof(unsigned long long) * (numShards + 1), cudaMemcpyDeviceToHost, stream));
            if (world > 1) CUDA_TRY(cudaMemcpyAsync(&myMicroTris, workloadDev, 8, cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaStreamSynchronize(stream));
            HostTrace::Mark("    histograms read (host sync 2)");
        }
    }

    // ---- sizes (ref: bake_cpu_impl.cpp:1763-1777) ----
    {
        const uint32_t bitCount = (uint32_t)d.format;
        uint32_t globalFormatDescs = 0;
        unsigned long long globalFormatBytes = 0;
        for (int l = 0; l <= kMaxLevel; ++l) {
            const uint32_t cnt = histHost[(d.format - 1) * 13 + l];
            globalFormatDescs += cnt;
            const unsigned long long bytes = ((1ull << (2 * l)) * bitCount) >> 3;
            globalFormatBytes += (unsigned long long)cnt * (bytes > 1 ? bytes : 1);
        }
        for (int i = 0; i < 26; ++i) numDescs += histHost[i];
        if (numDescs != globalFormatDescs) {
            // The SDK sizes its arrays from the global format only and then walks every item (undefined behaviour with mixed
            // per-triangle formats, SURVEY 7 "reference quirks").  Refused at staging already; kept as a guard.
            log.Log(ommMessageSeverity_Fatal, "[omm-b200] per-triangle formats that differ from desc.format are not supported");
            rc = ommResult_FAILURE;
            goto cleanup;
        }
        arrayBytes = globalFormatBytes;
        if (arrayBytes > 0xFFFFFFFFull) {  // ref: :1774-1775
            rc = ommResult_FAILURE;
            goto cleanup;
        }
    }

    // ---- K8: serialize ----
    {
        NvtxRange nvtxPack("pack + index buffer");
        res->device = baker->device;
        res->arrayDataSize = (uint32_t)arrayBytes;
        res->descCount = numDescs;
        res->indexCount = T;
        res->indexFormat = ifmt;
        if (!res->devIndexBuffer) CUDA_TRY(cudaMallocAsync(&res->devIndexBuffer, (size_t)(T ? T : 1) * 4, stream));
        if (numDescs) {
            const uint32_t packGridMax = (uint32_t)std::max(sms, 1) * 32u;
            if (!res->devArrayData) CUDA_TRY(cudaMallocAsync(&res->devArrayData, (size_t)arrayBytes + 16, stream));  // (a streamed bake has its worst-case buffer)
            if (!res->devDescArray) CUDA_TRY(cudaMallocAsync(&res->devDescArray, (size_t)numDescs * sizeof(ommCpuOpacityMicromapDesc), stream));
            if (resort) {
                // Compress lowered the level of some items and with it their sort keys: order the serialized items again
                uint32_t* resortVals = nullptr;
                unsigned long long* blockOffset = nullptr;
                CUDA_TRY(scratch.alloc(&resortVals, W));
                CUDA_TRY(scratch.alloc(&blockOffset, (size_t)numDescs + 1));
                ResortKeys<<<(W + TPB - 1) / TPB, TPB, 0, stream>>>(items, special, origOf, W, sortKeysIn, sortValsIn);
                size_t tmp = cubTempBytes;
                CUDA_TRY(cub::DeviceRadixSort::SortPairsDescending(cubTemp, tmp, sortKeysIn, sortKeysOut, sortValsIn, resortVals, (int)W, 0, 31, stream));
                ResortBlockSizes<<<(numDescs + 1 + TPB - 1) / TPB, TPB, 0, stream>>>(resortVals, items, numDescs, (int)d.format, blockBytes);
                tmp = cubTempBytes;
                CUDA_TRY(cub::DeviceScan::ExclusiveSum(cubTemp, tmp, blockBytes, blockOffset, (int)numDescs + 1, stream));
                ResortScatter<<<(numDescs + TPB - 1) / TPB, TPB, 0, stream>>>(resortVals, blockOffset, numDescs, descOfItem, offsetOfItem);
                launches += 13;
            }
            if (!(streaming && !streamFallback)) {
                WriteDescs<<<(W + TPB - 1) / TPB, TPB, 0, stream>>>(items, special, descOfItem, offsetOfItem, 0, W, (ommCpuOpacityMicromapDesc*)res->devDescArray);
                launches++;
            }
            if (streaming && !streamFallback) {
                // nothing left to pack or send: the last chunk's blocks are with the copy engine
            } else if (world > 1 && !hostPasses) {
                // ---- the exchange, part 2.  The survivors of a shard (a contiguous run of positions of the output order) occupy ONE
                // contiguous byte range of arrayData: every rank packs its own shards straight to their final place, and one group of
                // in-place broadcasts rooted at the shard owners completes the array on every rank.  No staging buffer, no host
                // synchronisation between packing and sending (the byte ranges came with the histograms above). ----
                NcclApi& nccl = Nccl();
                const ncclComm_t comm = (ncclComm_t)baker->shard.ncclComm;
                // ommCpuBake: the host copy of the array is assembled in a page-locked window shared by the ranks -- every rank sends
                // the blocks of its own shards there with its own copy engine over its own PCIe link, while NVLink completes the
                // device copies.  (One GPU pulling the whole array through one link was 5.3 of the 10.6 ms of an 8-GPU call.)
                if (earlyDownload && baker->shard.ctl && arrayBytes >= ((size_t)1 << 20)) {
                    windowProtocol = true;
                    baker->shard.bakeSeq++;
                    window = AcquireSharedWindow(baker->shard, (size_t)arrayBytes, log);
                    if (!window && baker->shard.ctl->winId >= 0) baker->shard.ctl->failSeq.store(baker->shard.bakeSeq, std::memory_order_release);
                    if (baker->shard.ctl->winId < 0) windowProtocol = false;  // the root has no window: every rank sees that
                    if (window) {
                        CUDA_TRY(PoolStreamCreate(&copyStream));
                        CUDA_TRY(PoolEventCreate(&copyEv[0]));
                        CUDA_TRY(PoolEventCreate(&copyEv[1]));
                        CUDA_TRY(PoolEventCreate(&sliceEv, false));
                        CUDA_TRY(cudaEventRecord(copyEv[0], copyStream));
                    }
                }
                for (int k = 0; k < owned.count; ++k) {
                    const uint32_t itemBegin = bounds[owned.shard[k]].item, itemEnd = bounds[owned.shard[k] + 1].item;
                    const unsigned long long b0 = shardOff[owned.shard[k]], b1 = shardOff[owned.shard[k] + 1];
                    if (itemEnd <= itemBegin || b1 == b0) continue;
                    PackItems<<<std::min(packGridMax, (itemEnd - itemBegin + 7) / 8), 256, 0, stream>>>(items, special, wordStart, stateWords, descOfItem, offsetOfItem, nullptr, nullptr,
                                                                                                 itemBegin, itemEnd, arrayBytes, (uint8_t*)res->devArrayData, nullptr);
                    launches++;
                    if (window) {
                        CUDA_TRY(cudaEventRecord(sliceEv, stream));
                        CUDA_TRY(cudaStreamWaitEvent(copyStream, sliceEv, 0));
                        CUDA_TRY(cudaMemcpyAsync((uint8_t*)window->ptr + b0, (const uint8_t*)res->devArrayData + b0, (size_t)(b1 - b0), cudaMemcpyDeviceToHost, copyStream));
                        sharedD2hBytes += b1 - b0;
                    }
                }
                if (window) CUDA_TRY(cudaEventRecord(copyEv[1], copyStream));
                if (gatherEv[1]) CUDA_TRY(cudaEventRecord(gatherEv[1], stream));
                const bool onRank0 = baker->shard.resultMode == 1;
                bool ncclOk = true;
                if (onRank0 && windowProtocol) {
                    // the array is being assembled in host memory by all ranks: nothing crosses NVLink
                    if (rank == 0) res->deviceArrayComplete = false;
                } else {
                    ncclOk = nccl.GroupStart() == ncclSuccess;
                    for (int r = 0; r < numShards && ncclOk; ++r) {
                        const size_t count = (size_t)(shardOff[r + 1] - shardOff[r]);
                        uint8_t* seg = (uint8_t*)res->devArrayData + shardOff[r];
                        const int owner = ShardOwner(r, world);
                        if (!count) continue;
                        if (!onRank0) ncclOk = nccl.Broadcast(seg, seg, count, ncclUint8, owner, comm, stream) == ncclSuccess;
                        else if (owner != 0 && rank == owner) ncclOk = nccl.Send(seg, count, ncclUint8, 0, comm, stream) == ncclSuccess;
                        else if (owner != 0 && rank == 0) ncclOk = nccl.Recv(seg, count, ncclUint8, owner, comm, stream) == ncclSuccess;
                    }
                    ncclOk = (nccl.GroupEnd() == ncclSuccess) && ncclOk;
                }
                if (onRank0 && rank != 0) res->arrayOnThisRank = false;
                if (gatherEv[2]) CUDA_TRY(cudaEventRecord(gatherEv[2], stream));
                if (!ncclOk) {
                    log.Log(ommMessageSeverity_Fatal, "[omm-b200] NCCL exchange of the packed blocks failed");
                    rc = ommResult_FAILURE;
                    goto cleanup;
                }
            } else {
                // The caller wants the result in host memory (ommCpuBake): pack the array in slices of work items and send each slice
                // over PCIe while the next one is packed; the slice boundaries (8 bytes each) are read back first.  Page-locked
                // destination only (a pageable one makes the copies synchronous).
                constexpr uint32_t kPackSlices = 8;
                unsigned long long sliceOffset[kPackSlices + 1];
                uint32_t sliceItem[kPackSlices + 1];
                uint32_t slices = 1;
                sliceItem[0] = 0; sliceOffset[0] = 0;
                if (earlyDownload && world == 1 && arrayBytes >= ((size_t)8 << 20) && W >= 64 * kPackSlices && AllocHostArrayData(res) && res->arrayDataFromPinnedPool) {
                    slices = kPackSlices;
                    for (uint32_t i = 1; i < slices; ++i) {
                        sliceItem[i] = (uint32_t)((unsigned long long)W * i / slices);
                        CUDA_TRY(cudaMemcpyAsync(&sliceOffset[i], offsetOfItem + sliceItem[i], 8, cudaMemcpyDeviceToHost, stream));
                    }
                    CUDA_TRY(cudaStreamSynchronize(stream));
                    CUDA_TRY(PoolStreamCreate(&copyStream));
                    CUDA_TRY(PoolEventCreate(&copyEv[0]));
                    CUDA_TRY(PoolEventCreate(&copyEv[1]));
                    CUDA_TRY(PoolEventCreate(&sliceEv, false));
                    CUDA_TRY(cudaEventRecord(copyEv[0], copyStream));
                }
                sliceItem[slices] = W; sliceOffset[slices] = arrayBytes;
                for (uint32_t i = 0; i < slices; ++i) {
                    const uint32_t i0 = sliceItem[i], i1 = sliceItem[i + 1];
                    if (i1 <= i0) continue;
                    PackItems<<<std::min(packGridMax, (i1 - i0 + 7) / 8), 256, 0, stream>>>(items, special, wordStart, stateWords, descOfItem, offsetOfItem, nullptr, nullptr, i0, i1,
                                                                                     arrayBytes, (uint8_t*)res->devArrayData, nullptr);
                    launches++;
                    if (copyStream && sliceOffset[i + 1] > sliceOffset[i]) {
                        CUDA_TRY(cudaEventRecord(sliceEv, stream));
                        CUDA_TRY(cudaStreamWaitEvent(copyStream, sliceEv, 0));
                        CUDA_TRY(cudaMemcpyAsync((uint8_t*)res->hostArrayData + sliceOffset[i], (const uint8_t*)res->devArrayData + sliceOffset[i],
                                                 (size_t)(sliceOffset[i + 1] - sliceOffset[i]), cudaMemcpyDeviceToHost, copyStream));
                    }
                }
                if (copyStream) CUDA_TRY(cudaEventRecord(copyEv[1], copyStream));
            }
        }
        if (T > 0 && !(streaming && !streamFallback)) {
            // no work item at all (every triangle invalid): the per-triangle item table was never written -> all unresolved
            // (found by the SDK's own LogTest.Validation_InvalidTriangles run against this library)
            if (W == 0) CUDA_TRY(cudaMemsetAsync(triFinal, 0xFF, sizeof(uint32_t) * (size_t)T, stream));
            WriteIndexBuffer<<<gridT, TPB, 0, stream>>>(triFinal, special, descOfItem, T, (int)d.unresolvedTriState, indexBytes, res->devIndexBuffer);
            launches++;
        }
        CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(cudaEventRecord(ev[3], stream));
    HostTrace::Mark("pack launched");

    // histograms (ref: bake_cpu_impl.cpp:1826-1852): non-zero entries, 2-state before 4-state, level ascending
    {
        uint32_t na = 0, ni = 0;
        for (uint32_t f = 1; f <= 2; ++f)
            for (uint32_t l = 0; l <= (uint32_t)kMaxLevel; ++l) {
                const uint32_t ca = histHost[(f - 1) * 13 + l], ci = histHost[26 + (f - 1) * 13 + l];
                if (ca) res->hostArrayHist[na++] = ommCpuOpacityMicromapUsageCount{ca, (uint16_t)l, (uint16_t)f};
                if (ci) res->hostIndexHist[ni++] = ommCpuOpacityMicromapUsageCount{ci, (uint16_t)l, (uint16_t)f};
            }
        res->desc.descArrayHistogramCount = na;
        res->desc.indexHistogramCount = ni;
    }

    scratch.freeAll();
    CUDA_TRY(cudaStreamSynchronize(stream));
    HostTrace::Mark("final sync");
    if (gatherEv[2]) {
        float a = 0.f, b = 0.f, c = 0.f;
        cudaEventElapsedTime(&a, ev[5], gatherEv[0]);
        cudaEventElapsedTime(&b, gatherEv[0], gatherEv[1]);
        cudaEventElapsedTime(&c, gatherEv[1], gatherEv[2]);
        fprintf(stderr, "[omm-b200 trace] rank %d exchange: digests + special indices %.3f ms, merge + pack of own shards %.3f ms, blocks %.3f ms\n", rank, a, b, c);
    }
    if (copyStream) {
        CUDA_TRY(cudaStreamSynchronize(copyStream));
        HostTrace::Mark("array data on the host");
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, copyEv[0], copyEv[1]) != cudaSuccess) cudaGetLastError();
        res->earlyD2hMs = ms;
        if (!windowProtocol) res->arrayDataDownloaded = true;
    }
    if (windowProtocol) {
        // every rank's part of the window is written once all ranks have passed this point
        const bool all = HostBarrier(baker->shard);
        const bool failed = !all || baker->shard.ctl->failSeq.load(std::memory_order_acquire) == baker->shard.bakeSeq;
        HostTrace::Mark("shared window complete (host barrier)");
        if (rank == 0 && window) {
            if (!window->unlinked) {  // every rank has mapped it by now
                char name[96];
                ShmName(name, sizeof(name), baker->shard.idHash, "w", window->id);
                shm_unlink(name);
                window->unlinked = true;
            }
            if (failed) {
                window->inUse = false;  // some rank could not write its part: the result is downloaded the ordinary way on request ...
                if (baker->shard.resultMode == 1) {  // ... which only works when this GPU holds the whole array
                    log.Log(ommMessageSeverity_Fatal, "[omm-b200] a rank could not write its shards to the shared host window");
                    rc = ommResult_FAILURE;
                    goto cleanup;
                }
            } else {
                res->hostArrayData = window->ptr;
                res->sharedWindowId = window->id;
                res->arrayDataDownloaded = true;
            }
        }
        tm->d2hBytes = sharedD2hBytes;
    }
    {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev[0], ev[1]); tm->setupMs = ms;
        if (W > 0) {
            cudaEventElapsedTime(&ms, ev[1], ev[4]); tm->classifyMs = ms;
            cudaEventElapsedTime(&ms, ev[4], ev[5]); tm->itemPostMs = ms;
            cudaEventElapsedTime(&ms, ev[5], ev[2]); tm->gatherMs = ms;
            cudaEventElapsedTime(&ms, ev[4], ev[3]); tm->postMs = ms;
        } else {
            tm->classifyMs = tm->itemPostMs = tm->gatherMs = 0.f;
            cudaEventElapsedTime(&ms, ev[1], ev[3]); tm->postMs = ms;
        }
        cudaEventElapsedTime(&ms, ev[0], ev[3]); tm->totalDeviceMs = ms;
        tm->workItems = W;
        tm->kernelLaunches = launches;
        tm->arrayDataBytes = numDescs ? arrayBytes : 0;
        tm->descCount = numDescs;
        tm->microTriangles = myMicroTris;
        tm->reserved = 0;
    }

cleanup:
    for (int l = 1; l < kMaxChunkLanes; ++l) {
        if (!laneStream[l]) continue;
        // (a complete bake has joined its lanes into the stream the host has synchronised since; a failed one may have left them running)
        if (rc != ommResult_SUCCESS) cudaStreamSynchronize(laneStream[l]);
        PoolStreamRelease(laneStream[l]);
        PoolEventRelease(laneEv[l], false);
    }
    PoolEventRelease(lanePrepEv, false);
    scratch.freeAll();
    for (int i = 0; i < 6; ++i)
        PoolEventRelease(ev[i]);
    if (cellTables) {
        if (rc != ommResult_SUCCESS) cudaStreamSynchronize(stream);  // kernels of a failed bake may still be reading the tables
        ReleaseCellTables(tex, cellTables);
    }
    if (ownStream) {
        cudaStreamSynchronize(stream);
        PoolStreamRelease(stream);
    }
    if (copyStream) {
        cudaStreamSynchronize(copyStream);
        PoolStreamRelease(copyStream);
    }
    if (pinnedReadback) PinnedPoolRelease(pinnedReadback);
    for (cudaEvent_t e : chunkEvs) PoolEventRelease(e, false);
    if (chunkEndHost) PinnedPoolRelease(chunkEndHost);
    for (cudaEvent_t e : {copyEv[0], copyEv[1], gatherEv[0], gatherEv[1], gatherEv[2]}) PoolEventRelease(e);
    PoolEventRelease(sliceEv, false);
    if (rc != ommResult_SUCCESS) {
        cudaGetLastError();
        DestroyResultDevice(res);
    }
    return rc;
}

// descriptors / index buffer of a result: page-locked like arrayData when they are big and the allocator is ours (a copy into pageable memory
// is staged by the driver at a fifth of the PCIe speed: 0.3 ms of a config-3 call)
static void* AllocHostSmallArray(BakeResultObject* res, size_t bytes, bool* fromPool) {
    *fromPool = false;
    if (res->usesDefaultAllocator && bytes >= ((size_t)1 << 20))
        if (void* p = PinnedPoolAcquire(bytes)) {
            *fromPool = true;
            return p;
        }
    return res->alloc.alloc(bytes, 64);
}

ommResult DownloadResult(BakeResultObject* res, float* d2hMs, uint64_t* d2hBytes) {
    const Logger& log = res->log;
    ommResult rc = ommResult_SUCCESS;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (res->downloaded) return ommResult_SUCCESS;
    const size_t idxBytes = (size_t)res->indexCount * (res->indexFormat == ommIndexFormat_UINT_32 ? 4 : (res->indexFormat == ommIndexFormat_UINT_16 ? 2 : 1));
    CUDA_TRY(cudaSetDevice(res->device));
    CUDA_TRY(PoolEventCreate(&e0));
    CUDA_TRY(PoolEventCreate(&e1));
    CUDA_TRY(cudaEventRecord(e0, 0));
    if (res->descCount) {
        res->hostDescArray = AllocHostSmallArray(res, (size_t)res->descCount * sizeof(ommCpuOpacityMicromapDesc), &res->descFromPinnedPool);
        if (!AllocHostArrayData(res) || !res->hostDescArray) { rc = ommResult_FAILURE; goto cleanup; }
        HostTrace::Mark("host result allocated");
        if (!res->arrayDataDownloaded)  // else: sent slice by slice while it was packed (BakeOnDevice, K8)
            CUDA_TRY(cudaMemcpyAsync(res->hostArrayData, res->devArrayData, res->arrayDataSize, cudaMemcpyDeviceToHost, 0));
        CUDA_TRY(cudaMemcpyAsync(res->hostDescArray, res->devDescArray, (size_t)res->descCount * sizeof(ommCpuOpacityMicromapDesc), cudaMemcpyDeviceToHost, 0));
    }
    res->hostIndexBuffer = AllocHostSmallArray(res, (size_t)res->indexCount * 4, &res->indexFromPinnedPool);
    if (!res->hostIndexBuffer) { rc = ommResult_FAILURE; goto cleanup; }
    CUDA_TRY(cudaMemcpyAsync(res->hostIndexBuffer, res->devIndexBuffer, idxBytes, cudaMemcpyDeviceToHost, 0));
    CUDA_TRY(cudaEventRecord(e1, 0));
    CUDA_TRY(cudaEventSynchronize(e1));
    HostTrace::Mark("D2H done");
    if (d2hMs) {
        CUDA_TRY(cudaEventElapsedTime(d2hMs, e0, e1));
        *d2hMs += res->earlyD2hMs;  // the overlapped part, as timed on the copy stream
    }
    if (d2hBytes) *d2hBytes = (uint64_t)res->arrayDataSize + (uint64_t)res->descCount * 8 + idxBytes;
    res->desc.arrayData = res->hostArrayData;
    res->desc.arrayDataSize = res->descCount ? res->arrayDataSize : 0;
    res->desc.descArray = (const ommCpuOpacityMicromapDesc*)res->hostDescArray;
    res->desc.descArrayCount = res->descCount;
    res->desc.descArrayHistogram = res->hostArrayHist;
    res->desc.indexBuffer = res->hostIndexBuffer;
    res->desc.indexCount = res->indexCount;
    res->desc.indexFormat = res->indexFormat;
    res->desc.indexHistogram = res->hostIndexHist;
    res->downloaded = true;
cleanup:
    PoolEventRelease(e0);
    PoolEventRelease(e1);
    return rc;
}

void DestroyResultDevice(BakeResultObject* res) {
    if (res->devArrayData || res->devDescArray || res->devIndexBuffer) cudaSetDevice(res->device);
    if (res->devArrayData) cudaFreeAsync(res->devArrayData, 0);
    if (res->devDescArray) cudaFreeAsync(res->devDescArray, 0);
    if (res->devIndexBuffer) cudaFreeAsync(res->devIndexBuffer, 0);
    res->devArrayData = res->devDescArray = res->devIndexBuffer = nullptr;
}

// ---------------------------------------------------------------------------------------------------------------------
// multi-GPU plumbing
// ---------------------------------------------------------------------------------------------------------------------
ommResult GetNcclUniqueId(void* out, size_t size) {
    if (!out || size < sizeof(ncclUniqueId)) return ommResult_INVALID_ARGUMENT;
    NcclApi& nccl = Nccl();
    if (!nccl.ok) return ommResult_FAILURE;
    ncclUniqueId id;
    if (nccl.GetUniqueId(&id) != ncclSuccess) return ommResult_FAILURE;
    memcpy(out, &id, sizeof(id));
    return ommResult_SUCCESS;
}
ommResult InitSharding(BakerObject* baker, int rank, int world, const void* idBytes, size_t idSize) {
    if (world < 1 || world > 64 || rank < 0 || rank >= world) return ommResult_INVALID_ARGUMENT;
    DestroySharding(baker);
    if (world == 1) return ommResult_SUCCESS;
    if (!idBytes || idSize < sizeof(ncclUniqueId)) return ommResult_INVALID_ARGUMENT;
    if (RequireDevice(baker->log, baker->device) != ommResult_SUCCESS) return ommResult_FAILURE;
    NcclApi& nccl = Nccl();
    if (!nccl.ok) {
        baker->log.Log(ommMessageSeverity_Fatal, "[omm-b200] libnccl.so.2 could not be loaded");
        return ommResult_FAILURE;
    }
    ncclUniqueId id;
    memcpy(&id, idBytes, sizeof(id));
    ncclComm_t comm = nullptr;
    if (nccl.CommInitRank(&comm, world, id, rank) != ncclSuccess) return ommResult_FAILURE;
    baker->shard.rank = rank;
    baker->shard.world = world;
    baker->shard.ncclComm = comm;
    // control block in shared memory (one box: the ranks are processes of one node).  Its absence only disables the parallel host
    // download of sharded ommCpuBake results.
    {
        unsigned long long h = 1469598103934665603ull;  // FNV-1a of the unique id
        for (size_t i = 0; i < sizeof(id); ++i) h = (h ^ (unsigned char)id.internal[i]) * 1099511628211ull;
        baker->shard.idHash = h;
        char name[96];
        ShmName(name, sizeof(name), h, "ctl", 0);
        baker->shard.ctl = (ShmControl*)MapShm(name, sizeof(ShmControl), true);  // every rank creates-or-opens; fresh objects are zero-filled
        baker->shard.bakeSeq = 0;
        if (baker->shard.ctl) {
            if (!HostBarrier(baker->shard, 60.0)) {  // everybody has mapped it: the name can go
                munmap(baker->shard.ctl, sizeof(ShmControl));
                baker->shard.ctl = nullptr;
            }
            if (rank == 0) shm_unlink(name);
        }
    }
    return ommResult_SUCCESS;
}
void DestroySharding(BakerObject* baker) {
    for (SharedHostWindow& w : baker->shard.windows) {
        if (!w.ptr) continue;
        cudaHostUnregister(w.ptr);
        munmap(w.ptr, w.capacity);
        if (baker->shard.rank == 0 && !w.unlinked) {
            char name[96];
            ShmName(name, sizeof(name), baker->shard.idHash, "w", w.id);
            shm_unlink(name);
        }
    }
    cudaGetLastError();
    baker->shard.windows.clear();
    if (baker->shard.ctl) munmap(baker->shard.ctl, sizeof(ShmControl));
    baker->shard.ctl = nullptr;
    if (baker->shard.ncclComm) {
        Nccl().CommDestroy((ncclComm_t)baker->shard.ncclComm);
        baker->shard.ncclComm = nullptr;
    }
    baker->shard.rank = 0;
    baker->shard.world = 1;
}

}  // namespace ommb200
