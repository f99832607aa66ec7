/*
 * omm_b200.h -- C ABI of libomm-b200.so, the B200-native Opacity Micro-Map baker.
 *
 * The first half of this header re-declares, layout-for-layout and value-for-value, the
 * subset of the Opacity Micro-Map SDK 1.9.0 C API that the CPU bake path uses, so that a
 * program compiled against the SDK's own omm.h can be linked against (or dlopen) this
 * library instead and get byte-identical results.  Each declaration cites the line of the
 * reference header it has to stay ABI-compatible with ("ref:" = /root/reference/
 * libraries/omm-lib/include/omm.h).  Do not include this header together with the SDK's
 * omm.h in one translation unit: the names are deliberately identical.
 *
 * The second half (ommB200*) is an extension for callers that want to keep inputs and/or
 * outputs resident in HBM, pick the CUDA device/stream, shard a bake over several GPUs, or
 * read the device-side timings that bench.py reports.
 */
#ifndef OMM_B200_H_
#define OMM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
#define OMMB200_EXTERN extern "C"
#else
#define OMMB200_EXTERN
#endif
#if defined(OMMB200_BUILDING_LIBRARY)
#define OMM_API OMMB200_EXTERN __attribute__((visibility("default")))
#elif !defined(OMM_API)
#define OMM_API OMMB200_EXTERN
#endif

/* ref: omm.h:17-19 -- version reported by ommGetLibraryDesc (test_basic.cpp:19-26 pins it). */
#define OMM_VERSION_MAJOR 1
#define OMM_VERSION_MINOR 9
#define OMM_VERSION_BUILD 0

/* ---- opaque handles (ref: omm.h:56-71) ------------------------------------------------ */
typedef uint8_t ommBool;
typedef struct _ommBaker* ommBaker;
typedef struct _ommCpuBakeResult* ommCpuBakeResult;
typedef struct _ommCpuTexture* ommCpuTexture;
typedef struct _ommCpuSerializedResult* ommCpuSerializedResult;
typedef struct _ommCpuDeserializedResult* ommCpuDeserializedResult;

/* ---- enums; all are 4-byte ints in the ABI (ref: omm.h:78-192, 282-334) ---------------- */
typedef enum ommResult {
    ommResult_SUCCESS = 0,
    ommResult_FAILURE = 1,
    ommResult_INVALID_ARGUMENT = 2,
    ommResult_INSUFFICIENT_SCRATCH_MEMORY = 3,
    ommResult_NOT_IMPLEMENTED = 4,
    ommResult_WORKLOAD_TOO_BIG = 5,
    ommResult_MAX_NUM = 6
} ommResult;

typedef enum ommMessageSeverity {
    ommMessageSeverity_Info = 0,
    ommMessageSeverity_PerfWarning = 1,
    ommMessageSeverity_Error = 2,
    ommMessageSeverity_Fatal = 3,
    ommMessageSeverity_MAX_NUM = 4
} ommMessageSeverity;

typedef enum ommOpacityState {
    ommOpacityState_Transparent = 0,
    ommOpacityState_Opaque = 1,
    ommOpacityState_UnknownTransparent = 2,
    ommOpacityState_UnknownOpaque = 3
} ommOpacityState;

typedef enum ommSpecialIndex {
    ommSpecialIndex_FullyTransparent = -1,
    ommSpecialIndex_FullyOpaque = -2,
    ommSpecialIndex_FullyUnknownTransparent = -3,
    ommSpecialIndex_FullyUnknownOpaque = -4
} ommSpecialIndex;

typedef enum ommFormat {
    ommFormat_INVALID = 0,
    ommFormat_OC1_2_State = 1, /* 1 bit / micro-triangle  */
    ommFormat_OC1_4_State = 2, /* 2 bits / micro-triangle */
    ommFormat_MAX_NUM = 3
} ommFormat;

typedef enum ommUnknownStatePromotion {
    ommUnknownStatePromotion_Nearest = 0,
    ommUnknownStatePromotion_ForceOpaque = 1,
    ommUnknownStatePromotion_ForceTransparent = 2,
    ommUnknownStatePromotion_MAX_NUM = 3
} ommUnknownStatePromotion;

typedef enum ommBakerType { ommBakerType_GPU = 0, ommBakerType_CPU = 1, ommBakerType_MAX_NUM = 2 } ommBakerType;

typedef enum ommTexCoordFormat {
    ommTexCoordFormat_UV16_UNORM = 0,
    ommTexCoordFormat_UV16_FLOAT = 1,
    ommTexCoordFormat_UV32_FLOAT = 2,
    ommTexCoordFormat_MAX_NUM = 3
} ommTexCoordFormat;

typedef enum ommIndexFormat {
    ommIndexFormat_UINT_16 = 0,
    ommIndexFormat_UINT_32 = 1,
    ommIndexFormat_UINT_8 = 2,
    ommIndexFormat_MAX_NUM = 3
} ommIndexFormat;

typedef enum ommTextureAddressMode {
    ommTextureAddressMode_Wrap = 0,
    ommTextureAddressMode_Mirror = 1,
    ommTextureAddressMode_Clamp = 2,
    ommTextureAddressMode_Border = 3,
    ommTextureAddressMode_MirrorOnce = 4,
    ommTextureAddressMode_MAX_NUM = 5
} ommTextureAddressMode;

typedef enum ommTextureFilterMode {
    ommTextureFilterMode_Nearest = 0,
    ommTextureFilterMode_Linear = 1,
    ommTextureFilterMode_MAX_NUM = 2
} ommTextureFilterMode;

typedef enum ommAlphaMode { ommAlphaMode_Test = 0, ommAlphaMode_Blend = 1, ommAlphaMode_MAX_NUM = 2 } ommAlphaMode;

typedef enum ommCpuSerializeFlags { ommCpuSerializeFlags_None = 0, ommCpuSerializeFlags_Compress = 1 } ommCpuSerializeFlags;

typedef enum ommCpuTextureFormat {
    ommCpuTextureFormat_UNORM8 = 0,
    ommCpuTextureFormat_FP32 = 1,
    ommCpuTextureFormat_MAX_NUM = 2
} ommCpuTextureFormat;

typedef enum ommCpuTextureFlags {
    ommCpuTextureFlags_None = 0,
    ommCpuTextureFlags_DisableZOrder = 1u << 0 /* host-layout hint in the SDK; a no-op here (HBM layout is ours) */
} ommCpuTextureFlags;

typedef enum ommCpuBakeFlags {
    ommCpuBakeFlags_None = 0,
    ommCpuBakeFlags_EnableInternalThreads = 1u << 0, /* accepted, ignored: the GPU grid is the parallelism */
    ommCpuBakeFlags_DisableSpecialIndices = 1u << 1,
    ommCpuBakeFlags_Force32BitIndices = 1u << 2,
    ommCpuBakeFlags_DisableDuplicateDetection = 1u << 3,
    ommCpuBakeFlags_EnableNearDuplicateDetection = 1u << 4,
    ommCpuBakeFlags_EnableValidation = 1u << 5,
    ommCpuBakeFlags_Allow8BitIndices = 1u << 6
} ommCpuBakeFlags;

/* ---- creation descs (ref: omm.h:194-274) ---------------------------------------------- */
typedef struct ommLibraryDesc {
    uint8_t versionMajor;
    uint8_t versionMinor;
    uint8_t versionBuild;
} ommLibraryDesc;

typedef struct ommSamplerDesc {
    ommTextureAddressMode addressingMode;
    ommTextureFilterMode filter;
    float borderAlpha;
} ommSamplerDesc;

typedef void* (*ommAllocate)(void* userArg, size_t size, size_t alignment);
typedef void* (*ommReallocate)(void* userArg, void* memory, size_t size, size_t alignment);
typedef void (*ommFree)(void* userArg, void* memory);

typedef struct ommMemoryAllocatorInterface {
    ommAllocate allocate;
    ommReallocate reallocate;
    ommFree free;
    void* userArg;
} ommMemoryAllocatorInterface;

typedef void (*ommMessageCallback)(ommMessageSeverity severity, const char* message, void* userArg);

typedef struct ommMessageInterface {
    ommMessageCallback messageCallback;
    void* userArg;
} ommMessageInterface;

typedef struct ommBakerCreationDesc {
    ommBakerType type;
    ommMemoryAllocatorInterface memoryAllocatorInterface;
    ommMessageInterface messageInterface;
} ommBakerCreationDesc;

/* ---- texture + bake descs (ref: omm.h:339-530) ---------------------------------------- */
typedef struct ommCpuTextureMipDesc {
    uint32_t width;
    uint32_t height;
    uint32_t rowPitch; /* 0 = tightly packed */
    const void* textureData;
} ommCpuTextureMipDesc;

typedef struct ommCpuTextureDesc {
    ommCpuTextureFormat format;
    ommCpuTextureFlags flags;
    const ommCpuTextureMipDesc* mips;
    uint32_t mipCount;
    float alphaCutoff; /* >= 0 embeds the cutoff and enables the summed-area coarse pass */
} ommCpuTextureDesc;

/* sizeof == 136 (ref: serialize_impl.cpp:86) */
typedef struct ommCpuBakeInputDesc {
    ommCpuBakeFlags bakeFlags;
    ommCpuTexture texture;
    ommSamplerDesc runtimeSamplerDesc;
    ommAlphaMode alphaMode;
    ommTexCoordFormat texCoordFormat;
    const void* texCoords;
    uint32_t texCoordStrideInBytes; /* 0 = packed */
    ommIndexFormat indexFormat;
    const void* indexBuffer;
    uint32_t indexCount;
    float dynamicSubdivisionScale; /* <= 0 disables the per-triangle level heuristic */
    float rejectionThreshold;
    float alphaCutoff;
    float nearDuplicateDeduplicationFactor;
    ommOpacityState alphaCutoffLessEqual;
    ommOpacityState alphaCutoffGreater;
    ommFormat format;
    const ommFormat* formats;
    ommUnknownStatePromotion unknownStatePromotion;
    ommSpecialIndex unresolvedTriState;
    uint8_t maxSubdivisionLevel; /* [0,12] */
    uint32_t maxArrayDataSize;   /* 0xFFFFFFFF = unlimited */
    const uint8_t* subdivisionLevels;
    uint64_t maxWorkloadSize;    /* 0xFFFF...F = unlimited */
} ommCpuBakeInputDesc;

typedef struct ommCpuOpacityMicromapDesc {
    uint32_t offset; /* byte offset into arrayData */
    uint16_t subdivisionLevel;
    uint16_t format;
} ommCpuOpacityMicromapDesc;

typedef struct ommCpuOpacityMicromapUsageCount {
    uint32_t count;
    uint16_t subdivisionLevel;
    uint16_t format;
} ommCpuOpacityMicromapUsageCount;

typedef struct ommCpuBakeResultDesc {
    const void* arrayData;
    uint32_t arrayDataSize;
    const ommCpuOpacityMicromapDesc* descArray;
    uint32_t descArrayCount;
    const ommCpuOpacityMicromapUsageCount* descArrayHistogram;
    uint32_t descArrayHistogramCount;
    const void* indexBuffer;
    uint32_t indexCount;
    ommIndexFormat indexFormat;
    const ommCpuOpacityMicromapUsageCount* indexHistogram;
    uint32_t indexHistogramCount;
} ommCpuBakeResultDesc;

/* ref: omm.h:1173-1184 */
typedef struct ommDebugStats {
    uint64_t totalOpaque;
    uint64_t totalTransparent;
    uint64_t totalUnknownTransparent;
    uint64_t totalUnknownOpaque;
    uint32_t totalFullyOpaque;
    uint32_t totalFullyTransparent;
    uint32_t totalFullyUnknownOpaque;
    uint32_t totalFullyUnknownTransparent;
    float knownAreaMetric;
} ommDebugStats;

/* defaults (ref: omm.h:203-210, 234-274, 346-382, 462-490) */
static inline ommSamplerDesc ommSamplerDescDefault(void) {
    ommSamplerDesc v = {ommTextureAddressMode_MAX_NUM, ommTextureFilterMode_MAX_NUM, 0.f};
    return v;
}
static inline ommBakerCreationDesc ommBakerCreationDescDefault(void) {
    ommBakerCreationDesc v = {ommBakerType_MAX_NUM, {NULL, NULL, NULL, NULL}, {NULL, NULL}};
    return v;
}
static inline ommCpuTextureMipDesc ommCpuTextureMipDescDefault(void) {
    ommCpuTextureMipDesc v = {0, 0, 0, NULL};
    return v;
}
static inline ommCpuTextureDesc ommCpuTextureDescDefault(void) {
    ommCpuTextureDesc v = {ommCpuTextureFormat_MAX_NUM, ommCpuTextureFlags_None, NULL, 0, -1.f};
    return v;
}
static inline ommCpuBakeInputDesc ommCpuBakeInputDescDefault(void) {
    ommCpuBakeInputDesc v;
    v.bakeFlags = ommCpuBakeFlags_None;
    v.texture = 0;
    v.runtimeSamplerDesc = ommSamplerDescDefault();
    v.alphaMode = ommAlphaMode_MAX_NUM;
    v.texCoordFormat = ommTexCoordFormat_MAX_NUM;
    v.texCoords = NULL;
    v.texCoordStrideInBytes = 0;
    v.indexFormat = ommIndexFormat_MAX_NUM;
    v.indexBuffer = NULL;
    v.indexCount = 0;
    v.dynamicSubdivisionScale = 2.f;
    v.rejectionThreshold = 0.f;
    v.alphaCutoff = 0.5f;
    v.nearDuplicateDeduplicationFactor = 0.15f;
    v.alphaCutoffLessEqual = ommOpacityState_Transparent;
    v.alphaCutoffGreater = ommOpacityState_Opaque;
    v.format = ommFormat_OC1_4_State;
    v.formats = NULL;
    v.unknownStatePromotion = ommUnknownStatePromotion_ForceOpaque;
    v.unresolvedTriState = ommSpecialIndex_FullyUnknownOpaque;
    v.maxSubdivisionLevel = 8;
    v.maxArrayDataSize = 0xFFFFFFFFu;
    v.subdivisionLevels = NULL;
    v.maxWorkloadSize = 0xFFFFFFFFFFFFFFFFull;
    return v;
}

/* ---- entry points of the bake path --------------------------------------------------- */
/* ref: omm.h:276   (bake.cpp:36)  */ OMM_API ommLibraryDesc ommGetLibraryDesc(void);
/* ref: omm.h:278   (bake.cpp:410) */ OMM_API ommResult ommCreateBaker(const ommBakerCreationDesc* desc, ommBaker* outBaker);
/* ref: omm.h:280   (bake.cpp:457) */ OMM_API ommResult ommDestroyBaker(ommBaker baker);
/* ref: omm.h:568   (bake.cpp:44)  */ OMM_API ommResult ommCpuCreateTexture(ommBaker baker, const ommCpuTextureDesc* desc, ommCpuTexture* outTexture);
/* ref: omm.h:570   (bake.cpp:71)  */ OMM_API ommResult ommCpuGetTextureDesc(ommCpuTexture texture, ommCpuTextureDesc* outDesc);
/* ref: omm.h:572   (bake.cpp:84)  */ OMM_API ommResult ommCpuDestroyTexture(ommBaker baker, ommCpuTexture texture);
/* ref: omm.h:574   (bake.cpp:103) */ OMM_API ommResult ommCpuBake(ommBaker baker, const ommCpuBakeInputDesc* bakeInputDesc, ommCpuBakeResult* outBakeResult);
/* ref: omm.h:576   (bake.cpp:118) */ OMM_API ommResult ommCpuDestroyBakeResult(ommCpuBakeResult bakeResult);
/* ref: omm.h:578   (bake.cpp:129) */ OMM_API ommResult ommCpuGetBakeResultDesc(ommCpuBakeResult bakeResult, const ommCpuBakeResultDesc** desc);
/* ref: omm.h:1201  (debug_impl.cpp:512-641): state counts of a result, as the SDK's known-answer tests read them. */
/* ---- serialization of bake inputs / results ---------------------------------------------------------------
 * ABI note: omm.h declares the two descs below as C++ references inside extern "C" (omm.h:583, 590); at the ABI level they are
 * pointers, which is what this C header says. */
typedef struct ommCpuBlobDesc { void* data; uint64_t size; } ommCpuBlobDesc;                           /* ref: omm.h:532-536 */
typedef struct ommCpuDeserializedDesc {                                                                 /* ref: omm.h:546-555 */
    ommCpuSerializeFlags flags;
    int numInputDescs;
    const ommCpuBakeInputDesc* inputDescs;
    int numResultDescs;
    const ommCpuBakeResultDesc* resultDescs;
} ommCpuDeserializedDesc;
/* ref: omm.h:583 (bake.cpp:137) */ OMM_API ommResult ommCpuSerialize(ommBaker baker, const ommCpuDeserializedDesc* desc, ommCpuSerializedResult* outResult);
/* ref: omm.h:585 (bake.cpp:169) */ OMM_API ommResult ommCpuGetSerializedResultDesc(ommCpuSerializedResult result, const ommCpuBlobDesc** desc);
/* ref: omm.h:587 (bake.cpp:182) */ OMM_API ommResult ommCpuDestroySerializedResult(ommCpuSerializedResult result);
/* ref: omm.h:590 (bake.cpp:195) */ OMM_API ommResult ommCpuDeserialize(ommBaker baker, const ommCpuBlobDesc* desc, ommCpuDeserializedResult* outResult);
/* ref: omm.h:592 (bake.cpp:224) */ OMM_API ommResult ommCpuGetDeserializedDesc(ommCpuDeserializedResult result, const ommCpuDeserializedDesc** desc);
/* ref: omm.h:594 (bake.cpp:240) */ OMM_API ommResult ommCpuDestroyDeserializedResult(ommCpuDeserializedResult result);

/* ---- out of scope, exported as NOT_IMPLEMENTED stubs so that programs built against omm.h link ------------------
 * ref: omm.h:1127-1141 (D3D12 / Vulkan command-list baker), omm.h:1199, 1204 (debug dumps).  Opaque pointers stand in for the SDK's
 * GPU structures, which this header does not declare. */
OMM_API ommResult ommGpuGetStaticResourceData(int resource, uint8_t* data, size_t* outByteSize);
OMM_API ommResult ommGpuCreatePipeline(ommBaker baker, const void* pipelineCfg, void** outPipeline);
OMM_API ommResult ommGpuDestroyPipeline(ommBaker baker, void* pipeline);
OMM_API ommResult ommGpuGetPipelineDesc(void* pipeline, const void** outPipelineDesc);
OMM_API ommResult ommGpuGetPreDispatchInfo(void* pipeline, const void* config, void* outPreDispatchInfo);
OMM_API ommResult ommGpuDispatch(void* pipeline, const void* config, const void** outDispatchDesc);
OMM_API ommResult ommDebugSaveAsImages(ommBaker baker, const ommCpuBakeInputDesc* bakeInputDesc, const ommCpuBakeResultDesc* res, const void* desc);
OMM_API ommResult ommDebugSaveBinaryToDisk(ommBaker baker, const ommCpuBlobDesc* data, const char* path);

OMM_API ommResult ommDebugGetStats(ommBaker baker, const ommCpuBakeResultDesc* res, ommDebugStats* out);

/* ======================================================================================
 * B200 extension.  Nothing below exists in the SDK.
 * ====================================================================================== */

/* Per-bake device timings in milliseconds, measured with CUDA events on the bake's stream. */
typedef struct ommB200BakeTimings {
    float h2dMs;        /* upload of index / UV / per-triangle arrays                 */
    float setupMs;      /* triangle fetch, level selection, UV pre-dedup               */
    float classifyMs;   /* coarse (SAT) + fine micro-triangle classification kernels   */
    float postMs;       /* special-index scan, XXH64, dedup, sort, scan, pack, indices */
    float d2hMs;        /* download of the result arrays (copy-stream time of the overlapped part included) */
    float totalDeviceMs;/* first event to last event                                   */
    uint64_t microTriangles;     /* sum over work items of 4^level (items classified on this rank) */
    uint32_t workItems;          /* unique UV triangles after pre-dedup (global)        */
    uint32_t kernelLaunches;     /* kernels launched by this library for the bake       */
    uint64_t h2dBytes;
    uint64_t d2hBytes;
    uint64_t arrayDataBytes;     /* arrayDataSize of the result                         */
    uint32_t descCount;
    uint32_t reserved;
    /* host wall clock of the same call, milliseconds (std::chrono::steady_clock) */
    float hostStageMs;    /* ommB200StageInputs part: index scan, allocations, H2D      */
    float hostBakeMs;     /* device pipeline incl. the small read-backs; in ommCpuBake also the array-data
                             download, which starts with the first packed slice (d2hMs has its copy time) */
    float hostDownloadMs; /* host allocation + D2H of what was not sent during the bake  */
    float hostTotalMs;    /* whole ommCpuBake call                                       */
    float itemPostMs;     /* special-index scan + XXH64 of this rank's items (part of postMs) */
    float gatherMs;       /* sharded bakes: NCCL all-gather of the per-item records (digest + special index), incl. the wait
                             for the slowest rank (0 on one GPU; part of postMs) */
} ommB200BakeTimings;

/* Select the CUDA device used by bakers created afterwards on this thread's process (default: current device). */
OMM_API ommResult ommB200SetDevice(int cudaDevice);

/* Number of CUDA devices visible; 0 means the library cannot run (there is no CPU fallback). */
OMM_API int ommB200GetDeviceCount(void);

/* Big result arrays are downloaded into page-locked host blocks that the library recycles across bakes (default allocator only).
 * At most OMM_B200_PINNED_CACHE_MB (environment, default 4096) stay cached; everything cached is freed when the last baker is destroyed.
 * ommB200TrimHostPool frees cached blocks until at most keepBytes remain and returns the bytes still cached. */
OMM_API size_t ommB200TrimHostPool(size_t keepBytes);

/* Timings of the most recent successful ommCpuBake / ommB200BakeResident on this baker. */
OMM_API ommResult ommB200GetLastBakeTimings(ommBaker baker, ommB200BakeTimings* out);

/*
 * Resident bake: identical to ommCpuBake except that the inputs named by the desc are uploaded
 * once by ommB200StageInputs and the result stays in HBM until ommB200DownloadResult is called.
 * bench.py's device-resident "value" times ommB200BakeResident alone.
 */
typedef struct _ommB200StagedInputs* ommB200StagedInputs;
OMM_API ommResult ommB200StageInputs(ommBaker baker, const ommCpuBakeInputDesc* desc, ommB200StagedInputs* outStaged);
OMM_API ommResult ommB200DestroyStagedInputs(ommB200StagedInputs staged);
/* cudaStream: a cudaStream_t (may be NULL for the baker's own stream). The call is asynchronous w.r.t. the host only
 * up to the points where sizes must be read back (two small synchronisations). */
OMM_API ommResult ommB200BakeResident(ommBaker baker, ommB200StagedInputs staged, void* cudaStream, ommCpuBakeResult* outBakeResult);
/* Device pointers of a resident result (valid until the result is destroyed). */
typedef struct ommB200DeviceResultDesc {
    const void* arrayData;   /* device pointer, arrayDataSize bytes       */
    const void* descArray;   /* device pointer, descArrayCount * 8 bytes  */
    const void* indexBuffer; /* device pointer, indexCount * index size   */
    uint32_t arrayDataSize;
    uint32_t descArrayCount;
    uint32_t indexCount;
    ommIndexFormat indexFormat;
} ommB200DeviceResultDesc;
OMM_API ommResult ommB200GetDeviceResultDesc(ommCpuBakeResult bakeResult, ommB200DeviceResultDesc* out);
/* Materialise the host-side ommCpuBakeResultDesc of a resident result (no-op if already downloaded). */
OMM_API ommResult ommB200DownloadResult(ommCpuBakeResult bakeResult);

/*
 * Multi-GPU: one process per GPU of one box.  Every rank calls the same bake with the same desc and the same input arrays (rank 0
 * uploads them, the others receive them over NVLink); the work items, already in output order, are cut into contiguous unit-balanced
 * runs ("shards", dealt in boustrophedon order when there are several per rank) which the ranks classify; ONE all-gather of 12 bytes
 * per work item (block digest + special index) precedes the replicated dedup / offset merge; every rank then packs its own shards
 * straight to their final byte range of arrayData, and the ranges are exchanged as ommB200SetShardedResultMode says.
 * ncclUniqueIdBytes is the 128-byte ncclUniqueId created by rank 0 and distributed by the launcher (bench.py uses torch.distributed).
 */
OMM_API ommResult ommB200InitSharding(ommBaker baker, int rank, int worldSize, const void* ncclUniqueIdBytes, size_t idSize);
OMM_API ommResult ommB200GetNcclUniqueId(void* outBytes, size_t idSize);
/* Where the complete arrayData of a sharded bake ends up (descriptors, index buffer and histograms are complete on every rank either way;
 * set it alike on every rank before baking):
 *   Replicated (default)  every rank: ommB200BakeResident leaves it in every GPU's HBM (one group of in-place NCCL broadcasts of the shards'
 *                         byte ranges); after ommCpuBake any rank may ask for the host copy.
 *   OnRank0               rank 0 only: ommB200BakeResident gathers the other ranks' byte ranges into rank 0's HBM (ncclSend / ncclRecv);
 *                         ommCpuBake assembles the host copy in a page-locked window every rank writes over its own PCIe link, with no
 *                         device-to-device traffic at all.  On the other ranks ommCpuGetBakeResultDesc, ommB200DownloadResult and
 *                         ommB200GetDeviceResultDesc report INVALID_ARGUMENT. */
typedef enum ommB200ShardedResultMode { ommB200ShardedResultMode_Replicated = 0, ommB200ShardedResultMode_OnRank0 = 1 } ommB200ShardedResultMode;
OMM_API ommResult ommB200SetShardedResultMode(ommBaker baker, ommB200ShardedResultMode mode);
/* The partition used by sharded bakes, exposed for tests: unitPrefix is the exclusive prefix sum (entries = items + 1,
 * last entry = total) of the per-item balancing weights -- warp units (max(4^level / 32, 1)), times a level-line density
 * factor estimated from the texture when the baker is sharded; outFirstItem receives worldSize + 1 item indices. */
OMM_API ommResult ommB200ComputeShardBounds(const uint64_t* unitPrefix, uint32_t entries, int worldSize, uint32_t* outFirstItem);
/* How the runs are dealt (exposed for tests): a sharded bake cuts the work items into worldSize x ommB200ShardsPerRank(worldSize)
 * runs with the partition above (called with that product as its worldSize); run s is classified by rank ommB200ShardOwner(s). */
OMM_API int ommB200ShardsPerRank(int worldSize);
OMM_API int ommB200ShardOwner(int shard, int worldSize);

#endif /* OMM_B200_H_ */
