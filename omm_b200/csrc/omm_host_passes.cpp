// omm_host_passes.cpp -- the two optional, inherently serial post-classification passes of ommCpuBake, on the host:
//
//   a17  near-duplicate merge   DeduplicateSimilarLSH (3 iterations) / DeduplicateSimilarBruteForce + MergeWorkItems
//                               (ref: bake_cpu_impl.cpp:1068-1132, 1134-1352, 1354-1430), enabled by
//                               ommCpuBakeFlags_EnableNearDuplicateDetection (LSH) or internal flag bit 10 (brute force)
//   a18  size-budget Compress   (ref: bake_cpu_impl.cpp:1474-1688), enabled by maxArrayDataSize != 0xFFFFFFFF
//
// plus the PromoteToSpecialIndices call that sits between them (ref: :1432-1472, :1965).  Both passes are greedy,
// order-dependent walks over the work-item list (first/nearest match wins, std::mt19937 stream, std::sort on float keys),
// so they run as host C++ over the packed 2-bit state words downloaded from HBM; the device pipeline continues with
// the second exact dedup afterwards (omm_bake.cu).  They are off by default and not on the benchmarked path.
//
// Libraries whose exact behaviour matters and is inherited by linking the same ones the SDK build uses: libstdc++
// std::mt19937 / std::sort (introsort; equal keys!), libm powf / logf.
#include <algorithm>
#include <cmath>
#include <limits>
#include <random>
#include <set>
#include <unordered_map>

#include "omm_internal.h"

namespace ommb200 {

namespace {

inline uint32_t NumMicroTris(uint32_t level) { return 1u << (level << 1); }

struct States {  // view of one item's 2-bit state block
    uint32_t* w;
    inline uint32_t get(uint32_t i) const { return (w[i >> 4] >> ((i & 15) * 2)) & 3u; }
    inline uint32_t get3(uint32_t i) const {  // UT folded into UO (ref: bake_cpu_impl.cpp:374-377)
        const uint32_t s = get(i);
        return s == ommOpacityState_UnknownTransparent ? (uint32_t)ommOpacityState_UnknownOpaque : s;
    }
    inline void set(uint32_t i, uint32_t s) {
        uint32_t& x = w[i >> 4];
        const uint32_t sh = (i & 15) * 2;
        x = (x & ~(3u << sh)) | (s << sh);
    }
};
inline bool IsKnown(uint32_t s) { return s == ommOpacityState_Opaque || s == ommOpacityState_Transparent; }
inline bool IsUnknown(uint32_t s) { return s == ommOpacityState_UnknownOpaque || s == ommOpacityState_UnknownTransparent; }

// number of micro-triangles whose 3-state values differ (ref: :1068-1083), word-parallel
uint32_t HammingDistance3State(const uint32_t* a, const uint32_t* b, uint32_t n) {
    uint32_t diff = 0;
    const uint32_t words = n >= 16 ? n >> 4 : 1;
    const uint32_t tailMask = n >= 16 ? 0xFFFFFFFFu : ((1u << (2 * n)) - 1u);
    for (uint32_t i = 0; i < words; ++i) {
        const uint32_t a3 = a[i] | ((a[i] >> 1) & 0x55555555u), b3 = b[i] | ((b[i] >> 1) & 0x55555555u);
        const uint32_t t = (a3 ^ b3) & (i + 1 == words ? tailMask : 0xFFFFFFFFu);
        diff += (uint32_t)__builtin_popcount((t | (t >> 1)) & 0x55555555u);
    }
    return diff;
}

// XXH64 (xxHash specification), used by the LSH layer hashes (ref: :1254)
constexpr uint64_t P1 = 0x9E3779B185EBCA87ull, P2 = 0xC2B2AE3D27D4EB4Full, P3 = 0x165667B19E3779F9ull, P4 = 0x85EBCA77C2B2AE63ull, P5 = 0x27D4EB2F165667C5ull;
inline uint64_t Rotl(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
inline uint64_t Round(uint64_t acc, uint64_t in) { return Rotl(acc + in * P2, 31) * P1; }
inline uint64_t Merge(uint64_t acc, uint64_t v) { return (acc ^ Round(0, v)) * P1 + P4; }
uint64_t Xxh64(const void* data, size_t len, uint64_t seed) {
    const uint8_t* p = (const uint8_t*)data;
    const uint8_t* end = p + len;
    uint64_t h;
    auto rd64 = [](const uint8_t* q) { uint64_t v; memcpy(&v, q, 8); return v; };
    auto rd32 = [](const uint8_t* q) { uint32_t v; memcpy(&v, q, 4); return v; };
    if (len >= 32) {
        uint64_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
        do {
            v1 = Round(v1, rd64(p)); v2 = Round(v2, rd64(p + 8)); v3 = Round(v3, rd64(p + 16)); v4 = Round(v4, rd64(p + 24));
            p += 32;
        } while (p + 32 <= end);
        h = Rotl(v1, 1) + Rotl(v2, 7) + Rotl(v3, 12) + Rotl(v4, 18);
        h = Merge(h, v1); h = Merge(h, v2); h = Merge(h, v3); h = Merge(h, v4);
    } else
        h = seed + P5;
    h += (uint64_t)len;
    while (p + 8 <= end) { h ^= Round(0, rd64(p)); h = Rotl(h, 27) * P1 + P4; p += 8; }
    if (p + 4 <= end) { h ^= (uint64_t)rd32(p) * P1; h = Rotl(h, 23) * P2 + P3; p += 4; }
    while (p < end) { h ^= (uint64_t)(*p) * P5; h = Rotl(h, 11) * P1; p++; }
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
}

struct Ctx {
    HostPassItem* items;
    uint32_t count;
    uint32_t* words;
    const unsigned long long* wordStart;
    States st(uint32_t i) const { return States{words + wordStart[i]}; }
};

// ref: :1093-1132
void MergeWorkItems(Ctx& c, uint32_t to, uint32_t from) {
    c.items[to].numPrims += c.items[from].numPrims;
    c.items[from].numPrims = 0;
    c.items[from].special = -1;
    c.items[from].mergedInto = to;
    c.items[to].statesChanged = true;
    States a = c.st(to), b = c.st(from);
    const uint32_t n = NumMicroTris(c.items[from].level);
    for (uint32_t u = 0; u < n; ++u) {
        const uint32_t ts = a.get(u), fs = b.get(u);
        if (ts != fs) {
            if (IsKnown(fs) && IsKnown(ts)) a.set(u, ommOpacityState_UnknownOpaque);
            else if (IsKnown(ts) && IsUnknown(fs)) a.set(u, fs);
        }
    }
}

// ref: :1134-1352
void DeduplicateSimilarLSH(Ctx& c, const ommCpuBakeInputDesc& desc, uint32_t iterations) {
    std::mt19937 mt(42);
    struct HashTable {
        std::vector<uint32_t> bitIndices;
        std::vector<uint64_t> workItemHashes;
        std::unordered_map<uint64_t, std::vector<uint32_t>> layerHashToWorkItem;
    };
    for (uint32_t attempts = 0; attempts < iterations; ++attempts) {
        std::vector<uint32_t> batch;
        batch.reserve(c.count);
        std::vector<HashTable> hashTables;
        std::vector<uint32_t> bitSamples;
        std::set<uint32_t> potentialMatches;
        for (uint32_t level = 1; level <= 12; ++level) {
            batch.clear();
            for (uint32_t i = 0; i < c.count; ++i) {
                const HostPassItem& it = c.items[i];
                if (it.special != 0 || it.format != ommFormat_OC1_4_State || it.level != level) continue;
                batch.push_back(i);
            }
            if (batch.empty()) continue;
            const uint32_t numMicroTriangles = NumMicroTris(level);
            const uint32_t n = (uint32_t)batch.size();
            const uint32_t d = numMicroTriangles;
            const float r = desc.nearDuplicateDeduplicationFactor * d;
            const float cc = 4.0f;
            const float p = 1.f / cc;
            const float Lf = std::ceil(std::pow((float)n, p));
            const uint32_t L = (uint32_t)Lf;
            if (L == 0) continue;
            const uint32_t k = uint32_t(std::ceil((std::log((float)n) * d) / (cc * r)));
            if (k == 0) continue;
            hashTables.resize(L);
            for (HashTable& ht : hashTables) {
                ht.workItemHashes.resize(c.count, 0);
                ht.bitIndices.resize(k);
                ht.layerHashToWorkItem.clear();
                for (uint32_t& bitIndex : ht.bitIndices) {
                    const uint32_t random = mt();
                    bitIndex = random & (numMicroTriangles - 1);
                }
            }
            bitSamples.resize(k);
            for (uint32_t idx : batch) {
                const States s = c.st(idx);
                for (HashTable& ht : hashTables) {
                    for (uint32_t kIt = 0; kIt < k; ++kIt) bitSamples[kIt] = s.get3(ht.bitIndices[kIt]);
                    const uint64_t hash = Xxh64(bitSamples.data(), sizeof(uint32_t) * bitSamples.size(), 42);
                    ht.workItemHashes[idx] = hash;
                    ht.layerHashToWorkItem[hash].push_back(idx);
                }
            }
            for (uint32_t idx : batch) {
                if (c.items[idx].special != 0) continue;  // merged away earlier in this pass
                potentialMatches.clear();
                for (const HashTable& ht : hashTables) {
                    const auto it = ht.layerHashToWorkItem.find(ht.workItemHashes[idx]);
                    if (it == ht.layerHashToWorkItem.end()) continue;
                    for (uint32_t cand : it->second) {
                        if (cand == idx) continue;
                        if (c.items[cand].special != 0) continue;
                        if (potentialMatches.size() > 3 * L) break;
                        potentialMatches.insert(cand);
                    }
                }
                float minDist = std::numeric_limits<float>::max();
                int32_t nearest = -1;
                for (uint32_t cand : potentialMatches) {
                    const float dist = float(HammingDistance3State(c.words + c.wordStart[idx], c.words + c.wordStart[cand], numMicroTriangles));
                    if (dist < r && dist < minDist) {
                        minDist = dist;
                        nearest = (int32_t)cand;
                    }
                }
                if (nearest >= 0) MergeWorkItems(c, idx, (uint32_t)nearest);
            }
        }
    }
}

// ref: :1354-1430
void DeduplicateSimilarBruteForce(Ctx& c) {
    if (c.count == 0) return;
    static constexpr float kMergeThreshold = 0.1f;
    static constexpr uint32_t kMaxComparsions = 2048;
    std::set<uint32_t> merged;
    for (uint32_t itA = 0; itA < c.count - 1; ++itA) {
        const HostPassItem& A = c.items[itA];
        if (A.special != 0 || A.format != ommFormat_OC1_4_State) continue;
        const uint32_t searchStart = itA + 1;
        const uint32_t searchEnd = std::min<uint32_t>(kMaxComparsions + searchStart, c.count);
        float minDist = std::numeric_limits<float>::max();
        int32_t nearest = -1;
        for (uint32_t itB = searchStart; itB < searchEnd; ++itB) {
            const HostPassItem& B = c.items[itB];
            if (B.special != 0 || B.format != ommFormat_OC1_4_State || B.numPrims == 0 || A.level != B.level) continue;
            if (merged.find(itB) != merged.end()) continue;
            const uint32_t n = NumMicroTris(A.level);
            const float dist = float(HammingDistance3State(c.words + c.wordStart[itA], c.words + c.wordStart[itB], n)) / n;
            if (dist < kMergeThreshold && dist < minDist) {
                minDist = dist;
                nearest = (int32_t)itB;
            }
        }
        if (nearest >= 0) {
            merged.insert(itA);
            merged.insert((uint32_t)nearest);
            MergeWorkItems(c, itA, (uint32_t)nearest);
        }
    }
}

// ref: :1432-1472
void PromoteToSpecialIndices(Ctx& c, const ommCpuBakeInputDesc& desc) {
    const bool disableSpecial = ((uint32_t)desc.bakeFlags & ommCpuBakeFlags_DisableSpecialIndices) != 0;
    for (uint32_t i = 0; i < c.count; ++i) {
        HostPassItem& it = c.items[i];
        if (it.special != 0) continue;
        const States s = c.st(i);
        const uint32_t n = NumMicroTris(it.level);
        bool allEqual = true;
        uint32_t common = s.get(0);
        for (uint32_t u = 1; u < n; ++u) allEqual &= common == s.get(u);
        if (!allEqual && desc.rejectionThreshold > 0.f) {
            uint32_t known = 0;
            for (uint32_t u = 0; u < n; ++u) known += IsKnown(s.get(u));
            const float knownFrac = known / (float)n;
            if (knownFrac < desc.rejectionThreshold) {
                allEqual = true;
                common = ommOpacityState_UnknownTransparent;
            }
        }
        if (allEqual && !disableSpecial) it.special = -int32_t(common) - 1;
    }
}

float Area2D(const HostPassItem& it) {  // ref: bake_cpu_impl.cpp:464-468 via util/geometry.h:130-138
    const float v0x = it.uv[4] - it.uv[0], v0y = it.uv[5] - it.uv[1], v1x = it.uv[2] - it.uv[0], v1y = it.uv[3] - it.uv[1];
    const float cz = v0x * v1y - v1x * v0y;
    return 0.5f * std::sqrt(cz * cz);
}

// ref: :1474-1555
float KnownRatio(const Ctx& c, uint32_t i) {
    const States s = c.st(i);
    const uint32_t total = NumMicroTris(c.items[i].level);
    uint32_t known = 0;
    for (uint32_t u = 0; u < total; ++u) known += IsKnown(s.get3(u));
    return (float)known / total;
}
float KnownRatioIfDownsampled(const Ctx& c, uint32_t i) {
    const States s = c.st(i);
    const size_t n = NumMicroTris(c.items[i].level - 1);
    uint32_t known = 0;
    for (size_t u = 0; u < n; ++u) {
        const uint32_t s0 = s.get3(4 * u), s1 = s.get3(4 * u + 1), s2 = s.get3(4 * u + 2), s3 = s.get3(4 * u + 3);
        if (IsKnown(s0) && s0 == s1 && s0 == s2 && s0 == s3) known++;
    }
    return known / (float)n;
}
void DownsampleOneLevel(Ctx& c, uint32_t i) {
    HostPassItem& it = c.items[i];
    it.level -= 1;
    it.statesChanged = true;
    States s = c.st(i);
    const size_t n = NumMicroTris(it.level);
    for (size_t u = 0; u < n; ++u) {
        const uint32_t s0 = s.get3(4 * u), s1 = s.get3(4 * u + 1), s2 = s.get3(4 * u + 2), s3 = s.get3(4 * u + 3);
        if (IsKnown(s0) && s0 == s1 && s0 == s2 && s0 == s3) s.set(u, s0);
        else s.set(u, ommOpacityState_UnknownOpaque);
    }
    // The SDK only resizes its byte vectors here (OmmArrayDataVector::ShrinkTo, ref: :413-421): the bytes beyond the new size
    // keep their old values AND the digest of the second exact dedup still runs over the ORIGINAL length
    // (_ommArrayDataSize is never updated, ref: :388, :1038-1040).  The stale fields are therefore left untouched; the
    // device hashes 4^originalLevel fields (ItemRec::hashLevel) and masks them out everywhere else.
}

// ref: :1557-1688
ommResult Compress(Ctx& c, const ommCpuBakeInputDesc& desc) {
    struct Info {
        float knownRatio = 0.f, knownRatioIfWeDownsample = 0.f, totalArea = 0.f;
        size_t totalMemory = 0, totalMemoryIfWeDownsample = 0;
        float coveragePerByte = 0.f;
    };
    auto compute = [&c](uint32_t i, Info& out) {
        const HostPassItem& it = c.items[i];
        out.knownRatio = KnownRatio(c, i);
        out.knownRatioIfWeDownsample = KnownRatioIfDownsampled(c, i);
        out.totalArea = 0;
        const float area = Area2D(it);
        for (uint32_t p = 0; p < it.numPrims; ++p) out.totalArea += area;
        out.totalMemory = std::max<size_t>(1, ((size_t)NumMicroTris(it.level) * 2) / 8);
        out.totalMemoryIfWeDownsample = std::max<size_t>(1, ((size_t)NumMicroTris(it.level - 1) * 2) / 8);
        const size_t memDelta = out.totalMemory - out.totalMemoryIfWeDownsample;
        const float coverageDelta = out.knownRatio - out.knownRatioIfWeDownsample;
        out.coveragePerByte = out.totalArea * coverageDelta / memDelta;
    };
    std::vector<std::pair<int, Info>> active;
    for (int i = 0; i < (int)c.count; ++i) {
        const HostPassItem& it = c.items[i];
        if (it.level == 0 || it.numPrims == 0 || it.special != 0) continue;
        Info info;
        compute((uint32_t)i, info);
        active.push_back(std::make_pair(i, info));
    }
    size_t totalMemory = 0;
    for (const auto& a : active) totalMemory += a.second.totalMemory;
    if (totalMemory < desc.maxArrayDataSize) return ommResult_SUCCESS;
    auto sortFn = [](const std::pair<int, Info>& ia, const std::pair<int, Info>& ib) { return ia.second.coveragePerByte < ib.second.coveragePerByte; };
    std::sort(active.begin(), active.end(), sortFn);
    while (totalMemory >= desc.maxArrayDataSize && active.size() != 0) {
        const int N = (int)active.size();
        for (int i = 0; i < N; ++i) {
            const uint32_t item = (uint32_t)active[i].first;
            totalMemory -= active[i].second.totalMemory;
            if (c.items[item].level == 0) return ommResult_FAILURE;
            DownsampleOneLevel(c, item);
            totalMemory += active[i].second.totalMemoryIfWeDownsample;
            if (c.items[item].level == 0) {
                active[i].first = -1;
                continue;
            }
            compute(item, active[i].second);
            if (totalMemory < desc.maxArrayDataSize) break;
            if (i + 1 != N) {
                if (active[i].second.coveragePerByte < active[i + 1].second.coveragePerByte) i--;
            }
        }
        for (int i = 0; i < (int)active.size(); ++i) {
            if (active[i].first == -1) {
                std::swap(active[i], active[active.size() - 1]);
                active.pop_back();
                i--;
            }
        }
        std::sort(active.begin(), active.end(), sortFn);
    }
    return ommResult_SUCCESS;
}

}  // namespace

bool HostPassesNeeded(const ommCpuBakeInputDesc& desc) {
    const uint32_t flags = (uint32_t)desc.bakeFlags;
    if (flags & ommCpuBakeFlags_DisableDuplicateDetection) return desc.maxArrayDataSize != 0xFFFFFFFFu;
    return (flags & ommCpuBakeFlags_EnableNearDuplicateDetection) != 0 || desc.maxArrayDataSize != 0xFFFFFFFFu;
}

// Steps 3-6 of the SDK's post-classification sequence (ref: bake_cpu_impl.cpp:1961-1967); steps 1-2 ran on the device
// before, steps 7-8 run on the device after.
ommResult RunHostPasses(const ommCpuBakeInputDesc& desc, HostPassItem* items, uint32_t count, uint32_t* words, const unsigned long long* wordStart) {
    const uint32_t flags = (uint32_t)desc.bakeFlags;
    const bool disableDup = (flags & ommCpuBakeFlags_DisableDuplicateDetection) != 0;
    const bool nearDup = (flags & ommCpuBakeFlags_EnableNearDuplicateDetection) != 0;
    const bool bruteForce = (flags & (1u << 10)) != 0;
    Ctx c{items, count, words, wordStart};
    if (!disableDup && nearDup && !bruteForce) DeduplicateSimilarLSH(c, desc, 3);
    if (!disableDup && nearDup && bruteForce) DeduplicateSimilarBruteForce(c);
    PromoteToSpecialIndices(c, desc);
    if (desc.maxArrayDataSize != 0xFFFFFFFFu) {
        const ommResult rc = Compress(c, desc);
        if (rc != ommResult_SUCCESS) return rc;
    }
    return ommResult_SUCCESS;
}

uint64_t HostXxh64(const void* data, size_t len, uint64_t seed) { return Xxh64(data, len, seed); }

}  // namespace ommb200
