// omm_post_passes.cuh -- the two optional passes of ommCpuBake that sit between the first exact dedup and the second one:
//
//   a17  near-duplicate merge: LSH by bit sampling (ommCpuBakeFlags_EnableNearDuplicateDetection) or a windowed exhaustive search
//        (internal flag bit 10)                                    -- SDK behaviour: bake_cpu_impl.cpp:1068-1132, 1134-1352, 1354-1430
//   a18  size-budget compression (maxArrayDataSize != 0xFFFFFFFF)  -- SDK behaviour: bake_cpu_impl.cpp:1474-1688
//
// Division of labour.  Everything that touches micro-triangle states runs on the GPU, on the 2-bit state words where the classifier left
// them: the LSH layer hashes (XXH64 of sampled states, one thread per item and table), 3-state Hamming distances of candidate pairs (one
// warp per pair), the state merge of two blocks, the "known" counts of every work item at every level it could be downsampled to (one
// pass: a group of 4^j micro-triangles collapses to a known state iff all of them hold that state), and the downsampling itself.  The
// host keeps what is inherently sequential in the SDK's definition -- which candidate is nearest *given the merges made so far*, the
// greedy order of the budget pass -- and works on hashes, distances and counts only; no state word crosses PCIe.
//
// The walk order of both passes is the SDK's first-seen work-item order (index w); device arrays are indexed by the position s of an
// item in output order (omm_bake.cu K3b), hence PassItem::pos.
#pragma once

#include <algorithm>
#include <cmath>
#include <random>
#include <unordered_map>
#include <vector>

#include "omm_xxh64.h"

namespace ommb200 {

struct PassItem {      // host view of one work item, indexed by w
    uint32_t pos;      // position in output order
    uint32_t prims;    // primitives referencing it
    int32_t special;   // 0 = none, -1 = gave its primitives away, -2..-5 special indices (as on the device)
    uint32_t root;     // w of the item that now holds this item's primitives
    float area;        // UV area of the triangle (budget pass)
    uint8_t level, format, levelAtStart;
};

// ---------------------------------------------------------------------------------------------------------------------
// device services
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t StateField(const uint32_t* __restrict__ words, uint32_t i) { return (words[i >> 4] >> ((i & 15u) * 2u)) & 3u; }
// 3-state view of a word: UnknownTransparent (2) folded into UnknownOpaque (3)   (ref: bake_cpu_impl.cpp:374-377)
__device__ __forceinline__ uint32_t Fold3(uint32_t w) { return w | ((w >> 1) & 0x55555555u); }

// LSH layer hashes: hash[t * n + i] = XXH64(seed 42) of the k sampled 3-state values (as uint32) of batch item i under table t
__global__ void LshHashKernel(const uint32_t* __restrict__ stateWords, const unsigned long long* __restrict__ wordStart, const uint32_t* __restrict__ batchPos,
                              uint32_t n, const uint32_t* __restrict__ bitIndices, uint32_t tables, uint32_t k, uint64_t* __restrict__ hash) {
    const unsigned long long id = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (unsigned long long)n * tables) return;
    const uint32_t i = (uint32_t)(id % n), t = (uint32_t)(id / n);
    const uint32_t* words = stateWords + wordStart[batchPos[i]];
    xxh::WordStream s(42);
    for (uint32_t j = 0; j < k; ++j) {
        const uint32_t f = StateField(words, __ldg(&bitIndices[t * k + j]));
        s.push(f == 2u ? 3u : f);
    }
    hash[id] = s.finish();
}

// 3-state Hamming distance of pairs of equally sized blocks, one warp per pair (positions in output order)
__global__ void PairDistanceKernel(const uint32_t* __restrict__ stateWords, const unsigned long long* __restrict__ wordStart, const uint2* __restrict__ pairs,
                                   const uint32_t* __restrict__ pairFields, uint32_t numPairs, uint32_t* __restrict__ dist) {
    const uint32_t p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (p >= numPairs) return;
    const uint32_t n = pairFields[p];
    const uint32_t* a = stateWords + wordStart[pairs[p].x];
    const uint32_t* b = stateWords + wordStart[pairs[p].y];
    const uint32_t words = n >= 16 ? n >> 4 : 1, tail = n >= 16 ? 0xFFFFFFFFu : ((1u << (2 * n)) - 1u);
    uint32_t diff = 0;
    for (uint32_t i = lane; i < words; i += 32) {
        const uint32_t x = (Fold3(a[i]) ^ Fold3(b[i])) & tail;
        diff += __popc((x | (x >> 1)) & 0x55555555u);
    }
    diff = __reduce_add_sync(0xFFFFFFFFu, diff);
    if (lane == 0) dist[p] = diff;
}

// Windowed exhaustive search: for item A = list[a] every later list entry within `reach` FIRST-SEEN indices of it and of the same
// level whose normalised distance is below the threshold is appended to `out` as (a, b, distance).  One warp per A.
struct NearPair {
    uint32_t a, b, dist;
};
__global__ void WindowSearchKernel(const uint32_t* __restrict__ stateWords, const unsigned long long* __restrict__ wordStart, const uint32_t* __restrict__ listPos,
                                   const uint32_t* __restrict__ listIndex, const uint8_t* __restrict__ listLevel, uint32_t count, uint32_t reach, float threshold,
                                   NearPair* __restrict__ out, uint32_t capacity, uint32_t* __restrict__ outCount) {
    const uint32_t a = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (a >= count) return;
    const uint32_t level = listLevel[a], n = 1u << (2 * level), words = n >= 16 ? n >> 4 : 1, tail = n >= 16 ? 0xFFFFFFFFu : ((1u << (2 * n)) - 1u);
    const uint32_t* wa = stateWords + wordStart[listPos[a]];
    for (uint32_t b = a + 1; b < count && listIndex[b] - listIndex[a] <= reach; ++b) {
        if (listLevel[b] != level) continue;
        const uint32_t* wb = stateWords + wordStart[listPos[b]];
        uint32_t diff = 0;
        for (uint32_t i = lane; i < words; i += 32) {
            const uint32_t x = (Fold3(wa[i]) ^ Fold3(wb[i])) & tail;
            diff += __popc((x | (x >> 1)) & 0x55555555u);
        }
        diff = __reduce_add_sync(0xFFFFFFFFu, diff);
        // the SDK compares float(diff) / n with 0.1f (bake_cpu_impl.cpp:1399-1401): the same expression here (IEEE division, -prec-div)
        if (lane == 0 && (float)diff / (float)n < threshold) {
            const uint32_t slot = atomicAdd(outCount, 1u);
            if (slot < capacity) out[slot] = NearPair{a, b, diff};
        }
    }
}

// `to` absorbs `from` (ref: bake_cpu_impl.cpp:1112-1130): where both are known and differ the micro-triangle becomes UnknownOpaque, where
// `to` is known and `from` unknown it takes `from`'s unknown state; anything else keeps `to`.
__global__ void MergeStatesKernel(uint32_t* __restrict__ stateWords, const unsigned long long* __restrict__ wordStart, uint32_t toPos, uint32_t fromPos,
                                  uint32_t fields) {
    const uint32_t words = fields >= 16 ? fields >> 4 : 1, valid = fields >= 16 ? 0xFFFFFFFFu : ((1u << (2 * fields)) - 1u);
    uint32_t* t = stateWords + wordStart[toPos];
    const uint32_t* f = stateWords + wordStart[fromPos];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < words; i += gridDim.x * blockDim.x) {
        const uint32_t tw = t[i], fw = f[i];
        const uint32_t lo = 0x55555555u;
        const uint32_t tUnknown = (tw >> 1) & lo, fUnknown = (fw >> 1) & lo;  // bit per field
        const uint32_t x = tw ^ fw;
        const uint32_t differ = (x | (x >> 1)) & lo;
        const uint32_t bothKnown = differ & ~tUnknown & ~fUnknown, takeFrom = differ & ~tUnknown & fUnknown;
        const uint32_t m3 = bothKnown * 3u, mf = takeFrom * 3u;  // two-bit masks
        const uint32_t merged = (tw & ~(m3 | mf)) | m3 | (fw & mf);
        t[i] = (merged & valid) | (tw & ~valid);
    }
}

// known[j] for j = 0..level: how many aligned groups of 4^j micro-triangles of the block hold one KNOWN state throughout, i.e. the number
// of known micro-triangles the block would have after j downsampling steps (ref: ComputeKnownRatio / DownsampleOneLevel,
// bake_cpu_impl.cpp:1474-1555).  One thread per listed item; the groups close like the digits of a base-4 counter.
__global__ void KnownPyramidKernel(const uint32_t* __restrict__ stateWords, const unsigned long long* __restrict__ wordStart, const uint32_t* __restrict__ listPos,
                                   const uint8_t* __restrict__ listLevel, uint32_t count, uint32_t* __restrict__ known /* [count][13] */) {
    const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= count) return;
    const uint32_t level = listLevel[id], n = 1u << (2 * level);
    const uint32_t* words = stateWords + wordStart[listPos[id]];
    uint32_t cnt[13];
    for (int j = 0; j < 13; ++j) cnt[j] = 0;
    // open group per height: children seen, whether all of them were uniform-known, their common state
    uint32_t seen[13], ok[13], val[13];
    for (int j = 0; j < 13; ++j) seen[j] = 0, ok[j] = 1, val[j] = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t s = StateField(words, i);
        uint32_t uniform = s < 2u ? 1u : 0u, v = s;  // a leaf is "uniform-known" iff it is known
        cnt[0] += uniform;
        for (uint32_t j = 1; j <= level; ++j) {  // the leaf closes its parent when it is the fourth child, and so on upwards
            if (seen[j] == 0) { ok[j] = uniform; val[j] = v; }
            else ok[j] = ok[j] & uniform & (val[j] == v ? 1u : 0u);
            if (++seen[j] < 4) break;
            seen[j] = 0;
            uniform = ok[j];
            v = val[j];
            cnt[j] += uniform;
        }
    }
    for (uint32_t j = 0; j < 13; ++j) known[(size_t)id * 13 + j] = j <= level ? cnt[j] : 0u;
}

// One downsampling step of the listed items (ref: DownsampleOneLevel, bake_cpu_impl.cpp:1499-1555): micro-triangle u of the coarser block is
// the common state of micro-triangles 4u..4u+3 when that is one known state, else UnknownOpaque.  Written to `scratch` first (a thread's
// output word overlaps other threads' input words), then copied over the head of the block.  The SDK shrinks its byte vectors without
// touching the bytes behind the new end and keeps hashing the ORIGINAL length afterwards (OmmArrayDataVector::ShrinkTo never updates
// _ommArrayDataSize): the stale fields are part of the second dedup's digest, so they are preserved here too.
__global__ void DownsampleComputeKernel(const uint32_t* __restrict__ stateWords, const unsigned long long* __restrict__ wordStart, const uint32_t* __restrict__ listPos,
                                        const uint8_t* __restrict__ listLevelNow, uint32_t count, uint32_t* __restrict__ scratch) {
    const uint32_t id = blockIdx.y;
    if (id >= count) return;
    const uint32_t fieldsOut = 1u << (2 * (listLevelNow[id] - 1u));
    const uint32_t wordsOut = fieldsOut >= 16 ? fieldsOut >> 4 : 1;
    const unsigned long long base = wordStart[listPos[id]];
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < wordsOut; w += gridDim.x * blockDim.x) {
        uint32_t out = 0;
        const uint32_t fieldsHere = fieldsOut >= 16 ? 16u : fieldsOut;
        for (uint32_t q = 0; q < fieldsHere; ++q) {
            const uint32_t u = w * 16u + q;
            const uint32_t in = (stateWords[base + (u >> 2)] >> ((u & 3u) * 8u)) & 0xFFu;  // the four children: one byte
            const uint32_t c = Fold3(in) & 0xFFu;
            const uint32_t s0 = c & 3u;
            const bool same = c == s0 * 0x55u;
            out |= ((same && s0 < 2u) ? s0 : 3u) << (2 * q);
        }
        scratch[base + w] = out;
    }
}
__global__ void DownsampleCommitKernel(uint32_t* __restrict__ stateWords, const unsigned long long* __restrict__ wordStart, const uint32_t* __restrict__ listPos,
                                       const uint8_t* __restrict__ listLevelNow, uint32_t count, const uint32_t* __restrict__ scratch) {
    const uint32_t id = blockIdx.y;
    if (id >= count) return;
    const uint32_t fieldsOut = 1u << (2 * (listLevelNow[id] - 1u));
    const uint32_t wordsOut = fieldsOut >= 16 ? fieldsOut >> 4 : 1, valid = fieldsOut >= 16 ? 0xFFFFFFFFu : ((1u << (2 * fieldsOut)) - 1u);
    const unsigned long long base = wordStart[listPos[id]];
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < wordsOut; w += gridDim.x * blockDim.x)
        stateWords[base + w] = (scratch[base + w] & valid) | (stateWords[base + w] & ~valid);
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
class PostPasses {
  public:
    PostPasses(cudaStream_t stream, std::vector<PassItem>& items, uint32_t* stateWords, const unsigned long long* wordStart, unsigned long long totalWords)
        : stream_(stream), items_(items), stateWords_(stateWords), wordStart_(wordStart), totalWords_(totalWords), dirty_(items.size(), 0), stamp_(items.size(), 0) {}
    ~PostPasses() {
        for (void* p : owned_) cudaFreeAsync(p, stream_);
    }
    uint32_t launches = 0;

    // ---- a17, LSH ----------------------------------------------------------------------------------------------------------------------
    // SDK contract (bake_cpu_impl.cpp:1134-1352): `iterations` passes; in each, for every level 1..12, the 4-state items without a special
    // index form a batch (first-seen order); L = ceil(n^(1/4)) tables of k = ceil(ln n * d / (4 r)) sampled positions drawn from one
    // std::mt19937(42) stream; every batch item, in order, gathers the unmerged members of its buckets (table by table, bucket members in
    // batch order, no more once it holds more than 3L) and absorbs the nearest one closer than r = factor * d (ties: lowest index).
    bool nearDuplicatesLsh(float factor, uint32_t iterations) {
        std::mt19937 rng(42);
        for (uint32_t pass = 0; pass < iterations; ++pass)
            for (uint32_t level = 1; level <= 12; ++level) {
                std::vector<uint32_t> batch;
                for (uint32_t w = 0; w < items_.size(); ++w)
                    if (items_[w].special == 0 && items_[w].format == ommFormat_OC1_4_State && items_[w].level == level) batch.push_back(w);
                if (batch.empty()) continue;
                const uint32_t n = (uint32_t)batch.size(), d = 1u << (2 * level);
                const float r = factor * d;
                const float c = 4.0f;
                const uint32_t L = (uint32_t)std::ceil(std::pow((float)n, 1.f / c));
                if (L == 0) continue;
                const uint32_t k = uint32_t(std::ceil((std::log((float)n) * d) / (c * r)));
                if (k == 0) continue;
                std::vector<uint32_t> sampled((size_t)L * k);
                for (uint32_t& s : sampled) s = (uint32_t)rng() & (d - 1);
                if (!lshRound(batch, d, r, L, k, sampled)) return false;
            }
        return true;
    }

    // ---- a17, windowed exhaustive search (internal flag) ---------------------------------------------------------------------------------
    // SDK contract (bake_cpu_impl.cpp:1354-1430): every 4-state item without a special index, in order, absorbs the nearest eligible item among
    // the next 2048 work items of its level whose normalised distance is below 0.1 (ties: lowest index).  A partner always comes later in the
    // order than the item absorbing it, and only absorbing items change, so every distance the walk needs is one between blocks as they
    // are NOW: one device search up front, the walk itself only filters by what has been absorbed meanwhile.
    bool nearDuplicatesWindowed() {
        constexpr uint32_t kReach = 2048;
        constexpr float kThreshold = 0.1f;
        std::vector<uint32_t> listW, listPos;
        std::vector<uint8_t> listLevel;
        for (uint32_t w = 0; w < items_.size(); ++w)
            if (items_[w].special == 0 && items_[w].format == ommFormat_OC1_4_State) {
                listW.push_back(w); listPos.push_back(items_[w].pos); listLevel.push_back(items_[w].level);
            }
        const uint32_t count = (uint32_t)listW.size();
        if (count < 2) return true;
        uint32_t *dPos = upload(listPos), *dIdx = upload(listW), *dCount = nullptr;
        uint8_t* dLevel = upload(listLevel);
        if (!dPos || !dIdx || !dLevel || !alloc(&dCount, 1)) return false;
        std::vector<NearPair> found;
        for (size_t capacity = std::max<size_t>(1u << 16, (size_t)count * 8);; capacity *= 4) {
            NearPair* dOut = nullptr;
            if (!alloc(&dOut, capacity)) return false;
            cudaMemsetAsync(dCount, 0, 4, stream_);
            WindowSearchKernel<<<(count + 7) / 8, 256, 0, stream_>>>(stateWords_, wordStart_, dPos, dIdx, dLevel, count, kReach, kThreshold, dOut, (uint32_t)capacity, dCount);
            launches++;
            uint32_t produced = 0;
            if (cudaMemcpyAsync(&produced, dCount, 4, cudaMemcpyDeviceToHost, stream_) != cudaSuccess || cudaStreamSynchronize(stream_) != cudaSuccess) return false;
            if (produced <= capacity) {
                found.resize(produced);
                if (produced && (cudaMemcpyAsync(found.data(), dOut, sizeof(NearPair) * produced, cudaMemcpyDeviceToHost, stream_) != cudaSuccess ||
                                 cudaStreamSynchronize(stream_) != cudaSuccess))
                    return false;
                break;
            }
        }
        std::sort(found.begin(), found.end(), [](const NearPair& x, const NearPair& y) { return x.a != y.a ? x.a < y.a : x.b < y.b; });
        size_t cursor = 0;
        for (uint32_t a = 0; a < count; ++a) {
            size_t end = cursor;
            while (end < found.size() && found[end].a == a) ++end;
            PassItem& A = items_[listW[a]];
            if (A.special == 0) {
                const float n = (float)(1u << (2 * A.level));
                float best = std::numeric_limits<float>::max();
                int64_t partner = -1;
                for (size_t q = cursor; q < end; ++q) {
                    const PassItem& B = items_[listW[found[q].b]];
                    if (B.special != 0 || B.prims == 0) continue;
                    const float dist = float(found[q].dist) / n;
                    if (dist < kThreshold && dist < best) { best = dist; partner = (int64_t)listW[found[q].b]; }
                }
                if (partner >= 0) absorb(listW[a], (uint32_t)partner);
            }
            cursor = end;
        }
        return true;
    }

    // ---- a18 ------------------------------------------------------------------------------------------------------------------------------
    // SDK contract (bake_cpu_impl.cpp:1557-1688): while the serialized blocks exceed the budget, the item with the least known coverage lost per
    // byte saved is downsampled one level; the candidates are kept sorted by that figure (std::sort: the order of equal keys is whatever
    // libstdc++'s introsort leaves, reproduced by sorting the same sequence with the same predicate).  The per-level known counts come from
    // the device in one pass, so a step costs a table lookup; the blocks themselves are downsampled on the device afterwards.
    ommResult compress(uint32_t budget) {
        struct Candidate {
            int w;                 // -1 once the item has reached level 0
            float lossPerByte;     // (coverage now - coverage one level down) * covered area / bytes saved
            size_t bytesNow, bytesNext;
        };
        std::vector<uint32_t> listPos;
        std::vector<uint8_t> listLevel;
        std::vector<uint32_t> slotOf(items_.size(), 0xFFFFFFFFu);
        for (uint32_t w = 0; w < items_.size(); ++w) {
            const PassItem& it = items_[w];
            if (it.level == 0 || it.prims == 0 || it.special != 0) continue;
            slotOf[w] = (uint32_t)listPos.size();
            listPos.push_back(it.pos);
            listLevel.push_back(it.level);
        }
        const uint32_t count = (uint32_t)listPos.size();
        if (count == 0) return ommResult_SUCCESS;
        std::vector<uint32_t> known((size_t)count * 13);
        {
            uint32_t *dPos = upload(listPos), *dKnown = nullptr;
            uint8_t* dLevel = upload(listLevel);
            if (!dPos || !dLevel || !alloc(&dKnown, known.size())) return ommResult_FAILURE;
            KnownPyramidKernel<<<(count + 127) / 128, 128, 0, stream_>>>(stateWords_, wordStart_, dPos, dLevel, count, dKnown);
            launches++;
            if (cudaMemcpyAsync(known.data(), dKnown, sizeof(uint32_t) * known.size(), cudaMemcpyDeviceToHost, stream_) != cudaSuccess ||
                cudaStreamSynchronize(stream_) != cudaSuccess)
                return ommResult_FAILURE;
        }
        auto blockBytes = [](uint32_t level) { return std::max<size_t>(1, ((size_t)(1u << (2 * level)) * 2) / 8); };
        auto rate = [&](uint32_t w) {
            const PassItem& it = items_[w];
            const uint32_t* kn = &known[(size_t)slotOf[w] * 13];
            const uint32_t steps = it.levelAtStart - it.level;   // downsampling steps taken so far
            const float ratioNow = (float)kn[steps] / (1u << (2 * it.level));
            const float ratioNext = kn[steps + 1] / (float)(size_t)(1u << (2 * (it.level - 1)));
            float coveredArea = 0;
            for (uint32_t p = 0; p < it.prims; ++p) coveredArea += it.area;   // the SDK adds the area once per primitive (float accumulation)
            Candidate c;
            c.w = (int)w;
            c.bytesNow = blockBytes(it.level);
            c.bytesNext = blockBytes(it.level - 1);
            c.lossPerByte = coveredArea * (ratioNow - ratioNext) / (c.bytesNow - c.bytesNext);
            return c;
        };
        std::vector<Candidate> queue;
        size_t total = 0;
        for (uint32_t w = 0; w < items_.size(); ++w)
            if (slotOf[w] != 0xFFFFFFFFu) {
                queue.push_back(rate(w));
                total += queue.back().bytesNow;
            }
        if (total < budget) return ommResult_SUCCESS;
        const auto cheaper = [](const Candidate& x, const Candidate& y) { return x.lossPerByte < y.lossPerByte; };
        std::sort(queue.begin(), queue.end(), cheaper);
        while (total >= budget && !queue.empty()) {
            // one sweep over the sorted queue; an item stays under the cursor for as long as another step on it is still cheaper than the next entry
            const int entries = (int)queue.size();
            for (int q = 0; q < entries; ++q) {
                PassItem& it = items_[(uint32_t)queue[q].w];
                if (it.level == 0) return ommResult_FAILURE;
                total = total - queue[q].bytesNow + queue[q].bytesNext;
                it.level -= 1;
                if (it.level == 0) {
                    queue[q].w = -1;
                    continue;
                }
                queue[q] = rate((uint32_t)queue[q].w);
                if (total < budget) break;
                if (q + 1 != entries && queue[q].lossPerByte < queue[q + 1].lossPerByte) --q;
            }
            // exhausted entries leave by swap-with-last, then the order is restored
            for (int q = 0; q < (int)queue.size(); ++q)
                if (queue[q].w == -1) {
                    std::swap(queue[q], queue[queue.size() - 1]);
                    queue.pop_back();
                    --q;
                }
            std::sort(queue.begin(), queue.end(), cheaper);
        }
        return downsampleBlocks();
    }

  private:
    cudaStream_t stream_;
    std::vector<PassItem>& items_;
    uint32_t* stateWords_;
    const unsigned long long* wordStart_;
    unsigned long long totalWords_;
    std::vector<uint32_t> dirty_;   // per item: epoch in which it last absorbed another one (its block changed)
    std::vector<uint32_t> stamp_;   // per item: membership mark of the candidate set being gathered
    std::vector<void*> owned_;

    template <class T>
    bool alloc(T** p, size_t count) {
        void* q = nullptr;
        if (cudaMallocAsync(&q, (count ? count : 1) * sizeof(T), stream_) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        owned_.push_back(q);
        *p = (T*)q;
        return true;
    }
    template <class T>
    T* upload(const std::vector<T>& v) {
        T* d = nullptr;
        if (!alloc(&d, v.size())) return nullptr;
        // the source vector may die before the copy is consumed: pageable copies are staged synchronously by the runtime, which is what is wanted here
        if (!v.empty() && cudaMemcpyAsync(d, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice, stream_) != cudaSuccess) return nullptr;
        return d;
    }

    void absorb(uint32_t to, uint32_t from) {
        PassItem& T = items_[to];
        PassItem& F = items_[from];
        T.prims += F.prims;
        F.prims = 0;
        F.special = -1;
        F.root = to;
        const uint32_t fields = 1u << (2 * F.level), words = fields >= 16 ? fields >> 4 : 1;
        MergeStatesKernel<<<std::max(1u, std::min(64u, (words + 255) / 256)), 256, 0, stream_>>>(stateWords_, wordStart_, T.pos, F.pos, fields);
        launches++;
        dirty_[to] = epoch_;
    }

    // distances of (a, b) pairs given by first-seen indices, as the blocks are after every merge issued so far
    bool distances(const std::vector<uint2>& pairsW, uint32_t fields, std::vector<uint32_t>& out) {
        out.resize(pairsW.size());
        if (pairsW.empty()) return true;
        std::vector<uint2> pairsPos(pairsW.size());
        std::vector<uint32_t> nf(pairsW.size(), fields);
        for (size_t i = 0; i < pairsW.size(); ++i) pairsPos[i] = make_uint2(items_[pairsW[i].x].pos, items_[pairsW[i].y].pos);
        uint2* dPairs = upload(pairsPos);
        uint32_t *dFields = upload(nf), *dDist = nullptr;
        if (!dPairs || !dFields || !alloc(&dDist, pairsW.size())) return false;
        PairDistanceKernel<<<(uint32_t)((pairsW.size() + 7) / 8), 256, 0, stream_>>>(stateWords_, wordStart_, dPairs, dFields, (uint32_t)pairsW.size(), dDist);
        launches++;
        return cudaMemcpyAsync(out.data(), dDist, sizeof(uint32_t) * out.size(), cudaMemcpyDeviceToHost, stream_) == cudaSuccess &&
               cudaStreamSynchronize(stream_) == cudaSuccess;
    }

    // ---- one (pass, level) round of the LSH merge ----
    uint32_t epoch_ = 0;
    uint32_t stampValue_ = 0;

    struct Buckets {                // one LSH table: batch slots grouped by layer hash, members in batch order
        std::vector<uint32_t> members;        // batch slots, grouped
        std::vector<uint32_t> begin, end;     // per batch slot: its group
    };
    static Buckets groupByHash(const uint64_t* hash, uint32_t n) {
        Buckets b;
        b.members.resize(n);
        b.begin.resize(n);
        b.end.resize(n);
        for (uint32_t i = 0; i < n; ++i) b.members[i] = i;
        std::stable_sort(b.members.begin(), b.members.end(), [hash](uint32_t x, uint32_t y) { return hash[x] < hash[y]; });
        for (uint32_t g0 = 0; g0 < n;) {
            uint32_t g1 = g0 + 1;
            while (g1 < n && hash[b.members[g1]] == hash[b.members[g0]]) ++g1;
            for (uint32_t q = g0; q < g1; ++q) b.begin[b.members[q]] = g0, b.end[b.members[q]] = g1;
            g0 = g1;
        }
        return b;
    }
    // the candidate set of batch slot i under the current special indices, ascending first-seen index (the order the SDK's std::set yields)
    void gather(const std::vector<uint32_t>& batch, const std::vector<Buckets>& tables, uint32_t i, uint32_t L, std::vector<uint32_t>& out) {
        out.clear();
        ++stampValue_;
        for (const Buckets& t : tables)
            for (uint32_t q = t.begin[i]; q < t.end[i]; ++q) {
                const uint32_t w = batch[t.members[q]];
                if (t.members[q] == i || items_[w].special != 0) continue;
                if (out.size() > 3 * (size_t)L) break;   // the SDK stops reading THIS bucket; the next table is still opened (and left at once)
                if (stamp_[w] != stampValue_) {
                    stamp_[w] = stampValue_;
                    out.push_back(w);
                }
            }
        std::sort(out.begin(), out.end());
    }
    bool lshRound(const std::vector<uint32_t>& batch, uint32_t d, float r, uint32_t L, uint32_t k, const std::vector<uint32_t>& sampled) {
        const uint32_t n = (uint32_t)batch.size();
        // layer hashes of the batch as the blocks are now (before any merge of this round), one table after the other
        std::vector<uint64_t> hash((size_t)n * L);
        {
            std::vector<uint32_t> batchPos(n);
            for (uint32_t i = 0; i < n; ++i) batchPos[i] = items_[batch[i]].pos;
            uint32_t *dBatch = upload(batchPos), *dBits = upload(sampled);
            uint64_t* dHash = nullptr;
            if (!dBatch || !dBits || !alloc(&dHash, hash.size())) return false;
            const unsigned long long threads = (unsigned long long)n * L;
            LshHashKernel<<<(uint32_t)((threads + 127) / 128), 128, 0, stream_>>>(stateWords_, wordStart_, dBatch, n, dBits, L, k, dHash);
            launches++;
            if (cudaMemcpyAsync(hash.data(), dHash, sizeof(uint64_t) * hash.size(), cudaMemcpyDeviceToHost, stream_) != cudaSuccess ||
                cudaStreamSynchronize(stream_) != cudaSuccess)
                return false;
        }
        std::vector<Buckets> tables;
        tables.reserve(L);
        for (uint32_t t = 0; t < L; ++t) tables.push_back(groupByHash(hash.data() + (size_t)t * n, n));

        // The walk, in windows: the candidate sets of a window's items are gathered as things stand at its start and their distances
        // computed in one launch; an item whose set has meanwhile gained a member, or holds one that absorbed something inside the window,
        // gets the missing distances on demand.
        constexpr uint32_t kWindow = 1024;
        std::vector<uint32_t> cand;
        std::vector<uint2> pairs;
        std::vector<uint32_t> dist, firstPair(kWindow + 1);
        for (uint32_t w0 = 0; w0 < n; w0 += kWindow) {
            const uint32_t w1 = std::min(n, w0 + kWindow);
            ++epoch_;
            pairs.clear();
            for (uint32_t i = w0; i < w1; ++i) {
                firstPair[i - w0] = (uint32_t)pairs.size();
                if (items_[batch[i]].special != 0) continue;
                gather(batch, tables, i, L, cand);
                for (uint32_t c : cand) pairs.push_back(make_uint2(batch[i], c));
            }
            firstPair[w1 - w0] = (uint32_t)pairs.size();
            if (!distances(pairs, d, dist)) return false;
            for (uint32_t i = w0; i < w1; ++i) {
                const uint32_t self = batch[i];
                if (items_[self].special != 0) continue;   // absorbed earlier in this round
                gather(batch, tables, i, L, cand);
                // distances known from the window's launch, unless the partner's block changed since
                std::vector<uint2> missing;
                std::vector<uint32_t> value(cand.size(), 0xFFFFFFFFu);
                for (size_t q = 0; q < cand.size(); ++q) {
                    bool have = false;
                    if (dirty_[cand[q]] != epoch_)
                        for (uint32_t p = firstPair[i - w0]; p < firstPair[i - w0 + 1]; ++p)
                            if (pairs[p].y == cand[q]) { value[q] = dist[p]; have = true; break; }
                    if (!have) missing.push_back(make_uint2(self, cand[q]));
                }
                if (!missing.empty()) {
                    std::vector<uint32_t> fresh;
                    if (!distances(missing, d, fresh)) return false;
                    size_t m = 0;
                    for (size_t q = 0; q < cand.size(); ++q)
                        if (value[q] == 0xFFFFFFFFu) value[q] = fresh[m++];
                }
                float best = std::numeric_limits<float>::max();
                int64_t partner = -1;
                for (size_t q = 0; q < cand.size(); ++q) {
                    const float dq = float(value[q]);
                    if (dq < r && dq < best) { best = dq; partner = (int64_t)cand[q]; }
                }
                if (partner >= 0) absorb(self, (uint32_t)partner);
            }
        }
        return true;
    }

    // the blocks of the items whose level the budget pass lowered, one level-synchronous step at a time
    ommResult downsampleBlocks() {
        uint32_t* scratch = nullptr;
        bool any = false;
        for (const PassItem& it : items_) any = any || it.level != it.levelAtStart;
        if (!any) return ommResult_SUCCESS;
        if (!alloc(&scratch, (size_t)totalWords_ + 4)) return ommResult_FAILURE;
        for (uint32_t step = 0; step < 12; ++step) {
            std::vector<uint32_t> listPos;
            std::vector<uint8_t> levelNow;
            uint32_t maxWords = 1;
            for (const PassItem& it : items_)
                if ((uint32_t)(it.levelAtStart - it.level) > step) {
                    listPos.push_back(it.pos);
                    levelNow.push_back((uint8_t)(it.levelAtStart - step));
                    const uint32_t fieldsOut = 1u << (2 * (it.levelAtStart - step - 1));
                    maxWords = std::max(maxWords, fieldsOut >= 16 ? fieldsOut >> 4 : 1u);
                }
            if (listPos.empty()) break;
            uint32_t* dPos = upload(listPos);
            uint8_t* dLevel = upload(levelNow);
            if (!dPos || !dLevel) return ommResult_FAILURE;
            const uint32_t count = (uint32_t)listPos.size();
            for (uint32_t first = 0; first < count; first += 65535) {   // gridDim.y limit
                const uint32_t part = std::min(65535u, count - first);
                const dim3 grid(std::min(256u, (maxWords + 127) / 128), part);
                DownsampleComputeKernel<<<grid, 128, 0, stream_>>>(stateWords_, wordStart_, dPos + first, dLevel + first, part, scratch);
                DownsampleCommitKernel<<<grid, 128, 0, stream_>>>(stateWords_, wordStart_, dPos + first, dLevel + first, part, scratch);
                launches += 2;
            }
        }
        return cudaGetLastError() == cudaSuccess ? ommResult_SUCCESS : ommResult_FAILURE;
    }
};

}  // namespace ommb200
