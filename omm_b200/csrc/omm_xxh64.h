// omm_xxh64.h -- XXH64 after the published specification (xxHash doc/xxhash_spec.md; the SDK vendors xxHash at submodule c961fbe6 and
// calls XXH64 at bake_cpu_impl.cpp:1039, 1254 and serialize_impl.cpp:272).  One byte-stream implementation for the host (blob digests) and
// the device (LSH layer hashes of the near-duplicate pass); the per-item block digests of the exact dedup have their own specialised kernel.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define OMM_XXH_HD __host__ __device__ inline
#else
#define OMM_XXH_HD inline
#endif

namespace ommb200 {
namespace xxh {
constexpr uint64_t kP1 = 0x9E3779B185EBCA87ull, kP2 = 0xC2B2AE3D27D4EB4Full, kP3 = 0x165667B19E3779F9ull, kP4 = 0x85EBCA77C2B2AE63ull, kP5 = 0x27D4EB2F165667C5ull;
OMM_XXH_HD uint64_t Rotl(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
OMM_XXH_HD uint64_t Round(uint64_t acc, uint64_t in) { return Rotl(acc + in * kP2, 31) * kP1; }
OMM_XXH_HD uint64_t MergeRound(uint64_t acc, uint64_t v) { return (acc ^ Round(0, v)) * kP1 + kP4; }
OMM_XXH_HD uint64_t Avalanche(uint64_t h) {
    h ^= h >> 33; h *= kP2; h ^= h >> 29; h *= kP3; h ^= h >> 32;
    return h;
}
// The accumulator chain of a long message,  acc <- Round(acc, in),  restated on  s = acc + in * P2  (the state one addition further):
//     s' = Rotl(s, 31) * P1 + x',   x' = in' * P2,
// so that a consumer which receives the pre-multiplied x' has one multiply-add per stripe on its dependent path.  ChainStep is that
// step with both halves of the rotation as one funnel shift each and the 64-bit product split by hand (the wide product of the low words
// takes x as its addend, the two cross products go into the high word): six machine instructions on sm_100a.  ChainStart(acc0) is the
// state whose step with the first x gives acc0 + x (P1 is odd, so the step is invertible): no special case for the first stripe;
// ChainEnd(s) = Rotl(s, 31) * P1 is the accumulator after the last one.  Used by the big-block digest kernel (omm_bake.cu).
OMM_XXH_HD uint32_t FunnelShiftR(uint32_t lo, uint32_t hi, uint32_t shift) {  // low word of (hi : lo) >> shift, shift < 32
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, shift);
#else
    return (uint32_t)((((uint64_t)hi << 32) | lo) >> (shift & 31u));
#endif
}
OMM_XXH_HD uint64_t ChainStep(uint64_t s, uint64_t x) {
    const uint32_t slo = (uint32_t)s, shi = (uint32_t)(s >> 32);
    const uint32_t rlo = FunnelShiftR(shi, slo, 1), rhi = FunnelShiftR(slo, shi, 1);  // the two halves of Rotl(s, 31)
    const uint64_t w = (uint64_t)rlo * (uint32_t)kP1 + x;
    const uint32_t hi = (uint32_t)(w >> 32) + rhi * (uint32_t)kP1 + rlo * (uint32_t)(kP1 >> 32);
    return ((uint64_t)hi << 32) | (uint32_t)w;
}
constexpr uint64_t MulInverse64(uint64_t a) {  // a odd; Newton's iteration doubles the correct low bits (3 to begin with)
    uint64_t x = a;
    for (int i = 0; i < 6; ++i) x *= 2ull - a * x;
    return x;
}
constexpr uint64_t kP1Inverse = MulInverse64(kP1);
static_assert(kP1 * kP1Inverse == 1ull, "modular inverse of PRIME64_1");
OMM_XXH_HD uint64_t ChainStart(uint64_t acc0) {
    const uint64_t pre = acc0 * kP1Inverse;
    return (pre >> 31) | (pre << 33);  // Rotl(result, 31) * P1 == acc0
}
OMM_XXH_HD uint64_t ChainEnd(uint64_t s) { return Rotl(s, 31) * kP1; }
// Streaming form over 32-bit little-endian words (all the library needs on the device: the hashed message is an array of uint32 samples).
struct WordStream {
    uint64_t v1, v2, v3, v4, seed;
    uint64_t pendingLo;   // first word of an incomplete 8-byte lane
    uint64_t lane[4];     // the current 32-byte stripe, filled lane by lane
    uint32_t words;       // words consumed so far
    OMM_XXH_HD explicit WordStream(uint64_t s) : v1(s + kP1 + kP2), v2(s + kP2), v3(s), v4(s - kP1), seed(s), pendingLo(0), lane{0, 0, 0, 0}, words(0) {}
    OMM_XXH_HD void push(uint32_t w) {
        if ((words & 1u) == 0) pendingLo = w;
        else {
            const uint32_t l = (words >> 1) & 3u;
            lane[l] = pendingLo | ((uint64_t)w << 32);
            if (l == 3) {
                v1 = Round(v1, lane[0]); v2 = Round(v2, lane[1]); v3 = Round(v3, lane[2]); v4 = Round(v4, lane[3]);
            }
        }
        ++words;
    }
    OMM_XXH_HD uint64_t finish() const {
        const uint64_t len = (uint64_t)words * 4ull;
        uint64_t h;
        if (len >= 32) {
            h = Rotl(v1, 1) + Rotl(v2, 7) + Rotl(v3, 12) + Rotl(v4, 18);
            h = MergeRound(h, v1); h = MergeRound(h, v2); h = MergeRound(h, v3); h = MergeRound(h, v4);
        } else
            h = seed + kP5;
        h += len;
        // complete 8-byte lanes of the unfinished stripe, then a lone 4-byte word
        const uint32_t tailWords = words & 7u, tailLanes = tailWords >> 1;
        for (uint32_t l = 0; l < tailLanes; ++l) {
            h ^= Round(0, lane[l]);
            h = Rotl(h, 27) * kP1 + kP4;
        }
        if (tailWords & 1u) {
            h ^= (pendingLo & 0xFFFFFFFFull) * kP1;
            h = Rotl(h, 23) * kP2 + kP3;
        }
        return Avalanche(h);
    }
};
}  // namespace xxh

// XXH64 of an arbitrary byte buffer (host side: digests of serialized blobs)
inline uint64_t HostXxh64(const void* data, size_t len, uint64_t seed) {
    using namespace xxh;
    const uint8_t* p = (const uint8_t*)data;
    const uint8_t* const end = p + len;
    auto rd64 = [](const uint8_t* q) { uint64_t v; memcpy(&v, q, 8); return v; };
    auto rd32 = [](const uint8_t* q) { uint32_t v; memcpy(&v, q, 4); return v; };
    uint64_t h;
    if (len >= 32) {
        uint64_t v1 = seed + kP1 + kP2, v2 = seed + kP2, v3 = seed, v4 = seed - kP1;
        for (; p + 32 <= end; p += 32) {
            v1 = Round(v1, rd64(p)); v2 = Round(v2, rd64(p + 8)); v3 = Round(v3, rd64(p + 16)); v4 = Round(v4, rd64(p + 24));
        }
        h = Rotl(v1, 1) + Rotl(v2, 7) + Rotl(v3, 12) + Rotl(v4, 18);
        h = MergeRound(h, v1); h = MergeRound(h, v2); h = MergeRound(h, v3); h = MergeRound(h, v4);
    } else
        h = seed + kP5;
    h += (uint64_t)len;
    for (; p + 8 <= end; p += 8) { h ^= Round(0, rd64(p)); h = Rotl(h, 27) * kP1 + kP4; }
    if (p + 4 <= end) { h ^= (uint64_t)rd32(p) * kP1; h = Rotl(h, 23) * kP2 + kP3; p += 4; }
    for (; p < end; ++p) { h ^= (uint64_t)(*p) * kP5; h = Rotl(h, 11) * kP1; }
    return Avalanche(h);
}
}  // namespace ommb200
