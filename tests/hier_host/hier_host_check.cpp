// hier_host_check.cpp -- TEST INFRASTRUCTURE.  Host build (g++, -ffp-contract=off) of the classifier arithmetic in
// omm_b200/csrc/omm_device_math.cuh + omm_hier.cuh.  For one work item it runs the hierarchical descent of
// HierClassifyKernel serially and compares every micro-triangle's state with the plain reference walk
// (ClassifyMicroTriangle, itself parity-tested against the SDK build on the GPU).  Used by tests/test_hier_host.py to fuzz
// the exact shortcuts (TestRegion) without a GPU.  Never linked into libomm-b200.so.
#include <cstdint>
#include <cstring>
#include <utility>
#include <vector>

#include "../../omm_b200/csrc/omm_hier.cuh"

using namespace ommb200;

struct HierCheckStats {
    uint64_t microTriangles;
    uint64_t mismatches;
    uint64_t tests[4];   // region tests at size exponent 3,2,1,0
    uint64_t passes[4];
    uint64_t fullEvals;
    uint64_t firstBadItem, firstBadIndex;
    int32_t firstBadGot, firstBadWant;
};

// optional sink for the plain reference walk's states, one byte per micro-triangle in item / bird-curve order
// (tests/test_hier_host.py::test_plain_walk_equals_the_sdk_build compares them with the SDK build's blocks)
static uint8_t* g_wantSink = nullptr;
static uint64_t g_wantPos = 0;
extern "C" __attribute__((visibility("default"))) void hier_host_set_state_sink(uint8_t* sink) { g_wantSink = sink; g_wantPos = 0; }

template <class Cfg>
static void Descend(const BakeParams& P, const DevMip& m, const HierItem* his, uint32_t nodeInItem, uint32_t nl, uint32_t e, uint32_t idx, uint8_t* states,
                    HierCheckStats* st, const float2* uv, bool degenerate, const ItemCellMap* map = nullptr) {
    const HierItem& hi = his[0];
    const int M = P.tex.mipCount;
    const uint32_t L = hi.level;
    if (e == 0) {  // leaves: the reference walk with the exact skips of LeafCell (slow path for items the shortcuts do not cover)
        st->fullEvals++;
        const uint32_t index = (nodeInItem << (2 * nl)) + idx;
#if !defined(OMM_HIER_STATS)  // the statistics build mirrors HierLeaves, which goes straight to the leaf walk
        if (hi.ok && M == 1) {
            st->tests[3]++;
            const int s = TestRegion<Cfg>(P, m, hi, index, L);  // quick single-micro-triangle proof (not on the GPU path: measured slower there; kept here as one more exactness check of TestRegion)
            if (s != 0) {
                st->passes[3]++;
                states[idx] = (uint8_t)(s > 0 ? P.stateGT : P.stateLE);
                return;
            }
        }
#endif
        if (hi.ok) {
            const int state = M == 1 ? LeafClassify<Cfg>(P, m, hi, index) : LeafClassifyMips<Cfg>(P, [&](int k) { return his[k]; }, index);
            states[idx] = (uint8_t)state;
        } else
            states[idx] = (uint8_t)ClassifyMicroTriangle<Cfg>(P, uv[0], uv[1], uv[2], degenerate, index, L);
        return;
    }
    int s = 0;
    st->tests[3 - e]++;
    if (hi.ok) {
        // every mip must pass with the same side (WarpTestRegionsAllMips)
        for (int k = 0; k < M; ++k) {
            RegionBox rb;
            int sk = 0;
            if (MakeRegionBox(P.tex.mips[k], his[k], (nodeInItem << (2 * (nl - e))) + idx, L - e, rb)) {
                if (map && M == 1) sk = LookupCellMap(*map, rb);  // initial regions only, like HierTestInitial
                if (sk == 0) sk = TestRegionBox<Cfg>(P, P.tex.mips[k], his[k], rb);
            }
            if (k == 0) s = sk;
            else if (sk != s) s = 0;
            if (s == 0) break;
        }
    }
    if (s != 0) {
        st->passes[3 - e]++;
        const uint32_t n = 1u << (2 * e);
        for (uint32_t i = 0; i < n; ++i) states[idx * n + i] = (uint8_t)(s > 0 ? P.stateGT : P.stateLE);
        return;
    }
    for (uint32_t k = 0; k < 4; ++k) Descend<Cfg>(P, m, his, nodeInItem, nl, e - 1, idx * 4 + k, states, st, uv, degenerate);
}

template <class Cfg>
static void CheckItems(const BakeParams& P, const float* uvs, const uint8_t* levels, uint32_t numItems, HierCheckStats* st) {
    const DevMip& m = P.tex.mips[0];
    std::vector<uint8_t> states(4096);
    for (uint32_t it = 0; it < numItems; ++it) {
        const float2 uv[3] = {make_float2(uvs[6 * it], uvs[6 * it + 1]), make_float2(uvs[6 * it + 2], uvs[6 * it + 3]), make_float2(uvs[6 * it + 4], uvs[6 * it + 5])};
        const uint32_t L = levels[it];
        const bool degenerate = TriIsDegenerate(uv[0], uv[1], uv[2]);
        std::vector<HierItem> his;
        for (int k = 0; k < P.tex.mipCount; ++k) his.push_back(MakeHierItemFor(P, P.tex.mips[k], uv[0], uv[1], uv[2], L, degenerate));
        const HierItem hi = his[0];
        const uint32_t nl = L < 6 ? L : 6;
        const uint32_t nodes = L > 6 ? 1u << (2 * (L - 6)) : 1u;
        const uint32_t e0 = nl < 3 ? nl : 3;
        for (uint32_t node = 0; node < nodes; ++node) {
            // whole-cell bitmap (F) of the item, or of the aligned node of 64 initial regions of a bigger item, as HierTestInitial builds it
            uint32_t plus[32] = {0}, minus[32] = {0};
            ItemCellMap map{0, 0, 0, 0, plus, minus};
            RegionBox box;
            const bool haveBox = P.tex.mipCount == 1 && hi.ok && L >= 3 && (L > 6 ? MakeNodeBox(m, hi, node, L - 6, box) : MakeItemBox(m, hi, box));
            // piece-level answers of HierTestInitial: constant area (H), then the item-independent tables (I)
            if (haveBox) {
                int sPiece = FlatRectSide<Cfg>(P, m, box.cx0, box.cy0, box.cx1, box.cy1);
                if (sPiece == 0 && ItemWithinStrongCaps(hi)) sPiece = StrongRectSide(P, m, hi, box, box.cx0, box.cy0, box.cx1, box.cy1);
                if (sPiece != 0) {
                    const uint32_t n = 1u << (2 * nl);
                    for (uint32_t i = 0; i < n; ++i) {
                        const int want = ClassifyMicroTriangle<Cfg>(P, uv[0], uv[1], uv[2], degenerate, (node << (2 * nl)) + i, L);
                        if (g_wantSink) g_wantSink[g_wantPos++] = (uint8_t)want;
                        st->microTriangles++;
                        if (want != (sPiece > 0 ? P.stateGT : P.stateLE)) {
                            if (st->mismatches == 0) { st->firstBadItem = it; st->firstBadIndex = (node << (2 * nl)) + i; st->firstBadGot = sPiece; st->firstBadWant = want; }
                            st->mismatches++;
                        }
                    }
                    st->passes[0] += 1u << (2 * (nl - e0));
                    st->tests[0] += 1u << (2 * (nl - e0));
                    continue;
                }
            }
            const bool strongCovers = ItemWithinStrongCaps(hi) && P.tex.strongPlus != nullptr && haveBox && box.cx0 >= 0 && box.cy0 >= 0 &&
                                      box.cx1 <= m.w - 2 && box.cy1 <= m.h - 2 && box.hix - box.lox + 1.f + hi.deltaEdge <= kStrongMaxExtent &&
                                      box.hiy - box.loy + 1.f + hi.deltaEdge <= kStrongMaxExtent;
            if (haveBox && !strongCovers && box.cx1 - box.cx0 < 32 && box.cy1 - box.cy0 < 32) {
                map.cx0 = box.cx0; map.cy0 = box.cy0; map.fw = box.cx1 - box.cx0 + 1; map.fh = box.cy1 - box.cy0 + 1;
                for (int y = 0; y < map.fh; ++y)
                    for (int x = 0; x < map.fw; ++x) {
                        const int s = WholeCellSide<Cfg>(P, m, hi, box, box.cx0 + x, box.cy0 + y);
                        if (s > 0) plus[y] |= 1u << x;
                        else if (s < 0) minus[y] |= 1u << x;
                    }
            }
            const uint32_t nInit = 1u << (2 * (nl - e0));
            for (uint32_t r = 0; r < nInit; ++r) Descend<Cfg>(P, m, his.data(), node, nl, e0, r, states.data(), st, uv, degenerate, e0 == 3 ? &map : nullptr);
            const uint32_t n = 1u << (2 * nl);
            for (uint32_t i = 0; i < n; ++i) {
                const uint32_t index = (node << (2 * nl)) + i;
                const int want = ClassifyMicroTriangle<Cfg>(P, uv[0], uv[1], uv[2], degenerate, index, L);
                if (g_wantSink) g_wantSink[g_wantPos++] = (uint8_t)want;
                st->microTriangles++;
                if (want != (int)states[i]) {
                    if (st->mismatches == 0) {
                        st->firstBadItem = it;
                        st->firstBadIndex = index;
                        st->firstBadGot = states[i];
                        st->firstBadWant = want;
                    }
                    st->mismatches++;
                }
            }
        }
    }
}

extern "C" __attribute__((visibility("default"))) int hier_host_check(const void* texels, int isFp32, int w, int h, int addrMode, float borderAlpha,
                                                                      float cutoff, int stateGT, int stateLE, int format, int promotion, const float* uvs,
                                                                      const uint8_t* levels, uint32_t numItems, HierCheckStats* st, int useSat, int numMips) {
    memset(st, 0, sizeof(*st));
    BakeParams P{};
    P.tex.texels = texels;
    P.tex.sat = nullptr;
    P.tex.isFp32 = isFp32;
    P.tex.mipCount = 1;
    // optional mip chain (2 x 2 box filter), texels of all mips back to back like the library stores them
    std::vector<uint8_t> chain;
    if (numMips > 1) {
        const size_t spp = isFp32 ? 4 : 1;
        std::vector<int> ws{w}, hs{h};
        std::vector<size_t> offs{0};
        size_t total = (size_t)w * h;
        for (int k = 1; k < numMips; ++k) {
            ws.push_back(std::max(1, ws.back() / 2)); hs.push_back(std::max(1, hs.back() / 2));
            offs.push_back(total);
            total += (size_t)ws.back() * hs.back();
        }
        chain.resize(total * spp);
        memcpy(chain.data(), texels, (size_t)w * h * spp);
        for (int k = 1; k < numMips; ++k)
            for (int y = 0; y < hs[k]; ++y)
                for (int x = 0; x < ws[k]; ++x) {
                    auto src = [&](int xx, int yy) { return offs[k - 1] + (size_t)std::min(yy, hs[k - 1] - 1) * ws[k - 1] + std::min(xx, ws[k - 1] - 1); };
                    const size_t dsti = offs[k] + (size_t)y * ws[k] + x;
                    if (isFp32) {
                        const float* f = (const float*)chain.data();
                        ((float*)chain.data())[dsti] = 0.25f * (f[src(2 * x, 2 * y)] + f[src(2 * x + 1, 2 * y)] + f[src(2 * x, 2 * y + 1)] + f[src(2 * x + 1, 2 * y + 1)]);
                    } else {
                        const uint8_t* b8 = chain.data();
                        chain[dsti] = (uint8_t)((b8[src(2 * x, 2 * y)] + b8[src(2 * x + 1, 2 * y)] + b8[src(2 * x, 2 * y + 1)] + b8[src(2 * x + 1, 2 * y + 1)]) / 4);
                    }
                }
        P.tex.texels = chain.data();
        P.tex.mipCount = numMips;
        for (int k = 1; k < numMips; ++k) {
            DevMip& mk = P.tex.mips[k];
            mk.w = ws[k]; mk.h = hs[k];
            mk.log2w = __builtin_ctz((unsigned)mk.w); mk.log2h = __builtin_ctz((unsigned)mk.h);
            mk.isPow2 = (mk.w & (mk.w - 1)) == 0 && (mk.h & (mk.h - 1)) == 0;
            mk.rcpw = 1.f / (float)mk.w; mk.rcph = 1.f / (float)mk.h;
            mk.texelOffset = offs[k]; mk.satOffset = offs[k];
        }
    }
    DevMip& m = P.tex.mips[0];
    m.w = w; m.h = h;
    m.log2w = __builtin_ctz((unsigned)w); m.log2h = __builtin_ctz((unsigned)h);
    m.isPow2 = (w & (w - 1)) == 0 && (h & (h - 1)) == 0;
    m.rcpw = 1.f / (float)w; m.rcph = 1.f / (float)h;
    m.texelOffset = 0; m.satOffset = 0;
    P.addrMode = addrMode;
    P.filterLinear = 1;
    P.borderAlpha = borderAlpha;
    P.cutoff = cutoff;
    P.stateGT = stateGT; P.stateLE = stateLE;
    P.globalFormat = format;
    P.promotion = promotion;
    P.pow2Mip0 = m.isPow2;
    // optional SAT pass with the texture's cutoff == the bake's (the only SAT configuration the hierarchical path takes)
    std::vector<uint32_t> sat;
    if (useSat) {
        sat.resize((size_t)w * h);
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                const size_t i = (size_t)y * w + x;
                const float a = isFp32 ? ((const float*)texels)[i] : (float)((const uint8_t*)texels)[i] * (1.f / 255.f);
                sat[i] = (a > cutoff ? 1u : 0u) + (x ? sat[i - 1] : 0u) + (y ? sat[i - w] : 0u) - ((x && y) ? sat[i - w - 1] : 0u);
            }
        P.tex.sat = sat.data();
        P.useCoarse = 1;
        P.coarseSameCutoff = 1;
    }
    // (H) constant-cell table of mip 0, as BuildFlatSat / FlatSatRows + SatCols compute it on the device
    std::vector<uint32_t> flat;
    if (w >= 2 && h >= 2 && numMips <= 1) {
        flat.resize((size_t)(w - 1) * (h - 1));
        for (int y = 0; y < h - 1; ++y)
            for (int x = 0; x < w - 1; ++x) {
                auto tx = [&](int xx, int yy) {
                    const size_t i = (size_t)yy * w + xx;
                    return isFp32 ? ((const float*)texels)[i] : (float)((const uint8_t*)texels)[i] * (1.f / 255.f);
                };
                const uint32_t bad = CellIsFlatGood(tx(x, y), tx(x + 1, y), tx(x, y + 1), tx(x + 1, y + 1), cutoff) ? 0u : 1u;
                const size_t i = (size_t)y * (w - 1) + x;
                flat[i] = bad + (x ? flat[i - 1] : 0u) + (y ? flat[i - (w - 1)] : 0u) - ((x && y) ? flat[i - (w - 1) - 1] : 0u);
            }
        P.tex.flatSat = flat.data();
    }
    // (I) item-independent whole-cell tables, as StrongSatRows + the column pass compute them
    std::vector<uint32_t> strongP, strongM;
    if (w >= 2 && h >= 2 && numMips <= 1) {
        strongP.resize((size_t)(w - 1) * (h - 1));
        strongM.resize((size_t)(w - 1) * (h - 1));
        for (int y = 0; y < h - 1; ++y)
            for (int x = 0; x < w - 1; ++x) {
                const int s = isFp32 ? StrongCellSide<KernelCfg<kAddrClamp, true>>(P, P.tex.mips[0], x, y) : StrongCellSide<KernelCfg<kAddrClamp, false>>(P, P.tex.mips[0], x, y);
                const size_t i = (size_t)y * (w - 1) + x;
                auto acc = [&](std::vector<uint32_t>& v, uint32_t bad) {
                    v[i] = bad + (x ? v[i - 1] : 0u) + (y ? v[i - (w - 1)] : 0u) - ((x && y) ? v[i - (w - 1) - 1] : 0u);
                };
                acc(strongP, s > 0 ? 0u : 1u);
                acc(strongM, s < 0 ? 0u : 1u);
            }
        P.tex.strongPlus = strongP.data();
        P.tex.strongMinus = strongM.data();
    }
    if (isFp32) CheckItems<KernelCfg<kAddrGeneric, true>>(P, uvs, levels, numItems, st);
    else CheckItems<KernelCfg<kAddrGeneric, false>>(P, uvs, levels, numItems, st);
    return st->mismatches == 0 ? 0 : 1;
}

#if defined(OMM_HIER_STATS)
extern "C" __attribute__((visibility("default"))) void hier_host_stats(unsigned long long* out) {
    for (int i = 0; i < 16; ++i) { out[i] = g_hierStats[i]; g_hierStats[i] = 0; }
}
#endif
