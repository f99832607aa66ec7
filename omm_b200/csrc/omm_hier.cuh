// omm_hier.cuh -- exact hierarchical shortcuts for the Linear / level-line classifier (default configuration).
//
// The reference decides the state of a micro-triangle t from three kinds of evidence (ref: bake_cpu_impl.cpp:861-908,
// bake_kernels_cpu.h:241-399):   (1) the bilinear sample at t.p0,   (2) the texel centres that lie inside t,
// (3) intersections of t's edges with the level line  alpha = cutoff  inside every bilinear cell the conservative raster
// visits.  A REGION is a sub-triangle of the bird curve (4^e consecutive micro-triangles, e = 0 is one micro-triangle).
// TestRegion() below proves, for a whole region at once, that all three kinds of evidence can only ever vote for ONE side
// s of the cutoff; then every micro-triangle of the region ends with coverage (s-counter > 0, other counter == 0) and its
// state is stateGT / stateLE whatever the promotion mode.  When the proof does not go through the region is split, and
// single micro-triangles that still fail are evaluated by the reference walk (ClassifyMicroTriangle).  The shortcut never
// changes a result: it only skips arithmetic whose outcome is known.
//
// Notation.  u = 2^-24 (unit round-off), W x H = texture size.  "r-space" = texel space after the raster offset:
// r(p) = fl(fl(p.x*W) - 0.5) -- the very expression of the rasterizer (cpu_raster.h:287-291) and of the bilinear sampler
// (texture_impl.cpp:263).  Cell (cx, cy) has its corners on the texel centres; cell-local coordinates are q = r - (cx, cy)
// (the kernel computes fl(fl(W*p.x) - (cx + 0.5)), bake_kernels_cpu.h:356-358), so the cell is [0,1]^2 and the bilinear
// patch is  h(x,y) = a' + b x + c y + d x y  with the float coefficients of bake_kernels_cpu.h:330-340 (a' = a - cutoff).
//
// (A) Enclosure.  Every lattice vertex of the region is a convex combination of the region's three corner vertices in real
//     arithmetic; the float evaluation of InterpolateTriangleUV (3 products, 2 sums) is off by <= 3.01 u Pmax, the products
//     with W and the offsets add <= 2 u (W Pmax + 1).  With eps = 16 u (max(W,H) Pmax + 2) every vertex of every
//     micro-triangle of the region has r within [lo, hi] = [min corner r - eps, max corner r + eps] (for a single
//     micro-triangle the corners ARE the vertices and eps = 4 u (...) only covers r-versus-q rounding).
//     Footprint F = cells floor(lo) .. floor(hi): a superset of the raster's cells [floor(min r), ceil(max r) - 1] and of the
//     cell floor(r(p0)) of the bilinear sample.
// (B) Edge test.  TestEdgeHyperbolaIntersection returns true only with a point (x^, y^) that passed InUnitSquare and
//     PointOnEdge.  PointOnEdge bounds |p-p0| + |p-p1| by L + 1e-5 (+ rounding), an ellipse that stays within
//     delta = sqrt(e(2L+e))/2 + e, e = 1.001e-5 + 16 u L, of the segment; hence (x^, y^) lies in
//     B = [lo - c - delta, hi - c + delta] intersected with [0,1]^2.  Backward error analysis of the three branches
//     (vertical / linear / quadratic, bake_kernels_cpu.h:157-236) gives |h(x^, y^)| <= R with
//        R = 4.01e-6 + 16 u T + 32 u (Gmax + |cutoff|) + 2.1 u max_K (c K + d M(K) + b)^2 / (d K),
//        T = |a'| + |b| + (|c| + |d|)(K + M + 1),   K = |k^| (edge slope),  M = |m^| <= Qy + Qx K,  Q = max |q|,
//     where the last term is only present when |d K| can reach 1e-6 (the quadratic branch) and K ranges over the slopes
//     the region's edges can have (three direction classes per work item, see MakeHierItem; the function of K is convex, so
//     the maximum is at an end of the range).  A bilinear function attains its extrema over the rectangle B at B's corners,
//     so  min over B's corners of s h > R  proves that no edge of the region can report an intersection in this cell.
//     The 3e-6 inside the constant also settles the "flat patch" branch (|b|,|c|,|d| < 1e-6 => sign(a') = s), and the
//     32 u (Gmax + |cutoff|) term the difference between h and the sampler's lerp form, so the bilinear sample at p0, whose
//     (wx, wy) lies in the B of its cell, votes s as well.
// (C) Texel centres.  A cell corner whose texel is on side s can only vote s.  A corner on the other side must be provably
//     outside every micro-triangle of the region as PointInTriangle evaluates it: each of its three edge functions is
//     computed with relative error <= 4.1 u of |e||P - A|, the three real values sum to twice the area, and for a point at
//     distance rho from the triangle the most negative one is >= kappa^2 emax rho / 2 (kappa = 2 Area / emax^2: the shortest edge
//     is >= kappa emax and every half-angle has sine >= kappa / 2), while a positive one is at least half as large; so
//     kappa^2 rho / 2 > 2 * 4.1 u (rho + diam), guaranteed by rho >= 5 * 4.1 u / kappa^2 * diam when kappa >= 0.01, makes two of
//     them reliably opposite in sign, and then the function returns false whatever the third one does (first test of
//     geometry.h:107 if they are s and t, the final comparison otherwise).  In r-space: the corner is separated from [lo, hi]
//     by pit = rho W + eps in x or in y.
//
// Everything here compiles for host and device (see omm_device_math.cuh); tests/hier_host_check.cpp fuzzes TestRegion
// against the reference walk on the CPU, the GPU parity suite does the same through the library.
#pragma once

#include "omm_device_math.cuh"

namespace ommb200 {

constexpr float kUnitRoundoff = 5.9604645e-8f;  // 2^-24
#if defined(OMM_HIER_STATS)
static unsigned long long g_hierStats[16];  // host-only instrumentation (tests/hier_host, scripts/leaf_stats.py): [0] leaf cells, [1] vertex values of one
                                            // sign, [2] closed by (D), [3] skipped by (E), [4]-[7] region-test failures, [8] cells that run the edge tests,
                                            // [9] of those with a hit, [10] leaves, [11] leaves crossed by the level line, [12]-[14] why (D) failed
#define OMM_STAT(i) (g_hierStats[i]++)
#else
#define OMM_STAT(i) ((void)0)
#endif
constexpr int kHierMaxCells = 16;               // a region whose footprint is larger is split instead

struct alignas(16) HierItem {
    float2 p0, p1, p2;
    uint32_t level;
    float epsRegion, epsSingle;  // r-space enclosure slack (A)
    float deltaEdge;             // PointOnEdge pad (B)
    float kmin[3], kmax[3];      // slope range of the three edge direction classes (B)
    float pitX, pitY;            // r-space separation that makes PointInTriangle provably false (C); +inf = unavailable
    int ok;                      // shortcuts applicable to this work item at all
    int pad;
};
static_assert(sizeof(HierItem) == 80, "HierItem is loaded with five 16-byte reads");

OMM_HD float UlpOf(float x) {  // x >= 2^-100
    return UintAsFloat((FloatAsUint(x) & 0x7F800000u) - (23u << 23));
}

OMM_HD HierItem MakeHierItem(const DevMip& m, float2 p0, float2 p1, float2 p2, uint32_t level, bool degenerate) {
    const float u = kUnitRoundoff;
    const float inf = UintAsFloat(0x7f800000u);
    HierItem it;
    it.pad = 0;
    it.p0 = p0; it.p1 = p1; it.p2 = p2;
    it.level = level;
    const float W = (float)m.w, H = (float)m.h;
    const float S = W > H ? W : H;
    const float Pmax = fmaxf(fmaxf(fmaxf(fabsf(p0.x), fabsf(p0.y)), fmaxf(fabsf(p1.x), fabsf(p1.y))), fmaxf(fabsf(p2.x), fabsf(p2.y)));
    const float ext = S * Pmax;
    it.ok = !degenerate && (ext < 1048576.f);  // NaN / Inf coordinates fail the comparison
    it.epsSingle = 4.f * u * (ext + 2.f);
    it.epsRegion = 16.f * u * (ext + 2.f);
    const float scale = UintAsFloat((127u - level) << 23);
    const float2 e[3] = {make_float2(p1.x - p0.x, p1.y - p0.y), make_float2(p2.x - p0.x, p2.y - p0.y), make_float2(p2.x - p1.x, p2.y - p1.y)};
    float gmax = 0.f, lmax = 0.f, emax2 = 0.f, emin2 = inf;
    float ax[3], ay[3];
    for (int j = 0; j < 3; ++j) {
        ax[j] = fabsf(W * e[j].x * scale);
        ay[j] = fabsf(H * e[j].y * scale);
        gmax = fmaxf(gmax, fmaxf(ax[j], ay[j]));
        lmax = fmaxf(lmax, sqrtf(ax[j] * ax[j] + ay[j] * ay[j]));
        const float l2 = e[j].x * e[j].x + e[j].y * e[j].y;
        emax2 = fmaxf(emax2, l2);
        emin2 = fminf(emin2, l2);
    }
    const float eta = it.epsRegion * 1.01f + 8.f * u * gmax;  // |computed edge component - ideal| (two vertices, float g)
    const float Lmax = lmax * (1.f + 8.f * u) + 2.f * eta;
    const float epsEff = 1.001e-5f + 16.f * u * Lmax;
    it.deltaEdge = (0.5f * sqrtf(epsEff * (2.f * Lmax + epsEff)) + epsEff) * 1.01f;
    // Floats of magnitude >= 16 are multiples of tau = ulp >= 2^-19 > 1e-6 and their differences with the cell centre and
    // with each other are exact (Sterbenz), so a non-zero k_denum of a steep edge is at least tau.
    const float x0 = W * p0.x, x1 = W * p1.x, x2 = W * p2.x;
    const float xlo = fminf(fminf(x0, x1), x2), xhi = fmaxf(fmaxf(x0, x1), x2);
    const float amin = (xlo > 0.f ? xlo : (xhi < 0.f ? -xhi : 0.f)) - 2.f;
    const float tau = amin >= 16.f ? UlpOf(amin) : 0.f;
    for (int j = 0; j < 3; ++j) {
        const float lowNum = ay[j] - eta;
        it.kmin[j] = (lowNum > 0.f ? lowNum : 0.f) / (ax[j] + eta) * (1.f - 8.f * u);
        if (ax[j] > 2.f * eta) it.kmax[j] = (ay[j] + eta) / (ax[j] - eta) * (1.f + 8.f * u);
        else it.kmax[j] = (ay[j] + eta) / fmaxf(9.9e-7f, tau) * (1.f + 8.f * u);
    }
    // (C): shape factor kappa = 2 Area / emax^2 of the micro-triangles.  They are the base triangle scaled by 2^-level with every
    // vertex moved by at most ev = 3.01 u Pmax (A): the area changes by <= 2 ev emax + 2 ev^2, the longest edge by <= 2 ev, so with
    // rr = ev / emax:  kappa_t >= (kappa - 4 rr - 4 rr^2) / (1 + 2 rr)^2.
    it.pitX = it.pitY = inf;
    const float area2 = fabsf(e[0].x * e[1].y - e[0].y * e[1].x);
    const float kappa = area2 / emax2;
    const float emaxT = sqrtf(emax2) * scale;
    const float rr = (3.02f * u * Pmax) / emaxT;
    const float kEff = (kappa * (1.f - 8.f * u) - 4.f * rr - 4.f * rr * rr) / ((1.f + 2.f * rr) * (1.f + 2.f * rr)) * (1.f - 8.f * u);
    if (kEff >= 0.01f) {
        const float rho0 = (5.f * 4.1f * u / (kEff * kEff)) * emaxT * (1.f + 2.f * rr) * 1.01f;
        it.pitX = rho0 * W + it.epsRegion;
        it.pitY = rho0 * H + it.epsRegion;
    }
    if (!(kappa >= 0.f)) it.ok = 0;  // NaN guard
    return it;
}

// Upper bound test of (B): returns true when `margin` (= min over B's corners of s*h, as computed in float) provably
// exceeds R for every edge slope the work item can have.  al..de = |a'|,|b|,|c|,|d|; qx,qy = max |cell-local coordinate|.
OMM_HD bool MarginBeatsEdgeBound(const HierItem& it, float margin, float al, float be, float ga, float de, float gmaxAbs, float cutoffAbs, float qx,
                                 float qy) {
    const float u = kUnitRoundoff;
    float kAll = fmaxf(fmaxf(it.kmax[0], it.kmax[1]), it.kmax[2]);
    const float mAll = (qy + qx * kAll) * (1.f + 4.f * u);
    const float T = al + be + (ga + de) * (kAll + mAll + 1.f);
    const float lin = 4.01e-6f + 16.f * u * T + 32.f * u * (gmaxAbs + cutoffAbs);
    const float rem = margin - lin;
    if (!(rem > 0.f)) return false;
    // quadratic branch: needs |d k| >= 1e-6.  2.1 u (ga K + de M(K) + be)^2 <= rem * de * K (1 - 4u)  at both ends of the range
    const float k0 = 9.9e-7f / (de * (1.f + 4.f * u));  // +inf when d == 0: no quadratic branch
    {
        // first the hull of the three slope ranges: the bound is convex in K, so passing at the two ends of the hull implies passing
        // for every class (the common case; three times cheaper than the per-class test below)
        const float klo = fmaxf(fminf(fminf(it.kmin[0], it.kmin[1]), it.kmin[2]), k0), khi = kAll;
        if (!(klo <= khi)) return true;  // no slope can reach the quadratic branch
        const float c1lo = (ga * klo + de * (qy + qx * klo) + be) * (1.f + 8.f * u);
        const float c1hi = (ga * khi + de * (qy + qx * khi) + be) * (1.f + 8.f * u);
        if ((2.1f * u * c1lo * c1lo < rem * (de * klo * (1.f - 8.f * u))) && (2.1f * u * c1hi * c1hi < rem * (de * khi * (1.f - 8.f * u)))) return true;
    }
    bool ok = true;
    for (int j = 0; j < 3; ++j) {
        const float klo = fmaxf(it.kmin[j], k0), khi = it.kmax[j];
        if (!(klo <= khi)) continue;
        const float c1lo = (ga * klo + de * (qy + qx * klo) + be) * (1.f + 8.f * u);
        const float c1hi = (ga * khi + de * (qy + qx * khi) + be) * (1.f + 8.f * u);
        ok = ok && (2.1f * u * c1lo * c1lo < rem * (de * klo * (1.f - 8.f * u)));
        ok = ok && (2.1f * u * c1hi * c1hi < rem * (de * khi * (1.f - 8.f * u)));
    }
    return ok;
}

// r-space enclosure [lo, hi] of a region (A).  Returns false when the coordinates are not moderate finite numbers.
struct RegionBox {
    float r0x, r0y, r1x, r1y, r2x, r2y;  // r-space corners of the region triangle
    float eps;                           // the enclosure slack that was applied
    float lox, loy, hix, hiy;
    int cx0, cy0, cx1, cy1;  // footprint cells
};
OMM_HD bool MakeRegionBox(const DevMip& m, const HierItem& it, uint32_t index, uint32_t regionLevel, RegionBox& rb) {
    const Tri rt = MicroTri(it.p0, it.p1, it.p2, index, regionLevel);
    const float W = (float)m.w, H = (float)m.h;
    const float eps = regionLevel == it.level ? it.epsSingle : it.epsRegion;
    const float r0x = rt.p0.x * W + -0.5f, r0y = rt.p0.y * H + -0.5f;
    const float r1x = rt.p1.x * W + -0.5f, r1y = rt.p1.y * H + -0.5f;
    const float r2x = rt.p2.x * W + -0.5f, r2y = rt.p2.y * H + -0.5f;
    rb.r0x = r0x; rb.r0y = r0y; rb.r1x = r1x; rb.r1y = r1y; rb.r2x = r2x; rb.r2y = r2y;
    rb.eps = eps;
    rb.lox = fminf(fminf(r0x, r1x), r2x) - eps; rb.loy = fminf(fminf(r0y, r1y), r2y) - eps;
    rb.hix = fmaxf(fmaxf(r0x, r1x), r2x) + eps; rb.hiy = fmaxf(fmaxf(r0y, r1y), r2y) + eps;
    if (!(rb.lox > -2097152.f && rb.loy > -2097152.f && rb.hix < 2097152.f && rb.hiy < 2097152.f)) return false;
    rb.cx0 = (int)floorf(rb.lox); rb.cy0 = (int)floorf(rb.loy); rb.cx1 = (int)floorf(rb.hix); rb.cy1 = (int)floorf(rb.hiy);
    return true;
}

// (F) Whole-cell test for a work item: +1 / -1 when the cell passes (B) and (C) with B = the whole cell [0,1]^2, all four
//     texels on side s, and Q = the largest cell-local coordinate any vertex of the ITEM can have (`box` is the item's
//     region box widened by one more eps, so it contains the region box of every sub-region).  Every sub-region's B is a
//     subset of the cell and its Q is smaller, and R grows with Q, so such a cell passes the test of EVERY region of the item
//     with the same s: TestRegion may take the answer from a per-item bitmap of these cells (HierTestInitial) instead of
//     evaluating it.
template <class Cfg>
OMM_HD int WholeCellSideQ(const BakeParams& P, const DevMip& m, const HierItem& it, float qx, float qy, int cx, int cy) {
    const int x0 = Addr1<Cfg>(P.addrMode, P.pow2Mip0, cx, m.w, m.log2w), x1 = Addr1<Cfg>(P.addrMode, P.pow2Mip0, cx + 1, m.w, m.log2w);
    const int y0 = Addr1<Cfg>(P.addrMode, P.pow2Mip0, cy, m.h, m.log2h), y1 = Addr1<Cfg>(P.addrMode, P.pow2Mip0, cy + 1, m.h, m.log2h);
    const float gx = TexFetch<Cfg>(P, m, x0, y0);
    const float gy = TexFetch<Cfg>(P, m, x0, y1);
    const float gz = TexFetch<Cfg>(P, m, x1, y1);
    const float gw = TexFetch<Cfg>(P, m, x1, y0);
    const bool o0 = P.cutoff < gx, o1 = P.cutoff < gy, o2 = P.cutoff < gz, o3 = P.cutoff < gw;
    if (o0 != o1 || o0 != o2 || o0 != o3) return 0;
    const float a = gx - P.cutoff;
    const float b = gw - gx;
    const float c = gy - gx;
    const float d = gx + gz - gy - gw;
    // h at the four cell corners, evaluated like TestRegion does for B = [0,1]^2
    const float h00 = a, h10 = a + b, h01 = a + c, h11 = (a + b) + (c + d);
    const float mn = fminf(fminf(h00, h01), fminf(h10, h11)), mx = fmaxf(fmaxf(h00, h01), fmaxf(h10, h11));
    int s;
    float margin;
    if (mn > 0.f) { s = 1; margin = mn; }
    else if (mx < 0.f) { s = -1; margin = -mx; }
    else return 0;
    if ((s > 0) != o0) return 0;
    const float gmaxAbs = fmaxf(fmaxf(fabsf(gx), fabsf(gy)), fmaxf(fabsf(gz), fabsf(gw)));
    if (!MarginBeatsEdgeBound(it, margin, fabsf(a), fabsf(b), fabsf(c), fabsf(d), gmaxAbs, fabsf(P.cutoff), qx, qy)) return 0;
    return s;
}
template <class Cfg>
OMM_HD int WholeCellSide(const BakeParams& P, const DevMip& m, const HierItem& it, const RegionBox& box, int cx, int cy) {
    const float fx = (float)cx, fy = (float)cy;
    const float qx = fmaxf(fabsf(box.lox - fx), fabsf(box.hix - fx)) + it.deltaEdge;
    const float qy = fmaxf(fabsf(box.loy - fy), fabsf(box.hiy - fy)) + it.deltaEdge;
    return WholeCellSideQ<Cfg>(P, m, it, qx, qy, cx, cy);
}

// (I) Item-independent whole-cell sides.  The bound R of (B) grows with the hull of the slope ranges, with the pad delta and with
//     the coordinate extent Q (MarginBeatsEdgeBound is monotone in all of them; the quadratic term is convex in K, so its maximum
//     over a sub-range is at most its maximum over the hull).  A cell that passes the whole-cell test (F) for the CAP item
//     -- every slope in [0, kStrongMaxSlope], delta = kStrongMaxDelta, Q = kStrongMaxExtent -- therefore passes it for every item
//     within the caps.  The two summed-area tables strongPlus / strongMinus count, over the interior cells of mip 0, the cells
//     that are NOT such a pass on the respective side (built per texture and cutoff, like (H)); a footprint inside the texture
//     with a zero count is answered with four loads instead of one texture gather and one bound evaluation per cell.
constexpr float kStrongMaxSlope = 64.f, kStrongMaxDelta = 0.05f, kStrongMaxExtent = 16.f;
OMM_HD HierItem StrongCapItem() {
    HierItem it;
    it.p0 = it.p1 = it.p2 = make_float2(0.f, 0.f);
    it.level = 0;
    it.epsRegion = it.epsSingle = 0.f;
    it.deltaEdge = kStrongMaxDelta;
    for (int j = 0; j < 3; ++j) { it.kmin[j] = 0.f; it.kmax[j] = kStrongMaxSlope; }
    it.pitX = it.pitY = 0.f;
    it.ok = 1;
    it.pad = 0;
    return it;
}
OMM_HD bool ItemWithinStrongCaps(const HierItem& it) {
    return it.ok && it.deltaEdge <= kStrongMaxDelta && it.kmax[0] <= kStrongMaxSlope && it.kmax[1] <= kStrongMaxSlope && it.kmax[2] <= kStrongMaxSlope;
}
// side of interior cell (cx, cy) for the cap item: +1 / -1 / 0   (0 <= cx <= w - 2, 0 <= cy <= h - 2: no addressing involved)
template <class Cfg>
OMM_HD int StrongCellSide(const BakeParams& P, const DevMip& m, int cx, int cy) {
    return WholeCellSideQ<Cfg>(P, m, StrongCapItem(), kStrongMaxExtent, kStrongMaxExtent, cx, cy);
}
// +1 / -1 when every cell of the rectangle is a cap pass of that side; `box` gives the coordinate extent the caller needs covered
OMM_HD int StrongRectSide(const BakeParams& P, const DevMip& m, const HierItem& it, const RegionBox& box, int cx0, int cy0, int cx1, int cy1) {
    const uint32_t* sp = P.tex.strongPlus;
    const uint32_t* sm = P.tex.strongMinus;
    if (!sp || !sm || cx0 < 0 || cy0 < 0 || cx1 > m.w - 2 || cy1 > m.h - 2 || cx1 < cx0 || cy1 < cy0) return 0;
    // Q of any cell of the box: at most the box size plus delta
    // (a footprint cell is at most one cell beyond either end of the box)
    if (!(box.hix - box.lox + 1.f + it.deltaEdge <= kStrongMaxExtent && box.hiy - box.loy + 1.f + it.deltaEdge <= kStrongMaxExtent)) return 0;
    const size_t sw = (size_t)(m.w - 1);
    const size_t iA = (size_t)(cy0 - 1) * sw + (size_t)(cx0 - 1), iB = (size_t)(cy0 - 1) * sw + (size_t)cx1, iC = (size_t)cy1 * sw + (size_t)(cx0 - 1),
                 iD = (size_t)cy1 * sw + (size_t)cx1;
    const bool hasA = cx0 > 0 && cy0 > 0, hasB = cy0 > 0, hasC = cx0 > 0;
    const uint32_t badPlus = LoadRO(sp + iD) + (hasA ? LoadRO(sp + iA) : 0u) - (hasB ? LoadRO(sp + iB) : 0u) - (hasC ? LoadRO(sp + iC) : 0u);
    if (badPlus == 0u) return 1;
    const uint32_t badMinus = LoadRO(sm + iD) + (hasA ? LoadRO(sm + iA) : 0u) - (hasB ? LoadRO(sm + iB) : 0u) - (hasC ? LoadRO(sm + iC) : 0u);
    if (badMinus == 0u) return -1;
    return 0;
}

// Bitmap of whole-cell sides over the footprint of an item (at most 32 x 32 cells): bit x of plus[y] / minus[y].
struct ItemCellMap {
    int cx0, cy0, fw, fh;  // fw == 0: no map
    const uint32_t* plus;
    const uint32_t* minus;
};
// Item box: region box of the whole item (bird index 0 at level 0) widened by one more eps.
OMM_HD bool MakeItemBox(const DevMip& m, const HierItem& it, RegionBox& box) {
    HierItem whole = it;
    whole.level = 0xFFFFFFFFu;  // forces epsRegion in MakeRegionBox
    if (!MakeRegionBox(m, whole, 0, 0, box)) return false;
    box.lox -= it.epsRegion; box.loy -= it.epsRegion; box.hix += it.epsRegion; box.hiy += it.epsRegion;
    box.cx0 = (int)floorf(box.lox); box.cy0 = (int)floorf(box.loy); box.cx1 = (int)floorf(box.hix); box.cy1 = (int)floorf(box.hiy);
    return true;
}
// (S) The SAT pass and the region proofs.  With a summed-area table (texture created with an alpha cutoff, one mip, Linear) the reference
//     first decides every micro-triangle whose SAT rectangle is uniform (CoarseState; bake_cpu_impl.cpp:749-801) and only walks the
//     others.  The rectangle spans the address-mapped texels of the micro-triangle's two bounding-box corners, floor(r_min) and
//     floor(r_max) + 1.  Where the address mapping is monotone over that texel range the rectangle is exactly the set of texels the
//     footprint cells use; it then holds the four texels of p0's cell, so a decisive SAT answer t means h has the sign of t all over
//     that cell and agrees with any region proof (whose side is the sign of h at the region's points in it).  Where the mapping FOLDS
//     inside the range (Mirror / MirrorOnce across a mirror axis) or wraps all the way round (Wrap across a period boundary: the two
//     corners can map to one texel column), the rectangle is not that set, the reference's answer is whatever the rectangle holds,
//     and only the micro-triangle-level evaluation (LeafClassify applies CoarseState first) reproduces it.  A work item is therefore
//     left to the exact walk (ok = 0) unless its whole texel range lies on one monotone piece of the mapping:
//       Clamp, Border  always (Clamp is monotone; Border rejects every rectangle that leaves the texture);
//       Wrap, Mirror   within one period in x and in y (a flipped Mirror period maps decreasingly: the reference rejects ex < sx);
//       MirrorOnce     entirely at texels >= 0 or entirely at texels <= -1 in x and in y.
//     Found by the randomized host campaign (MirrorOnce, SAT, UVs straddling 0): the region proofs said "opaque", the reference's
//     folded rectangle held a single transparent texel column.
OMM_HD int FloorDivInt(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }
OMM_HD bool SatPassCompatible(const BakeParams& P, const DevMip& m, const HierItem& it) {
    if (!P.useCoarse) return true;
    if (P.addrMode == ommTextureAddressMode_Clamp || P.addrMode == ommTextureAddressMode_Border) return true;
    RegionBox box;
    if (!MakeItemBox(m, it, box)) return false;
    const int tx0 = box.cx0, tx1 = box.cx1 + 1, ty0 = box.cy0, ty1 = box.cy1 + 1;  // superset of every micro-triangle's texel range
    if (P.addrMode == ommTextureAddressMode_MirrorOnce) return (tx0 >= 0 || tx1 <= -1) && (ty0 >= 0 || ty1 <= -1);
    if (P.addrMode == ommTextureAddressMode_Wrap || P.addrMode == ommTextureAddressMode_Mirror)
        return FloorDivInt(tx0, m.w) == FloorDivInt(tx1, m.w) && FloorDivInt(ty0, m.h) == FloorDivInt(ty1, m.h);
    return false;
}
// The per-item constants as the kernels use them: MakeHierItem plus (S).
OMM_HD HierItem MakeHierItemFor(const BakeParams& P, const DevMip& m, float2 p0, float2 p1, float2 p2, uint32_t level, bool degenerate) {
    HierItem it = MakeHierItem(m, p0, p1, p2, level, degenerate);
    if (it.ok && !SatPassCompatible(P, m, it)) it.ok = 0;
    return it;
}
// Node box: the same for the sub-triangle `index` at `nodeLevel` of the item (an aligned group of initial regions of a large item).
OMM_HD bool MakeNodeBox(const DevMip& m, const HierItem& it, uint32_t index, uint32_t nodeLevel, RegionBox& box) {
    HierItem whole = it;
    whole.level = 0xFFFFFFFFu;  // forces epsRegion in MakeRegionBox
    if (!MakeRegionBox(m, whole, index, nodeLevel, box)) return false;
    box.lox -= it.epsRegion; box.loy -= it.epsRegion; box.hix += it.epsRegion; box.hiy += it.epsRegion;
    box.cx0 = (int)floorf(box.lox); box.cy0 = (int)floorf(box.loy); box.cx1 = (int)floorf(box.hix); box.cy1 = (int)floorf(box.hiy);
    return true;
}
// +1 / -1 when every footprint cell of the region box is a whole-cell pass of that side, else 0
OMM_HD int LookupCellMap(const ItemCellMap& map, const RegionBox& rb) {
    if (map.fw == 0) return 0;
    const int x0 = rb.cx0 - map.cx0, x1 = rb.cx1 - map.cx0, y0 = rb.cy0 - map.cy0, y1 = rb.cy1 - map.cy0;
    if (x0 < 0 || y0 < 0 || x1 >= map.fw || y1 >= map.fh) return 0;
    const uint32_t cols = (x1 - x0 == 31 ? 0xFFFFFFFFu : ((1u << (x1 - x0 + 1)) - 1u)) << x0;
    uint32_t allPlus = cols, allMinus = cols;
    for (int y = y0; y <= y1; ++y) {
        allPlus &= map.plus[y];
        allMinus &= map.minus[y];
    }
    if (allPlus == cols) return 1;
    if (allMinus == cols) return -1;
    return 0;
}

// (H) Constant areas.  A cell is FLAT-GOOD for a cutoff when its four texels are the same float g and |g - cutoff| exceeds
//     1e-5 (1 + |g| + |cutoff|).  There b = c = d = 0 exactly, so the kernel takes its "all points on the same level" branch
//     (bake_kernels_cpu.h:343-354) and never tests an edge; texel centres vote the side of g; the bilinear sample is a lerp of four
//     equal values, off by <= 7 u |g| from g.  A rectangle of cells inside the texture (texel columns cx0 .. cx1 + 1 < w, no
//     addressing involved) that contains only flat-good cells holds a single value, because neighbouring cells share texels: every
//     region whose footprint lies in it is on the side of that value, whatever its size.  flatSat is the inclusive summed-area
//     table of the NOT-flat-good flags of mip 0, built per texture and cutoff (BuildFlatSat); this makes large triangles in fully
//     opaque / fully transparent parts of a texture O(1) instead of O(cells).
OMM_HD bool CellIsFlatGood(float g00, float g10, float g01, float g11, float cutoff) {
    if (!(g00 == g10 && g00 == g01 && g00 == g11)) return false;
    return fabsf(g00 - cutoff) > 1e-5f * (1.f + fabsf(g00) + fabsf(cutoff));
}
template <class Cfg>
OMM_HD int FlatRectSide(const BakeParams& P, const DevMip& m, int cx0, int cy0, int cx1, int cy1) {
    const uint32_t* sat = P.tex.flatSat;
    if (!sat || cx0 < 0 || cy0 < 0 || cx1 > m.w - 2 || cy1 > m.h - 2 || cx1 < cx0 || cy1 < cy0) return 0;
    const size_t sw = (size_t)(m.w - 1);
    const uint32_t A = (cx0 > 0 && cy0 > 0) ? LoadRO(sat + (size_t)(cy0 - 1) * sw + (cx0 - 1)) : 0u;
    const uint32_t B = cy0 > 0 ? LoadRO(sat + (size_t)(cy0 - 1) * sw + cx1) : 0u;
    const uint32_t C = cx0 > 0 ? LoadRO(sat + (size_t)cy1 * sw + (cx0 - 1)) : 0u;
    const uint32_t D = LoadRO(sat + (size_t)cy1 * sw + cx1);
    if (D + A - B - C != 0u) return 0;
    return P.cutoff < TexLoad<Cfg>(P.tex, m, cx0, cy0) ? 1 : -1;
}

// Returns +1 / -1 when every micro-triangle of the region (bird index `index` at subdivision level `regionLevel` of the
// work item; regionLevel == it.level means a single micro-triangle) is provably on that side of the cutoff, 0 otherwise.
// One footprint cell of a region: +1 / -1 when the cell passes (B), (C) [and (G) when `single`: the region lies in this one cell]
// on that side, 0 when it does not.  Only the enclosure, the corners and eps of `rb` are used, so HierTestList can evaluate the
// cells of thirty-two regions side by side, one cell per lane.
template <class Cfg>
OMM_HD int TestRegionCell(const BakeParams& P, const DevMip& m, const HierItem& it, const RegionBox& rb, int cx, int cy, bool single) {
    const float lox = rb.lox, loy = rb.loy, hix = rb.hix, hiy = rb.hiy;
    const float delta = it.deltaEdge;
    const int y0 = Addr1<Cfg>(P.addrMode, P.pow2Mip0, cy, m.h, m.log2h), y1 = Addr1<Cfg>(P.addrMode, P.pow2Mip0, cy + 1, m.h, m.log2h);
    const float fy = (float)cy;
    const float by0 = fmaxf(0.f, loy - fy - delta), by1 = fminf(1.f, hiy - fy + delta);
    const float qy = fmaxf(fabsf(loy - fy), fabsf(hiy - fy)) + delta;
    const int x0 = Addr1<Cfg>(P.addrMode, P.pow2Mip0, cx, m.w, m.log2w), x1 = Addr1<Cfg>(P.addrMode, P.pow2Mip0, cx + 1, m.w, m.log2w);
    // (c00, c01, c11, c10) as the reference gathers them
    const float gx = TexFetch<Cfg>(P, m, x0, y0);
    const float gy = TexFetch<Cfg>(P, m, x0, y1);
    const float gz = TexFetch<Cfg>(P, m, x1, y1);
    const float gw = TexFetch<Cfg>(P, m, x1, y0);
    const float a = gx - P.cutoff;
    const float b = gw - gx;
    const float c = gy - gx;
    const float d = gx + gz - gy - gw;
    const float fx = (float)cx;
    const float bx0 = fmaxf(0.f, lox - fx - delta), bx1 = fminf(1.f, hix - fx + delta);
    const float qx = fmaxf(fabsf(lox - fx), fabsf(hix - fx)) + delta;
    // h at the four corners of B
    const float e0 = a + b * bx0, e1 = a + b * bx1;
    const float f0 = c + d * bx0, f1 = c + d * bx1;
    const float h00 = e0 + f0 * by0, h01 = e0 + f0 * by1, h10 = e1 + f1 * by0, h11 = e1 + f1 * by1;
    const float mn = fminf(fminf(h00, h01), fminf(h10, h11)), mx = fmaxf(fmaxf(h00, h01), fmaxf(h10, h11));
    int s;
    float margin;
    if (mn > 0.f) { s = 1; margin = mn; }
    else if (mx < 0.f) { s = -1; margin = -mx; }
    else if (single) {
        // (G) The whole region lies in this one cell: bound h over the region TRIANGLE instead of its bounding box.  Every
        // vertex of the region is within eps of the triangle T spanned by the three corner positions q_i (A), and a
        // reported intersection within delta of such an edge (B).  Along a segment h deviates from the chord by at most
        // |d| |dx dy| / 4; a point of T lies on a segment from q_0 to a point of the opposite edge, so over T
        //     s h >= min_i s h(q_i) - |d| wx wy / 2,
        // and moving eps + delta away costs at most (eps + delta)(|b| + |c| + |d|(Qx + Qy)).  The float evaluation of
        // h(q_i) is off by <= 8 u (|a'| + |b| Qx + |c| Qy + |d| Qx Qy).
        const float q0x = rb.r0x - fx, q0y = rb.r0y - fy, q1x = rb.r1x - fx, q1y = rb.r1y - fy, q2x = rb.r2x - fx, q2y = rb.r2y - fy;
        const float v0 = (a + b * q0x) + (c + d * q0x) * q0y;
        const float v1 = (a + b * q1x) + (c + d * q1x) * q1y;
        const float v2 = (a + b * q2x) + (c + d * q2x) * q2y;
        if (v0 > 0.f && v1 > 0.f && v2 > 0.f) s = 1;
        else if (v0 < 0.f && v1 < 0.f && v2 < 0.f) s = -1;
        else { OMM_STAT(4); return 0; }
        const float al = fabsf(a), be = fabsf(b), ga = fabsf(c), de = fabsf(d);
        const float wx = hix - lox, wy = hiy - loy;
        const float pad = 0.505f * de * wx * wy + 2.f * (rb.eps + delta) * (be + ga + de * (qx + qy)) +
                          8.f * kUnitRoundoff * (al + be * qx + ga * qy + de * qx * qy);
        margin = fminf(fminf(fabsf(v0), fabsf(v1)), fabsf(v2)) - pad;
    } else { OMM_STAT(4); return 0; }
    // (C) texel centres on the other side must be out of reach of PointInTriangle
    const bool want = s > 0;
    const bool o0 = P.cutoff < gx, o1 = P.cutoff < gy, o2 = P.cutoff < gz, o3 = P.cutoff < gw;
    if (o0 != want || o1 != want || o2 != want || o3 != want) {
        const bool farL = lox - fx >= it.pitX;             // corners with local x = 0 are left of the region
        const bool farR = (fx + 1.f) - hix >= it.pitX;     // corners with local x = 1 are right of it
        const bool farB = loy - fy >= it.pitY;
        const bool farT = (fy + 1.f) - hiy >= it.pitY;
        // corner 0 = (0,0), 1 = (0,1), 2 = (1,1), 3 = (1,0); a corner at local x = 1 is also "left of the region" when the
        // region starts beyond it, i.e. lox - (fx + 1) >= pit, which implies farL; the four flags below are the cheap subset.
        if (o0 != want && !(farL || farB)) { OMM_STAT(6); return 0; }
        if (o1 != want && !(farL || farT)) { OMM_STAT(6); return 0; }
        if (o2 != want && !(farR || farT)) { OMM_STAT(6); return 0; }
        if (o3 != want && !(farR || farB)) { OMM_STAT(6); return 0; }
    }
    const float gmaxAbs = fmaxf(fmaxf(fabsf(gx), fabsf(gy)), fmaxf(fabsf(gz), fabsf(gw)));
    if (!MarginBeatsEdgeBound(it, margin, fabsf(a), fabsf(b), fabsf(c), fabsf(d), gmaxAbs, fabsf(P.cutoff), qx, qy)) { OMM_STAT(7); return 0; }
    return s;
}

// All footprint cells must pass on the same side.
template <class Cfg>
OMM_HD int TestRegionBox(const BakeParams& P, const DevMip& m, const HierItem& it, const RegionBox& rb) {
    const int cx0 = rb.cx0, cy0 = rb.cy0, cx1 = rb.cx1, cy1 = rb.cy1;
    if (ItemWithinStrongCaps(it)) {
        const int s = StrongRectSide(P, m, it, rb, cx0, cy0, cx1, cy1);  // (I)
        if (s != 0) return s;
    }
    if ((cx1 - cx0 + 1) * (cy1 - cy0 + 1) > kHierMaxCells) return FlatRectSide<Cfg>(P, m, cx0, cy0, cx1, cy1);  // (H) or split
    const bool single = cx0 == cx1 && cy0 == cy1;
    int sAll = 0;
    for (int cy = cy0; cy <= cy1; ++cy)
        for (int cx = cx0; cx <= cx1; ++cx) {
            const int s = TestRegionCell<Cfg>(P, m, it, rb, cx, cy, single);
            if (s == 0) return 0;
            if (sAll != 0 && sAll != s) { OMM_STAT(5); return 0; }
            sAll = s;
        }
    return sAll;
}
template <class Cfg>
OMM_HD int TestRegion(const BakeParams& P, const DevMip& m, const HierItem& it, uint32_t index, uint32_t regionLevel) {
    RegionBox rb;
    if (!MakeRegionBox(m, it, index, regionLevel, rb)) return 0;
    return TestRegionBox<Cfg>(P, m, it, rb);
}


// ---- leaf: one micro-triangle, the reference walk with two exact skips ------------------------------------------------
// (D) Per-cell edge filter.  q0,q1,q2 are the reference's own cell-local vertex coordinates.  A reported intersection lies
//     within delta of one of the three segments (B), at a point of the unit square where |h| <= R.  Along a segment h is a
//     quadratic g(t) with second-order coefficient A2 = d ex ey, so g stays within |A2|/4 of the chord between its end
//     values; moving delta away from the segment changes h by at most delta (|b| + |c| + |d|(Qx + Qy)).  Hence, when the
//     three vertex values have one sign s and
//        min_i |h(q_i)| - max_j |A2_j| / 4 - 3 delta (|b| + |c| + |d|(Qx + Qy)) - 8 u (|a'| + |b| Qx + |c| Qy + |d| Qx Qy)  >  R
//     (the last term covers the float evaluation of h(q_i)), none of the three tests can succeed and they are skipped.
// (E) Votes that cannot matter.  Unless the promotion is Nearest the state depends only on which counters are non-zero
//     (bake_kernels_cpu.h:27-50).  In a cell whose four texels are on side s, where side s has been voted already and the
//     edge filter holds, the corner and flat-patch branches can only vote s again: the whole cell is skipped.
// (Evaluating the edge tests elsewhere -- queued for a kernel of their own, or parked in a per-warp shared-memory queue and run one
// edge per lane after the walk -- was built twice and measured slower both times; they stay in place.)
template <class Cfg>
OMM_HD void LeafCell(const BakeParams& P, const DevMip& m, const HierItem& it, const Tri& tri, int px, int py, Coverage& cov, bool countsMatter) {
    const float pfx = (float)px + 0.5f, pfy = (float)py + 0.5f;
    const int x0 = Addr1<Cfg>(P.addrMode, P.pow2Mip0, px, m.w, m.log2w), y0 = Addr1<Cfg>(P.addrMode, P.pow2Mip0, py, m.h, m.log2h);
    const int x1 = Addr1<Cfg>(P.addrMode, P.pow2Mip0, px + 1, m.w, m.log2w), y1 = Addr1<Cfg>(P.addrMode, P.pow2Mip0, py + 1, m.h, m.log2h);
    const float gx = TexFetch<Cfg>(P, m, x0, y0);
    const float gy = TexFetch<Cfg>(P, m, x0, y1);
    const float gz = TexFetch<Cfg>(P, m, x1, y1);
    const float gw = TexFetch<Cfg>(P, m, x1, y0);
    const float a = gx;
    const float b = gw - gx;
    const float c = gy - gx;
    const float d = gx + gz - gy - gw;
    const float sx = (float)m.w, sy = (float)m.h;
    const float h0 = a - P.cutoff;
    const float2 q0 = make_float2(sx * tri.p0.x - pfx, sy * tri.p0.y - pfy);
    const float2 q1 = make_float2(sx * tri.p1.x - pfx, sy * tri.p1.y - pfy);
    const float2 q2 = make_float2(sx * tri.p2.x - pfx, sy * tri.p2.y - pfy);
    const bool o0 = P.cutoff < gx, o1 = P.cutoff < gy, o2 = P.cutoff < gz, o3 = P.cutoff < gw;

    // (D) edge filter
    bool edgesCannotHit = false;
    int s = 0;
    const float v0 = (h0 + b * q0.x) + (c + d * q0.x) * q0.y;
    const float v1 = (h0 + b * q1.x) + (c + d * q1.x) * q1.y;
    const float v2 = (h0 + b * q2.x) + (c + d * q2.x) * q2.y;
    {
        if (v0 > 0.f && v1 > 0.f && v2 > 0.f) s = 1;
        else if (v0 < 0.f && v1 < 0.f && v2 < 0.f) s = -1;
        if (s != 0 && it.ok) {
            const float al = fabsf(h0), be = fabsf(b), ga = fabsf(c), de = fabsf(d);
            const float qx = fmaxf(fmaxf(fabsf(q0.x), fabsf(q1.x)), fabsf(q2.x)) + it.deltaEdge;
            const float qy = fmaxf(fmaxf(fabsf(q0.y), fabsf(q1.y)), fabsf(q2.y)) + it.deltaEdge;
            const float e01 = fabsf((q1.x - q0.x) * (q1.y - q0.y)), e12 = fabsf((q2.x - q1.x) * (q2.y - q1.y)), e20 = fabsf((q0.x - q2.x) * (q0.y - q2.y));
            const float sag = 0.2525f * de * fmaxf(fmaxf(e01, e12), e20);
            const float pad = 3.f * it.deltaEdge * (be + ga + de * (qx + qy)) + 8.f * kUnitRoundoff * (al + be * qx + ga * qy + de * qx * qy);
            const float margin = fminf(fminf(fabsf(v0), fabsf(v1)), fabsf(v2)) - sag - pad;
            const float gmaxAbs = fmaxf(fmaxf(fabsf(gx), fabsf(gy)), fmaxf(fabsf(gz), fabsf(gw)));
            edgesCannotHit = MarginBeatsEdgeBound(it, margin, al, be, ga, de, gmaxAbs, fabsf(P.cutoff), qx, qy);
#if defined(OMM_HIER_STATS)
            if (!edgesCannotHit) g_hierStats[!(margin > 1e-5f) ? 12 : (de < 1e-6f ? 13 : 14)]++;  // near the line / d is rounding noise / other
#endif
        }
    }
#if defined(OMM_HIER_STATS)
    g_hierStats[0]++;
    if (s != 0) g_hierStats[1]++;
    if (edgesCannotHit) g_hierStats[2]++;
#endif
    // (E)
    if (edgesCannotHit && !countsMatter) {
        const bool want = s > 0;
        if (o0 == want && o1 == want && o2 == want && o3 == want && (want ? cov.above : cov.below) != 0) {
#if defined(OMM_HIER_STATS)
            g_hierStats[3]++;
#endif
            return;
        }
    }
    {
        // (C) at the leaf: a texel centre separated from the micro-triangle's box by pit in x or in y is provably outside for
        // PointInTriangle (the q are the cell-local vertex coordinates, epsSingle covers their rounding against r - cell)
        const float lox = fminf(fminf(q0.x, q1.x), q2.x) - it.epsSingle, hix = fmaxf(fmaxf(q0.x, q1.x), q2.x) + it.epsSingle;
        const float loy = fminf(fminf(q0.y, q1.y), q2.y) - it.epsSingle, hiy = fmaxf(fmaxf(q0.y, q1.y), q2.y) + it.epsSingle;
        const bool farL = lox >= it.pitX, farR = 1.f - hix >= it.pitX, farB = loy >= it.pitY, farT = 1.f - hiy >= it.pitY;
        const float ipx = pfx * m.rcpw, ipy = pfy * m.rcph;
        const bool in0 = !(farL || farB) && PointInTri(tri, ipx, ipy);
        const bool in1 = !(farL || farT) && PointInTri(tri, ipx + 0.0f, ipy + m.rcph);
        const bool in2 = !(farR || farT) && PointInTri(tri, ipx + m.rcpw, ipy + m.rcph);
        const bool in3 = !(farR || farB) && PointInTri(tri, ipx + m.rcpw, ipy + 0.0f);
        const bool isOpaque = (in0 && o0) || (in1 && o1) || (in2 && o2) || (in3 && o3);
        const bool isTransparent = (in0 && !o0) || (in1 && !o1) || (in2 && !o2) || (in3 && !o3);
        if (isOpaque) cov.above += 1;
        if (isTransparent) cov.below += 1;
        if (isOpaque && isTransparent) return;
    }
    if (IsZero(b, 1e-6f) && IsZero(c, 1e-6f) && IsZero(d, 1e-6f)) {
        if (P.cutoff < a) cov.above += 1;
        else cov.below += 1;
        return;
    }
    if (edgesCannotHit) return;
    // (Testing the edges whose end points straddle the level line first was measured: the per-lane order makes the calls diverge
    // and costs more than the skipped tests save.)
    OMM_STAT(8);
    const bool hit = EdgeHyperbola(q0, q1, h0, b, c, d) || EdgeHyperbola(q1, q2, h0, b, c, d) || EdgeHyperbola(q2, q0, h0, b, c, d);
    if (hit) {
        OMM_STAT(9);
        cov.above += 1;
        cov.below += 1;
    }
}

// One micro-triangle of a non-degenerate work item, Linear filter, level-line test, single mip, no SAT pass
// (ref: bake_cpu_impl.cpp:861-908 with ResampleFine's Normal path).
template <class Cfg>
OMM_HD int LeafClassify(const BakeParams& P, const DevMip& m, const HierItem& it, uint32_t index) {
    const Tri st = MicroTri(it.p0, it.p1, it.p2, index, it.level);
    if (P.useCoarse) {
        // SAT pass of the reference first (bake_cpu_impl.cpp:749-801, 861-864).  The hierarchical path is only taken when the
        // texture's cutoff equals the bake's (SelectHierKernels) and for work items whose texel range the address mapping does not
        // fold (S): a decisive SAT answer then agrees with any region proof, because the SAT rectangle holds the four texels of
        // p0's cell and a patch whose corners are all on one side cannot have h of the other sign with margin.
        const int cs = CoarseState<Cfg>(P, st);
        if (cs >= 0 && cs != ommOpacityState_UnknownOpaque) return cs;
    }
    if (P.tex.flatSat) {
        // (H) a micro-triangle of many texels over a constant area: its footprint (A) instead of a walk over every cell
        const float W = (float)m.w, H = (float)m.h, eps = it.epsSingle;
        const float r0x = st.p0.x * W + -0.5f, r0y = st.p0.y * H + -0.5f, r1x = st.p1.x * W + -0.5f, r1y = st.p1.y * H + -0.5f;
        const float r2x = st.p2.x * W + -0.5f, r2y = st.p2.y * H + -0.5f;
        const float lox = fminf(fminf(r0x, r1x), r2x) - eps, loy = fminf(fminf(r0y, r1y), r2y) - eps;
        const float hix = fmaxf(fmaxf(r0x, r1x), r2x) + eps, hiy = fmaxf(fmaxf(r0y, r1y), r2y) + eps;
        if (lox > -2097152.f && loy > -2097152.f && hix < 2097152.f && hiy < 2097152.f && (hix - lox) * (hiy - loy) > 16.f) {
            const int s = FlatRectSide<Cfg>(P, m, (int)floorf(lox), (int)floorf(loy), (int)floorf(hix), (int)floorf(hiy));
            if (s != 0) return s > 0 ? P.stateGT : P.stateLE;
        }
    }
    Coverage cov{0u, 0u};
    const bool countsMatter = P.promotion == ommUnknownStatePromotion_Nearest;
    if (P.cutoff < TexBilinear<Cfg>(P, m, st.p0)) cov.above++;
    else cov.below++;
    const RasterSetup rs = MakeRasterSetup(st, m.w, m.h, -0.5f);
    RasterCursor cur = RasterBegin(rs);
    int x, y;
    while (RasterNext(rs, cur, x, y)) {
        LeafCell<Cfg>(P, m, it, st, x, y, cov, countsMatter);
        if (!countsMatter && cov.above != 0 && cov.below != 0) break;  // exact early-out, see ClassifyMicroTriangle
    }
    OMM_STAT(10);
    if (cov.above != 0 && cov.below != 0) OMM_STAT(11);
    return StateFromCoverage(P, cov.above, cov.below);
}
// Several mips (ref: bake_cpu_impl.cpp:866-908): the reference classifies against mip 0, 1, ... and stops as soon as the state is an
// Unknown one; votes accumulate over the mips.  `itemOfMip(k)` returns the HierItem of the work item for mip k (the constants of the
// exact skips depend on the mip's size).  The region tests demand the same side on EVERY mip, which gives state(s) whether or not the
// reference would have stopped early.
template <class Cfg, class ItemOfMip>
OMM_HD int LeafClassifyMips(const BakeParams& P, const ItemOfMip& itemOfMip, uint32_t index) {
    const HierItem it0 = itemOfMip(0);
    const Tri st = MicroTri(it0.p0, it0.p1, it0.p2, index, it0.level);
    Coverage cov{0u, 0u};
    const bool countsMatter = P.promotion == ommUnknownStatePromotion_Nearest;
    for (int mip = 0; mip < P.tex.mipCount; ++mip) {
        const DevMip& m = P.tex.mips[mip];
        const HierItem it = mip == 0 ? it0 : itemOfMip(mip);
        if (P.cutoff < TexBilinear<Cfg>(P, m, st.p0)) cov.above++;
        else cov.below++;
        const RasterSetup rs = MakeRasterSetup(st, m.w, m.h, -0.5f);
        RasterCursor cur = RasterBegin(rs);
        int x, y;
        bool stop = false;
        while (RasterNext(rs, cur, x, y)) {
            LeafCell<Cfg>(P, m, it, st, x, y, cov, countsMatter);
            if (!countsMatter && cov.above != 0 && cov.below != 0) { stop = true; break; }  // exact early-out, see ClassifyMicroTriangle
        }
        if (stop) break;
        if (IsUnknownState(StateFromCoverage(P, cov.above, cov.below))) break;
    }
    return StateFromCoverage(P, cov.above, cov.below);
}

}  // namespace ommb200
