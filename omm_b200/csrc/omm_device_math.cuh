// omm_device_math.cuh -- device-side arithmetic of the micro-triangle classifier.
//
// Bit-exactness contract: the SDK's CPU baker is built with SSE scalar float math, no FMA, no fast-math
// (libraries/omm-lib/CMakeLists.txt:137-146).  This translation unit is compiled with -fmad=false and default
// (IEEE) division / sqrt, and every expression keeps the operation ORDER of the reference expression it
// stands for; the citations name that expression ("ref:" paths are under libraries/omm-lib/src).
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "omm_internal.h"

// Every function of this header compiles for the device AND for the host: the host build (g++ or nvcc's host pass,
// -ffp-contract=off) is what tests/hier_host_check.cpp links to fuzz the exact shortcuts of omm_hier.cuh against the
// plain reference walk without a GPU.  It is never part of the product path (the library has no CPU fallback).
#if defined(__CUDACC__)
#define OMM_HD __host__ __device__ __forceinline__
#define OMM_HD_NOINLINE __host__ __device__ __noinline__
#else
#define OMM_HD inline
#define OMM_HD_NOINLINE inline
#endif

namespace ommb200 {

OMM_HD float UintAsFloat(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}
OMM_HD uint32_t FloatAsUint(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}
template <class T>
OMM_HD T LoadRO(const T* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

// ---- bake parameters visible to the kernels ----------------------------------------------------------------------
struct BakeParams {
    DevTexture tex;
    int addrMode;       // ommTextureAddressMode
    int filterLinear;   // 1 = Linear, 0 = Nearest
    float borderAlpha;
    float cutoff;
    int stateGT, stateLE;  // ommOpacityState for alpha > cutoff / <= cutoff
    int globalFormat;      // desc.format (used by GetStateFromCoverage, ref: bake_cpu_impl.cpp:907)
    int promotion;         // ommUnknownStatePromotion
    int pow2Mip0;          // template parameter bTexIsPow2 of the SDK (ref: bake_cpu_impl.cpp:299)
    int useCoarse;         // SAT pass enabled (ref: bake_cpu_impl.cpp:723-727, 746)
    int disableFine;       // internal flag bit 9
    int disableLevelLine;  // internal flag bit 8
    int aabbTesting;       // internal flag bit 7
    int coarseSameCutoff;  // the texture's SAT was built for the bake's alpha cutoff
    int skipUniformFill;   // hierarchical classifier: items proved uniform as a whole need no state words (they become special indices)
};

struct Tri {
    float2 p0, p1, p2;
    float2 p0p2, p1p0, p2p1;
    float2 aabb_s, aabb_e;
};

constexpr int kTexCoordBorder = 0x7FFFFFFE;  // ref: util/texture.h:22

// (int)float as x86-64 cvttss2si does it: out-of-range and NaN give INT_MIN ("integer indefinite").
OMM_HD int f2i(float f) {
#if defined(__CUDA_ARCH__)
    return (f >= -2147483648.f && f < 2147483648.f) ? __float2int_rz(f) : (int)0x80000000;
#else
    return (f >= -2147483648.f && f < 2147483648.f) ? (int)f : (int)0x80000000;
#endif
}
OMM_HD float fminStd(float a, float b) { return b < a ? b : a; }  // std::min
OMM_HD float fmaxStd(float a, float b) { return a < b ? b : a; }  // std::max
OMM_HD int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

OMM_HD Tri MakeTri(float2 p0, float2 p1, float2 p2) {  // ref: util/geometry.h:63-75
    Tri t;
    t.p0 = p0; t.p1 = p1; t.p2 = p2;
    t.p0p2 = make_float2(p0.x - p2.x, p0.y - p2.y);
    t.p1p0 = make_float2(p1.x - p0.x, p1.y - p0.y);
    t.p2p1 = make_float2(p2.x - p1.x, p2.y - p1.y);
    t.aabb_s = make_float2(fminStd(fminStd(p0.x, p1.x), p2.x), fminStd(fminStd(p0.y, p1.y), p2.y));
    t.aabb_e = make_float2(fmaxStd(fmaxStd(p0.x, p1.x), p2.x), fmaxStd(fmaxStd(p0.y, p1.y), p2.y));
    return t;
}

OMM_HD bool TriIsDegenerate(float2 p0, float2 p1, float2 p2) {  // ref: util/geometry.h:44-47
    const float area = 0.5f * fabsf(p0.x * (p1.y - p2.y) + p1.x * (p2.y - p0.y) + p2.x * (p0.y - p1.y));
    return (double)area < 1e-9;
}
OMM_HD bool TriIsCCW(float2 p0, float2 p1, float2 p2) {  // ref: util/geometry.h:49-55
    const double ax = (double)(p2.x - p0.x), ay = (double)(p2.y - p0.y);
    const double bx = (double)(p1.x - p0.x), by = (double)(p1.y - p0.y);
#if defined(__CUDA_ARCH__)
    const double nz = __dsub_rn(__dmul_rn(ax, by), __dmul_rn(bx, ay));
#else
    const double nz = ax * by - bx * ay;  // both products are exact in double (24-bit x 24-bit significands)
#endif
    return nz < 0;
}
OMM_HD bool PointInTri(const Tri& t, float px, float py) {  // ref: util/geometry.h:101-114
    const float ptp2x = px - t.p2.x, ptp2y = py - t.p2.y;
    const float ptp0x = px - t.p0.x, ptp0y = py - t.p0.y;
    const float s = t.p0p2.x * ptp2y - t.p0p2.y * ptp2x;
    const float tt = t.p1p0.x * ptp0y - t.p1p0.y * ptp0x;
    if ((s < 0) != (tt < 0) && s != 0 && tt != 0) return false;
    const float ptp1x = px - t.p1.x, ptp1y = py - t.p1.y;
    const float d = t.p2p1.x * ptp1y - t.p2p1.y * ptp1x;
    return d == 0 || (d < 0) == (s + tt <= 0);
}

// ---- bird curve (ref: util/bird.h:36-118, 170-182) ---------------------------------------------------------------
OMM_HD uint32_t ExtractEvenBits(uint32_t x) {
    x &= 0x55555555u;
    x = (x | (x >> 1)) & 0x33333333u;
    x = (x | (x >> 2)) & 0x0f0f0f0fu;
    x = (x | (x >> 4)) & 0x00ff00ffu;
    x = (x | (x >> 8)) & 0x0000ffffu;
    return x;
}
OMM_HD uint32_t PrefixEor(uint32_t x) {
    x ^= x >> 1; x ^= x >> 2; x ^= x >> 4; x ^= x >> 8;
    return x;
}
// Discrete barycentrics of micro-triangle `index`: lattice vertex (iu, iv) of its first corner and orientation.
OMM_HD void Index2DBary(uint32_t index, uint32_t level, uint32_t& iu, uint32_t& iv, bool& upright) {
    const uint32_t b0 = ExtractEvenBits(index), b1 = ExtractEvenBits(index >> 1);
    const uint32_t fx = PrefixEor(b0), fy = PrefixEor(b0 & ~b1);
    const uint32_t t = fy ^ b1;
    uint32_t u = (fx & ~t) | (b0 & ~t) | (~b0 & ~fx & t);
    uint32_t v = fy ^ b0;
    uint32_t w = (~fx & ~t) | (b0 & ~t) | (~b0 & fx & t);
    const uint32_t mask = (1u << level) - 1u;
    u &= mask; v &= mask; w &= mask;
    upright = ((u & 1) ^ (v & 1) ^ (w & 1)) != 0;
    if (!upright) { u += 1; v += 1; }
    iu = u; iv = v;
}
// ref: util/geometry.h:241-248 -- (p0*b.x + p1*b.y) + p2*b.z with b = (1-u-v, u, v)
OMM_HD float2 InterpUV(float u, float v, float2 p0, float2 p1, float2 p2) {
    const float bx = 1.f - u - v, by = u, bz = v;
    return make_float2(p0.x * bx + p1.x * by + p2.x * bz, p0.y * bx + p1.y * by + p2.y * bz);
}
OMM_HD Tri MicroTri(float2 p0, float2 p1, float2 p2, uint32_t index, uint32_t level) {
    if (level == 0) {
        return MakeTri(InterpUV(0.f, 0.f, p0, p1, p2), InterpUV(1.f, 0.f, p0, p1, p2), InterpUV(0.f, 1.f, p0, p1, p2));
    }
    uint32_t iu, iv;
    bool upright;
    Index2DBary(index, level, iu, iv, upright);
    const float levelScale = UintAsFloat((127u - level) << 23);
    float du = 1.f * levelScale, dv = 1.f * levelScale;
    const float u = (float)iu * levelScale, v = (float)iv * levelScale;
    if (!upright) { du = -du; dv = -dv; }
    return MakeTri(InterpUV(u, v, p0, p1, p2), InterpUV(u + du, v, p0, p1, p2), InterpUV(u, v + dv, p0, p1, p2));
}

// ---- texture addressing (ref: util/texture.h:35-91) ----------------------------------------------------------------
OMM_HD int Addr1Generic(int mode, int pow2, int c, int size, int sizeLog2) {
    switch (mode) {
    case ommTextureAddressMode_Wrap:
        return pow2 ? (int)((uint32_t)c & (uint32_t)(size - 1)) : (int)((uint32_t)c % (uint32_t)size);
    case ommTextureAddressMode_Mirror:
        if (pow2) {
            const int a = (c < 0 ? -c : c) - (c < 0);
            const int flipped = (a >> sizeLog2) & 1;
            const int wrapped = (int)((uint32_t)a & (uint32_t)(size - 1));
            return flipped ? size - wrapped - 1 : wrapped;
        } else {
            const int a = f2i(fabsf((float)c + 0.5f));
            const uint32_t flipped = ((uint32_t)(a / size)) % 2u;
            const int wrapped = (int)((uint32_t)a % (uint32_t)size);
            return flipped ? size - wrapped - 1 : wrapped;
        }
    case ommTextureAddressMode_Clamp:
        return clampi(c, 0, size - 1);
    case ommTextureAddressMode_Border:
        return (c >= size || c < 0) ? kTexCoordBorder : c;
    case ommTextureAddressMode_MirrorOnce:
        return clampi(f2i(fabsf((float)c + 0.5f)), 0, size - 1);
    default:
        return 0x7FFFFFFF;
    }
}

// Compile-time specialisations of the kernel for the common sampler / texture combinations; kAddrGeneric keeps every
// combination available through the run-time switch above.
enum AddrSel { kAddrGeneric = 0, kAddrWrapPow2 = 1, kAddrClamp = 2 };
template <int kAddr_, bool kFp32_>
struct KernelCfg {
    static constexpr int kAddr = kAddr_;
    static constexpr bool kFp32 = kFp32_;
};

template <class Cfg>
OMM_HD int Addr1(int mode, int pow2, int c, int size, int sizeLog2) {
    if (Cfg::kAddr == kAddrWrapPow2) return (int)((uint32_t)c & (uint32_t)(size - 1));
    if (Cfg::kAddr == kAddrClamp) return clampi(c, 0, size - 1);
    return Addr1Generic(mode, pow2, c, size, sizeLog2);
}

template <class Cfg>
OMM_HD float TexLoad(const DevTexture& t, const DevMip& m, int x, int y) {  // ref: texture_impl.h:178-202
    const unsigned long long idx = m.texelOffset + (unsigned long long)((unsigned)x + (unsigned)y * (unsigned)m.w);
    if (Cfg::kFp32) return LoadRO((const float*)t.texels + idx);
    return (float)LoadRO((const uint8_t*)t.texels + idx) * (1.f / 255.f);
}
// texel (x,y) through address mode + border colour
template <class Cfg>
OMM_HD float TexFetch(const BakeParams& P, const DevMip& m, int cx, int cy) {
    if (Cfg::kAddr == kAddrGeneric && (cx == kTexCoordBorder || cy == kTexCoordBorder)) return P.borderAlpha;
    return TexLoad<Cfg>(P.tex, m, cx, cy);
}
OMM_HD float GlmLerp(float x, float y, float a) { return x * (1.f - a) + y * a; }

// ref: texture_impl.cpp:261-278 -- run-time bilinear point sample (per-mip pow2 flag).  The SDK reads out of bounds for
// Border addressing when the footprint leaves the texture; here such texels are borderAlpha (documented deviation on
// an input the SDK itself cannot process).
template <class Cfg>
OMM_HD float TexBilinear(const BakeParams& P, const DevMip& m, float2 p) {
    const float px = p.x * (float)m.w - 0.5f, py = p.y * (float)m.h - 0.5f;
    const float fx = floorf(px), fy = floorf(py);
    const int ix = f2i(fx), iy = f2i(fy);
    const int x0 = Addr1<Cfg>(P.addrMode, m.isPow2, ix, m.w, m.log2w), y0 = Addr1<Cfg>(P.addrMode, m.isPow2, iy, m.h, m.log2h);
    const int x1 = Addr1<Cfg>(P.addrMode, m.isPow2, ix + 1, m.w, m.log2w), y1 = Addr1<Cfg>(P.addrMode, m.isPow2, iy + 1, m.h, m.log2h);
    const float a = TexFetch<Cfg>(P, m, x0, y0);
    const float b = TexFetch<Cfg>(P, m, x0, y1);
    const float c = TexFetch<Cfg>(P, m, x1, y0);
    const float d = TexFetch<Cfg>(P, m, x1, y1);
    const float wx = px - fx, wy = py - fy;
    const float ac = GlmLerp(a, c, wx);
    const float bd = GlmLerp(b, d, wx);
    return GlmLerp(ac, bd, wy);
}

// ---- coverage -> state (ref: bake_kernels_cpu.h:25-61) -------------------------------------------------------------
OMM_HD int StateFromCoverage(const BakeParams& P, uint32_t above, uint32_t below) {
    if (above != 0 && below != 0) {
        if (P.globalFormat == ommFormat_OC1_4_State) {
            if (P.promotion == ommUnknownStatePromotion_ForceOpaque) return ommOpacityState_UnknownOpaque;
            if (P.promotion == ommUnknownStatePromotion_ForceTransparent) return ommOpacityState_UnknownTransparent;
            return (above >= below ? P.stateGT : P.stateLE) | 2;
        }
        if (P.promotion == ommUnknownStatePromotion_ForceOpaque) return ommOpacityState_Opaque;
        if (P.promotion == ommUnknownStatePromotion_ForceTransparent) return ommOpacityState_Transparent;
        return above >= below ? P.stateGT : P.stateLE;
    }
    if (above == 0) return P.stateLE;
    return P.stateGT;
}
OMM_HD bool IsUnknownState(int s) { return s == ommOpacityState_UnknownOpaque || s == ommOpacityState_UnknownTransparent; }

// ---- level-line test (ref: bake_kernels_cpu.h:115-238) -------------------------------------------------------------
OMM_HD bool IsZero(float v, float eps) { return v < eps && v > -eps; }
OMM_HD float Len2(float x, float y) { return sqrtf(x * x + y * y); }
OMM_HD bool InUnitSquare(float x, float y) { return x >= 0.f && x <= 1.f && y >= 0.f && y <= 1.f; }

// |p - p0| + |p - p1| - |p1 - p0| within 1e-5 (ref: bake_kernels_cpu.h:115-133).  The segment length is only ever used
// here, so it is computed on demand (the reference computes it eagerly; the value is the same).
OMM_HD bool PointOnEdge(float2 p0, float2 p1, float x, float y) {
    const float length = Len2(p1.x - p0.x, p1.y - p0.y);
    const float l = Len2(x - p0.x, y - p0.y) + Len2(x - p1.x, y - p1.y) - length;
    return IsZero(l, 1e-5f);
}

// Exact pre-filter for "x = RN(n / c0) lies in [0,1]" that avoids the division when the answer is certainly no:
//   * x <= 1  <=>  n/c0 <= 1 exactly: if n/c0 > 1 then |n| >= nextafter(|c0|), so n/c0 >= 1 + 2^-23/mant(c0) > 1 + 2^-24 and the
//     correctly rounded quotient is >= 1 + 2^-23 > 1; if n/c0 <= 1, RN is monotonic and RN(1) = 1.
//   * x >= 0 fails when n and c0 have strictly opposite signs, except when the quotient underflows to -0 (which compares
//     >= 0): that needs |n| <= 2^-150 |c0|, excluded here by requiring |n| > 2^-100.
// "true" means "cannot be decided cheaply or is inside": the caller then evaluates the reference expression itself.
OMM_HD bool QuotientMayBeInUnitRange(float n, float c0) {
    const bool le1 = c0 > 0.f ? (n <= c0) : (n >= c0);
    if (!le1) return false;                       // also rejects NaN numerators, like the reference's (x <= 1.f)
    const bool oppositeSigns = (n > 0.f && c0 < 0.f) || (n < 0.f && c0 > 0.f);
    if (oppositeSigns && fabsf(n) > 7.888609052e-31f) return false;
    return true;
}

// h = (a - cutoff, b, c, d); locals named as in the reference (ref: bake_kernels_cpu.h:144-238).
OMM_HD_NOINLINE bool EdgeHyperbola(float2 p0, float2 p1, float hx, float hy, float hz, float hw) {
    if (p0.x > p1.x) { const float2 t = p0; p0 = p1; p1 = t; }
    const float a = hx, b = hy, c = hz, d = hw;
    const float k_denum = p1.x - p0.x;
    if (IsZero(k_denum, 1e-6f)) {
        const float x = p0.x;
        const float n = x;
        const float c0 = d * n + c;
        const float c1 = a + b * n;
        if (IsZero(c0, 1e-6f)) return false;
        if (!(x >= 0.f && x <= 1.f)) return false;  // InUnitSquare would fail on x whatever y is
        const float y = -c1 / c0;
        return InUnitSquare(x, y) && PointOnEdge(p0, p1, x, y);
    }
    const float k_enum = p1.y - p0.y;
    const float k = k_enum / k_denum;
    const float m = p1.y - p1.x * k;
    const float c0 = d * k;
    const float c1 = c * k + d * m + b;
    const float c2 = a + c * m;
    if (IsZero(c0, 1e-6f)) {
        if (IsZero(c1, 1e-6f)) return false;
        if (!QuotientMayBeInUnitRange(-c2, c1)) return false;
        const float x = -c2 / c1;
        const float y = k * x + m;
        return InUnitSquare(x, y) && PointOnEdge(p0, p1, x, y);
    }
    const float innerRoot = c1 * c1 - 4.f * c0 * c2;
    if (innerRoot > 0.f) {
        const float root = sqrtf(innerRoot);
        const float n0 = 0.5f * (-c1 + root);
        const float n1 = 0.5f * (-c1 - root);
        bool hit = false;
        if (QuotientMayBeInUnitRange(n0, c0)) {
            const float x0 = n0 / c0;
            const float y0 = k * x0 + m;
            hit = InUnitSquare(x0, y0) && PointOnEdge(p0, p1, x0, y0);
        }
        if (!hit && QuotientMayBeInUnitRange(n1, c0)) {
            const float x1 = n1 / c0;
            const float y1 = k * x1 + m;
            hit = InUnitSquare(x1, y1) && PointOnEdge(p0, p1, x1, y1);
        }
        return hit;
    }
    return false;
}

struct Coverage {
    uint32_t above, below;
};

// ref: bake_kernels_cpu.h:241-399.  `tri` is the micro-triangle in UV space (original winding).
template <class Cfg, bool kDegenerate>
OMM_HD void LevelLineCell(const BakeParams& P, const DevMip& m, const Tri& tri, int px, int py, Coverage& cov) {
    const float pfx = (float)px + 0.5f, pfy = (float)py + 0.5f;
    const int x0 = Addr1<Cfg>(P.addrMode, P.pow2Mip0, px, m.w, m.log2w), y0 = Addr1<Cfg>(P.addrMode, P.pow2Mip0, py, m.h, m.log2h);
    const int x1 = Addr1<Cfg>(P.addrMode, P.pow2Mip0, px + 1, m.w, m.log2w), y1 = Addr1<Cfg>(P.addrMode, P.pow2Mip0, py + 1, m.h, m.log2h);
    // gatherRed = (c00, c01, c11, c10)
    const float gx = TexFetch<Cfg>(P, m, x0, y0);
    const float gy = TexFetch<Cfg>(P, m, x0, y1);
    const float gz = TexFetch<Cfg>(P, m, x1, y1);
    const float gw = TexFetch<Cfg>(P, m, x1, y0);
    if (!kDegenerate) {
        const float ipx = pfx * m.rcpw, ipy = pfy * m.rcph;
        const bool o0 = P.cutoff < gx, o1 = P.cutoff < gy, o2 = P.cutoff < gz, o3 = P.cutoff < gw;
        const bool in0 = PointInTri(tri, ipx, ipy);
        const bool in1 = PointInTri(tri, ipx + 0.0f, ipy + m.rcph);
        const bool in2 = PointInTri(tri, ipx + m.rcpw, ipy + m.rcph);
        const bool in3 = PointInTri(tri, ipx + m.rcpw, ipy + 0.0f);
        const bool isOpaque = (in0 && o0) || (in1 && o1) || (in2 && o2) || (in3 && o3);
        const bool isTransparent = (in0 && !o0) || (in1 && !o1) || (in2 && !o2) || (in3 && !o3);
        if (isOpaque) cov.above += 1;
        if (isTransparent) cov.below += 1;
        if (isOpaque && isTransparent) return;
    }
    const float a = gx;
    const float b = gw - gx;
    const float c = gy - gx;
    const float d = gx + gz - gy - gw;
    if (IsZero(b, 1e-6f) && IsZero(c, 1e-6f) && IsZero(d, 1e-6f)) {
        if (P.cutoff < a) cov.above += 1;
        else cov.below += 1;
        return;
    }
    const float sx = (float)m.w, sy = (float)m.h;
    const float h0 = a - P.cutoff;
    if (kDegenerate) {
        const float2 e0 = make_float2(sx * tri.aabb_s.x - pfx, sy * tri.aabb_s.y - pfy);
        const float2 e1 = make_float2(sx * tri.aabb_e.x - pfx, sy * tri.aabb_e.y - pfy);
        if (EdgeHyperbola(e0, e1, h0, b, c, d)) { cov.above += 1; cov.below += 1; }
    } else {
        const float2 q0 = make_float2(sx * tri.p0.x - pfx, sy * tri.p0.y - pfy);
        const float2 q1 = make_float2(sx * tri.p1.x - pfx, sy * tri.p1.y - pfy);
        const float2 q2 = make_float2(sx * tri.p2.x - pfx, sy * tri.p2.y - pfy);
        if (EdgeHyperbola(q0, q1, h0, b, c, d) || EdgeHyperbola(q1, q2, h0, b, c, d) || EdgeHyperbola(q2, q0, h0, b, c, d)) {
            cov.above += 1; cov.below += 1;
        }
    }
}

// ref: bake_kernels_cpu.h:404-452 (only reachable through internal flag bits 7/8)
template <class Cfg>
OMM_HD void ConservativeBilinearCell(const BakeParams& P, const DevMip& m, int px, int py, Coverage& cov) {
    const float pfx = (float)px + 0.5f, pfy = (float)py + 0.5f;
    const int ix = f2i(pfx), iy = f2i(pfy);
    const int x0 = Addr1<Cfg>(P.addrMode, P.pow2Mip0, ix, m.w, m.log2w), y0 = Addr1<Cfg>(P.addrMode, P.pow2Mip0, iy, m.h, m.log2h);
    const int x1 = Addr1<Cfg>(P.addrMode, P.pow2Mip0, ix + 1, m.w, m.log2w), y1 = Addr1<Cfg>(P.addrMode, P.pow2Mip0, iy + 1, m.h, m.log2h);
    const float gx = TexFetch<Cfg>(P, m, x0, y0), gy = TexFetch<Cfg>(P, m, x0, y1), gz = TexFetch<Cfg>(P, m, x1, y1), gw = TexFetch<Cfg>(P, m, x1, y0);
    const float mn = fminStd(fminStd(fminStd(gx, gy), gz), gw);
    const float mx = fmaxStd(fmaxStd(fmaxStd(gx, gy), gz), gw);
    if (P.cutoff < mx) cov.above += 1;
    if (P.cutoff > mn) cov.below += 1;
}
// ref: bake_cpu_impl.cpp:994-1009
template <class Cfg>
OMM_HD void NearestCell(const BakeParams& P, const DevMip& m, int px, int py, Coverage& cov) {
    const int cx = Addr1<Cfg>(P.addrMode, P.pow2Mip0, px, m.w, m.log2w), cy = Addr1<Cfg>(P.addrMode, P.pow2Mip0, py, m.h, m.log2h);
    const float alpha = TexFetch<Cfg>(P, m, cx, cy);
    if (P.cutoff < alpha) cov.above += 1;
    else cov.below += 1;
}

// ---- rasterizers (ref: util/cpu_raster.h) ----------------------------------------------------------------------------
struct EdgeFn {
    float nx, ny, c;
};
OMM_HD EdgeFn MakeEdgeFn(float2 p, float2 q) {  // ref: util/cpu_raster.h:26-29
    EdgeFn e;
    e.nx = q.y - p.y;
    e.ny = p.x - q.x;
    e.c = -(e.nx * p.x + e.ny * p.y);
    return e;
}
OMM_HD float EvalEdgeCons(const EdgeFn& e, float sx, float sy) {  // ref: util/cpu_raster.h:46-51, ext=(1,1)
    const float ev = (e.nx * sx + e.ny * sy) + e.c;
    const float bx = e.nx > 0 ? 0.f : e.nx;
    const float by = e.ny > 0 ? 0.f : e.ny;
    return ev + bx * 1.f + by * 1.f;
}

// Raster-space set-up of the conservative triangle rasterizer (ref: util/cpu_raster.h:278-306).
struct RasterSetup {
    EdgeFn e0, e1, e2;
    int minx, miny, maxx, maxy;
};
OMM_HD RasterSetup MakeRasterSetup(const Tri& t_, int rw, int rh, float off) {
    const bool ccw = TriIsCCW(t_.p0, t_.p1, t_.p2);
    const float rfx = (float)rw, rfy = (float)rh;
    const float2 a = make_float2(t_.p0.x * rfx + off, t_.p0.y * rfy + off);
    const float2 b = make_float2(t_.p1.x * rfx + off, t_.p1.y * rfy + off);
    const float2 c = make_float2(t_.p2.x * rfx + off, t_.p2.y * rfy + off);
    const float2 q0 = ccw ? a : c, q1 = b, q2 = ccw ? c : a;
    RasterSetup r;
    r.minx = f2i(floorf(fminStd(fminStd(q0.x, q1.x), q2.x)));
    r.miny = f2i(floorf(fminStd(fminStd(q0.y, q1.y), q2.y)));
    r.maxx = f2i(ceilf(fmaxStd(fmaxStd(q0.x, q1.x), q2.x)));
    r.maxy = f2i(ceilf(fmaxStd(fmaxStd(q0.y, q1.y), q2.y)));
    r.e0 = MakeEdgeFn(q0, q1);
    r.e1 = MakeEdgeFn(q1, q2);
    r.e2 = MakeEdgeFn(q2, q0);
    return r;
}
OMM_HD bool CellInside(const RasterSetup& r, int x, int y) {
    const float sx = (float)x, sy = (float)y;
    return EvalEdgeCons(r.e0, sx, sy) < 0.f && EvalEdgeCons(r.e1, sx, sy) < 0.f && EvalEdgeCons(r.e2, sx, sy) < 0.f;
}

// Serial over-conservative raster with the reference's row scan ("stop the row at the first exit after an entry").
// f(x, y) returns true to abort the whole raster (used for the exact early-out, see ClassifyMicroTriangle).
template <class F>
OMM_HD bool RasterTriConservative(const Tri& t, int rw, int rh, float off, F&& f) {
    const RasterSetup r = MakeRasterSetup(t, rw, rh, off);
    for (int y = r.miny; y < r.maxy; ++y) {
        bool wasInside = false;
        for (int x = r.minx; x < r.maxx; ++x) {
            if (CellInside(r, x, y)) {
                if (f(x, y)) return true;
                wasInside = true;
            } else if (wasInside)
                break;
        }
    }
    return false;
}

// The same walk as RasterTriConservative, as a resumable iterator: yields the covered cells one at a time in the
// reference's visiting order (rows bottom-up, each row left to right until the first exit after an entry).
struct RasterCursor {
    int x, y;
    bool wasInside;
};
OMM_HD RasterCursor RasterBegin(const RasterSetup& r) { return RasterCursor{r.minx, r.miny, false}; }
OMM_HD bool RasterNext(const RasterSetup& r, RasterCursor& c, int& ox, int& oy) {
    while (c.y < r.maxy) {
        while (c.x < r.maxx) {
            if (CellInside(r, c.x, c.y)) {
                ox = c.x;
                oy = c.y;
                c.wasInside = true;
                ++c.x;
                return true;
            }
            if (c.wasInside) break;
            ++c.x;
        }
        ++c.y;
        c.x = r.minx;
        c.wasInside = false;
    }
    return false;
}

// ref: util/cpu_raster.h:486-555 (conservative DDA along a segment)
template <class F>
OMM_HD bool RasterLineConservative(float2 lp0, float2 lp1, int rw, int rh, float off, F&& f) {
    const float rfx = (float)rw, rfy = (float)rh;
    float2 p0 = make_float2(lp0.x * rfx + off, lp0.y * rfy + off);
    float2 p1 = make_float2(lp1.x * rfx + off, lp1.y * rfy + off);
    if (p0.x > p1.x) { const float2 t = p0; p0 = p1; p1 = t; }
    const float dx = p1.x - p0.x, dy = p1.y - p0.y;
    int x = f2i(floorf(p0.x)), y = f2i(floorf(p0.y));
    const int stepX = (dx > 0) ? 1 : ((dx < 0) ? -1 : 0);
    const int stepY = (dy > 0) ? 1 : ((dy < 0) ? -1 : 0);
    const float inf = UintAsFloat(0x7f800000u);
    const float tDeltaX = (stepX != 0) ? 1.f / fabsf(dx) : inf;
    const float tDeltaY = (stepY != 0) ? 1.f / fabsf(dy) : inf;
    float tMaxX = inf, tMaxY = inf;
    if (stepX != 0) tMaxX = (((float)x + (stepX > 0 ? 1.f : 0.f)) - p0.x) / dx;
    if (stepY != 0) tMaxY = (((float)y + (stepY > 0 ? 1.f : 0.f)) - p0.y) / dy;
    if (stepX == 0 && stepY == 0) return f(x, y);
    const int yMin = f2i(fminStd(floorf(p0.y), floorf(p1.y))), yMax = f2i(fmaxStd(ceilf(p0.y), ceilf(p1.y)));
    const int xMin = f2i(fminStd(floorf(p0.x), floorf(p1.x))), xMax = f2i(fmaxStd(ceilf(p0.x), ceilf(p1.x)));
    while (x >= xMin && x <= xMax && y >= yMin && y <= yMax) {
        if (f(x, y)) return true;
        if (tMaxX < tMaxY) { x += stepX; tMaxX += tDeltaX; }
        else { y += stepY; tMaxY += tDeltaY; }
    }
    return false;
}

// ---- coarse SAT classification of one micro-triangle (ref: bake_cpu_impl.cpp:749-801) ----------------------------
// returns -1 when the coarse pass leaves the micro-triangle untouched, else the state it sets.
template <class Cfg>
OMM_HD int CoarseState(const BakeParams& P, const Tri& st) {
    const DevMip& m = P.tex.mips[0];
    if (f2i(st.aabb_s.x) != f2i(st.aabb_e.x) || f2i(st.aabb_s.y) != f2i(st.aabb_e.y)) return -1;
    const float fsx = st.aabb_s.x * (float)m.w - 0.5f, fsy = st.aabb_s.y * (float)m.h - 0.5f;
    const float fex = st.aabb_e.x * (float)m.w - 0.5f, fey = st.aabb_e.y * (float)m.h - 0.5f;
    const int sx = Addr1<Cfg>(P.addrMode, P.pow2Mip0, f2i(floorf(fsx)), m.w, m.log2w);
    const int sy = Addr1<Cfg>(P.addrMode, P.pow2Mip0, f2i(floorf(fsy)), m.h, m.log2h);
    const int ex = Addr1<Cfg>(P.addrMode, P.pow2Mip0, f2i(floorf(fex)) + 1, m.w, m.log2w);
    const int ey = Addr1<Cfg>(P.addrMode, P.pow2Mip0, f2i(floorf(fey)) + 1, m.h, m.log2h);
    if (ex < sx || ey < sy) return -1;
    if (sx < 0 || sy < 0 || sx >= m.w || sy >= m.h) return -1;
    if (ex < 0 || ey < 0 || ex >= m.w || ey >= m.h) return -1;
    const uint32_t area = (uint32_t)((ex - sx + 1) * (ey - sy + 1));
    const uint32_t* sat = P.tex.sat + m.satOffset;
    const int sx1 = sx - 1, sy1 = sy - 1;  // ref: texture_impl.h:108-125
    const uint32_t A = (sx1 >= 0 && sy1 >= 0) ? LoadRO(sat + sx1 + (size_t)sy1 * m.w) : 0;
    const uint32_t B = (sy1 >= 0) ? LoadRO(sat + ex + (size_t)sy1 * m.w) : 0;
    const uint32_t C = (sx1 >= 0) ? LoadRO(sat + sx1 + (size_t)ey * m.w) : 0;
    const uint32_t D = LoadRO(sat + ex + (size_t)ey * m.w);
    const uint32_t sa = D + A - B - C;
    if (sa == 0) return P.stateLE;
    if (sa == area) return P.stateGT;
    return -1;
}

// ---- one micro-triangle, start to finish (ref: bake_cpu_impl.cpp:716-1029) ------------------------------------------
// Exact early-out: under ForceOpaque / ForceTransparent the final state depends only on whether both counters are
// non-zero (bake_kernels_cpu.h:27-50), counters never decrease, and every later mip only adds to them, so the walk can
// stop the moment both are non-zero.  Under Nearest promotion the counts matter and the full walk is done.
template <class Cfg>
OMM_HD int ClassifyMicroTriangle(const BakeParams& P, float2 b0, float2 b1, float2 b2, bool baseDegenerate, uint32_t index,
                                                    uint32_t level) {
    const Tri st = MicroTri(b0, b1, b2, index, level);
    int state = ommOpacityState_UnknownOpaque;
    if (P.useCoarse) {
        const int cs = CoarseState<Cfg>(P, st);
        if (cs >= 0) state = cs;
    }
    if (P.disableFine) return state;
    Coverage cov{0u, 0u};
    const bool earlyOut = P.promotion != ommUnknownStatePromotion_Nearest;
    if (P.filterLinear) {
        if (state != ommOpacityState_UnknownOpaque) return state;
        if (!P.disableLevelLine) {
            for (int mip = 0; mip < P.tex.mipCount; ++mip) {
                const DevMip& m = P.tex.mips[mip];
                if (P.cutoff < TexBilinear<Cfg>(P, m, st.p0)) cov.above++;
                else cov.below++;
                bool stop;
                if (!baseDegenerate) {
                    stop = RasterTriConservative(st, m.w, m.h, -0.5f, [&](int x, int y) {
                        LevelLineCell<Cfg, false>(P, m, st, x, y, cov);
                        return earlyOut && cov.above != 0 && cov.below != 0;
                    });
                } else {
                    stop = RasterLineConservative(st.aabb_s, st.aabb_e, m.w, m.h, -0.5f, [&](int x, int y) {
                        LevelLineCell<Cfg, true>(P, m, st, x, y, cov);
                        return earlyOut && cov.above != 0 && cov.below != 0;
                    });
                }
                if (stop) break;
                if (IsUnknownState(StateFromCoverage(P, cov.above, cov.below))) break;
            }
        } else if (P.aabbTesting) {
            const DevMip& m = P.tex.mips[0];
            const float2 c1 = make_float2(st.aabb_e.x, st.aabb_s.y), c2 = make_float2(st.aabb_s.x, st.aabb_e.y);
            const Tri t0 = MakeTri(st.aabb_s, c1, c2), t1 = MakeTri(st.aabb_e, c1, c2);
            RasterTriConservative(t0, m.w, m.h, -0.5f, [&](int x, int y) { ConservativeBilinearCell<Cfg>(P, m, x, y, cov); return false; });
            RasterTriConservative(t1, m.w, m.h, -0.5f, [&](int x, int y) { ConservativeBilinearCell<Cfg>(P, m, x, y, cov); return false; });
        } else {
            const DevMip& m = P.tex.mips[0];
            RasterTriConservative(st, m.w, m.h, -0.5f, [&](int x, int y) { ConservativeBilinearCell<Cfg>(P, m, x, y, cov); return false; });
        }
    } else {
        for (int mip = 0; mip < P.tex.mipCount; ++mip) {
            const DevMip& m = P.tex.mips[mip];
            const bool stop = RasterTriConservative(st, m.w, m.h, 0.f, [&](int x, int y) {
                NearestCell<Cfg>(P, m, x, y, cov);
                return earlyOut && cov.above != 0 && cov.below != 0;
            });
            if (stop) break;
            if (IsUnknownState(StateFromCoverage(P, cov.above, cov.below))) break;
        }
    }
    return StateFromCoverage(P, cov.above, cov.below);
}

}  // namespace ommb200
