/*
 * omm_oracle.c -- TEST INFRASTRUCTURE ONLY.  Single-threaded plain-C restatement of the Opacity
 * Micro-Map SDK 1.9.0 CPU bake path (ommCpuBake and the texture object it reads).
 *
 * This file is the parity oracle ("port") for libomm-b200.so.  It is never linked into, imported
 * by, or called from the product; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  It exports the same C ABI as the SDK (include/omm_b200.h) so
 * one harness can drive the SDK build (oracle/_ref/libomm-lib.so), this port and the CUDA library.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py byte-compares this port with the unmodified
 * SDK built from /root/reference (oracle/Makefile -> oracle/_ref/libomm-lib.so) on every workload the
 * suite uses, and tests/test_kat_reference.py replays the SDK's own known-answer counts
 * (support/tests/test_omm_bake_cpu.cpp) against it through tests/golden/.
 *
 * Every function cites the reference lines it restates ("ref:" paths are relative to
 * /root/reference/libraries/omm-lib/src).  Floating point: the SDK is built -O3 -msse4.1 without FMA
 * or fast-math, so every expression below keeps the reference's operation order and this file must
 * be compiled with -ffp-contract=off (see oracle/Makefile).
 *
 * Not restated (the port returns ommResult_NOT_IMPLEMENTED): near-duplicate merge
 * (DeduplicateSimilarLSH / BruteForce, ref: bake_cpu_impl.cpp:1134-1430) and the size-budget
 * Compress pass (ref: :1557-1688).  They are "next" rows in SURVEY.md section 8f.
 */
#include "../include/omm_b200.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define OAPI __attribute__((visibility("default")))

typedef struct { float x, y; } f2;
typedef struct { int x, y; } i2;

/* ------------------------------------------------------------------------------------------ */
/* Texture object.  ref: texture_impl.cpp:77-224 (Create), texture_impl.h:178-223 (Load/Bilinear) */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int w, h;
    int log2w, log2h;   /* ctz(size), ref: util/bit_tricks.h:66-78 (only meaningful for pow2 sizes) */
    int isPow2;
    float rcpw, rcph;   /* 1.f / size, ref: texture_impl.cpp:102 */
    void* texels;       /* row-major copy, FP32 or UNORM8 */
    uint32_t* sat;      /* inclusive summed-area table of (alpha > cutoff), row-major, or NULL */
} OMip;

typedef struct {
    uint32_t magic;
    ommCpuTextureFormat format;
    ommCpuTextureFlags flags;
    float alphaCutoff;
    uint32_t mipCount;
    OMip* mips;
} OTexture;

typedef struct {
    uint32_t magic;
    ommMessageInterface log;
} OBaker;

#define OBAKER_MAGIC 0x0b4ce501u
#define OTEX_MAGIC 0x07e87001u

static int ctz_slow(uint32_t n) { /* ref: util/bit_tricks.h:66-78 */
    if (n == 0) return 32;
    int c = 0;
    while ((n & 1u) == 0) { c++; n >>= 1; }
    return c;
}
static int is_pow2(int x) { return x > 0 && !(x & (x - 1)); } /* ref: util/bit_tricks.h:36-38 */

static float tex_load(const OTexture* t, int mip, int x, int y) { /* ref: texture_impl.h:178-202 */
    const OMip* m = &t->mips[mip];
    size_t idx = (size_t)x + (size_t)y * (size_t)m->w;
    if (t->format == ommCpuTextureFormat_FP32) return ((const float*)m->texels)[idx];
    return (float)((const uint8_t*)m->texels)[idx] * (1.f / 255.f);
}

/* ------------------------------------------------------------------------------------------ */
/* Address modes.  ref: util/texture.h:35-91                                                    */
/* ------------------------------------------------------------------------------------------ */
#define TEXCOORD_BORDER 0x7FFFFFFE

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

static int addr1(ommTextureAddressMode mode, int pow2, int c, int size, int sizeLog2) {
    switch (mode) {
    case ommTextureAddressMode_Wrap:
        if (pow2) return (int)((uint32_t)c & (uint32_t)(size - 1));
        return (int)((uint32_t)c % (uint32_t)size);
    case ommTextureAddressMode_Mirror:
        if (pow2) {
            int a = abs(c) - (c < 0);
            int flipped = (a >> sizeLog2) & 1;
            int wrapped = (int)((uint32_t)a & (uint32_t)(size - 1));
            return flipped ? size - wrapped - 1 : wrapped;
        } else {
            int a = (int)fabsf((float)c + 0.5f);
            uint32_t flipped = ((uint32_t)(a / size)) % 2u;
            int wrapped = (int)((uint32_t)a % (uint32_t)size);
            return flipped ? size - wrapped - 1 : wrapped;
        }
    case ommTextureAddressMode_Clamp:
        return clampi(c, 0, size - 1);
    case ommTextureAddressMode_Border:
        return (c >= size || c < 0) ? TEXCOORD_BORDER : c;
    case ommTextureAddressMode_MirrorOnce: {
        int a = (int)fabsf((float)c + 0.5f);
        return clampi(a, 0, size - 1);
    }
    default:
        return 0x7FFFFFFF;
    }
}

/* ref: util/texture.h:123-153 -- coords of texel (x,y) and (x+1,y+1) through the address mode. */
static void gather4(ommTextureAddressMode mode, int pow2, int x, int y, const OMip* m, i2* c00, i2* c10, i2* c01, i2* c11) {
    int ox = addr1(mode, pow2, x, m->w, m->log2w), oy = addr1(mode, pow2, y, m->h, m->log2h);
    int px = addr1(mode, pow2, x + 1, m->w, m->log2w), py = addr1(mode, pow2, y + 1, m->h, m->log2h);
    c00->x = ox; c00->y = oy;
    c10->x = px; c10->y = oy;
    c01->x = ox; c01->y = py;
    c11->x = px; c11->y = py;
}

static int is_border(i2 c) { return c.x == TEXCOORD_BORDER || c.y == TEXCOORD_BORDER; }

static float glm_lerp(float x, float y, float a) { return x * (1.f - a) + y * a; } /* ref: glm func_common.inl:160 */

/* ref: texture_impl.cpp:261-278 -- run-time (non-template) bilinear point sample; uses the per-mip pow2 flag.
 * The SDK has no border handling here (Border + out-of-range footprint reads out of bounds there);
 * the port substitutes borderAlpha so that it stays defined.  Such inputs are excluded from parity configs. */
static float tex_bilinear(const OTexture* t, ommTextureAddressMode mode, float borderAlpha, f2 p, int mip) {
    const OMip* m = &t->mips[mip];
    float px = p.x * (float)m->w - 0.5f, py = p.y * (float)m->h - 0.5f;
    float fx = floorf(px), fy = floorf(py);
    i2 c00, c10, c01, c11;
    gather4(mode, m->isPow2, (int)fx, (int)fy, m, &c00, &c10, &c01, &c11);
    float a = is_border(c00) ? borderAlpha : tex_load(t, mip, c00.x, c00.y);
    float b = is_border(c01) ? borderAlpha : tex_load(t, mip, c01.x, c01.y);
    float c = is_border(c10) ? borderAlpha : tex_load(t, mip, c10.x, c10.y);
    float d = is_border(c11) ? borderAlpha : tex_load(t, mip, c11.x, c11.y);
    float wx = px - fx, wy = py - fy; /* glm::fract = x - floor(x) */
    float ac = glm_lerp(a, c, wx);
    float bd = glm_lerp(b, d, wx);
    return glm_lerp(ac, bd, wy);
}

/* ref: texture_impl.h:108-125 */
static uint32_t tex_sat(const OMip* m, i2 s, i2 e) {
    int sx1 = s.x - 1, sy1 = s.y - 1;
    uint32_t A = (sx1 >= 0 && sy1 >= 0) ? m->sat[sx1 + sy1 * m->w] : 0;
    uint32_t B = (sy1 >= 0) ? m->sat[e.x + sy1 * m->w] : 0;
    uint32_t C = (sx1 >= 0) ? m->sat[sx1 + e.y * m->w] : 0;
    uint32_t D = m->sat[e.x + e.y * m->w];
    return (uint32_t)(int32_t)(D + A - B - C);
}
static int in_texture(const OMip* m, i2 c) { return c.x >= 0 && c.y >= 0 && c.x < m->w && c.y < m->h; }

/* ------------------------------------------------------------------------------------------ */
/* Triangle + bird curve.  ref: util/geometry.h:59-116, util/bird.h:36-182                      */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    f2 p0, p1, p2;
    f2 p0p2, p1p0, p2p1;
    f2 aabb_s, aabb_e;
} Tri;

static float minf2(float a, float b) { return b < a ? b : a; } /* std::min */
static float maxf2(float a, float b) { return a < b ? b : a; } /* std::max */

static Tri make_tri(f2 p0, f2 p1, f2 p2) { /* ref: util/geometry.h:63-75 */
    Tri t;
    t.p0 = p0; t.p1 = p1; t.p2 = p2;
    t.p0p2.x = p0.x - p2.x; t.p0p2.y = p0.y - p2.y;
    t.p1p0.x = p1.x - p0.x; t.p1p0.y = p1.y - p0.y;
    t.p2p1.x = p2.x - p1.x; t.p2p1.y = p2.y - p1.y;
    t.aabb_s.x = minf2(minf2(p0.x, p1.x), p2.x); t.aabb_s.y = minf2(minf2(p0.y, p1.y), p2.y);
    t.aabb_e.x = maxf2(maxf2(p0.x, p1.x), p2.x); t.aabb_e.y = maxf2(maxf2(p0.y, p1.y), p2.y);
    return t;
}

static int tri_is_invalid(const Tri* t) { /* ref: util/geometry.h:37-42 */
    float v[6] = {t->p0.x, t->p0.y, t->p1.x, t->p1.y, t->p2.x, t->p2.y};
    for (int i = 0; i < 6; ++i)
        if (isnan(v[i]) || isinf(v[i])) return 1;
    return 0;
}
static int tri_is_degenerate(const Tri* t) { /* ref: util/geometry.h:44-47 (float expression vs double 1e-9) */
    float area = 0.5f * fabsf(t->p0.x * (t->p1.y - t->p2.y) + t->p1.x * (t->p2.y - t->p0.y) + t->p2.x * (t->p0.y - t->p1.y));
    return area < 1e-9;
}
static int tri_is_ccw(f2 p0, f2 p1, f2 p2) { /* ref: util/geometry.h:49-55 (double cross of float differences) */
    double ax = (double)(p2.x - p0.x), ay = (double)(p2.y - p0.y);
    double bx = (double)(p1.x - p0.x), by = (double)(p1.y - p0.y);
    double nz = ax * by - bx * ay;
    return nz < 0;
}
static int point_in_tri(const Tri* t, f2 pt) { /* ref: util/geometry.h:101-114 */
    float ptp2x = pt.x - t->p2.x, ptp2y = pt.y - t->p2.y;
    float ptp0x = pt.x - t->p0.x, ptp0y = pt.y - t->p0.y;
    float s = t->p0p2.x * ptp2y - t->p0p2.y * ptp2x;
    float tt = t->p1p0.x * ptp0y - t->p1p0.y * ptp0x;
    if ((s < 0) != (tt < 0) && s != 0 && tt != 0) return 0;
    float ptp1x = pt.x - t->p1.x, ptp1y = pt.y - t->p1.y;
    float d = t->p2p1.x * ptp1y - t->p2p1.y * ptp1x;
    return d == 0 || (d < 0) == (s + tt <= 0);
}

static uint32_t extract_even_bits(uint32_t x) { /* ref: util/bird.h:36-44 */
    x &= 0x55555555u;
    x = (x | (x >> 1)) & 0x33333333u;
    x = (x | (x >> 2)) & 0x0f0f0f0fu;
    x = (x | (x >> 4)) & 0x00ff00ffu;
    x = (x | (x >> 8)) & 0x0000ffffu;
    return x;
}
static uint32_t prefix_eor(uint32_t x) { /* ref: util/bird.h:47-54 */
    x ^= x >> 1; x ^= x >> 2; x ^= x >> 4; x ^= x >> 8;
    return x;
}
/* ref: util/bird.h:57-118 -- bird-curve index -> barycentrics of the three micro-vertices */
static void index2bary(uint32_t index, uint32_t level, f2* uv0, f2* uv1, f2* uv2) {
    if (level == 0) {
        uv0->x = 0; uv0->y = 0; uv1->x = 1; uv1->y = 0; uv2->x = 0; uv2->y = 1;
        return;
    }
    uint32_t b0 = extract_even_bits(index), b1 = extract_even_bits(index >> 1);
    uint32_t fx = prefix_eor(b0), fy = prefix_eor(b0 & ~b1);
    uint32_t t = fy ^ b1;
    uint32_t iu = (fx & ~t) | (b0 & ~t) | (~b0 & ~fx & t);
    uint32_t iv = fy ^ b0;
    uint32_t iw = (~fx & ~t) | (b0 & ~t) | (~b0 & fx & t);
    uint32_t mask = (1u << level) - 1u;
    iu &= mask; iv &= mask; iw &= mask;
    int upright = (iu & 1) ^ (iv & 1) ^ (iw & 1);
    if (!upright) { iu += 1; iv += 1; }
    union { uint32_t u; float f; } sc;
    sc.u = (127u - level) << 23;
    float levelScale = sc.f;
    float du = 1.f * levelScale, dv = 1.f * levelScale;
    float u = (float)iu * levelScale, v = (float)iv * levelScale;
    if (!upright) { du = -du; dv = -dv; }
    uv0->x = u; uv0->y = v;
    uv1->x = u + du; uv1->y = v;
    uv2->x = u; uv2->y = v + dv;
}
/* ref: util/geometry.h:241-248 */
static f2 interp_uv(f2 uv, const Tri* t) {
    float bx = 1.f - uv.x - uv.y, by = uv.x, bz = uv.y;
    f2 r;
    r.x = t->p0.x * bx + t->p1.x * by + t->p2.x * bz;
    r.y = t->p0.y * bx + t->p1.y * by + t->p2.y * bz;
    return r;
}
static Tri micro_tri(const Tri* t, uint32_t index, uint32_t level) { /* ref: util/bird.h:170-182 */
    f2 uv0, uv1, uv2;
    index2bary(index, level, &uv0, &uv1, &uv2);
    return make_tri(interp_uv(uv0, t), interp_uv(uv1, t), interp_uv(uv2, t));
}

/* ------------------------------------------------------------------------------------------ */
/* Classification kernels.  ref: bake_kernels_cpu.h                                             */
/* ------------------------------------------------------------------------------------------ */
typedef struct { uint32_t above, below; } Coverage;

static ommOpacityState state_from_coverage(ommFormat fmt, ommUnknownStatePromotion mode, ommOpacityState gt, ommOpacityState le, Coverage c) {
    /* ref: bake_kernels_cpu.h:25-61 */
    if (c.above != 0 && c.below != 0) {
        if (fmt == ommFormat_OC1_4_State) {
            if (mode == ommUnknownStatePromotion_ForceOpaque) return ommOpacityState_UnknownOpaque;
            if (mode == ommUnknownStatePromotion_ForceTransparent) return ommOpacityState_UnknownTransparent;
            return (ommOpacityState)((c.above >= c.below ? gt : le) | 2u);
        }
        if (mode == ommUnknownStatePromotion_ForceOpaque) return ommOpacityState_Opaque;
        if (mode == ommUnknownStatePromotion_ForceTransparent) return ommOpacityState_Transparent;
        return c.above >= c.below ? gt : le;
    }
    if (c.above == 0) return le;
    return gt;
}
static int is_unknown(ommOpacityState s) { return s == ommOpacityState_UnknownOpaque || s == ommOpacityState_UnknownTransparent; }
static int is_known(ommOpacityState s) { return s == ommOpacityState_Opaque || s == ommOpacityState_Transparent; }

static int is_zero(float v, float eps) { return v < eps && v > -eps; } /* ref: bake_kernels_cpu.h:135-137 */
static float len2(float x, float y) { return sqrtf(x * x + y * y); }   /* glm::length(vec2) */
static int in_unit_square(float x, float y) { return x >= 0.f && x <= 1.f && y >= 0.f && y <= 1.f; }

typedef struct { f2 p0, p1; float length; } Edge;
static int point_on_edge(const Edge* e, float x, float y) { /* ref: bake_kernels_cpu.h:125-128 */
    float l = len2(x - e->p0.x, y - e->p0.y) + len2(x - e->p1.x, y - e->p1.y) - e->length;
    return is_zero(l, 1e-5f);
}

/* ref: bake_kernels_cpu.h:144-238.  h = (a - cutoff, b, c, d) of the bilinear patch; the function's local names
 * a,b,c,d are h.x,h.y,h.z,h.w exactly as in the reference. */
static int edge_hyperbola(f2 p0, f2 p1, float hx, float hy, float hz, float hw) {
    if (p0.x > p1.x) { f2 tmp = p0; p0 = p1; p1 = tmp; }
    Edge edge;
    edge.p0 = p0; edge.p1 = p1; edge.length = len2(p1.x - p0.x, p1.y - p0.y);
    const float a = hx, b = hy, c = hz, d = hw;
    const float k_denum = p1.x - p0.x;
    if (is_zero(k_denum, 1e-6f)) {
        const float x = p0.x;
        const float n = x;
        const float c0 = d * n + c;
        const float c1 = a + b * n;
        if (is_zero(c0, 1e-6f)) return 0;
        const float y = -c1 / c0;
        return in_unit_square(x, y) && point_on_edge(&edge, x, y);
    } else {
        const float k_enum = p1.y - p0.y;
        const float k = k_enum / k_denum;
        const float m = p1.y - p1.x * k;
        const float c0 = d * k;
        const float c1 = c * k + d * m + b;
        const float c2 = a + c * m;
        if (is_zero(c0, 1e-6f)) {
            if (is_zero(c1, 1e-6f)) return 0;
            const float x = -c2 / c1;
            const float y = k * x + m;
            return in_unit_square(x, y) && point_on_edge(&edge, x, y);
        } else {
            const float innerRoot = c1 * c1 - 4 * c0 * c2;
            if (innerRoot > 0.f) {
                const float root = sqrtf(innerRoot);
                const float x0 = 0.5f * (-c1 + root) / c0;
                const float x1 = 0.5f * (-c1 - root) / c0;
                const float y0 = k * x0 + m;
                const float y1 = k * x1 + m;
                const int i0 = in_unit_square(x0, y0) && point_on_edge(&edge, x0, y0);
                const int i1 = in_unit_square(x1, y1) && point_on_edge(&edge, x1, y1);
                return i0 || i1;
            }
            return 0;
        }
    }
}

typedef struct {
    const OTexture* tex;
    const Tri* tri; /* micro-triangle in UV space, original winding */
    int mip;
    int pow2;       /* pow2 flag of mip 0 (the SDK's template parameter, ref: bake_cpu_impl.cpp:299) */
    ommTextureAddressMode mode;
    float cutoff, borderAlpha;
    Coverage* cov;
} KParams;

/* ref: bake_kernels_cpu.h:241-399 */
static void level_line_kernel(int px, int py, int degenerate, KParams* p) {
    const OMip* m = &p->tex->mips[p->mip];
    const float pfx = (float)px + 0.5f, pfy = (float)py + 0.5f;
    const float ipx = pfx * m->rcpw, ipy = pfy * m->rcph;
    i2 c00, c10, c01, c11;
    gather4(p->mode, p->pow2, px, py, m, &c00, &c10, &c01, &c11);
    float gx, gy, gz, gw;
    if (p->mode == ommTextureAddressMode_Border) {
        gx = is_border(c00) ? p->borderAlpha : tex_load(p->tex, p->mip, c00.x, c00.y);
        gy = is_border(c01) ? p->borderAlpha : tex_load(p->tex, p->mip, c01.x, c01.y);
        gz = is_border(c11) ? p->borderAlpha : tex_load(p->tex, p->mip, c11.x, c11.y);
        gw = is_border(c10) ? p->borderAlpha : tex_load(p->tex, p->mip, c10.x, c10.y);
    } else {
        gx = tex_load(p->tex, p->mip, c00.x, c00.y);
        gy = tex_load(p->tex, p->mip, c01.x, c01.y);
        gz = tex_load(p->tex, p->mip, c11.x, c11.y);
        gw = tex_load(p->tex, p->mip, c10.x, c10.y);
    }
    if (!degenerate) {
        const int o0 = p->cutoff < gx, o1 = p->cutoff < gy, o2 = p->cutoff < gz, o3 = p->cutoff < gw;
        f2 q0 = {ipx, ipy};
        f2 q1 = {ipx + 0.0f, ipy + m->rcph};
        f2 q2 = {ipx + m->rcpw, ipy + m->rcph};
        f2 q3 = {ipx + m->rcpw, ipy + 0.0f};
        const int in0 = point_in_tri(p->tri, q0), in1 = point_in_tri(p->tri, q1);
        const int in2 = point_in_tri(p->tri, q2), in3 = point_in_tri(p->tri, q3);
        const int isOpaque = (in0 && o0) || (in1 && o1) || (in2 && o2) || (in3 && o3);
        const int isTransparent = (in0 && !o0) || (in1 && !o1) || (in2 && !o2) || (in3 && !o3);
        if (isOpaque) p->cov->above += 1;
        if (isTransparent) p->cov->below += 1;
        if (isOpaque && isTransparent) return;
    }
    const float a = gx;
    const float b = gw - gx;
    const float c = gy - gx;
    const float d = gx + gz - gy - gw;
    if (is_zero(b, 1e-6f) && is_zero(c, 1e-6f) && is_zero(d, 1e-6f)) {
        if (p->cutoff < a) p->cov->above += 1;
        else p->cov->below += 1;
        return;
    }
    const float sx = (float)m->w, sy = (float)m->h;
    const float h0 = a - p->cutoff;
    if (degenerate) {
        f2 e0 = {sx * p->tri->aabb_s.x - pfx, sy * p->tri->aabb_s.y - pfy};
        f2 e1 = {sx * p->tri->aabb_e.x - pfx, sy * p->tri->aabb_e.y - pfy};
        if (edge_hyperbola(e0, e1, h0, b, c, d)) { p->cov->above += 1; p->cov->below += 1; }
    } else {
        const f2 v[3] = {p->tri->p0, p->tri->p1, p->tri->p2};
        for (int e = 0; e < 3; ++e) {
            const f2 A = v[e % 3], B = v[(e + 1) % 3];
            f2 e0 = {sx * A.x - pfx, sy * A.y - pfy};
            f2 e1 = {sx * B.x - pfx, sy * B.y - pfy};
            if (edge_hyperbola(e0, e1, h0, b, c, d)) { p->cov->above += 1; p->cov->below += 1; break; }
        }
    }
}

/* ref: bake_kernels_cpu.h:404-452 (internal flags 7/8 only) */
static void conservative_bilinear_kernel(int px, int py, KParams* p) {
    const OMip* m = &p->tex->mips[p->mip];
    const float pfx = (float)px + 0.5f, pfy = (float)py + 0.5f;
    i2 c00, c10, c01, c11;
    gather4(p->mode, p->pow2, (int)pfx, (int)pfy, m, &c00, &c10, &c01, &c11);
    const int brd = p->mode == ommTextureAddressMode_Border;
    float gx = (brd && is_border(c00)) ? p->borderAlpha : tex_load(p->tex, p->mip, c00.x, c00.y);
    float gy = (brd && is_border(c01)) ? p->borderAlpha : tex_load(p->tex, p->mip, c01.x, c01.y);
    float gz = (brd && is_border(c11)) ? p->borderAlpha : tex_load(p->tex, p->mip, c11.x, c11.y);
    float gw = (brd && is_border(c10)) ? p->borderAlpha : tex_load(p->tex, p->mip, c10.x, c10.y);
    const float mn = minf2(minf2(minf2(gx, gy), gz), gw);
    const float mx = maxf2(maxf2(maxf2(gx, gy), gz), gw);
    if (p->cutoff < mx) p->cov->above += 1;
    if (p->cutoff > mn) p->cov->below += 1;
}

static void nearest_kernel(int px, int py, KParams* p) { /* ref: bake_cpu_impl.cpp:994-1009 */
    const OMip* m = &p->tex->mips[p->mip];
    int cx = addr1(p->mode, p->pow2, px, m->w, m->log2w), cy = addr1(p->mode, p->pow2, py, m->h, m->log2h);
    const int border = p->mode == ommTextureAddressMode_Border && (cx == TEXCOORD_BORDER || cy == TEXCOORD_BORDER);
    const float alpha = border ? p->borderAlpha : tex_load(p->tex, p->mip, cx, cy);
    if (p->cutoff < alpha) p->cov->above += 1;
    else p->cov->below += 1;
}

enum { K_LEVEL_LINE, K_LEVEL_LINE_DEGEN, K_CONS_BILINEAR, K_NEAREST };
static void run_kernel(int which, int x, int y, KParams* p) {
    switch (which) {
    case K_LEVEL_LINE: level_line_kernel(x, y, 0, p); break;
    case K_LEVEL_LINE_DEGEN: level_line_kernel(x, y, 1, p); break;
    case K_CONS_BILINEAR: conservative_bilinear_kernel(x, y, p); break;
    default: nearest_kernel(x, y, p); break;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Rasterizers.  ref: util/cpu_raster.h                                                         */
/* ------------------------------------------------------------------------------------------ */
typedef struct { float nx, ny, c; } EdgeFn;
static EdgeFn edge_fn(f2 p, f2 q) { /* ref: util/cpu_raster.h:26-29 */
    EdgeFn e;
    e.nx = q.y - p.y; e.ny = p.x - q.x;
    e.c = -(e.nx * p.x + e.ny * p.y);
    return e;
}
static float eval_edge_cons(const EdgeFn* e, float sx, float sy) { /* ref: util/cpu_raster.h:46-51, ext = (1,1) */
    const float ev = (e->nx * sx + e->ny * sy) + e->c;
    const float bx = e->nx > 0 ? 0.f : e->nx;
    const float by = e->ny > 0 ? 0.f : e->ny;
    return ev + bx * 1.f + by * 1.f;
}
/* ref: util/cpu_raster.h:278-341 (OverConservative, serial, no barycentrics) */
static void raster_tri_conservative(const Tri* _t, int rw, int rh, float offx, float offy, int which, KParams* kp) {
    const int ccw = tri_is_ccw(_t->p0, _t->p1, _t->p2);
    const float rfx = (float)rw, rfy = (float)rh;
    f2 a = {_t->p0.x * rfx + offx, _t->p0.y * rfy + offy};
    f2 b = {_t->p1.x * rfx + offx, _t->p1.y * rfy + offy};
    f2 c = {_t->p2.x * rfx + offx, _t->p2.y * rfy + offy};
    Tri t = ccw ? make_tri(a, b, c) : make_tri(c, b, a);
    const int minx = (int)floorf(t.aabb_s.x), miny = (int)floorf(t.aabb_s.y);
    const int maxx = (int)ceilf(t.aabb_e.x), maxy = (int)ceilf(t.aabb_e.y);
    const EdgeFn e0 = edge_fn(t.p0, t.p1), e1 = edge_fn(t.p1, t.p2), e2 = edge_fn(t.p2, t.p0);
    for (int y = miny; y < maxy; ++y) {
        int wasInside = 0;
        for (int x = minx; x < maxx; ++x) {
            const float sx = (float)x, sy = (float)y;
            const float v0 = eval_edge_cons(&e0, sx, sy), v1 = eval_edge_cons(&e1, sx, sy), v2 = eval_edge_cons(&e2, sx, sy);
            if (v0 < 0.f && v1 < 0.f && v2 < 0.f) {
                run_kernel(which, x, y, kp);
                wasInside = 1;
            } else if (wasInside)
                break;
        }
    }
}
/* ref: util/cpu_raster.h:486-555 (conservative DDA line) */
static void raster_line_conservative(f2 lp0, f2 lp1, int rw, int rh, float offx, float offy, int which, KParams* kp) {
    const float rfx = (float)rw, rfy = (float)rh;
    f2 p0 = {lp0.x * rfx + offx, lp0.y * rfy + offy};
    f2 p1 = {lp1.x * rfx + offx, lp1.y * rfy + offy};
    if (p0.x > p1.x) { f2 tmp = p0; p0 = p1; p1 = tmp; }
    const float dx = p1.x - p0.x, dy = p1.y - p0.y;
    int x = (int)floorf(p0.x), y = (int)floorf(p0.y);
    const int stepX = (dx > 0) ? 1 : ((dx < 0) ? -1 : 0);
    const int stepY = (dy > 0) ? 1 : ((dy < 0) ? -1 : 0);
    const float tDeltaX = (stepX != 0) ? 1.f / fabsf(dx) : INFINITY;
    const float tDeltaY = (stepY != 0) ? 1.f / fabsf(dy) : INFINITY;
    float tMaxX, tMaxY;
    if (stepX != 0) {
        float next = ((float)x + (stepX > 0 ? 1.f : 0.f));
        tMaxX = (next - p0.x) / dx;
    } else
        tMaxX = INFINITY;
    if (stepY != 0) {
        float next = ((float)y + (stepY > 0 ? 1.f : 0.f));
        tMaxY = (next - p0.y) / dy;
    } else
        tMaxY = INFINITY;
    if (stepX == 0 && stepY == 0) {
        run_kernel(which, x, y, kp);
        return;
    }
    const int yMin = (int)minf2(floorf(p0.y), floorf(p1.y)), yMax = (int)maxf2(ceilf(p0.y), ceilf(p1.y));
    const int xMin = (int)minf2(floorf(p0.x), floorf(p1.x)), xMax = (int)maxf2(ceilf(p0.x), ceilf(p1.x));
    while (x >= xMin && x <= xMax && y >= yMin && y <= yMax) {
        run_kernel(which, x, y, kp);
        if (tMaxX < tMaxY) { x += stepX; tMaxX += tDeltaX; }
        else { y += stepY; tMaxY += tDeltaY; }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Hashes: libstdc++ std::hash<float> (murmur _Hash_bytes), glm hash_combine, XXH64             */
/* ------------------------------------------------------------------------------------------ */
/* libstdc++ (GCC 13) libsupc++/hash_bytes.cc, 64-bit variant, specialised for len == 4 and the default seed.
 * std::hash<float> returns 0 for +-0.0f (functional_hash.h). */
static uint64_t std_hash_float(float v) {
    if (v == 0.0f) return 0;
    const uint64_t mul = (((uint64_t)0xc6a4a793UL) << 32) + (uint64_t)0x5bd1e995UL;
    uint32_t bits;
    memcpy(&bits, &v, 4);
    uint64_t hash = (uint64_t)0xc70f6907UL ^ (4 * mul);
    hash ^= (uint64_t)bits; /* load_bytes(end, 4): little-endian tail */
    hash *= mul;
    hash = (hash ^ (hash >> 47)) * mul;
    hash = hash ^ (hash >> 47);
    return hash;
}
static void glm_hash_combine(uint64_t* seed, uint64_t hash) { /* ref: external/glm/glm/gtx/hash.inl:6-10 */
    hash += 0x9e3779b9 + (*seed << 6) + (*seed >> 2);
    *seed ^= hash;
}
static uint64_t hash_f2(f2 v) { /* ref: external/glm/glm/gtx/hash.inl:22-29 */
    uint64_t seed = 0;
    glm_hash_combine(&seed, std_hash_float(v.x));
    glm_hash_combine(&seed, std_hash_float(v.y));
    return seed;
}
static void omm_hash_combine(uint64_t* seed, uint64_t h) { /* ref: util/geometry.h:141-146 */
    *seed ^= h + 0x9e3779b9 + (*seed << 6) + (*seed >> 2);
}

/* XXH64 as published in the xxHash specification (doc/xxhash_spec.md of Cyan4973/xxHash, pinned by the SDK at
 * submodule commit c961fbe6); call site ref: bake_cpu_impl.cpp:1039 with seed 42. */
static const uint64_t XP1 = 0x9E3779B185EBCA87ULL, XP2 = 0xC2B2AE3D27D4EB4FULL, XP3 = 0x165667B19E3779F9ULL,
                      XP4 = 0x85EBCA77C2B2AE63ULL, XP5 = 0x27D4EB2F165667C5ULL;
static uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static uint64_t xxh_round(uint64_t acc, uint64_t in) { return rotl64(acc + in * XP2, 31) * XP1; }
static uint64_t xxh_merge(uint64_t acc, uint64_t v) { return (acc ^ xxh_round(0, v)) * XP1 + XP4; }
static uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
OAPI uint64_t omm_oracle_xxh64(const void* data, size_t len, uint64_t seed) {
    const uint8_t* p = (const uint8_t*)data;
    const uint8_t* end = p + len;
    uint64_t h;
    if (len >= 32) {
        uint64_t v1 = seed + XP1 + XP2, v2 = seed + XP2, v3 = seed, v4 = seed - XP1;
        do {
            v1 = xxh_round(v1, rd64(p)); v2 = xxh_round(v2, rd64(p + 8));
            v3 = xxh_round(v3, rd64(p + 16)); v4 = xxh_round(v4, rd64(p + 24));
            p += 32;
        } while (p + 32 <= end);
        h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
        h = xxh_merge(h, v1); h = xxh_merge(h, v2); h = xxh_merge(h, v3); h = xxh_merge(h, v4);
    } else
        h = seed + XP5;
    h += (uint64_t)len;
    while (p + 8 <= end) { h ^= xxh_round(0, rd64(p)); h = rotl64(h, 27) * XP1 + XP4; p += 8; }
    if (p + 4 <= end) { h ^= (uint64_t)rd32(p) * XP1; h = rotl64(h, 23) * XP2 + XP3; p += 4; }
    while (p < end) { h ^= (uint64_t)(*p) * XP5; h = rotl64(h, 11) * XP1; p++; }
    h ^= h >> 33; h *= XP2; h ^= h >> 29; h *= XP3; h ^= h >> 32;
    return h;
}

/* tiny open-addressing u64 -> u32 map, "first insert wins" (stands in for std::unordered_map) */
typedef struct { uint64_t* keys; uint32_t* vals; uint8_t* used; size_t cap; } Map;
static int map_init(Map* m, size_t n) {
    size_t cap = 16;
    while (cap < n * 2 + 1) cap <<= 1;
    m->cap = cap;
    m->keys = (uint64_t*)malloc(cap * sizeof(uint64_t));
    m->vals = (uint32_t*)malloc(cap * sizeof(uint32_t));
    m->used = (uint8_t*)calloc(cap, 1);
    return m->keys && m->vals && m->used;
}
static void map_free(Map* m) { free(m->keys); free(m->vals); free(m->used); }
static uint32_t* map_find(Map* m, uint64_t k) {
    size_t i = (size_t)((k * 0x9E3779B97F4A7C15ULL) >> 20) & (m->cap - 1);
    while (m->used[i]) {
        if (m->keys[i] == k) return &m->vals[i];
        i = (i + 1) & (m->cap - 1);
    }
    return NULL;
}
static void map_insert(Map* m, uint64_t k, uint32_t v) {
    size_t i = (size_t)((k * 0x9E3779B97F4A7C15ULL) >> 20) & (m->cap - 1);
    while (m->used[i]) i = (i + 1) & (m->cap - 1);
    m->used[i] = 1; m->keys[i] = k; m->vals[i] = v;
}

/* ------------------------------------------------------------------------------------------ */
/* Input fetch.  ref: util/geometry.h:148-239                                                   */
/* ------------------------------------------------------------------------------------------ */
static float half_to_float(uint16_t h) { /* glm detail::toFloat32, ref: external/glm/glm/detail/type_half.inl:31-103 */
    int s = (h >> 15) & 1, e = (h >> 10) & 0x1f, m = h & 0x3ff;
    union { uint32_t u; float f; } r;
    if (e == 0) {
        if (m == 0) { r.u = (uint32_t)s << 31; return r.f; }
        while (!(m & 0x400)) { m <<= 1; e -= 1; }
        e += 1; m &= ~0x400;
    } else if (e == 31) {
        r.u = ((uint32_t)s << 31) | 0x7f800000u | ((uint32_t)m << 13);
        return r.f;
    }
    e = e + (127 - 15);
    r.u = ((uint32_t)s << 31) | ((uint32_t)e << 23) | ((uint32_t)m << 13);
    return r.f;
}
static f2 fetch_uv(const void* texCoords, uint32_t stride, ommTexCoordFormat fmt, uint32_t index) {
    const uint8_t* base = (const uint8_t*)texCoords + (size_t)stride * index;
    f2 r = {0, 0};
    if (fmt == ommTexCoordFormat_UV16_UNORM) {
        uint16_t v[2]; memcpy(v, base, 4);
        r.x = (float)v[0] * 1.5259021896696421759365224689097e-5f;
        r.y = (float)v[1] * 1.5259021896696421759365224689097e-5f;
    } else if (fmt == ommTexCoordFormat_UV16_FLOAT) {
        uint16_t v[2]; memcpy(v, base, 4);
        r.x = half_to_float(v[0]); r.y = half_to_float(v[1]);
    } else if (fmt == ommTexCoordFormat_UV32_FLOAT) {
        memcpy(&r, base, 8);
    }
    return r;
}
static uint32_t texcoord_size(ommTexCoordFormat fmt) { return fmt == ommTexCoordFormat_UV32_FLOAT ? 8u : (fmt < ommTexCoordFormat_UV32_FLOAT ? 4u : 0u); }
static void fetch_indices(ommIndexFormat fmt, const void* idx, size_t first, uint32_t out[3]) {
    for (int i = 0; i < 3; ++i) {
        if (fmt == ommIndexFormat_UINT_8) out[i] = ((const uint8_t*)idx)[first + i];
        else if (fmt == ommIndexFormat_UINT_16) out[i] = ((const uint16_t*)idx)[first + i];
        else out[i] = ((const uint32_t*)idx)[first + i];
    }
}
static Tri get_triangle(const ommCpuBakeInputDesc* d, uint32_t prim) { /* ref: bake_cpu_impl.cpp:579-587 */
    uint32_t stride = d->texCoordStrideInBytes == 0 ? texcoord_size(d->texCoordFormat) : d->texCoordStrideInBytes;
    uint32_t ix[3];
    fetch_indices(d->indexFormat, d->indexBuffer, (size_t)3 * prim, ix);
    return make_tri(fetch_uv(d->texCoords, stride, d->texCoordFormat, ix[0]), fetch_uv(d->texCoords, stride, d->texCoordFormat, ix[1]),
                    fetch_uv(d->texCoords, stride, d->texCoordFormat, ix[2]));
}

/* ------------------------------------------------------------------------------------------ */
/* Subdivision level selection.  ref: bake_cpu_impl.cpp:464-560                                 */
/* ------------------------------------------------------------------------------------------ */
static float area2d(f2 p0, f2 p1, f2 p2) { /* ref: :464-468; glm cross/length on (v,0) vectors */
    const float v0x = p2.x - p0.x, v0y = p2.y - p0.y, v1x = p1.x - p0.x, v1y = p1.y - p0.y;
    /* cross((v0,0),(v1,0)) = (v0y*0 - v1y*0, 0*v1x - 0*v0x, v0x*v1y - v1x*v0y); length = sqrt(dot) */
    const float cx = v0y * 0.f - v1y * 0.f, cy = 0.f * v1x - 0.f * v0x, cz = v0x * v1y - v1x * v0y;
    return 0.5f * sqrtf(cx * cx + cy * cy + cz * cz);
}
static uint32_t area_heuristic(const ommCpuBakeInputDesc* d, const Tri* t, int w, int h) { /* ref: :470-509 */
    const float sx = (float)(uint32_t)w, sy = (float)(uint32_t)h;
    f2 a = {t->p0.x * sx, t->p0.y * sy}, b = {t->p1.x * sx, t->p1.y * sy}, c = {t->p2.x * sx, t->p2.y * sy};
    const float pixelUvArea = area2d(a, b, c);
    const float target = d->dynamicSubdivisionScale * d->dynamicSubdivisionScale;
    const uint32_t ratio = (uint32_t)(int64_t)(pixelUvArea / target);
    uint32_t v = ratio;
    v--; v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16; v++;
    static const uint32_t bm[5] = {0xAAAAAAAAu, 0xCCCCCCCCu, 0xF0F0F0F0u, 0xFF00FF00u, 0xFFFF0000u};
    uint32_t r = (v & bm[0]) != 0;
    for (uint32_t i = 4; i > 0; i--) r |= (uint32_t)((v & bm[i]) != 0) << i;
    const uint32_t lvl = r >> 1;
    return lvl < d->maxSubdivisionLevel ? lvl : d->maxSubdivisionLevel;
}
static uint32_t edge_heuristic(const ommCpuBakeInputDesc* d, const Tri* t, int w, int h) { /* ref: :511-528 */
    const float sx = (float)(uint32_t)w, sy = (float)(uint32_t)h;
    const float e0x = sx * (t->p1.x - t->p0.x), e0y = sy * (t->p1.y - t->p0.y);
    const float e1x = sx * (t->p2.x - t->p0.x), e1y = sy * (t->p2.y - t->p0.y);
    const float e2x = sx * (t->p2.x - t->p1.x), e2y = sy * (t->p2.y - t->p1.y);
    const float l0 = e0x * e0x + e0y * e0y, l1 = e1x * e1x + e1y * e1y, l2 = e2x * e2x + e2y * e2y;
    float eMax = l0; /* std::max({l0,l1,l2}) keeps the first maximum */
    if (eMax < l1) eMax = l1;
    if (eMax < l2) eMax = l2;
    const float n = eMax < 1e-6 ? 0 : log2f(eMax) / 2.f - log2f(d->dynamicSubdivisionScale);
    const int lvl = (int)ceilf(n);
    return (uint32_t)clampi(lvl, 0, d->maxSubdivisionLevel);
}
static int32_t level_for_primitive(const ommCpuBakeInputDesc* d, int edgeHeuristic, uint32_t i, const Tri* t, int w, int h) { /* ref: :542-560 */
    if (d->subdivisionLevels && d->subdivisionLevels[i] <= 12) return d->subdivisionLevels[i];
    if (d->dynamicSubdivisionScale > 0) {
        if (tri_is_degenerate(t) || edgeHeuristic) return (int32_t)edge_heuristic(d, t, w, h);
        return (int32_t)area_heuristic(d, t, w, h);
    }
    return d->maxSubdivisionLevel;
}

/* ------------------------------------------------------------------------------------------ */
/* Bake.  ref: bake_cpu_impl.cpp:589-1985                                                       */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    uint32_t level;
    ommFormat format;
    Tri uvTri;
    uint32_t* prims; uint32_t numPrims, capPrims;
    uint32_t descOffset;
    uint32_t specialIndex; /* 0 = none */
    uint8_t* states;       /* 4^level bytes, 4/2-state */
    uint8_t* states3;      /* 4^level bytes, UT folded into UO */
} Item;

typedef struct {
    uint32_t magic;
    ommCpuBakeResultDesc desc;
    uint8_t* arrayData;
    ommCpuOpacityMicromapDesc* descArray;
    ommCpuOpacityMicromapUsageCount *arrayHist, *indexHist;
    int32_t* indexBuffer;
} OResult;
#define ORESULT_MAGIC 0x0e5017aau

static void item_push_prim(Item* it, uint32_t p) {
    if (it->numPrims == it->capPrims) {
        it->capPrims = it->capPrims ? it->capPrims * 2 : 2;
        it->prims = (uint32_t*)realloc(it->prims, it->capPrims * sizeof(uint32_t));
    }
    it->prims[it->numPrims++] = p;
}
static void set_state(Item* it, uint32_t i, ommOpacityState s) { /* ref: :374-377 */
    it->states[i] = (uint8_t)s;
    it->states3[i] = (uint8_t)(s == ommOpacityState_UnknownTransparent ? ommOpacityState_UnknownOpaque : s);
}

enum {
    FLAG_AABB_TESTING = 1u << 7, FLAG_DISABLE_LEVEL_LINE = 1u << 8, FLAG_DISABLE_FINE = 1u << 9,
    FLAG_NEAR_DUP_BRUTE = 1u << 10, FLAG_EDGE_HEURISTIC = 1u << 11
}; /* ref: bake_cpu_impl.cpp:44-48 */

static void promote_special(const ommCpuBakeInputDesc* d, Item* items, uint32_t n) { /* ref: :1432-1472 */
    const int disableSpecial = (d->bakeFlags & ommCpuBakeFlags_DisableSpecialIndices) != 0;
    for (uint32_t w = 0; w < n; ++w) {
        Item* it = &items[w];
        if (it->specialIndex != 0) continue;
        const uint32_t N = 1u << (it->level << 1);
        int allEqual = 1;
        ommOpacityState common = (ommOpacityState)it->states[0];
        for (uint32_t i = 1; i < N; ++i) allEqual &= (common == (ommOpacityState)it->states[i]);
        if (!allEqual && d->rejectionThreshold > 0.f) {
            uint32_t known = 0;
            for (uint32_t i = 0; i < N; ++i) known += is_known((ommOpacityState)it->states[i]);
            const float frac = known / (float)N;
            if (frac < d->rejectionThreshold) { allEqual = 1; common = ommOpacityState_UnknownTransparent; }
        }
        if (allEqual && !disableSpecial) it->specialIndex = (uint32_t)(-(int32_t)common - 1);
    }
}
static void dedup_exact(const ommCpuBakeInputDesc* d, Item* items, uint32_t n) { /* ref: :1031-1066 */
    if (d->bakeFlags & ommCpuBakeFlags_DisableDuplicateDetection) return;
    Map m;
    map_init(&m, n);
    for (uint32_t i = 0; i < n; ++i) {
        Item* it = &items[i];
        const uint64_t digest = omm_oracle_xxh64(it->states3, (size_t)1 << (it->level << 1), 42);
        uint32_t* found = map_find(&m, digest);
        if (!found) map_insert(&m, digest, i);
        else {
            Item* dst = &items[*found];
            for (uint32_t k = 0; k < it->numPrims; ++k) item_push_prim(dst, it->prims[k]);
            it->numPrims = 0;
            it->specialIndex = (uint32_t)-1;
        }
    }
    map_free(&m);
}

typedef struct { uint64_t key; uint32_t idx; } SortKey;
static int sortkey_desc(const void* a, const void* b) { /* std::greater<pair<u64,u32>>, ref: :1751 */
    const SortKey* x = (const SortKey*)a; const SortKey* y = (const SortKey*)b;
    if (x->key != y->key) return x->key > y->key ? -1 : 1;
    if (x->idx != y->idx) return x->idx > y->idx ? -1 : 1;
    return 0;
}
static uint32_t bit_interleave16(uint32_t x, uint32_t y) { /* ref: util/bit_tricks.h:40-64 */
    x = (x | (x << 8)) & 0x00FF00FFu; x = (x | (x << 4)) & 0x0F0F0F0Fu; x = (x | (x << 2)) & 0x33333333u; x = (x | (x << 1)) & 0x55555555u;
    y = (y | (y << 8)) & 0x00FF00FFu; y = (y | (y << 4)) & 0x0F0F0F0Fu; y = (y | (y << 2)) & 0x33333333u; y = (y | (y << 1)) & 0x55555555u;
    return x | (y << 1);
}

static void free_items(Item* items, uint32_t n) {
    for (uint32_t i = 0; i < n; ++i) { free(items[i].prims); free(items[i].states); free(items[i].states3); }
    free(items);
}

static ommResult bake_impl(const ommCpuBakeInputDesc* d, OResult* res) {
    const OTexture* tex = (const OTexture*)d->texture;
    const uint32_t flags = (uint32_t)d->bakeFlags;
    const int edgeHeur = (flags & FLAG_EDGE_HEURISTIC) != 0;
    const int disableLevelLine = (flags & FLAG_DISABLE_LEVEL_LINE) != 0;
    const int aabbTesting = (flags & FLAG_AABB_TESTING) != 0;
    const int disableDup = (flags & ommCpuBakeFlags_DisableDuplicateDetection) != 0;
    const ommTextureAddressMode mode = d->runtimeSamplerDesc.addressingMode;
    const int pow2 = tex->mips[0].isPow2;
    const int32_t T = (int32_t)(d->indexCount / 3u);

    /* ---- SetupWorkItems, ref: :589-660 ---- */
    Item* items = (Item*)calloc((size_t)(T > 0 ? T : 1), sizeof(Item));
    uint32_t W = 0;
    Map idmap;
    map_init(&idmap, (size_t)T);
    for (int32_t i = 0; i < T; ++i) {
        const Tri uv = get_triangle(d, (uint32_t)i);
        const int32_t lvl = level_for_primitive(d, edgeHeur, (uint32_t)i, &uv, tex->mips[0].w, tex->mips[0].h);
        if (lvl == 0xE || tri_is_invalid(&uv) || (disableLevelLine && tri_is_degenerate(&uv))) continue;
        const ommFormat fmt = (!d->formats || d->formats[i] == ommFormat_INVALID) ? d->format : d->formats[i];
        uint64_t seed = 42;
        omm_hash_combine(&seed, hash_f2(uv.p0));
        omm_hash_combine(&seed, hash_f2(uv.p1));
        omm_hash_combine(&seed, hash_f2(uv.p2));
        omm_hash_combine(&seed, (uint64_t)(int64_t)lvl);
        omm_hash_combine(&seed, (uint64_t)(int64_t)(int32_t)fmt);
        uint32_t* found = map_find(&idmap, seed);
        if (!found || disableDup) {
            if (lvl > 12) { map_free(&idmap); free_items(items, W); return ommResult_INVALID_ARGUMENT; }
            if (!found) map_insert(&idmap, seed, W);
            Item* it = &items[W++];
            it->level = (uint32_t)lvl; it->format = fmt; it->uvTri = uv;
            const size_t N = (size_t)1 << (lvl << 1);
            it->states = (uint8_t*)malloc(N); it->states3 = (uint8_t*)malloc(N);
            memset(it->states, ommOpacityState_UnknownOpaque, N);
            memset(it->states3, ommOpacityState_UnknownOpaque, N);
            item_push_prim(it, (uint32_t)i);
        } else
            item_push_prim(&items[*found], (uint32_t)i);
    }
    map_free(&idmap);

    /* ---- ValidateWorkloadSize, ref: :662-713 ---- */
    if (d->maxWorkloadSize != 0xFFFFFFFFFFFFFFFFull) {
        uint64_t workload = 0;
        const float sx = (float)tex->mips[0].w, sy = (float)tex->mips[0].h;
        for (uint32_t w = 0; w < W; ++w) {
            const int ax = (int)((items[w].uvTri.aabb_e.x - items[w].uvTri.aabb_s.x) * sx);
            const int ay = (int)((items[w].uvTri.aabb_e.y - items[w].uvTri.aabb_s.y) * sy);
            workload += (uint64_t)(int64_t)(ax * ay);
        }
        if (workload > d->maxWorkloadSize) { free_items(items, W); return ommResult_WORKLOAD_TOO_BIG; }
    }

    if (aabbTesting && !disableLevelLine) { free_items(items, W); return ommResult_INVALID_ARGUMENT; } /* ref: :718-719 */

    /* ---- ResampleCoarse, ref: :716-808 ---- */
    if (tex->mips[0].sat && tex->mipCount == 1 && d->runtimeSamplerDesc.filter == ommTextureFilterMode_Linear) {
        const OMip* m = &tex->mips[0];
        for (uint32_t w = 0; w < W; ++w) {
            Item* it = &items[w];
            const uint32_t N = 1u << (it->level << 1);
            for (uint32_t u = 0; u < N; ++u) {
                const Tri st = micro_tri(&it->uvTri, u, it->level);
                if ((int)st.aabb_s.x != (int)st.aabb_e.x || (int)st.aabb_s.y != (int)st.aabb_e.y) continue;
                const float fsx = st.aabb_s.x * (float)m->w - 0.5f, fsy = st.aabb_s.y * (float)m->h - 0.5f;
                const float fex = st.aabb_e.x * (float)m->w - 0.5f, fey = st.aabb_e.y * (float)m->h - 0.5f;
                i2 s00, s10, s01, s11, e00, e10, e01, e11;
                gather4(mode, pow2, (int)floorf(fsx), (int)floorf(fsy), m, &s00, &s10, &s01, &s11);
                gather4(mode, pow2, (int)floorf(fex), (int)floorf(fey), m, &e00, &e10, &e01, &e11);
                const i2 as = s00, ae = e11;
                if (ae.x < as.x || ae.y < as.y) continue;
                if (!in_texture(m, as) || !in_texture(m, ae)) continue;
                const uint32_t area = (uint32_t)((ae.x - as.x + 1) * (ae.y - as.y + 1));
                const uint32_t sa = tex_sat(m, as, ae);
                if (sa == 0) set_state(it, u, d->alphaCutoffLessEqual);
                else if (sa == area) set_state(it, u, d->alphaCutoffGreater);
            }
        }
    }

    /* ---- ResampleFine (Normal then Degenerate), ref: :817-1029, 1953-1955 ---- */
    if (!(flags & FLAG_DISABLE_FINE)) {
        for (int cls = 0; cls < 2; ++cls) {
            for (uint32_t w = 0; w < W; ++w) {
                Item* it = &items[w];
                const int degen = tri_is_degenerate(&it->uvTri);
                if (degen != cls) continue;
                const uint32_t N = 1u << (it->level << 1);
                for (uint32_t u = 0; u < N; ++u) {
                    Coverage cov = {0, 0};
                    KParams kp;
                    kp.tex = tex; kp.pow2 = pow2; kp.mode = mode; kp.cutoff = d->alphaCutoff;
                    kp.borderAlpha = d->runtimeSamplerDesc.borderAlpha; kp.cov = &cov;
                    if (d->runtimeSamplerDesc.filter == ommTextureFilterMode_Linear) {
                        if (it->states[u] != ommOpacityState_UnknownOpaque) continue;
                        const Tri st = micro_tri(&it->uvTri, u, it->level);
                        kp.tri = &st;
                        if (!disableLevelLine) {
                            for (uint32_t mip = 0; mip < tex->mipCount; ++mip) {
                                kp.mip = (int)mip;
                                const OMip* m = &tex->mips[mip];
                                if (d->alphaCutoff < tex_bilinear(tex, mode, kp.borderAlpha, st.p0, (int)mip)) cov.above++;
                                else cov.below++;
                                if (!cls) raster_tri_conservative(&st, m->w, m->h, -0.5f, -0.5f, K_LEVEL_LINE, &kp);
                                else raster_line_conservative(st.aabb_s, st.aabb_e, m->w, m->h, -0.5f, -0.5f, K_LEVEL_LINE_DEGEN, &kp);
                                if (is_unknown(state_from_coverage(d->format, d->unknownStatePromotion, d->alphaCutoffGreater, d->alphaCutoffLessEqual, cov))) break;
                            }
                        } else if (aabbTesting) {
                            kp.mip = 0;
                            const OMip* m = &tex->mips[0];
                            f2 c1 = {st.aabb_e.x, st.aabb_s.y}, c2 = {st.aabb_s.x, st.aabb_e.y};
                            const Tri t0 = make_tri(st.aabb_s, c1, c2), t1 = make_tri(st.aabb_e, c1, c2);
                            raster_tri_conservative(&t0, m->w, m->h, -0.5f, -0.5f, K_CONS_BILINEAR, &kp);
                            raster_tri_conservative(&t1, m->w, m->h, -0.5f, -0.5f, K_CONS_BILINEAR, &kp);
                        } else {
                            kp.mip = 0;
                            const OMip* m = &tex->mips[0];
                            raster_tri_conservative(&st, m->w, m->h, -0.5f, -0.5f, K_CONS_BILINEAR, &kp);
                        }
                        set_state(it, u, state_from_coverage(d->format, d->unknownStatePromotion, d->alphaCutoffGreater, d->alphaCutoffLessEqual, cov));
                    } else {
                        const Tri st = micro_tri(&it->uvTri, u, it->level);
                        kp.tri = &st;
                        for (uint32_t mip = 0; mip < tex->mipCount; ++mip) {
                            kp.mip = (int)mip;
                            const OMip* m = &tex->mips[mip];
                            raster_tri_conservative(&st, m->w, m->h, 0.f, 0.f, K_NEAREST, &kp);
                            if (is_unknown(state_from_coverage(d->format, d->unknownStatePromotion, d->alphaCutoffGreater, d->alphaCutoffLessEqual, cov))) break;
                        }
                        set_state(it, u, state_from_coverage(d->format, d->unknownStatePromotion, d->alphaCutoffGreater, d->alphaCutoffLessEqual, cov));
                    }
                }
            }
        }
    }

    /* ---- promote / dedup sequence, ref: :1957-1971 (near-dup and compress are not restated) ---- */
    promote_special(d, items, W);
    dedup_exact(d, items, W);
    promote_special(d, items, W);
    dedup_exact(d, items, W);
    promote_special(d, items, W);

    /* ---- CreateUsageHistograms, ref: :1690-1705 ---- */
    uint32_t arrayHist[3][13], indexHist[3][13];
    memset(arrayHist, 0, sizeof(arrayHist)); memset(indexHist, 0, sizeof(indexHist));
    for (uint32_t w = 0; w < W; ++w)
        if (items[w].specialIndex == 0) {
            arrayHist[items[w].format][items[w].level] += 1;
            indexHist[items[w].format][items[w].level] += items[w].numPrims;
        }

    /* ---- MicromapSpatialSort, ref: :1707-1754 ---- */
    SortKey* keys = (SortKey*)malloc(sizeof(SortKey) * (W ? W : 1));
    for (uint32_t w = 0; w < W; ++w) {
        const Item* it = &items[w];
        keys[w].idx = w;
        if (it->specialIndex != 0) keys[w].key = (1ull << 63) | (uint64_t)w;
        else {
            const float cx = (it->uvTri.p0.x + it->uvTri.p1.x + it->uvTri.p2.x) / 3.f;
            const float cy = (it->uvTri.p0.y + it->uvTri.p1.y + it->uvTri.p2.y) / 3.f;
            const int qx = (int)(8192.f * cx), qy = (int)(8192.f * cy);
            const int mx = clampi((int)fabsf((float)qx + 0.5f), 0, 8191), my = clampi((int)fabsf((float)qy + 0.5f), 0, 8191);
            keys[w].key = ((uint64_t)it->level << 60) | (uint64_t)bit_interleave16((uint32_t)mx, (uint32_t)my);
        }
    }
    qsort(keys, W, sizeof(SortKey), sortkey_desc);

    /* ---- Serialize, ref: :1756-1920 ---- */
    ommResult rc = ommResult_SUCCESS;
    {
        const uint32_t bitCount = (uint32_t)d->format;
        uint32_t descCount = 0;
        size_t arraySize = 0;
        for (uint32_t l = 0; l < 13; ++l) {
            const uint32_t cnt = arrayHist[d->format][l];
            descCount += cnt;
            size_t bits = ((size_t)1 << (l << 1)) * bitCount;
            size_t bytes = bits >> 3;
            arraySize += (size_t)cnt * (bytes > 1 ? bytes : 1);
        }
        if (arraySize > 0xFFFFFFFFull) rc = ommResult_FAILURE;
        if (rc == ommResult_SUCCESS && descCount != 0) {
            res->arrayData = (uint8_t*)calloc(arraySize, 1);
            res->descArray = (ommCpuOpacityMicromapDesc*)calloc(descCount, sizeof(ommCpuOpacityMicromapDesc));
            uint32_t off = 0, descOff = 0;
            for (uint32_t k = 0; k < W && rc == ommResult_SUCCESS; ++k) {
                Item* it = &items[keys[k].idx];
                if (it->specialIndex != 0) continue;
                if (off >= arraySize || descOff >= descCount) { rc = ommResult_FAILURE; break; }
                res->descArray[descOff].subdivisionLevel = (uint16_t)it->level;
                res->descArray[descOff].format = (uint16_t)it->format;
                res->descArray[descOff].offset = off;
                it->descOffset = descOff++;
                const uint32_t N = 1u << (it->level << 1);
                const uint32_t is2 = it->format == ommFormat_OC1_2_State;
                uint8_t* dst = res->arrayData + off;
                for (uint32_t u = 0; u < N; ++u) {
                    const uint32_t s = it->states[u];
                    const uint8_t val = is2 ? (uint8_t)(s << (u & 7)) : (uint8_t)(s << ((u & 3) << 1));
                    dst[u >> (2 + is2)] |= val;
                }
                const uint32_t adv = (N * bitCount) >> 3;
                off += adv > 1 ? adv : 1;
            }
            res->desc.arrayDataSize = (uint32_t)arraySize;
            res->desc.descArrayCount = descCount;
        }
    }
    if (rc == ommResult_SUCCESS) {
        res->arrayHist = (ommCpuOpacityMicromapUsageCount*)calloc(26, sizeof(ommCpuOpacityMicromapUsageCount));
        res->indexHist = (ommCpuOpacityMicromapUsageCount*)calloc(26, sizeof(ommCpuOpacityMicromapUsageCount));
        uint32_t na = 0, ni = 0;
        for (uint32_t f = 1; f <= 2; ++f)
            for (uint32_t l = 0; l < 13; ++l) {
                if (arrayHist[f][l]) { res->arrayHist[na].count = arrayHist[f][l]; res->arrayHist[na].subdivisionLevel = (uint16_t)l; res->arrayHist[na].format = (uint16_t)f; na++; }
                if (indexHist[f][l]) { res->indexHist[ni].count = indexHist[f][l]; res->indexHist[ni].subdivisionLevel = (uint16_t)l; res->indexHist[ni].format = (uint16_t)f; ni++; }
            }
        res->desc.descArrayHistogramCount = na;
        res->desc.indexHistogramCount = ni;

        res->indexBuffer = (int32_t*)malloc(sizeof(int32_t) * (size_t)(T > 0 ? T : 1));
        for (int32_t i = 0; i < T; ++i) res->indexBuffer[i] = (int32_t)d->unresolvedTriState;
        for (uint32_t w = 0; w < W; ++w)
            for (uint32_t k = 0; k < items[w].numPrims; ++k)
                res->indexBuffer[items[w].prims[k]] = items[w].specialIndex != 0 ? (int32_t)items[w].specialIndex : (int32_t)items[w].descOffset;

        ommIndexFormat ifmt = ommIndexFormat_UINT_32;
        const int allow8 = (flags & ommCpuBakeFlags_Allow8BitIndices) != 0, force32 = (flags & ommCpuBakeFlags_Force32BitIndices) != 0;
        if (allow8 && T <= 127 && !force32) {
            int8_t* b8 = (int8_t*)res->indexBuffer;
            for (int32_t i = 0; i < T; ++i) b8[i] = (int8_t)res->indexBuffer[i];
            ifmt = ommIndexFormat_UINT_8;
        } else if (T <= 32767 && !force32) {
            int16_t* b16 = (int16_t*)res->indexBuffer;
            for (int32_t i = 0; i < T; ++i) b16[i] = (int16_t)res->indexBuffer[i];
            ifmt = ommIndexFormat_UINT_16;
        }
        res->desc.arrayData = res->arrayData;
        res->desc.descArray = res->descArray;
        res->desc.descArrayHistogram = res->arrayHist;
        res->desc.indexBuffer = res->indexBuffer;
        res->desc.indexCount = (uint32_t)T;
        res->desc.indexFormat = ifmt;
        res->desc.indexHistogram = res->indexHist;
    }
    free(keys);
    free_items(items, W);
    return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* C ABI                                                                                        */
/* ------------------------------------------------------------------------------------------ */
OAPI ommLibraryDesc ommGetLibraryDesc(void) {
    ommLibraryDesc d = {OMM_VERSION_MAJOR, OMM_VERSION_MINOR, OMM_VERSION_BUILD};
    return d;
}
OAPI ommResult ommCreateBaker(const ommBakerCreationDesc* desc, ommBaker* outBaker) {
    if (!desc || desc->type != ommBakerType_CPU) return ommResult_INVALID_ARGUMENT;
    OBaker* b = (OBaker*)calloc(1, sizeof(OBaker));
    b->magic = OBAKER_MAGIC;
    b->log = desc->messageInterface;
    *outBaker = (ommBaker)b;
    return ommResult_SUCCESS;
}
OAPI ommResult ommDestroyBaker(ommBaker baker) {
    if (!baker) return ommResult_INVALID_ARGUMENT;
    free(baker);
    return ommResult_SUCCESS;
}
static void free_texture(OTexture* t) {
    if (!t) return;
    for (uint32_t i = 0; t->mips && i < t->mipCount; ++i) { free(t->mips[i].texels); free(t->mips[i].sat); }
    free(t->mips);
    free(t);
}
OAPI ommResult ommCpuCreateTexture(ommBaker baker, const ommCpuTextureDesc* desc, ommCpuTexture* outTexture) {
    if (!baker || !desc) return ommResult_INVALID_ARGUMENT;
    if (desc->mipCount == 0 || desc->format == ommCpuTextureFormat_MAX_NUM) return ommResult_INVALID_ARGUMENT; /* ref: texture_impl.cpp:43-62 */
    for (uint32_t i = 0; i < desc->mipCount; ++i) {
        const ommCpuTextureMipDesc* m = &desc->mips[i];
        if (!m->textureData || m->width == 0 || m->height == 0 || m->width > 65536 || m->height > 65536) return ommResult_INVALID_ARGUMENT;
    }
    OTexture* t = (OTexture*)calloc(1, sizeof(OTexture));
    t->magic = OTEX_MAGIC; t->format = desc->format; t->flags = desc->flags; t->alphaCutoff = desc->alphaCutoff; t->mipCount = desc->mipCount;
    t->mips = (OMip*)calloc(desc->mipCount, sizeof(OMip));
    const size_t spp = desc->format == ommCpuTextureFormat_FP32 ? 4 : 1;
    const int linear = (desc->flags & ommCpuTextureFlags_DisableZOrder) != 0;
    const int enableSAT = desc->alphaCutoff >= 0; /* ref: texture_impl.cpp:91 (numElements is still 0 there) */
    for (uint32_t i = 0; i < desc->mipCount; ++i) {
        const ommCpuTextureMipDesc* s = &desc->mips[i];
        OMip* m = &t->mips[i];
        m->w = (int)s->width; m->h = (int)s->height;
        m->log2w = ctz_slow((uint32_t)m->w); m->log2h = ctz_slow((uint32_t)m->h);
        m->isPow2 = is_pow2(m->w) && is_pow2(m->h);
        m->rcpw = 1.f / (float)m->w; m->rcph = 1.f / (float)m->h;
        m->texels = malloc(spp * (size_t)m->w * (size_t)m->h);
        const uint8_t* src = (const uint8_t*)s->textureData;
        for (int y = 0; y < m->h; ++y) {
            /* ref: texture_impl.cpp:137-184 -- the linear path takes rowPitch in bytes, the Morton path in texels */
            const size_t rowBytes = linear ? (s->rowPitch == 0 ? spp * s->width : s->rowPitch) : spp * (s->rowPitch == 0 ? s->width : s->rowPitch);
            memcpy((uint8_t*)m->texels + spp * (size_t)m->w * (size_t)y, src + rowBytes * (size_t)y, spp * (size_t)m->w);
        }
        if (enableSAT) { /* ref: texture_impl.cpp:191-220 */
            m->sat = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)m->w * (size_t)m->h);
            for (int y = 0; y < m->h; ++y)
                for (int x = 0; x < m->w; ++x) m->sat[x + y * m->w] = tex_load(t, (int)i, x, y) > t->alphaCutoff;
            for (int y = 0; y < m->h; ++y)
                for (int x = 1; x < m->w; ++x) m->sat[x + y * m->w] += m->sat[x - 1 + y * m->w];
            for (int y = 1; y < m->h; ++y)
                for (int x = 0; x < m->w; ++x) m->sat[x + y * m->w] += m->sat[x + (y - 1) * m->w];
        }
    }
    *outTexture = (ommCpuTexture)t;
    return ommResult_SUCCESS;
}
OAPI ommResult ommCpuGetTextureDesc(ommCpuTexture texture, ommCpuTextureDesc* outDesc) { /* ref: texture_impl.cpp:280-325 */
    if (!texture || !outDesc) return ommResult_INVALID_ARGUMENT;
    const OTexture* t = (const OTexture*)texture;
    outDesc->format = t->format; outDesc->flags = t->flags; outDesc->alphaCutoff = t->alphaCutoff; outDesc->mipCount = t->mipCount;
    if (!outDesc->mips) return ommResult_SUCCESS;
    const size_t spp = t->format == ommCpuTextureFormat_FP32 ? 4 : 1;
    for (uint32_t i = 0; i < t->mipCount; ++i) {
        ommCpuTextureMipDesc* m = (ommCpuTextureMipDesc*)&outDesc->mips[i];
        m->width = (uint32_t)t->mips[i].w; m->height = (uint32_t)t->mips[i].h; m->rowPitch = (uint32_t)t->mips[i].w;
        if (m->textureData) memcpy((void*)m->textureData, t->mips[i].texels, spp * (size_t)m->width * m->height);
    }
    return ommResult_SUCCESS;
}
OAPI ommResult ommCpuDestroyTexture(ommBaker baker, ommCpuTexture texture) {
    (void)baker;
    if (!texture) return ommResult_INVALID_ARGUMENT;
    free_texture((OTexture*)texture);
    return ommResult_SUCCESS;
}
static void free_result(OResult* r) {
    if (!r) return;
    free(r->arrayData); free(r->descArray); free(r->arrayHist); free(r->indexHist); free(r->indexBuffer); free(r);
}
static int state_compatible(ommOpacityState s, ommFormat f) { return f == ommFormat_OC1_2_State ? (s == ommOpacityState_Opaque || s == ommOpacityState_Transparent) : 1; }
OAPI ommResult ommCpuBake(ommBaker baker, const ommCpuBakeInputDesc* d, ommCpuBakeResult* outBakeResult) {
    if (!baker || !d) return ommResult_INVALID_ARGUMENT;
    /* ref: bake_cpu_impl.cpp:235-290 (messages are not restated; the product's strings are tested against the SDK's directly) */
    if (!d->texture || ((const OTexture*)d->texture)->magic != OTEX_MAGIC) return ommResult_INVALID_ARGUMENT;
    if (d->alphaMode == ommAlphaMode_MAX_NUM || d->runtimeSamplerDesc.addressingMode == ommTextureAddressMode_MAX_NUM ||
        d->runtimeSamplerDesc.filter == ommTextureFilterMode_MAX_NUM || d->texCoordFormat == ommTexCoordFormat_MAX_NUM || !d->texCoords ||
        d->indexFormat == ommIndexFormat_MAX_NUM || !d->indexBuffer || d->indexCount == 0 || d->maxSubdivisionLevel > 12)
        return ommResult_INVALID_ARGUMENT;
    const uint32_t flags = (uint32_t)d->bakeFlags;
    if ((flags & (ommCpuBakeFlags_EnableNearDuplicateDetection | FLAG_NEAR_DUP_BRUTE)) && (flags & ommCpuBakeFlags_DisableDuplicateDetection)) return ommResult_INVALID_ARGUMENT;
    if ((flags & ommCpuBakeFlags_EnableValidation) && !((OBaker*)baker)->log.messageCallback) return ommResult_INVALID_ARGUMENT;
    const OTexture* tex = (const OTexture*)d->texture;
    if (tex->alphaCutoff >= 0.f && tex->alphaCutoff != d->alphaCutoff) return ommResult_INVALID_ARGUMENT;
    if (!state_compatible(d->alphaCutoffGreater, d->format) || !state_compatible(d->alphaCutoffLessEqual, d->format)) return ommResult_INVALID_ARGUMENT;
    if (flags & (ommCpuBakeFlags_EnableNearDuplicateDetection | FLAG_NEAR_DUP_BRUTE)) return ommResult_NOT_IMPLEMENTED;
    if (d->maxArrayDataSize != 0xFFFFFFFFu) return ommResult_NOT_IMPLEMENTED;
    OResult* r = (OResult*)calloc(1, sizeof(OResult));
    r->magic = ORESULT_MAGIC;
    ommResult rc = bake_impl(d, r);
    if (rc != ommResult_SUCCESS) { free_result(r); return rc; }
    *outBakeResult = (ommCpuBakeResult)r;
    return ommResult_SUCCESS;
}
OAPI ommResult ommCpuDestroyBakeResult(ommCpuBakeResult bakeResult) {
    if (!bakeResult) return ommResult_INVALID_ARGUMENT;
    free_result((OResult*)bakeResult);
    return ommResult_SUCCESS;
}
OAPI ommResult ommCpuGetBakeResultDesc(ommCpuBakeResult bakeResult, const ommCpuBakeResultDesc** desc) {
    if (!bakeResult || !desc) return ommResult_INVALID_ARGUMENT;
    *desc = &((OResult*)bakeResult)->desc;
    return ommResult_SUCCESS;
}

/* helpers exported for unit tests of the restated pieces (golden vectors live in tests/golden/) */
OAPI uint64_t omm_oracle_std_hash_float(float v) { return std_hash_float(v); }
OAPI void omm_oracle_index2bary(uint32_t index, uint32_t level, float* out6) {
    f2 a, b, c;
    index2bary(index, level, &a, &b, &c);
    out6[0] = a.x; out6[1] = a.y; out6[2] = b.x; out6[3] = b.y; out6[4] = c.x; out6[5] = c.y;
}
OAPI int omm_oracle_texcoord(int mode, int pow2, int c, int size) { return addr1((ommTextureAddressMode)mode, pow2, c, size, ctz_slow((uint32_t)size)); }
OAPI uint32_t omm_oracle_morton(uint32_t x, uint32_t y) { return bit_interleave16(x, y); }
