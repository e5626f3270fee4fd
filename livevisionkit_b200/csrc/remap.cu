// K8 (exact build; the default build is remap_fast.cu) — FSR-EASU warp/remap for sm_100a, bit-exact twin of
// oracle/easu_ref.c.  Selected by LVKB200_REMAP_EXACT=1 / lvkb200_set_remap_exact(1), and used for sources that are not
// 4-byte aligned.
//
// Replaces lvk::remap (LiveVisionKit/Functions/Image.cpp:28-151) and its OpenCL kernels
// easu_remap / easu_remap_homography / easu (Functions/OpenCL/Sources/FSR.cl:98-318,362-452), and lvk::upscale
// (Image.cpp:155-201) with its kernel easu_scale (FSR.cl:326-358): the same kernel, MODE 2.
//
// Design (not a translation of the 8x8 OpenCL work-groups):
//   * one CTA = a 32x8 destination tile, one destination pixel per thread;
//   * the warp is near-identity, so the source footprint of a tile is the tile plus a small margin:
//     the CTA computes the exact source bounding box of its EASU pixels, stages it ONCE into shared
//     memory as float4 {c0,c1,c2,luma}/255 (each source byte is converted once per tile instead of
//     12x per destination pixel) and every thread then gathers its 12 taps with LDS.128;
//   * tiles whose footprint does not fit (extreme warps) fall back to direct global reads;
//   * border band -> nearest neighbour, outside -> background, exactly as FSR.cl:387-399/436-448.
// Arithmetic: IEEE float32.  Every a*b+c that FSR.cl writes as one expression is ONE fused multiply-add
// (__fmaf_rn), exactly as in the CPU restatement (oracle/easu_ref.c, fmaf); the translation unit is compiled with
// --fmad=false so that nothing ELSE is contracted -> results are bit-identical to the oracle (tests assert 0 LSB).
// Roofline: 6 B/px algorithmic (3 read + 3 written); see DESIGN.md.

#include <cstdlib>

#include "common.hpp"
#include "remap_common.cuh"

namespace lvkb200
{
namespace
{

constexpr int TILE_W = 32;
constexpr int TILE_H = 8;

// FSR.cl:131-176
template <int CORNER>
__device__ __forceinline__ void easu_accumulate(float& dirx, float& diry, float& len, float ppx, float ppy, float lA,
                                                float lB, float lC, float lD, float lE)
{
    float w;
    if (CORNER == 0) w = (1.0f - ppx) * (1.0f - ppy);
    if (CORNER == 1) w = ppx * (1.0f - ppy);
    if (CORNER == 2) w = (1.0f - ppx) * ppy;
    if (CORNER == 3) w = ppx * ppy;

    float dc = lD - lC;
    float cb = lC - lB;
    float lenX = aprx_lo_rcp(fmaxf(fabsf(dc), fabsf(cb)));
    float dirX = lD - lB;
    dirx = __fmaf_rn(dirX, w, dirx);
    lenX = sat01(fabsf(dirX) * lenX);
    lenX *= lenX;
    len = __fmaf_rn(lenX, w, len);

    float ec = lE - lC;
    float ca = lC - lA;
    float lenY = aprx_lo_rcp(fmaxf(fabsf(ec), fabsf(ca)));
    float dirY = lE - lA;
    diry = __fmaf_rn(dirY, w, diry);
    lenY = sat01(fabsf(dirY) * lenY);
    lenY *= lenY;
    len = __fmaf_rn(lenY, w, len);
}

// FSR.cl:98-126
__device__ __forceinline__ void easu_tap(float& aCx, float& aCy, float& aCz, float& aW, float offx, float offy,
                                         float dirx, float diry, float lenx, float leny, float lob, float clp,
                                         const float4& c)
{
    float vx = __fmaf_rn(offx, dirx, offy * diry);
    float vy = __fmaf_rn(offx, -diry, offy * dirx);
    vx *= lenx;
    vy *= leny;
    float d2 = fminf(__fmaf_rn(vx, vx, vy * vy), clp);
    float wA = __fmaf_rn(lob, d2, -1.0f);
    float wB = __fmaf_rn(2.0f / 5.0f, d2, -1.0f);
    wA *= wA;
    wB = __fmaf_rn(25.0f / 16.0f, wB * wB, -(25.0f / 16.0f - 1.0f));
    float w = wB * wA;
    aCx = __fmaf_rn(c.x, w, aCx);
    aCy = __fmaf_rn(c.y, w, aCy);
    aCz = __fmaf_rn(c.z, w, aCz);
    aW += w;
}

// FSR.cl:181-318.  TAP(dx,dy) returns {c0,c1,c2,luma} (already /255) of source pixel f+(dx,dy).
template <typename Tap>
__device__ __forceinline__ uchar3 easu(const Tap& tap, float ppx, float ppy)
{
    const float4 b = tap(0, -1), c = tap(1, -1);
    const float4 e = tap(-1, 0), f = tap(0, 0), g = tap(1, 0), h = tap(2, 0);
    const float4 i = tap(-1, 1), j = tap(0, 1), k = tap(1, 1), l = tap(2, 1);
    const float4 n = tap(0, 2), o = tap(1, 2);

    float len = 0.0f, dirx = 0.0f, diry = 0.0f;
    easu_accumulate<0>(dirx, diry, len, ppx, ppy, b.w, e.w, f.w, g.w, j.w);
    easu_accumulate<1>(dirx, diry, len, ppx, ppy, c.w, f.w, g.w, h.w, k.w);
    easu_accumulate<2>(dirx, diry, len, ppx, ppy, f.w, i.w, j.w, k.w, n.w);
    easu_accumulate<3>(dirx, diry, len, ppx, ppy, g.w, j.w, k.w, l.w, o.w);

    float dirR = __fmaf_rn(dirx, dirx, diry * diry);
    const bool zro = dirR < (1.0f / 32768.0f);
    dirR = aprx_lo_rsq(dirR);
    dirR = zro ? 1.0f : dirR;
    dirx = zro ? 1.0f : dirx;
    dirx *= dirR;
    diry *= dirR;

    len = len * 0.5f;
    len *= len;

    const float stretch = __fmaf_rn(dirx, dirx, diry * diry) * aprx_lo_rcp(fmaxf(fabsf(dirx), fabsf(diry)));
    const float len2x = __fmaf_rn(stretch - 1.0f, len, 1.0f);
    const float len2y = __fmaf_rn(-0.5f, len, 1.0f);
    const float lob = __fmaf_rn((1.0f / 4.0f - 0.04f) - 0.5f, len, 0.5f);
    const float clp = aprx_lo_rcp(lob);

    const float mi0 = fminf(f.x, fminf(g.x, fminf(j.x, k.x)));
    const float mi1 = fminf(f.y, fminf(g.y, fminf(j.y, k.y)));
    const float mi2 = fminf(f.z, fminf(g.z, fminf(j.z, k.z)));
    const float ma0 = fmaxf(f.x, fmaxf(g.x, fmaxf(j.x, k.x)));
    const float ma1 = fmaxf(f.y, fmaxf(g.y, fmaxf(j.y, k.y)));
    const float ma2 = fmaxf(f.z, fmaxf(g.z, fmaxf(j.z, k.z)));

    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, aW = 0.0f;
    easu_tap(a0, a1, a2, aW, 0.0f - ppx, -1.0f - ppy, dirx, diry, len2x, len2y, lob, clp, b);
    easu_tap(a0, a1, a2, aW, 1.0f - ppx, -1.0f - ppy, dirx, diry, len2x, len2y, lob, clp, c);
    easu_tap(a0, a1, a2, aW, -1.0f - ppx, 1.0f - ppy, dirx, diry, len2x, len2y, lob, clp, i);
    easu_tap(a0, a1, a2, aW, 0.0f - ppx, 1.0f - ppy, dirx, diry, len2x, len2y, lob, clp, j);
    easu_tap(a0, a1, a2, aW, 0.0f - ppx, 0.0f - ppy, dirx, diry, len2x, len2y, lob, clp, f);
    easu_tap(a0, a1, a2, aW, -1.0f - ppx, 0.0f - ppy, dirx, diry, len2x, len2y, lob, clp, e);
    easu_tap(a0, a1, a2, aW, 1.0f - ppx, 1.0f - ppy, dirx, diry, len2x, len2y, lob, clp, k);
    easu_tap(a0, a1, a2, aW, 2.0f - ppx, 1.0f - ppy, dirx, diry, len2x, len2y, lob, clp, l);
    easu_tap(a0, a1, a2, aW, 2.0f - ppx, 0.0f - ppy, dirx, diry, len2x, len2y, lob, clp, h);
    easu_tap(a0, a1, a2, aW, 1.0f - ppx, 0.0f - ppy, dirx, diry, len2x, len2y, lob, clp, g);
    easu_tap(a0, a1, a2, aW, 0.0f - ppx, 2.0f - ppy, dirx, diry, len2x, len2y, lob, clp, n);
    easu_tap(a0, a1, a2, aW, 1.0f - ppx, 2.0f - ppy, dirx, diry, len2x, len2y, lob, clp, o);

    const float rcpW = 1.0f / aW;  // native_recip
    const float v0 = fminf(ma0, fmaxf(mi0, a0 * rcpW));
    const float v1 = fminf(ma1, fmaxf(mi1, a1 * rcpW));
    const float v2 = fminf(ma2, fmaxf(mi2, a2 * rcpW));
    uchar3 out;  // convert_uchar3: truncation
    out.x = (unsigned char)__float2int_rz(v0 * 255.0f);
    out.y = (unsigned char)__float2int_rz(v1 * 255.0f);
    out.z = (unsigned char)__float2int_rz(v2 * 255.0f);
    return out;
}

template <bool YUV>
__device__ __forceinline__ float4 load_texel(const uint8_t* __restrict__ p)
{
    const float norm = 0.00392156862f;
    float4 t;
    t.x = u8_to_float(__ldg(p)) * norm;
    t.y = u8_to_float(__ldg(p + 1)) * norm;
    t.z = u8_to_float(__ldg(p + 2)) * norm;
    // FSR.cl:229-241 (the #ifndef is inverted relative to its comments; reproduced as written)
    t.w = YUV ? __fmaf_rn(t.z, 0.5f, __fmaf_rn(t.x, 0.5f, t.y)) : t.x;
    return t;
}

template <bool YUV>
struct GlobalTap
{
    const uint8_t* base;  // &src[sy * pitch + 3 * sx]
    size_t pitch;
    __device__ __forceinline__ float4 operator()(int dx, int dy) const
    {
        return load_texel<YUV>(base + (ptrdiff_t)dy * (ptrdiff_t)pitch + 3 * dx);
    }
};

template <bool YUV>
__device__ __noinline__ uchar3 easu_global(const uint8_t* __restrict__ base, size_t pitch, float ppx, float ppy)
{
    GlobalTap<YUV> tap{base, pitch};
    return easu(tap, ppx, ppy);
}

// FSR.cl:131-176, the part that depends only on the SOURCE pixel C and its cross A(up) B(left) D(right) E(down):
// {dirX, dirY, sat(|dirX|*rcp(max(|D-C|,|C-B|)))^2, sat(|dirY|*rcp(max(|E-C|,|C-A|)))^2}.  Evaluated once per source
// pixel of the staged tile instead of once per (destination pixel, corner): same operations, same results.
__device__ __forceinline__ float4 easu_direction_terms(float lA, float lB, float lC, float lD, float lE)
{
    const f2 de = pk(lD, lE), cc = pk1(lC), ba = pk(lB, lA);
    const f2 dc_ec = add2(de, neg2(cc));  // (D-C, E-C)
    const f2 cb_ca = add2(cc, neg2(ba));  // (C-B, C-A)
    const f2 dir = add2(de, neg2(ba));    // (D-B, E-A)
    float lenX = aprx_lo_rcp(fmaxf(fabsf(dc_ec.x), fabsf(cb_ca.x)));
    float lenY = aprx_lo_rcp(fmaxf(fabsf(dc_ec.y), fabsf(cb_ca.y)));
    f2 len = pk(sat01(fabsf(dir.x) * lenX), sat01(fabsf(dir.y) * lenY));
    len = mul2(len, len);
    return make_float4(dir.x, dir.y, len.x, len.y);
}

// FSR.cl:98-126 for the two pixels of a pair at once.  t1 = offy*diry, t2 = offy*dirx (shared by the taps of one row).
__device__ __forceinline__ void easu_tap2(float (&aA)[3], float (&aB)[3], f2& aW, f2 offx, f2 t1, f2 t2, f2 dirx,
                                          f2 ndiry, f2 lenx, f2 leny, f2 lob, f2 clp, const float4& cA,
                                          const float4& cB)
{
    f2 vx = fma2(offx, dirx, t1);
    f2 vy = fma2(offx, ndiry, t2);
    vx = mul2(vx, lenx);
    vy = mul2(vy, leny);
    f2 d2 = fma2(vx, vx, mul2(vy, vy));
    d2.x = fminf(d2.x, clp.x);
    d2.y = fminf(d2.y, clp.y);
    f2 wA = fma2(lob, d2, pk1(-1.0f));
    f2 wB = fma2(pk1(2.0f / 5.0f), d2, pk1(-1.0f));
    wA = mul2(wA, wA);
    wB = fma2(pk1(25.0f / 16.0f), mul2(wB, wB), pk1(-(25.0f / 16.0f - 1.0f)));
    const f2 w = mul2(wB, wA);
    aA[0] = __fmaf_rn(cA.x, w.x, aA[0]);
    aA[1] = __fmaf_rn(cA.y, w.x, aA[1]);
    aA[2] = __fmaf_rn(cA.z, w.x, aA[2]);
    aB[0] = __fmaf_rn(cB.x, w.y, aB[0]);
    aB[1] = __fmaf_rn(cB.y, w.y, aB[1]);
    aB[2] = __fmaf_rn(cB.z, w.y, aB[2]);
    aW = add2(aW, w);
}

__device__ __forceinline__ uchar3 easu_resolve(const float (&a)[3], float aW, const float4& f, const float4& g,
                                               const float4& j, const float4& k)
{
    const float mi0 = fminf(f.x, fminf(g.x, fminf(j.x, k.x)));
    const float mi1 = fminf(f.y, fminf(g.y, fminf(j.y, k.y)));
    const float mi2 = fminf(f.z, fminf(g.z, fminf(j.z, k.z)));
    const float ma0 = fmaxf(f.x, fmaxf(g.x, fmaxf(j.x, k.x)));
    const float ma1 = fmaxf(f.y, fmaxf(g.y, fmaxf(j.y, k.y)));
    const float ma2 = fmaxf(f.z, fmaxf(g.z, fmaxf(j.z, k.z)));
    const float rcpW = 1.0f / aW;  // native_recip
    const float v0 = fminf(ma0, fmaxf(mi0, a[0] * rcpW));
    const float v1 = fminf(ma1, fmaxf(mi1, a[1] * rcpW));
    const float v2 = fminf(ma2, fmaxf(mi2, a[2] * rcpW));
    uchar3 out;  // convert_uchar3: truncation
    out.x = (unsigned char)__float2int_rz(v0 * 255.0f);
    out.y = (unsigned char)__float2int_rz(v1 * 255.0f);
    out.z = (unsigned char)__float2int_rz(v2 * 255.0f);
    return out;
}

// FSR.cl:181-318 for a pixel pair, taps and per-source-pixel direction terms read from the staged tile.
// tA / tB = &tile[f], dA / dB = &terms[f] of pixel A / B; pp* = fractional source position.
template <int STRIDE>
__device__ __forceinline__ void easu_pair(const float4* __restrict__ tA, const float4* __restrict__ tB,
                                          const float4* __restrict__ dA, const float4* __restrict__ dB, f2 ppx, f2 ppy,
                                          uchar3& outA, uchar3& outB)
{
    // ---- direction / length: bilinear blend of the four corners' terms (order f, g, j, k as FSR.cl:254-257)
    const f2 omx = add2(pk1(1.0f), neg2(ppx)), omy = add2(pk1(1.0f), neg2(ppy));
    const f2 w0 = mul2(omx, omy), w1 = mul2(ppx, omy), w2 = mul2(omx, ppy), w3 = mul2(ppx, ppy);
    f2 dirx, diry, len;
    {
        float lenA = 0.0f, dxA = 0.0f, dyA = 0.0f, lenB = 0.0f, dxB = 0.0f, dyB = 0.0f;
#define LVKB_ACC(OFF, WA, WB)                                                                                         \
    {                                                                                                                 \
        const float4 qa = dA[OFF], qb = dB[OFF];                                                                      \
        dxA = __fmaf_rn(qa.x, WA, dxA); lenA = __fmaf_rn(qa.z, WA, lenA);                                             \
        dyA = __fmaf_rn(qa.y, WA, dyA); lenA = __fmaf_rn(qa.w, WA, lenA);                                             \
        dxB = __fmaf_rn(qb.x, WB, dxB); lenB = __fmaf_rn(qb.z, WB, lenB);                                             \
        dyB = __fmaf_rn(qb.y, WB, dyB); lenB = __fmaf_rn(qb.w, WB, lenB);                                             \
    }
        LVKB_ACC(0, w0.x, w0.y)
        LVKB_ACC(1, w1.x, w1.y)
        LVKB_ACC(STRIDE, w2.x, w2.y)
        LVKB_ACC(STRIDE + 1, w3.x, w3.y)
#undef LVKB_ACC
        dirx = pk(dxA, dxB); diry = pk(dyA, dyB); len = pk(lenA, lenB);
    }

    f2 dirR = fma2(dirx, dirx, mul2(diry, diry));
    const bool zA = dirR.x < (1.0f / 32768.0f), zB = dirR.y < (1.0f / 32768.0f);
    dirR = pk(zA ? 1.0f : aprx_lo_rsq(dirR.x), zB ? 1.0f : aprx_lo_rsq(dirR.y));
    dirx = pk(zA ? 1.0f : dirx.x, zB ? 1.0f : dirx.y);
    dirx = mul2(dirx, dirR);
    diry = mul2(diry, dirR);

    len = mul2(len, pk1(0.5f));
    len = mul2(len, len);

    f2 stretch = fma2(dirx, dirx, mul2(diry, diry));
    stretch = mul2(stretch, pk(aprx_lo_rcp(fmaxf(fabsf(dirx.x), fabsf(diry.x))),
                               aprx_lo_rcp(fmaxf(fabsf(dirx.y), fabsf(diry.y)))));
    // (stretch - 1) in scalar: a packed add fed by a packed mul would be contracted by ptxas (see source_position)
    const f2 len2x = fma2(pk(__fadd_rn(stretch.x, -1.0f), __fadd_rn(stretch.y, -1.0f)), len, pk1(1.0f));
    const f2 len2y = fma2(pk1(-0.5f), len, pk1(1.0f));
    const f2 lob = fma2(pk1((1.0f / 4.0f - 0.04f) - 0.5f), len, pk1(0.5f));
    const f2 clp = pk(aprx_lo_rcp(lob.x), aprx_lo_rcp(lob.y));
    const f2 ndiry = neg2(diry);

    // ---- the 12 taps, in the accumulation order of FSR.cl:291-302 (b c i j f e k l h g n o)
    const f2 nppx = neg2(ppx), nppy = neg2(ppy);
    const f2 oxm = add2(pk1(-1.0f), nppx), ox0 = add2(pk1(0.0f), nppx), ox1 = add2(pk1(1.0f), nppx),
             ox2 = add2(pk1(2.0f), nppx);
    const f2 oym = add2(pk1(-1.0f), nppy), oy0 = add2(pk1(0.0f), nppy), oy1 = add2(pk1(1.0f), nppy),
             oy2 = add2(pk1(2.0f), nppy);
    const f2 t1m = mul2(oym, diry), t10 = mul2(oy0, diry), t11 = mul2(oy1, diry), t12 = mul2(oy2, diry);
    const f2 t2m = mul2(oym, dirx), t20 = mul2(oy0, dirx), t21 = mul2(oy1, dirx), t22 = mul2(oy2, dirx);

    float aA[3] = {0.0f, 0.0f, 0.0f}, aB[3] = {0.0f, 0.0f, 0.0f};
    f2 aW = pk1(0.0f);
#define LVKB_TAP(DX, DY, OX, T1, T2)                                                                                  \
    easu_tap2(aA, aB, aW, OX, T1, T2, dirx, ndiry, len2x, len2y, lob, clp, tA[(DY) * STRIDE + (DX)],                  \
              tB[(DY) * STRIDE + (DX)]);
    LVKB_TAP(0, -1, ox0, t1m, t2m)   // b
    LVKB_TAP(1, -1, ox1, t1m, t2m)   // c
    LVKB_TAP(-1, 1, oxm, t11, t21)   // i
    LVKB_TAP(0, 1, ox0, t11, t21)    // j
    LVKB_TAP(0, 0, ox0, t10, t20)    // f
    LVKB_TAP(-1, 0, oxm, t10, t20)   // e
    LVKB_TAP(1, 1, ox1, t11, t21)    // k
    LVKB_TAP(2, 1, ox2, t11, t21)    // l
    LVKB_TAP(2, 0, ox2, t10, t20)    // h
    LVKB_TAP(1, 0, ox1, t10, t20)    // g
    LVKB_TAP(0, 2, ox0, t12, t22)    // n
    LVKB_TAP(1, 2, ox1, t12, t22)    // o
#undef LVKB_TAP

    outA = easu_resolve(aA, aW.x, tA[0], tA[1], tA[STRIDE], tA[STRIDE + 1]);
    outB = easu_resolve(aB, aW.y, tB[0], tB[1], tB[STRIDE], tB[STRIDE + 1]);
}

// Destination tile of one CTA: 32 x 16 pixels, 256 threads, thread (tx, ty) owns the pair (x, y) and (x, y + 8).
constexpr int PAIR_DY = 8;
constexpr int CTA_H = 2 * PAIR_DY;
constexpr int STG_W = 44;  // staged source tile capacity (pixels): 32 + 3 taps + warp slack, <= 2 * TILE_W
constexpr int STG_H = 28;  // 16 + 3 taps + warp slack, <= 4 * TILE_H

// ceil(65536 / d) for d = 1 .. STG_W: row = (i * c_inv16[width]) >> 16 == i / width for i < 65536 / width
struct Inv16Table
{
    unsigned short v[STG_W + 1];
    constexpr Inv16Table() : v{}
    {
        v[0] = 0;
        for (int d = 1; d <= STG_W; d++) v[d] = (unsigned short)((65536 + d - 1) / d > 65535 ? 65535 : (65536 + d - 1) / d);
    }
};
__constant__ Inv16Table c_inv16_table = Inv16Table();
#define c_inv16 c_inv16_table.v

template <int MODE, bool YUV, int OCC>
__global__ void __launch_bounds__(TILE_W* TILE_H, OCC)
    k_easu_remap(const uint8_t* __restrict__ src, size_t src_pitch, uint8_t* __restrict__ dst, size_t dst_pitch, int W,
                 int H, int dW, int dH, Transform T, MeshArgs M, uchar3 bg)
{
    // W x H = source image (border classification), dW x dH = destination image (== source except for MODE 2)
    __shared__ float4 tile[STG_H * STG_W];   // {c0, c1, c2, luma} / 255 of the staged source pixels
    __shared__ float4 terms[STG_H * STG_W];  // easu_direction_terms of the same pixels (interior only)
    __shared__ float luma[STG_H * STG_W];    // tile[].w again, contiguous: the 5-point cross reads are conflict-free
    __shared__ int bbox[4];                  // minx, miny, maxx, maxy of f over the EASU pixels of this tile

    const int tx = threadIdx.x & (TILE_W - 1);
    const int ty = threadIdx.x / TILE_W;
    const int x = blockIdx.x * TILE_W + tx;
    const int yA = blockIdx.y * CTA_H + ty, yB = yA + PAIR_DY;
    const bool insideA = (x < dW) && (yA < dH), insideB = (x < dW) && (yB < dH);

    if (threadIdx.x == 0)
    {
        bbox[0] = INT_MAX; bbox[1] = INT_MAX; bbox[2] = INT_MIN; bbox[3] = INT_MIN;
    }

    float sxA, syA, sxB, syB;
    // scalar on purpose: ptxas fuses mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (even with explicit .rn and --fmad=false),
    // which would change the rounding of these mul-then-add expressions
    source_position<MODE>(x, yA, W, H, T, M, sxA, syA);
    source_position<MODE>(x, yB, W, H, T, M, sxB, syB);
    const PixelClass A = classify(sxA, syA, W, H, insideA), B = classify(sxB, syB, W, H, insideB);

    __syncthreads();
    {
        // warp-level reduce, one shared atomic per warp
        const int lminx = min(A.do_easu ? A.sx : INT_MAX, B.do_easu ? B.sx : INT_MAX);
        const int lminy = min(A.do_easu ? A.sy : INT_MAX, B.do_easu ? B.sy : INT_MAX);
        const int lmaxx = max(A.do_easu ? A.sx : INT_MIN, B.do_easu ? B.sx : INT_MIN);
        const int lmaxy = max(A.do_easu ? A.sy : INT_MIN, B.do_easu ? B.sy : INT_MIN);
        const int wminx = __reduce_min_sync(0xffffffffu, lminx), wminy = __reduce_min_sync(0xffffffffu, lminy);
        const int wmaxx = __reduce_max_sync(0xffffffffu, lmaxx), wmaxy = __reduce_max_sync(0xffffffffu, lmaxy);
        if ((threadIdx.x & 31) == 0 && wminx != INT_MAX)
        {
            atomicMin(&bbox[0], wminx); atomicMin(&bbox[1], wminy);
            atomicMax(&bbox[2], wmaxx); atomicMax(&bbox[3], wmaxy);
        }
    }
    __syncthreads();

    const bool any_easu = bbox[0] != INT_MAX;
    const int x0 = bbox[0] - 1, y0 = bbox[1] - 1;                    // taps reach f-1 .. f+2
    const int bw = bbox[2] + 2 - x0 + 1, bh = bbox[3] + 2 - y0 + 1;  // all inside the image (border band excluded)
    const bool staged = any_easu && bw <= STG_W && bh <= STG_H;

    if (staged)
    {
        // The window's bw x bh texels as ONE linear index space, 256 consecutive texels per pass: every pass but the
        // last keeps all lanes busy (a fixed row/column slot grid ran 8 slots per thread for ~2.8 texels of work,
        // with all of their instructions issued predicated-off).  row = i / bw by multiply-shift, exact for
        // i < 65536 / (bw - 1) (i < 1232, bw <= 44).  Lanes read consecutive pixels of a row.
        static_assert(STG_W * STG_H <= 5 * TILE_W * TILE_H && STG_W * STG_H * (STG_W - 1) < 65536, "flat staging index");
        constexpr int CTA = TILE_W * TILE_H;
        const int n_tex = bw * bh;
        const unsigned inv_bw = c_inv16[bw];  // ceil(65536 / bw): a table, not a per-thread integer division
        const uint8_t* const p00 = src + (size_t)y0 * src_pitch + 3 * x0;
        auto stage_texel = [&](int i) {
            const int r = (int)(((unsigned)i * inv_bw) >> 16), c = i - r * bw;
            const float4 v = load_texel<YUV>(p00 + (size_t)r * src_pitch + 3 * c);
            tile[r * STG_W + c] = v;
            luma[r * STG_W + c] = v.w;
        };
        // a full tile's window has > 2 * 256 texels: two unconditional-shaped passes, then the (short) rest
        if ((int)threadIdx.x < n_tex) stage_texel((int)threadIdx.x);
        if ((int)threadIdx.x + CTA < n_tex) stage_texel((int)threadIdx.x + CTA);
        for (int i = (int)threadIdx.x + 2 * CTA; i < n_tex; i += CTA) stage_texel(i);
        __syncthreads();
        // direction terms of the texels that can be a corner f/g/j/k: columns 1 .. bw-2, rows 1 .. bh-2, same flat walk
        const int iw = bw - 2, n_int = iw * (bh - 2);
        const unsigned inv_iw = c_inv16[iw];
        auto terms_texel = [&](int i) {
            const int r = (int)(((unsigned)i * inv_iw) >> 16), c = i - r * iw;
            const float* l = &luma[(r + 1) * STG_W + c + 1];
            terms[(r + 1) * STG_W + c + 1] = easu_direction_terms(l[-STG_W], l[-1], l[0], l[1], l[STG_W]);
        };
        if ((int)threadIdx.x < n_int) terms_texel((int)threadIdx.x);
        if ((int)threadIdx.x + CTA < n_int) terms_texel((int)threadIdx.x + CTA);
        for (int i = (int)threadIdx.x + 2 * CTA; i < n_int; i += CTA) terms_texel(i);
    }
    __syncthreads();

    uchar3 outA = bg, outB = bg;
    if (staged && (A.do_easu || B.do_easu))
    {
        // a lane whose pixel is not an EASU pixel computes on its partner's taps and discards the result
        const int iA = A.do_easu ? (A.sy - y0) * STG_W + (A.sx - x0) : (B.sy - y0) * STG_W + (B.sx - x0);
        const int iB = B.do_easu ? (B.sy - y0) * STG_W + (B.sx - x0) : iA;
        uchar3 eA, eB;
        easu_pair<STG_W>(&tile[iA], &tile[iB], &terms[iA], &terms[iB], pk(A.ppx, B.ppx), pk(A.ppy, B.ppy), eA, eB);
        if (A.do_easu) outA = eA;
        if (B.do_easu) outB = eB;
    }
    else
    {
        // extreme warp: the footprint does not fit the staging tile -> direct global reads, one pixel at a time
        if (A.do_easu) outA = easu_global<YUV>(src + (size_t)A.sy * src_pitch + 3 * A.sx, src_pitch, A.ppx, A.ppy);
        if (B.do_easu) outB = easu_global<YUV>(src + (size_t)B.sy * src_pitch + 3 * B.sx, src_pitch, B.ppx, B.ppy);
    }
    // border band (FSR.cl:387-399): nearest neighbour.  Only tiles on the frame's edge have any -> one warp-wide test
    if (__any_sync(0xffffffffu, (A.border && A.in_src) || (B.border && B.in_src)))
    {
        if (A.border && A.in_src)
        {
            const uint8_t* p = src + (size_t)A.sy * src_pitch + 3 * A.sx;
            outA.x = __ldg(p); outA.y = __ldg(p + 1); outA.z = __ldg(p + 2);
        }
        if (B.border && B.in_src)
        {
            const uint8_t* p = src + (size_t)B.sy * src_pitch + 3 * B.sx;
            outB.x = __ldg(p); outB.y = __ldg(p + 1); outB.z = __ldg(p + 2);
        }
    }

    if (insideA)
    {
        uint8_t* q = dst + (size_t)yA * dst_pitch + 3 * x;
        q[0] = outA.x; q[1] = outA.y; q[2] = outA.z;
    }
    if (insideB)
    {
        uint8_t* q = dst + (size_t)yB * dst_pitch + 3 * x;
        q[0] = outB.x; q[1] = outB.y; q[2] = outB.z;
    }
}

// =====================================================================================================================
// Persistent, software-pipelined variant (LVKB200_REMAP_KERNEL=3; measured slower than the default): one CTA per (SM x OCC) walks over destination tiles; while tile
// n is converted / filtered, the raw source bytes of tile n+1 are already in flight as TMA bulk copies
// (cp.async.bulk global -> shared, one per source row, completion counted by an mbarrier), so no warp ever waits on a
// global load.  The source window of a tile is PLANNED from the tile's four corners (a projective map is monotone
// along rows and columns, so the corners bound the footprint); it is only a prefetch hint: every pixel re-checks that
// its 12 taps lie inside the staged window and otherwise takes the direct-global path, so the result never depends
// on the plan.
// =====================================================================================================================

constexpr int RAW_PITCH = 160;               // bytes per staged raw row: 3 * STG_W = 132, + 15 alignment, rounded to 16
constexpr int RAW_BYTES = STG_H * RAW_PITCH;  // one raw buffer
constexpr int V3_SMEM = STG_H * STG_W * (2 * (int)sizeof(float4) + (int)sizeof(float)) + 2 * RAW_BYTES;

struct TilePlan
{
    int x0, y0;   // top-left source pixel of the staged window
    int bw, bh;   // window size in pixels (bw == 0: nothing staged)
    int off;      // byte offset of pixel x0 inside a raw row (alignment of the bulk copy)
    int row_bytes;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t done;
    do
    {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Executed by warp 0: plans the source window of destination tile `t` into plan[b] and, on the bulk path, starts its
// row copies into raw buffer b.
template <int MODE>
__device__ __forceinline__ void plan_and_fetch(int t, int b, int tiles_x, int W, int H, const Transform& T,
                                               const MeshArgs& M, const uint8_t* __restrict__ src, size_t src_pitch,
                                               bool bulk, TilePlan* plan, uint8_t* raw, uint64_t* full)
{
    const int lane = threadIdx.x & 31;
    const int tyi = t / tiles_x, txi = t - tyi * tiles_x;
    const int X0 = txi * TILE_W, Y0 = tyi * CTA_H;
    const int X1 = min(X0 + TILE_W - 1, W - 1), Y1 = min(Y0 + CTA_H - 1, H - 1);
    float fxs, fys;
    source_position<MODE>((lane & 1) ? X1 : X0, (lane & 2) ? Y1 : Y0, W, H, T, M, fxs, fys);
    // clamp before the conversion so that NaN / huge coordinates of a degenerate transform stay harmless
    const int cx = __float2int_rz(fminf(fmaxf(fxs, -8.0f), (float)W + 8.0f));
    const int cy = __float2int_rz(fminf(fmaxf(fys, -8.0f), (float)H + 8.0f));
    const bool corner = lane < 4;
    int minx = __reduce_min_sync(0xffffffffu, corner ? cx : INT_MAX);
    int miny = __reduce_min_sync(0xffffffffu, corner ? cy : INT_MAX);
    int maxx = __reduce_max_sync(0xffffffffu, corner ? cx : INT_MIN);
    int maxy = __reduce_max_sync(0xffffffffu, corner ? cy : INT_MIN);
    // +-1 px of slack for float rounding (mesh mode: the bilinear offsets are not monotone, the slack is a heuristic);
    // f of an EASU pixel lies in [1, W-5] x [1, H-5] (FSR.cl:387-399)
    minx = max(minx - 1, 1); miny = max(miny - 1, 1);
    maxx = min(maxx + 1, W - 5); maxy = min(maxy + 1, H - 5);
    TilePlan pl;
    pl.x0 = minx - 1; pl.y0 = miny - 1;  // taps reach f-1 .. f+2
    pl.bw = (minx <= maxx && miny <= maxy) ? min(maxx + 2 - pl.x0 + 1, STG_W) : 0;
    pl.bh = min(maxy + 2 - pl.y0 + 1, STG_H);
    pl.off = bulk ? (int)((3u * (unsigned)pl.x0) & 15u) : 0;
    pl.row_bytes = (pl.off + 3 * pl.bw + 15) & ~15;
    if (lane == 0) plan[b] = pl;
    if (bulk && pl.bw > 0)
    {
        if (lane == 0) mbar_arrive_expect_tx(&full[b], (uint32_t)(pl.row_bytes * pl.bh));
        __syncwarp();
        if (lane < pl.bh)
            bulk_g2s(raw + b * RAW_BYTES + lane * RAW_PITCH,
                     src + (size_t)(pl.y0 + lane) * src_pitch + 3 * pl.x0 - pl.off, (uint32_t)pl.row_bytes, &full[b]);
    }
}

template <int MODE, bool YUV, int OCC>
__global__ void __launch_bounds__(TILE_W* TILE_H, OCC)
    k_easu_remap_pipelined(const uint8_t* __restrict__ src, size_t src_pitch, uint8_t* __restrict__ dst,
                           size_t dst_pitch, int W, int H, Transform T, MeshArgs M, uchar3 bg, int tiles_x, int n_tiles,
                           int bulk)
{
    extern __shared__ __align__(128) unsigned char smem[];
    float4* const tile = reinterpret_cast<float4*>(smem);   // {c0, c1, c2, luma} / 255 of the staged source pixels
    float4* const terms = tile + STG_H * STG_W;             // easu_direction_terms of the same pixels (interior only)
    float* const luma = reinterpret_cast<float*>(terms + STG_H * STG_W);  // tile[].w again, contiguous (conflict-free)
    uint8_t* const raw = reinterpret_cast<uint8_t*>(luma + STG_H * STG_W);  // 2 raw (byte) windows, bulk-copy targets
    __shared__ TilePlan plan[2];
    __shared__ __align__(8) uint64_t full[2];

    const int tx = threadIdx.x & (TILE_W - 1);
    const int ty = threadIdx.x / TILE_W;
    const bool warp0 = threadIdx.x < 32;

    if (threadIdx.x == 0)
    {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    int t = blockIdx.x;
    if (warp0 && t < n_tiles) plan_and_fetch<MODE>(t, 0, tiles_x, W, H, T, M, src, src_pitch, bulk != 0, plan, raw, full);
    __syncthreads();

    uint32_t phase0 = 0, phase1 = 0;  // mbarrier phase parity of full[0] / full[1]
    for (int it = 0; t < n_tiles; t += gridDim.x, it++)
    {
        const int b = it & 1;
        const int tn = t + (int)gridDim.x;
        // raw[b^1] was last read by the conversion of iteration it-1, which every thread left through a barrier
        if (warp0 && tn < n_tiles)
            plan_and_fetch<MODE>(tn, b ^ 1, tiles_x, W, H, T, M, src, src_pitch, bulk != 0, plan, raw, full);

        const TilePlan pl = plan[b];
        const int x0 = pl.x0, y0 = pl.y0, bw = pl.bw, bh = pl.bh;
        uint8_t* const rawb = raw + b * RAW_BYTES;

        if (bw > 0)
        {
            if (bulk)
            {
                mbar_wait(&full[b], b ? phase1 : phase0);
                if (b) phase1 ^= 1u; else phase0 ^= 1u;
            }
            else
            {
                // source rows are not 16-byte aligned: plain byte copy, one warp per row (no prefetch on this path)
                for (int r = ty; r < bh; r += TILE_H)
                {
                    const uint8_t* g = src + (size_t)(y0 + r) * src_pitch + 3 * x0;
                    for (int i = tx; i < 3 * bw; i += TILE_W) rawb[r * RAW_PITCH + i] = __ldg(g + i);
                }
                __syncthreads();
            }
            // ---- raw bytes -> float4 texels.  thread (tx, ty): columns tx, tx+32 of rows ty, ty+8, ty+16, ty+24
            static_assert(STG_W <= 2 * TILE_W && STG_H <= 4 * TILE_H, "staging pattern covers the window");
            const float norm = 0.00392156862f;
#pragma unroll
            for (int rr = 0; rr < 4; rr++)
            {
                const int r = ty + TILE_H * rr;
                if (r < bh)
                {
#pragma unroll
                    for (int cc = 0; cc < 2; cc++)
                    {
                        const int c = tx + TILE_W * cc;
                        if (c < bw)
                        {
                            const uint8_t* q = rawb + r * RAW_PITCH + pl.off + 3 * c;
                            float4 v;
                            v.x = u8_to_float(q[0]) * norm;
                            v.y = u8_to_float(q[1]) * norm;
                            v.z = u8_to_float(q[2]) * norm;
                            // FSR.cl:229-241 (the #ifndef is inverted relative to its comments; reproduced as written)
                            v.w = YUV ? __fmaf_rn(v.z, 0.5f, __fmaf_rn(v.x, 0.5f, v.y)) : v.x;
                            tile[r * STG_W + c] = v;
                            luma[r * STG_W + c] = v.w;
                        }
                    }
                }
            }
            __syncthreads();
            // ---- direction terms of the pixels that can be a corner f/g/j/k: columns 1 .. bw-2, rows 1 .. bh-2
#pragma unroll
            for (int rr = 0; rr < 4; rr++)
            {
                const int r = 1 + ty + TILE_H * rr;
                if (r < bh - 1)
                {
#pragma unroll
                    for (int cc = 0; cc < 2; cc++)
                    {
                        const int c = 1 + tx + TILE_W * cc;
                        if (c < bw - 1)
                        {
                            const float* l = &luma[r * STG_W + c];
                            terms[r * STG_W + c] = easu_direction_terms(l[-STG_W], l[-1], l[0], l[1], l[STG_W]);
                        }
                    }
                }
            }
            __syncthreads();
        }

        // ---- the tile's pixels: thread (tx, ty) owns (x, yA) and (x, yA + 8)
        const int tyi = t / tiles_x, txi = t - tyi * tiles_x;
        const int x = txi * TILE_W + tx;
        const int yA = tyi * CTA_H + ty, yB = yA + PAIR_DY;
        const bool insideA = (x < W) && (yA < H), insideB = (x < W) && (yB < H);
        float sxA, syA, sxB, syB;
        source_position<MODE>(x, yA, W, H, T, M, sxA, syA);
        source_position<MODE>(x, yB, W, H, T, M, sxB, syB);
        const PixelClass A = classify(sxA, syA, W, H, insideA), B = classify(sxB, syB, W, H, insideB);
        // all 12 taps (f-1 .. f+2) and the 4 corner terms inside the staged window?
        const bool fitA = A.do_easu && bw > 0 && A.sx - 1 >= x0 && A.sx + 2 < x0 + bw && A.sy - 1 >= y0 && A.sy + 2 < y0 + bh;
        const bool fitB = B.do_easu && bw > 0 && B.sx - 1 >= x0 && B.sx + 2 < x0 + bw && B.sy - 1 >= y0 && B.sy + 2 < y0 + bh;

        uchar3 outA = bg, outB = bg;
        if (fitA || fitB)
        {
            // a lane whose pixel is not a staged EASU pixel computes on its partner's taps and discards the result
            const int iA = fitA ? (A.sy - y0) * STG_W + (A.sx - x0) : (B.sy - y0) * STG_W + (B.sx - x0);
            const int iB = fitB ? (B.sy - y0) * STG_W + (B.sx - x0) : iA;
            uchar3 eA, eB;
            easu_pair<STG_W>(&tile[iA], &tile[iB], &terms[iA], &terms[iB], pk(A.ppx, B.ppx), pk(A.ppy, B.ppy), eA, eB);
            if (fitA) outA = eA;
            if (fitB) outB = eB;
        }
        // outside the planned window (extreme warps): direct global reads, one pixel at a time
        // rare pixels, behind one warp-wide test: outside the planned window -> direct global reads; border band
        // (FSR.cl:387-399) -> nearest neighbour
        if (__any_sync(0xffffffffu, (A.do_easu && !fitA) || (B.do_easu && !fitB) || (A.border && A.in_src) ||
                                        (B.border && B.in_src)))
        {
            if (A.do_easu && !fitA)
                outA = easu_global<YUV>(src + (size_t)A.sy * src_pitch + 3 * A.sx, src_pitch, A.ppx, A.ppy);
            if (B.do_easu && !fitB)
                outB = easu_global<YUV>(src + (size_t)B.sy * src_pitch + 3 * B.sx, src_pitch, B.ppx, B.ppy);
            if (A.border && A.in_src)
            {
                const uint8_t* p = src + (size_t)A.sy * src_pitch + 3 * A.sx;
                outA.x = __ldg(p); outA.y = __ldg(p + 1); outA.z = __ldg(p + 2);
            }
            if (B.border && B.in_src)
            {
                const uint8_t* p = src + (size_t)B.sy * src_pitch + 3 * B.sx;
                outB.x = __ldg(p); outB.y = __ldg(p + 1); outB.z = __ldg(p + 2);
            }
        }
        if (insideA)
        {
            uint8_t* q = dst + (size_t)yA * dst_pitch + 3 * x;
            q[0] = outA.x; q[1] = outA.y; q[2] = outA.z;
        }
        if (insideB)
        {
            uint8_t* q = dst + (size_t)yB * dst_pitch + 3 * x;
            q[0] = outB.x; q[1] = outB.y; q[2] = outB.z;
        }
        __syncthreads();  // tile / terms / plan[b] are free for the next iteration
    }
}

}  // namespace

static int remap_occupancy()
{
    static const int occ = [] {
        const char* e = getenv("LVKB200_REMAP_OCC");  // tuning knob: resident CTAs per SM the kernel is compiled for
        const int v = e ? atoi(e) : 4;
        return (v == 2 || v == 3 || v == 5) ? v : 4;
    }();
    return occ;
}

static int remap_kernel_version()
{
    static const int v = [] {
        // tuning knob: 2 = one CTA per tile (default: measured 49 us at 1080p), 3 = persistent CTAs with TMA bulk
        // prefetch (53 us: the kernel is issue / shared-memory bound, not load-latency bound, so the prefetch buys
        // nothing and the extra barrier per tile costs)
        const char* e = getenv("LVKB200_REMAP_KERNEL");
        return (e && atoi(e) == 3) ? 3 : 2;
    }();
    return v;
}

template <int MODE, bool YUV, int OCC>
static void launch_easu_occ(cudaStream_t cs, const RemapParams& p, const Transform& T, const MeshArgs& M, uchar3 bg)
{
    const int dw = MODE == 2 ? p.dst_width : p.width, dh = MODE == 2 ? p.dst_height : p.height;
    const int tiles_x = div_up(dw, TILE_W), tiles_y = div_up(dh, CTA_H);
    if (MODE == 2 || remap_kernel_version() == 2)
    {
        k_easu_remap<MODE, YUV, OCC><<<dim3(tiles_x, tiles_y), TILE_W * TILE_H, 0, cs>>>(
            p.src, p.src_pitch, p.dst, p.dst_pitch, p.width, p.height, dw, dh, T, M, bg);
        return;
    }
    if constexpr (MODE != 2)
    {
    static const int ctas = [] {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(k_easu_remap_pipelined<MODE, YUV, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, V3_SMEM);
        return sms * OCC;
    }();
    const int n_tiles = tiles_x * tiles_y;
    // TMA bulk copies need 16-byte aligned global addresses: the frame ring always is, arbitrary caller buffers may not
    const int bulk = ((reinterpret_cast<uintptr_t>(p.src) & 15u) == 0 && (p.src_pitch & 15u) == 0) ? 1 : 0;
    k_easu_remap_pipelined<MODE, YUV, OCC><<<min(ctas, n_tiles), TILE_W * TILE_H, V3_SMEM, cs>>>(
        p.src, p.src_pitch, p.dst, p.dst_pitch, p.width, p.height, T, M, bg, tiles_x, n_tiles, bulk);
    }
}

template <int MODE, bool YUV>
static void launch_easu(cudaStream_t cs, const RemapParams& p, const Transform& T, const MeshArgs& M, uchar3 bg)
{
    switch (remap_occupancy())
    {
    case 2: launch_easu_occ<MODE, YUV, 2>(cs, p, T, M, bg); break;
    case 3: launch_easu_occ<MODE, YUV, 3>(cs, p, T, M, bg); break;
    case 5: launch_easu_occ<MODE, YUV, 5>(cs, p, T, M, bg); break;
    default: launch_easu_occ<MODE, YUV, 4>(cs, p, T, M, bg);
    }
}

cudaError_t launch_remap_homography_exact(cudaStream_t cs, const RemapParams& p, const float t[9])
{
    const Transform T{t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7], t[8]};
    const uchar3 bg = make_uchar3(p.bg[0], p.bg[1], p.bg[2]);
    if (p.yuv)
        launch_easu<0, true>(cs, p, T, MeshArgs{}, bg);
    else
        launch_easu<0, false>(cs, p, T, MeshArgs{}, bg);
    count_launches(1);
    return cudaGetLastError();
}

cudaError_t launch_upscale_exact(cudaStream_t cs, const RemapParams& p)
{
    // Image.cpp:191-194: rscale = (float)src / (float)dst per axis
    Transform T{};
    T.r1x = static_cast<float>(p.width) / static_cast<float>(p.dst_width);
    T.r2x = static_cast<float>(p.height) / static_cast<float>(p.dst_height);
    const uchar3 bg = make_uchar3(0, 0, 0);  // never used: every source position lies inside the source
    if (p.yuv)
        launch_easu<2, true>(cs, p, T, MeshArgs{}, bg);
    else
        launch_easu<2, false>(cs, p, T, MeshArgs{}, bg);
    count_launches(1);
    return cudaGetLastError();
}

cudaError_t launch_remap_mesh_exact(cudaStream_t cs, const RemapParams& p, const float* mesh, int mesh_cols, int mesh_rows)
{
    const uchar3 bg = make_uchar3(p.bg[0], p.bg[1], p.bg[2]);
    // cv::resize: scale = 1 / (dsize / ssize), in double
    const double sx = 1.0 / ((double)p.width / (double)mesh_cols), sy = 1.0 / ((double)p.height / (double)mesh_rows);
    const MeshArgs M{reinterpret_cast<const float2*>(mesh), mesh_cols, mesh_rows, sx, sy};
    if (p.yuv)
        launch_easu<1, true>(cs, p, Transform{}, M, bg);
    else
        launch_easu<1, false>(cs, p, Transform{}, M, bg);
    count_launches(1);
    return cudaGetLastError();
}

}  // namespace lvkb200
