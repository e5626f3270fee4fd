// K8 — FSR-EASU warp/remap for sm_100a.
//
// Replaces lvk::remap (LiveVisionKit/Functions/Image.cpp:28-151) and its OpenCL kernels
// easu_remap / easu_remap_homography / easu (Functions/OpenCL/Sources/FSR.cl:98-318,362-452).
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

#include "common.hpp"

namespace lvkb200
{
namespace
{

constexpr int TILE_W = 32;
constexpr int TILE_H = 8;
constexpr int SRC_W = 48;  // staged source tile capacity (pixels)
constexpr int SRC_H = 20;

struct Transform
{
    float r1x, r1y, r1z, r2x, r2y, r2z, r3x, r3y, r3z;
};

__device__ __forceinline__ float aprx_lo_rsq(float a) { return __uint_as_float(0x5f347d74u - (__float_as_uint(a) >> 1)); }
__device__ __forceinline__ float aprx_lo_rcp(float a) { return __uint_as_float(0x7ef07ebbu - __float_as_uint(a)); }
__device__ __forceinline__ float sat01(float x) { return fmaxf(0.0f, fminf(1.0f, x)); }

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
    t.x = (float)__ldg(p) * norm;
    t.y = (float)__ldg(p + 1) * norm;
    t.z = (float)__ldg(p + 2) * norm;
    // FSR.cl:229-241 (the #ifndef is inverted relative to its comments; reproduced as written)
    t.w = YUV ? __fmaf_rn(t.z, 0.5f, __fmaf_rn(t.x, 0.5f, t.y)) : t.x;
    return t;
}

struct SmemTap
{
    const float4* base;  // &tile[(sy - y0) * SRC_W + (sx - x0)]
    __device__ __forceinline__ float4 operator()(int dx, int dy) const { return base[dy * SRC_W + dx]; }
};

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

// MODE 0: homography (FSR.cl:407-452).  MODE 1: mesh offsets, bilinear upsample fused (WarpMesh.cpp:190-191 + FSR.cl:362-403).
template <int MODE, bool YUV>
__global__ void __launch_bounds__(TILE_W* TILE_H)
    k_easu_remap(const uint8_t* __restrict__ src, size_t src_pitch, uint8_t* __restrict__ dst, size_t dst_pitch, int W,
                 int H, Transform T, const float2* __restrict__ mesh, int mesh_cols, int mesh_rows, double mesh_sx,
                 double mesh_sy, uchar3 bg)
{
    __shared__ float4 tile[SRC_H * SRC_W];
    __shared__ int bbox[4];  // minx, miny, maxx, maxy over the EASU pixels of this tile

    const int tx = threadIdx.x & (TILE_W - 1);
    const int ty = threadIdx.x / TILE_W;
    const int x = blockIdx.x * TILE_W + tx;
    const int y = blockIdx.y * TILE_H + ty;
    const bool inside = (x < W) && (y < H);

    if (threadIdx.x == 0)
    {
        bbox[0] = INT_MAX; bbox[1] = INT_MAX; bbox[2] = INT_MIN; bbox[3] = INT_MIN;
    }

    // ---- source coordinate of this destination pixel
    float subx, suby;
    {
        const float fx = (float)x, fy = (float)y;
        float offx, offy;
        if (MODE == 0)
        {
            const float dz = 1.0f / (T.r3x * fx + T.r3y * fy + T.r3z);
            offx = (T.r1x * fx + T.r1y * fy + T.r1z) * dz - fx;
            offy = (T.r2x * fx + T.r2y * fy + T.r2z) * dz - fy;
        }
        else
        {
            // cv::resize(mesh -> WxH, INTER_LINEAR) on CV_32FC2, then cv::multiply by (W, H).
            float mx = (float)(((double)x + 0.5) * mesh_sx - 0.5);
            float my = (float)(((double)y + 0.5) * mesh_sy - 0.5);
            int cx = (int)floorf(mx), cy = (int)floorf(my);
            mx -= (float)cx;
            my -= (float)cy;
            if (cx < 0) { mx = 0.0f; cx = 0; }
            if (cx >= mesh_cols - 1) { mx = 0.0f; cx = mesh_cols - 1; }
            if (cy < 0) { my = 0.0f; cy = 0; }
            if (cy >= mesh_rows - 1) { my = 0.0f; cy = mesh_rows - 1; }
            const int cx1 = min(cx + 1, mesh_cols - 1), cy1 = min(cy + 1, mesh_rows - 1);
            const float2 m00 = __ldg(&mesh[cy * mesh_cols + cx]), m01 = __ldg(&mesh[cy * mesh_cols + cx1]);
            const float2 m10 = __ldg(&mesh[cy1 * mesh_cols + cx]), m11 = __ldg(&mesh[cy1 * mesh_cols + cx1]);
            const float ax0 = 1.0f - mx, ax1 = mx, ay0 = 1.0f - my, ay1 = my;
            const float h0x = m00.x * ax0 + m01.x * ax1, h0y = m00.y * ax0 + m01.y * ax1;
            const float h1x = m10.x * ax0 + m11.x * ax1, h1y = m10.y * ax0 + m11.y * ax1;
            offx = (h0x * ay0 + h1x * ay1) * (float)W;
            offy = (h0y * ay0 + h1y * ay1) * (float)H;
        }
        subx = fx + offx;
        suby = fy + offy;
    }
    const int sx = __float2int_rz(subx);  // convert_int2_rtz
    const int sy = __float2int_rz(suby);
    const float ppx = subx - floorf(subx);
    const float ppy = suby - floorf(suby);

    // ---- classify (FSR.cl:387-399)
    const bool border = (sx < 1) || (sy < 1) || (sx >= W - 4) || (sy >= H - 4);
    const bool in_src = (sx >= 0) && (sx < W) && (sy >= 0) && (sy < H);
    const bool do_easu = inside && !border;

    __syncthreads();
    {
        // warp-level reduce, one shared atomic per warp
        const int lminx = do_easu ? sx : INT_MAX, lminy = do_easu ? sy : INT_MAX;
        const int lmaxx = do_easu ? sx : INT_MIN, lmaxy = do_easu ? sy : INT_MIN;
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
    const bool staged = any_easu && bw <= SRC_W && bh <= SRC_H;

    if (staged)
    {
        // thread (tx, ty) stages columns tx and tx+32 of rows ty, ty+8, ty+16 (SRC_W <= 64, SRC_H <= 24): no integer
        // division, lanes read consecutive pixels of one row
        static_assert(SRC_W <= 2 * TILE_W && SRC_H <= 3 * TILE_H, "staging pattern covers the tile");
        const uint8_t* p0 = src + (size_t)(y0 + ty) * src_pitch + 3 * (x0 + tx);
#pragma unroll
        for (int rr = 0; rr < 3; rr++)
        {
            const int r = ty + TILE_H * rr;
            if (r < bh)
            {
                const uint8_t* p = p0 + (size_t)(TILE_H * rr) * src_pitch;
                if (tx < bw) tile[r * SRC_W + tx] = load_texel<YUV>(p);
                if (tx + TILE_W < bw) tile[r * SRC_W + tx + TILE_W] = load_texel<YUV>(p + 3 * TILE_W);
            }
        }
    }
    __syncthreads();

    if (!inside) return;

    uchar3 out = bg;
    if (border)
    {
        if (in_src)
        {
            const uint8_t* p = src + (size_t)sy * src_pitch + 3 * sx;
            out.x = __ldg(p); out.y = __ldg(p + 1); out.z = __ldg(p + 2);
        }
    }
    else if (staged)
    {
        SmemTap tap{&tile[(sy - y0) * SRC_W + (sx - x0)]};
        out = easu(tap, ppx, ppy);
    }
    else
    {
        GlobalTap<YUV> tap{src + (size_t)sy * src_pitch + 3 * sx, src_pitch};
        out = easu(tap, ppx, ppy);
    }

    uint8_t* q = dst + (size_t)y * dst_pitch + 3 * x;
    q[0] = out.x; q[1] = out.y; q[2] = out.z;
}

}  // namespace

cudaError_t launch_remap_homography(cudaStream_t cs, const RemapParams& p, const float t[9])
{
    const dim3 grid(div_up(p.width, TILE_W), div_up(p.height, TILE_H));
    const Transform T{t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7], t[8]};
    const uchar3 bg = make_uchar3(p.bg[0], p.bg[1], p.bg[2]);
    if (p.yuv)
        k_easu_remap<0, true><<<grid, TILE_W * TILE_H, 0, cs>>>(p.src, p.src_pitch, p.dst, p.dst_pitch, p.width,
                                                               p.height, T, nullptr, 0, 0, 0.0, 0.0, bg);
    else
        k_easu_remap<0, false><<<grid, TILE_W * TILE_H, 0, cs>>>(p.src, p.src_pitch, p.dst, p.dst_pitch, p.width,
                                                                p.height, T, nullptr, 0, 0, 0.0, 0.0, bg);
    count_launches(1);
    return cudaGetLastError();
}

cudaError_t launch_remap_mesh(cudaStream_t cs, const RemapParams& p, const float* mesh, int mesh_cols, int mesh_rows)
{
    const dim3 grid(div_up(p.width, TILE_W), div_up(p.height, TILE_H));
    const Transform T{};
    const uchar3 bg = make_uchar3(p.bg[0], p.bg[1], p.bg[2]);
    // cv::resize: scale = 1 / (dsize / ssize), in double
    const double sx = 1.0 / ((double)p.width / (double)mesh_cols), sy = 1.0 / ((double)p.height / (double)mesh_rows);
    const float2* m = reinterpret_cast<const float2*>(mesh);
    if (p.yuv)
        k_easu_remap<1, true><<<grid, TILE_W * TILE_H, 0, cs>>>(p.src, p.src_pitch, p.dst, p.dst_pitch, p.width,
                                                               p.height, T, m, mesh_cols, mesh_rows, sx, sy, bg);
    else
        k_easu_remap<1, false><<<grid, TILE_W * TILE_H, 0, cs>>>(p.src, p.src_pitch, p.dst, p.dst_pitch, p.width,
                                                                p.height, T, m, mesh_cols, mesh_rows, sx, sy, bg);
    count_launches(1);
    return cudaGetLastError();
}

}  // namespace lvkb200
