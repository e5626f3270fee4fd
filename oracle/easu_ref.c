/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into, imported by or called from the
 * product path (livevisionkit_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use it, and only as the checker / the CPU baseline.
 *
 * Scalar float32 CPU restatement of LiveVisionKit's FSR-EASU remap kernels
 *   LiveVisionKit/Functions/OpenCL/Sources/FSR.cl:55-73   (APrxLo* bit-trick approximations)
 *   LiveVisionKit/Functions/OpenCL/Sources/FSR.cl:98-126  (easu_tap)
 *   LiveVisionKit/Functions/OpenCL/Sources/FSR.cl:131-176 (easu_accumulate)
 *   LiveVisionKit/Functions/OpenCL/Sources/FSR.cl:181-318 (easu)
 *   LiveVisionKit/Functions/OpenCL/Sources/FSR.cl:362-403 (easu_remap, per-pixel offset map)
 *   LiveVisionKit/Functions/OpenCL/Sources/FSR.cl:407-452 (easu_remap_homography)
 *   LiveVisionKit/Functions/OpenCL/Sources/FSR.cl:326-358 (easu_scale), :460-535 (rcas) — lvk::upscale / lvk::sharpen
 * and of the host side that launches them
 *   LiveVisionKit/Functions/Image.cpp:28-81, 85-151.
 *
 * Semantics chosen where OpenCL leaves latitude: IEEE float32; multiply-add contraction — which an OpenCL compiler is
 * free to apply (FP_CONTRACT is ON by default in OpenCL C) — is made EXPLICIT by ONE rule: a product whose only use
 * is one addition/subtraction is fused into it (FMA(a,b,c) below = fmaf); when both operands of the addition are such
 * products the LEFT one is fused (a*b + c*d -> fma(a, b, c*d)); nothing else is contracted (-ffp-contract=off).
 * The rule applies to the whole kernel text: easu_tap, easu_accumulate, easu AND the position arithmetic of the
 * __kernel wrappers (FSR.cl:423-427).  native_recip(x) := 1.0f/x; convert_*_rtz / convert_uchar := C truncation.
 * It is the arithmetic the exact build of the CUDA kernel implements with __fmaf_rn (0 LSB between the two).
 * Parity status: PINNED against the reference's own kernel source compiled for the CPU (oracle/_ref/, built by
 * oracle/ref_build/build_ref.sh from FSR.cl where it lies, strict = no contraction and contract = gcc's
 * -ffp-contract=fast): tests/test_fsr_ref_cpu.py and the fixtures tests/golden/fsr_ref_golden.npz, which those
 * libraries generated.  Two legal builds of the reference differ from EACH OTHER by up to 2 LSB on ~1.3e-4 of the
 * bytes; this restatement lies inside that spread (<= 1 LSB from either, on ~1e-5 of the bytes).
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off -pthread -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <pthread.h>
#include <unistd.h>

/* Row-parallel driver (no OpenMP runtime in this image): splits [0, rows) over `threads` pthreads. */
typedef void (*row_fn)(int y0, int y1, void* ctx);
typedef struct { row_fn fn; void* ctx; int y0, y1; } row_job;
static void* row_trampoline(void* p) { row_job* j = (row_job*)p; j->fn(j->y0, j->y1, j->ctx); return 0; }
static void parallel_rows(row_fn fn, void* ctx, int rows, int threads)
{
    if (threads <= 0) threads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (threads > 256) threads = 256;
    if (threads > rows) threads = rows;
    if (threads <= 1) { fn(0, rows, ctx); return; }
    pthread_t tid[256];
    row_job jobs[256];
    for (int t = 0; t < threads; t++)
    {
        jobs[t].fn = fn; jobs[t].ctx = ctx;
        jobs[t].y0 = (int)((long long)rows * t / threads);
        jobs[t].y1 = (int)((long long)rows * (t + 1) / threads);
        pthread_create(&tid[t], 0, row_trampoline, &jobs[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(tid[t], 0);
}

#define FMA(a, b, c) fmaf((a), (b), (c))

static inline float as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

/* FSR.cl:60,65 */
static inline float aprx_lo_rsq(float a) { return as_float(0x5f347d74u - (as_uint(a) >> 1)); }
static inline float aprx_lo_rcp(float a) { return as_float(0x7ef07ebbu - as_uint(a)); }

static inline float fmax_cl(float a, float b) { return (a < b) ? b : a; } /* OpenCL max(): y if x<y */
static inline float fmin_cl(float a, float b) { return (b < a) ? b : a; } /* OpenCL min(): y if y<x */
static inline float saturate(float x) { return fmaxf(0.0f, fminf(1.0f, x)); } /* FSR.cl:79 */

/* FSR.cl:98-126 */
static inline void easu_tap(float aC[3], float* aW, float offx, float offy, float dirx, float diry,
                            float lenx, float leny, float lob, float clp, const float c[3])
{
    float vx = FMA(offx, dirx, offy * diry);
    float vy = FMA(offx, -diry, offy * dirx);
    vx *= lenx;
    vy *= leny;
    float d2 = fmin_cl(FMA(vx, vx, vy * vy), clp);
    float wA = FMA(lob, d2, -1.0f);
    float wB = FMA(2.0f / 5.0f, d2, -1.0f);
    wA *= wA;
    wB = FMA(25.0f / 16.0f, wB * wB, -(25.0f / 16.0f - 1.0f));
    float w = wB * wA;
    aC[0] = FMA(c[0], w, aC[0]);
    aC[1] = FMA(c[1], w, aC[1]);
    aC[2] = FMA(c[2], w, aC[2]);
    *aW += w;
}

/* FSR.cl:131-176. corner: 0=S 1=T 2=U 3=V */
static inline void easu_accumulate(float dir[2], float* len, float ppx, float ppy, int corner,
                                   float lA, float lB, float lC, float lD, float lE)
{
    float w = 0.0f;
    if (corner == 3) w = ppx * ppy;
    if (corner == 2) w = (1.0f - ppx) * ppy;
    if (corner == 1) w = ppx * (1.0f - ppy);
    if (corner == 0) w = (1.0f - ppx) * (1.0f - ppy);

    float dc = lD - lC;
    float cb = lC - lB;
    float lenX = aprx_lo_rcp(fmax_cl(fabsf(dc), fabsf(cb)));
    float dirX = lD - lB;
    dir[0] = FMA(dirX, w, dir[0]);
    lenX = saturate(fabsf(dirX) * lenX);
    lenX *= lenX;
    *len = FMA(lenX, w, *len);

    float ec = lE - lC;
    float ca = lC - lA;
    float lenY = aprx_lo_rcp(fmax_cl(fabsf(ec), fabsf(ca)));
    float dirY = lE - lA;
    dir[1] = FMA(dirY, w, dir[1]);
    lenY = saturate(fabsf(dirY) * lenY);
    lenY *= lenY;
    *len = FMA(lenY, w, *len);
}

/* FSR.cl:181-318.  src points at pixel (0,0); step in bytes; (sx,sy) = 'f'. yuv selects the
 * YUV_INPUT build (Image.cpp:100-113 defines it iff src.format == YUV).  Note the #ifndef at
 * FSR.cl:229 is inverted relative to its comments: WITHOUT YUV_INPUT luma = channel 0, WITH
 * YUV_INPUT luma = 0.5*ch2 + (0.5*ch0 + ch1).  Reproduced as written. */
static inline void easu(const uint8_t* src, int step, int sx, int sy, float ppx, float ppy, int yuv,
                        uint8_t out[3])
{
    const float norm = 0.00392156862f;
    const uint8_t* r0 = src + (size_t)(sy - 1) * step + 3 * sx;
    const uint8_t* r1 = r0 + step - 3;
    const uint8_t* r2 = r1 + step;
    const uint8_t* r3 = r0 + 3 * (size_t)step;

    float b[3], c[3], e[3], f[3], g[3], h[3], i[3], j[3], k[3], l[3], n[3], o[3];
    for (int ch = 0; ch < 3; ch++)
    {
        b[ch] = (float)r0[ch] * norm;     c[ch] = (float)r0[3 + ch] * norm;
        e[ch] = (float)r1[ch] * norm;     f[ch] = (float)r1[3 + ch] * norm;
        g[ch] = (float)r1[6 + ch] * norm; h[ch] = (float)r1[9 + ch] * norm;
        i[ch] = (float)r2[ch] * norm;     j[ch] = (float)r2[3 + ch] * norm;
        k[ch] = (float)r2[6 + ch] * norm; l[ch] = (float)r2[9 + ch] * norm;
        n[ch] = (float)r3[ch] * norm;     o[ch] = (float)r3[3 + ch] * norm;
    }

#define LUMA(p) (yuv ? FMA((p)[2], 0.5f, FMA((p)[0], 0.5f, (p)[1])) : (p)[0])
    const float bL = LUMA(b), cL = LUMA(c), eL = LUMA(e), fL = LUMA(f), gL = LUMA(g), hL = LUMA(h);
    const float iL = LUMA(i), jL = LUMA(j), kL = LUMA(k), lL = LUMA(l), nL = LUMA(n), oL = LUMA(o);
#undef LUMA

    float len = 0.0f, dir[2] = {0.0f, 0.0f};
    easu_accumulate(dir, &len, ppx, ppy, 0, bL, eL, fL, gL, jL);
    easu_accumulate(dir, &len, ppx, ppy, 1, cL, fL, gL, hL, kL);
    easu_accumulate(dir, &len, ppx, ppy, 2, fL, iL, jL, kL, nL);
    easu_accumulate(dir, &len, ppx, ppy, 3, gL, jL, kL, lL, oL);

    float dirR = FMA(dir[0], dir[0], dir[1] * dir[1]);
    int zro = dirR < (1.0f / 32768.0f);
    dirR = aprx_lo_rsq(dirR);
    dirR = zro ? 1.0f : dirR;
    dir[0] = zro ? 1.0f : dir[0];
    dir[0] *= dirR;
    dir[1] *= dirR;

    len = len * 0.5f;
    len *= len;

    float stretch = FMA(dir[0], dir[0], dir[1] * dir[1]) * aprx_lo_rcp(fmax_cl(fabsf(dir[0]), fabsf(dir[1])));
    float len2x = FMA(stretch - 1.0f, len, 1.0f);
    float len2y = FMA(-0.5f, len, 1.0f);
    float lob = FMA((1.0f / 4.0f - 0.04f) - 0.5f, len, 0.5f);
    float clp = aprx_lo_rcp(lob);

    float mi4[3], ma4[3];
    for (int ch = 0; ch < 3; ch++)
    {
        mi4[ch] = fmin_cl(f[ch], fmin_cl(g[ch], fmin_cl(j[ch], k[ch])));
        ma4[ch] = fmax_cl(f[ch], fmax_cl(g[ch], fmax_cl(j[ch], k[ch])));
    }

    float aC[3] = {0.0f, 0.0f, 0.0f}, aW = 0.0f;
    easu_tap(aC, &aW, 0.0f - ppx, -1.0f - ppy, dir[0], dir[1], len2x, len2y, lob, clp, b);
    easu_tap(aC, &aW, 1.0f - ppx, -1.0f - ppy, dir[0], dir[1], len2x, len2y, lob, clp, c);
    easu_tap(aC, &aW, -1.0f - ppx, 1.0f - ppy, dir[0], dir[1], len2x, len2y, lob, clp, i);
    easu_tap(aC, &aW, 0.0f - ppx, 1.0f - ppy, dir[0], dir[1], len2x, len2y, lob, clp, j);
    easu_tap(aC, &aW, 0.0f - ppx, 0.0f - ppy, dir[0], dir[1], len2x, len2y, lob, clp, f);
    easu_tap(aC, &aW, -1.0f - ppx, 0.0f - ppy, dir[0], dir[1], len2x, len2y, lob, clp, e);
    easu_tap(aC, &aW, 1.0f - ppx, 1.0f - ppy, dir[0], dir[1], len2x, len2y, lob, clp, k);
    easu_tap(aC, &aW, 2.0f - ppx, 1.0f - ppy, dir[0], dir[1], len2x, len2y, lob, clp, l);
    easu_tap(aC, &aW, 2.0f - ppx, 0.0f - ppy, dir[0], dir[1], len2x, len2y, lob, clp, h);
    easu_tap(aC, &aW, 1.0f - ppx, 0.0f - ppy, dir[0], dir[1], len2x, len2y, lob, clp, g);
    easu_tap(aC, &aW, 0.0f - ppx, 2.0f - ppy, dir[0], dir[1], len2x, len2y, lob, clp, n);
    easu_tap(aC, &aW, 1.0f - ppx, 2.0f - ppy, dir[0], dir[1], len2x, len2y, lob, clp, o);

    float rcpW = 1.0f / aW; /* native_recip */
    for (int ch = 0; ch < 3; ch++)
    {
        float v = fmin_cl(ma4[ch], fmax_cl(mi4[ch], aC[ch] * rcpW));
        out[ch] = (uint8_t)(int)(v * 255.0f); /* convert_uchar3: truncation */
    }
}

/* Shared tail of both remap kernels, FSR.cl:383-402 / 432-451. */
static inline void remap_pixel(const uint8_t* src, int src_step, int src_rows, int src_cols, uint8_t* dst_px,
                               float subx, float suby, const uint8_t bg[3], int yuv)
{
    int sx = (int)subx; /* convert_int2_rtz */
    int sy = (int)suby;
    subx -= floorf(subx);
    suby -= floorf(suby);

    if (sx < 1 || sy < 1 || sx >= src_cols - 4 || sy >= src_rows - 4)
    {
        if (sx >= 0 && sx < src_cols && sy >= 0 && sy < src_rows)
        {
            const uint8_t* p = src + (size_t)sy * src_step + 3 * sx;
            dst_px[0] = p[0]; dst_px[1] = p[1]; dst_px[2] = p[2];
        }
        else
        {
            dst_px[0] = bg[0]; dst_px[1] = bg[1]; dst_px[2] = bg[2];
        }
        return;
    }
    easu(src, src_step, sx, sy, subx, suby, yuv, dst_px);
}

/* FSR.cl:407-452 + Image.cpp:85-151.  t = already-inverted (dst->src) homography, row-major,
 * given in double and narrowed to float exactly as cv::Vec4f(t.at<double>(..)) does. */
typedef struct
{
    const uint8_t* src; int src_step, rows, cols; uint8_t* dst; int dst_step;
    float r[9]; const uint8_t* bg; int yuv;
} homog_ctx;

static void homog_rows(int y0, int y1, void* p)
{
    const homog_ctx* c = (const homog_ctx*)p;
    const float r1x = c->r[0], r1y = c->r[1], r1z = c->r[2];
    const float r2x = c->r[3], r2y = c->r[4], r2z = c->r[5];
    const float r3x = c->r[6], r3y = c->r[7], r3z = c->r[8];
    for (int y = y0; y < y1; y++)
    {
        for (int x = 0; x < c->cols; x++)
        {
            float fx = (float)x, fy = (float)y;
            /* FSR.cl:423-427, contracted by the rule in the header: (a*x + b*y) + c -> fma(a, x, b*y) + c, n*dz - f -> fma */
            float dz = 1.0f / (FMA(r3x, fx, r3y * fy) + r3z);
            float offx = FMA(FMA(r1x, fx, r1y * fy) + r1z, dz, -fx);
            float offy = FMA(FMA(r2x, fx, r2y * fy) + r2z, dz, -fy);
            float subx = (float)x + offx;
            float suby = (float)y + offy;
            remap_pixel(c->src, c->src_step, c->rows, c->cols, c->dst + (size_t)y * c->dst_step + 3 * x, subx, suby,
                        c->bg, c->yuv);
        }
    }
}

void oracle_easu_remap_homography(const uint8_t* src, int src_step, int rows, int cols, uint8_t* dst, int dst_step,
                                  const double t[9], const uint8_t bg[3], int yuv, int threads)
{
    homog_ctx c = {src, src_step, rows, cols, dst, dst_step, {0}, bg, yuv};
    for (int k = 0; k < 9; k++) c.r[k] = (float)t[k];
    parallel_rows(homog_rows, &c, rows, threads);
}

/* FSR.cl:362-403 + Image.cpp:28-81.  map = CV_32FC2 per-pixel offsets (pixels), map_step in bytes.
 * dst has the size of the map (map_rows x map_cols). */
typedef struct
{
    const uint8_t* src; int src_step, src_rows, src_cols; uint8_t* dst; int dst_step;
    const float* map; int map_step, map_cols; const uint8_t* bg; int yuv;
} map_ctx;

static void map_rows_fn(int y0, int y1, void* p)
{
    const map_ctx* c = (const map_ctx*)p;
    for (int y = y0; y < y1; y++)
    {
        const float* mrow = (const float*)((const uint8_t*)c->map + (size_t)y * c->map_step);
        for (int x = 0; x < c->map_cols; x++)
        {
            float subx = (float)x + mrow[2 * x];
            float suby = (float)y + mrow[2 * x + 1];
            remap_pixel(c->src, c->src_step, c->src_rows, c->src_cols, c->dst + (size_t)y * c->dst_step + 3 * x,
                        subx, suby, c->bg, c->yuv);
        }
    }
}

void oracle_easu_remap_map(const uint8_t* src, int src_step, int src_rows, int src_cols, uint8_t* dst, int dst_step,
                           const float* map, int map_step, int map_rows, int map_cols, const uint8_t bg[3], int yuv,
                           int threads)
{
    map_ctx c = {src, src_step, src_rows, src_cols, dst, dst_step, map, map_step, map_cols, bg, yuv};
    parallel_rows(map_rows_fn, &c, map_rows, threads);
}

/* ---- lvk::upscale: FSR.cl:326-358 (easu_scale) + Image.cpp:155-201 ------------------------------------------------
 * rscale = {(float)src.cols / (float)dst.cols, (float)src.rows / (float)dst.rows} (Image.cpp:191-194).
 * Source pixels with src_coord.x == 0 || src_coord.y == 0 || src_coord.x >= src_cols-4 || src_coord.y >= src_rows-4
 * are copied (nearest), everything else runs EASU.  size == src.size() is a plain copy (Image.cpp:162-166). */
typedef struct
{
    const uint8_t* src; int src_step, src_rows, src_cols; uint8_t* dst; int dst_step, dst_cols;
    float rsx, rsy; int yuv;
} scale_ctx;

static void scale_rows_fn(int y0, int y1, void* p)
{
    const scale_ctx* c = (const scale_ctx*)p;
    for (int y = y0; y < y1; y++)
    {
        for (int x = 0; x < c->dst_cols; x++)
        {
            float subx = (float)x * c->rsx, suby = (float)y * c->rsy;
            int sx = (int)subx, sy = (int)suby; /* convert_int2_rtz */
            subx -= floorf(subx);
            suby -= floorf(suby);
            uint8_t* q = c->dst + (size_t)y * c->dst_step + 3 * x;
            if (sx == 0 || sy == 0 || sx >= c->src_cols - 4 || sy >= c->src_rows - 4)
            {
                const uint8_t* s = c->src + (size_t)sy * c->src_step + 3 * sx;
                q[0] = s[0]; q[1] = s[1]; q[2] = s[2];
                continue;
            }
            easu(c->src, c->src_step, sx, sy, subx, suby, c->yuv, q);
        }
    }
}

void oracle_easu_scale(const uint8_t* src, int src_step, int src_rows, int src_cols, uint8_t* dst, int dst_step,
                       int dst_rows, int dst_cols, int yuv, int threads)
{
    if (dst_rows == src_rows && dst_cols == src_cols)
    {
        for (int y = 0; y < src_rows; y++) memcpy(dst + (size_t)y * dst_step, src + (size_t)y * src_step, 3 * (size_t)src_cols);
        return;
    }
    scale_ctx c = {src, src_step, src_rows, src_cols, dst, dst_step, dst_cols,
                   (float)src_cols / (float)dst_cols, (float)src_rows / (float)dst_rows, yuv};
    parallel_rows(scale_rows_fn, &c, dst_rows, threads);
}

/* ---- lvk::sharpen: FSR.cl:460-535 (rcas) + Image.cpp:205-233 --------------------------------------------------------
 * `sharpness` is the KERNEL argument exp2(-2*(1-s)) (Image.cpp:227; computed by the caller in float).
 * Latitude taken (the reference leaves these undefined):
 *   - ScalingFilter runs it IN PLACE (sharpen(output, output), ScalingFilter.cpp:57), a data race between work-groups
 *     in the reference; here src and dst are distinct images (every tap reads the unsharpened input);
 *   - the border test copies `coord.x <= cols || coord.y <= rows` (an always-true OR that also writes the padding
 *     threads' pixels out of bounds); here exactly the in-image border pixels are copied;
 *   - min()/max() on NaN (0 * inf when a ring is all 0 or all 1): the non-NaN operand wins (fminf/fmaxf, what GPU
 *     min/max instructions do);
 *   - convert_uchar3 of an out-of-range value: truncation then saturation to [0, 255]. */
static inline float aprx_med_rcp(float a) /* FSR.cl:70 */
{
    float b = as_float(0x7ef19fffu - as_uint(a));
    return b * FMA(-b, a, 2.0f);
}

typedef struct
{
    const uint8_t* src; int src_step, rows, cols; uint8_t* dst; int dst_step; float sharp;
} rcas_ctx;

static void rcas_rows_fn(int y0, int y1, void* p)
{
    const rcas_ctx* c = (const rcas_ctx*)p;
    const float norm = 0.00392156862f;
    for (int y = y0; y < y1; y++)
    {
        for (int x = 0; x < c->cols; x++)
        {
            const uint8_t* s = c->src + (size_t)y * c->src_step + 3 * x;
            uint8_t* q = c->dst + (size_t)y * c->dst_step + 3 * x;
            if (x == 0 || x >= c->cols - 1 || y == 0 || y >= c->rows - 1)
            {
                q[0] = s[0]; q[1] = s[1]; q[2] = s[2];
                continue;
            }
            float lobe_c[3], sum[3], e[3];
            for (int ch = 0; ch < 3; ch++)
            {
                const float b = (float)s[ch - c->src_step] * norm, h = (float)s[ch + c->src_step] * norm;
                const float d = (float)s[ch - 3] * norm, f = (float)s[ch + 3] * norm;
                e[ch] = (float)s[ch] * norm;
                const float mn4 = fminf(b, fminf(d, fminf(f, h)));
                const float mx4 = fmaxf(b, fmaxf(d, fmaxf(f, h)));
                const float hit_min = fminf(mn4, e[ch]) * (1.0f / (4.0f * mx4));
                const float hit_max = (1.0f - fmaxf(mx4, e[ch])) * (1.0f / FMA(4.0f, mn4, -4.0f));
                lobe_c[ch] = fmaxf(-hit_min, hit_max);
                sum[ch] = ((b + d) + h) + f;
            }
            /* lobeR = ch2, lobeG = ch1, lobeB = ch0: max(lobeR, max(lobeG, lobeB)) */
            float lobe = fmaxf(lobe_c[2], fmaxf(lobe_c[1], lobe_c[0]));
            lobe = fminf(fmaxf(lobe, -0.1875f), 0.0f) * c->sharp; /* clamp(x, lo, hi) = min(max(x, lo), hi) */
            const float rcpL = aprx_med_rcp(FMA(4.0f, lobe, 1.0f));
            for (int ch = 0; ch < 3; ch++)
            {
                const float v = FMA(sum[ch], lobe, e[ch]) * rcpL;
                int iv = (int)(v * 255.0f);
                q[ch] = (uint8_t)(iv < 0 ? 0 : (iv > 255 ? 255 : iv));
            }
        }
    }
}

/* Image.cpp:227: std::exp2(-2.0f * (1.0f - sharpness)), the float overload = the platform's exp2f (NumPy's exp2 differs
 * from glibc's in the last bit for 20% of the settings, so the oracle calls libm like the reference does). */
float oracle_rcas_kernel_sharpness(float sharpness) { return exp2f(-2.0f * (1.0f - sharpness)); }

void oracle_rcas(const uint8_t* src, int src_step, int rows, int cols, uint8_t* dst, int dst_step, float kernel_sharpness,
                 int threads)
{
    rcas_ctx c = {src, src_step, rows, cols, dst, dst_step, kernel_sharpness};
    parallel_rows(rcas_rows_fn, &c, rows, threads);
}

int oracle_max_threads(void) { return (int)sysconf(_SC_NPROCESSORS_ONLN); }
