// K1 + K4 — image pyramid, Scharr derivatives and sparse pyramidal Lucas-Kanade for sm_100a.
//
// Replaces cv::SparsePyrLKOpticalFlow::calc as configured and called by FrameTracker
// (LiveVisionKit/Vision/FrameTracker.cpp:33-35,41-48,140-146): window 11x11, 3 pyramid levels above the base,
// TermCriteria(COUNT+EPS, 5, 0.01), flags 0, minEigThreshold 1e-4, no error output.
// OpenCV's CPU algorithm (upstream video/lkpyramid.cpp, imgproc pyrDown — not under /root/reference) restated:
//   * pyramid: pyrDown = separable [1 4 6 4 1], (sum+128)>>8, BORDER_REFLECT_101, size (n+1)/2; every level is
//     stored with an 11-px REFLECT_101 border so window reads never need clamping;
//   * Scharr derivatives of every level as int16 (dx, dy) pairs, zero border;
//   * per point, coarse to fine: 14-bit fixed-point bilinear patch of the previous image (5 fractional bits) and of
//     its derivatives, 2x2 gradient matrix, min-eigenvalue test, <= 5 Newton iterations with the same stopping rules.
// Window sums are accumulated EXACTLY in integers (OpenCV accumulates the same integer products in float32 SIMD
// lanes), converted to float once: positions agree to ~1e-4 px, status identically (tests/test_lk_gpu.py).
// One warp tracks one feature through all levels in a single launch: 121 window pixels = 4 per lane, patch kept in
// registers, warp-shuffle reductions for A11/A12/A22/b1/b2.  Compiled with --fmad=false (the CPU path has no FMA).

#include <cooperative_groups.h>

#include <cstdlib>
#include <cstring>

#include "common.hpp"
#include "lk.hpp"

namespace cg = cooperative_groups;

namespace lvkb200
{
namespace
{

constexpr int P = LK_PAD;  // 11
constexpr int WIN = 11;

__device__ __forceinline__ int reflect101(int p, int n)
{
    // cv::borderInterpolate(BORDER_REFLECT_101) for |overshoot| < n
    if (n == 1) return 0;
    while (p < 0 || p >= n)
    {
        if (p < 0) p = -p;
        else p = 2 * (n - 1) - p;
    }
    return p;
}

// the same for any overshoot, without a loop (n >= 2)
__device__ __forceinline__ int reflect101_closed(int p, int n)
{
    const int period = 2 * (n - 1);
    int m = p % period;
    if (m < 0) m += period;
    return (m < n) ? m : period - m;
}

// Level 0: copy the detection image into the padded layout.
__global__ void k_pyr_pad0(const uint8_t* __restrict__ src, size_t src_pitch, int w, int h, uint8_t* __restrict__ dst,
                           size_t dst_pitch)
{
    const int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y * blockDim.y + threadIdx.y;
    if (px >= w + 2 * P || py >= h + 2 * P) return;
    const int x = reflect101(px - P, w), y = reflect101(py - P, h);
    dst[(size_t)py * dst_pitch + px] = __ldg(src + (size_t)y * src_pitch + x);
}

// Level l (padded) from level l-1 (padded): pyrDown evaluated at every padded pixel's reflected coordinate.
__global__ void k_pyr_down(const uint8_t* __restrict__ prev, size_t prev_pitch, uint8_t* __restrict__ dst,
                           size_t dst_pitch, int w, int h)
{
    const int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y * blockDim.y + threadIdx.y;
    if (px >= w + 2 * P || py >= h + 2 * P) return;
    const int x = reflect101(px - P, w), y = reflect101(py - P, h);
    const uint8_t* base = prev + (size_t)(2 * y - 2 + P) * prev_pitch + (2 * x - 2 + P);
    int sum = 0;
#pragma unroll
    for (int j = 0; j < 5; j++)
    {
        const uint8_t* r = base + (size_t)j * prev_pitch;
        const int row = (int)r[0] + 4 * (int)r[1] + 6 * (int)r[2] + 4 * (int)r[3] + (int)r[4];
        const int kj = (j == 0 || j == 4) ? 1 : ((j == 2) ? 6 : 4);
        sum += kj * row;
    }
    dst[(size_t)py * dst_pitch + px] = (uint8_t)((sum + 128) >> 8);
}

struct ScharrArg
{
    const uint8_t* img[LK_MAX_LEVELS];
    short2* deriv[LK_MAX_LEVELS];
    size_t img_pitch[LK_MAX_LEVELS];
    size_t deriv_pitch[LK_MAX_LEVELS];  // in short2 elements
    int w[LK_MAX_LEVELS], h[LK_MAX_LEVELS];
};

// calcScharrDeriv for every level in one launch (blockIdx.z = level); zero border.
__global__ void k_scharr(ScharrArg a)
{
    const int l = blockIdx.z;
    const int w = a.w[l], h = a.h[l];
    const int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y * blockDim.y + threadIdx.y;
    if (px >= w + 2 * P || py >= h + 2 * P) return;
    short2 out = make_short2(0, 0);
    const int x = px - P, y = py - P;
    if (x >= 0 && x < w && y >= 0 && y < h)
    {
        const uint8_t* c = a.img[l] + (size_t)py * a.img_pitch[l] + px;
        const size_t s = a.img_pitch[l];
        const int u0 = c[-(ptrdiff_t)s - 1], u1 = c[-(ptrdiff_t)s], u2 = c[-(ptrdiff_t)s + 1];
        const int m0 = c[-1], m2 = c[1];
        const int d0 = c[s - 1], d1 = c[s], d2 = c[s + 1];
        const int t0l = (u0 + d0) * 3 + m0 * 10, t0r = (u2 + d2) * 3 + m2 * 10;
        const int t1l = d0 - u0, t1c = d1 - u1, t1r = d2 - u2;
        out.x = (short)(t0r - t0l);
        out.y = (short)((t1r + t1l) * 3 + t1c * 10);
    }
    a.deriv[l][(size_t)py * a.deriv_pitch[l] + px] = out;
}

// Whole pyramid + derivative planes of one detection image in ONE cooperative launch (grid-wide barriers between
// the dependent levels) instead of levels+2 tiny dependent launches: the chain is latency-, not bandwidth-bound.
struct PyrArg
{
    const uint8_t* det;
    size_t det_pitch;
    uint8_t* img[LK_MAX_LEVELS];
    short2* deriv[LK_MAX_LEVELS];
    size_t img_pitch[LK_MAX_LEVELS];
    size_t deriv_pitch[LK_MAX_LEVELS];
    int w[LK_MAX_LEVELS], h[LK_MAX_LEVELS];
    int levels;
};

__global__ void __launch_bounds__(256) k_pyramid_fused(PyrArg a)
{
    cg::grid_group grid = cg::this_grid();
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    {
        const int pw = a.w[0] + 2 * P, ph = a.h[0] + 2 * P;
        for (int i = tid; i < pw * ph; i += nth)
        {
            const int py = i / pw, px = i - py * pw;
            const int x = reflect101(px - P, a.w[0]), y = reflect101(py - P, a.h[0]);
            a.img[0][(size_t)py * a.img_pitch[0] + px] = __ldg(a.det + (size_t)y * a.det_pitch + x);
        }
    }
    grid.sync();
    for (int l = 1; l < a.levels; l++)
    {
        const int w = a.w[l], h = a.h[l], pw = w + 2 * P, ph = h + 2 * P;
        const uint8_t* prev = a.img[l - 1];
        const size_t pp = a.img_pitch[l - 1];
        for (int i = tid; i < pw * ph; i += nth)
        {
            const int py = i / pw, px = i - py * pw;
            const int x = reflect101(px - P, w), y = reflect101(py - P, h);
            const uint8_t* base = prev + (size_t)(2 * y - 2 + P) * pp + (2 * x - 2 + P);
            int sum = 0;
#pragma unroll
            for (int j = 0; j < 5; j++)
            {
                const uint8_t* r = base + (size_t)j * pp;
                const int row = (int)r[0] + 4 * (int)r[1] + 6 * (int)r[2] + 4 * (int)r[3] + (int)r[4];
                const int kj = (j == 0 || j == 4) ? 1 : ((j == 2) ? 6 : 4);
                sum += kj * row;
            }
            a.img[l][(size_t)py * a.img_pitch[l] + px] = (uint8_t)((sum + 128) >> 8);
        }
        grid.sync();
    }
    for (int l = 0; l < a.levels; l++)
    {
        const int w = a.w[l], h = a.h[l], pw = w + 2 * P, ph = h + 2 * P;
        const size_t s = a.img_pitch[l];
        for (int i = tid; i < pw * ph; i += nth)
        {
            const int py = i / pw, px = i - py * pw;
            short2 out = make_short2(0, 0);
            const int x = px - P, y = py - P;
            if (x >= 0 && x < w && y >= 0 && y < h)
            {
                const uint8_t* c = a.img[l] + (size_t)py * s + px;
                const int u0 = c[-(ptrdiff_t)s - 1], u1 = c[-(ptrdiff_t)s], u2 = c[-(ptrdiff_t)s + 1];
                const int m0 = c[-1], m2 = c[1];
                const int d0 = c[s - 1], d1 = c[s], d2 = c[s + 1];
                const int t0l = (u0 + d0) * 3 + m0 * 10, t0r = (u2 + d2) * 3 + m2 * 10;
                const int t1l = d0 - u0, t1c = d1 - u1, t1r = d2 - u2;
                out.x = (short)(t0r - t0l);
                out.y = (short)((t1r + t1l) * 3 + t1c * 10);
            }
            a.deriv[l][(size_t)py * a.deriv_pitch[l] + px] = out;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// The same pyramid WITHOUT grid-wide barriers: one CTA owns a 32x32 tile of level 0 and the matching 16x16 / 8x8 / 4x4
// tiles of levels 1..3, and recomputes the halo it needs in shared memory (a 69x69 patch of level 0 feeds a 33x33
// patch of level 1, a 15x15 patch of level 2 and a 6x6 patch of level 3).  The levels are tiny (130 K pixels at
// 480x270), so the cooperative version above spends its time in four grid.sync()s and, needing all its CTAs
// co-resident, cannot start while another stream's kernel (the previous output's remap) fills the machine.
//
// Patch entry (cx, cy) of level l holds the PADDED image value P_l(c) = L_l(reflect101(c)), evaluated at the reflected
// coordinate exactly like k_pyr_down, so image-border pixels see the same neighbours as in the padded global layout.
// Each interior pixel is written by the CTA that owns it, together with its mirror images in the 11-pixel border
// (single reflection: every level is wider than the window); the derivative planes' zero border is written once, when
// the planes are allocated.
constexpr int PT_TILE = 32;
constexpr int PT_S3 = 6, PT_S2 = 2 * PT_S3 + 3, PT_S1 = 2 * PT_S2 + 3, PT_S0 = 2 * PT_S1 + 3;  // 6, 15, 33, 69
static_assert(LK_MAX_LEVELS == 4, "patch sizes are laid out for four levels");

// writes value v of interior pixel (x, y) to the padded plane and to its mirror images inside the border
__device__ __forceinline__ void store_mirrored(uint8_t* img, size_t pitch, int w, int h, int x, int y, uint8_t v)
{
    int xs[3], ys[3], nx = 1, ny = 1;
    xs[0] = x; ys[0] = y;
    if (x >= 1 && x <= P) xs[nx++] = -x;
    if (x >= w - 1 - P && x <= w - 2) xs[nx++] = 2 * (w - 1) - x;
    if (y >= 1 && y <= P) ys[ny++] = -y;
    if (y >= h - 1 - P && y <= h - 2) ys[ny++] = 2 * (h - 1) - y;
    for (int j = 0; j < ny; j++)
        for (int i = 0; i < nx; i++) img[(size_t)(ys[j] + P) * pitch + (xs[i] + P)] = v;
}

// pyrDown of the finer patch `f` (size fs, origin fo) evaluated at coarse coordinate (rx, ry)
__device__ __forceinline__ int pyr_down_at(const uint8_t* f, int fs, int fox, int foy, int rx, int ry)
{
    int bx = 2 * rx - 2 - fox, by = 2 * ry - 2 - foy;
    bx = max(0, min(fs - 5, bx));  // only entries nobody reads can be clamped (see above)
    by = max(0, min(fs - 5, by));
    const uint8_t* base = f + by * fs + bx;
    int sum = 0;
#pragma unroll
    for (int j = 0; j < 5; j++)
    {
        const uint8_t* r = base + j * fs;
        const int row = (int)r[0] + 4 * (int)r[1] + 6 * (int)r[2] + 4 * (int)r[3] + (int)r[4];
        const int kj = (j == 0 || j == 4) ? 1 : ((j == 2) ? 6 : 4);
        sum += kj * row;
    }
    return (sum + 128) >> 8;
}

// interior pixels of this CTA's tile of one level: padded image (+ mirrors) and Scharr derivative from the patch
template <int TILE>
__device__ __forceinline__ void emit_level(const PyrArg& a, int l, const uint8_t* patch, int ps, int ox, int oy, int tx0,
                                           int ty0)
{
    const int w = a.w[l], h = a.h[l];
    for (int i = threadIdx.x; i < TILE * TILE; i += blockDim.x)
    {
        const int ly = i / TILE, lx = i - ly * TILE;
        const int x = tx0 + lx, y = ty0 + ly;
        if (x >= w || y >= h) continue;
        const uint8_t* c = patch + (y - oy) * ps + (x - ox);
        store_mirrored(a.img[l], a.img_pitch[l], w, h, x, y, c[0]);
        const int u0 = c[-ps - 1], u1 = c[-ps], u2 = c[-ps + 1];
        const int m0 = c[-1], m2 = c[1];
        const int d0 = c[ps - 1], d1 = c[ps], d2 = c[ps + 1];
        const int t0l = (u0 + d0) * 3 + m0 * 10, t0r = (u2 + d2) * 3 + m2 * 10;
        const int t1l = d0 - u0, t1c = d1 - u1, t1r = d2 - u2;
        a.deriv[l][(size_t)(y + P) * a.deriv_pitch[l] + (x + P)] =
            make_short2((short)(t0r - t0l), (short)((t1r + t1l) * 3 + t1c * 10));
    }
}

__global__ void __launch_bounds__(256) k_pyramid_tiles(PyrArg a)
{
    __shared__ uint8_t p0[PT_S0 * PT_S0], p1[PT_S1 * PT_S1], p2[PT_S2 * PT_S2], p3[PT_S3 * PT_S3];
    // patch origins (level coordinates): the level-3 patch is the 4x4 tile plus the Scharr halo, each finer patch the
    // 5-tap footprint of the coarser one
    const int o3x = (PT_TILE >> 3) * blockIdx.x - 1, o3y = (PT_TILE >> 3) * blockIdx.y - 1;
    const int o2x = 2 * o3x - 2, o2y = 2 * o3y - 2;
    const int o1x = 2 * o2x - 2, o1y = 2 * o2y - 2;
    const int o0x = 2 * o1x - 2, o0y = 2 * o1y - 2;

    // reflected coordinates of every patch column and row, once per CTA (an integer modulo per ELEMENT was a third of
    // this kernel's instructions): rx[l][i] = reflect101(o_l.x + i), same for ry
    __shared__ short rx0[PT_S0], ry0[PT_S0], rx1[PT_S1], ry1[PT_S1], rx2[PT_S2], ry2[PT_S2], rx3[PT_S3], ry3[PT_S3];
    {
        const int t = threadIdx.x;
        if (t < PT_S0) { rx0[t] = (short)reflect101_closed(o0x + t, a.w[0]); ry0[t] = (short)reflect101_closed(o0y + t, a.h[0]); }
        else if (t >= 96 && t < 96 + PT_S1 && a.levels > 1)
        { rx1[t - 96] = (short)reflect101_closed(o1x + t - 96, a.w[1]); ry1[t - 96] = (short)reflect101_closed(o1y + t - 96, a.h[1]); }
        else if (t >= 160 && t < 160 + PT_S2 && a.levels > 2)
        { rx2[t - 160] = (short)reflect101_closed(o2x + t - 160, a.w[2]); ry2[t - 160] = (short)reflect101_closed(o2y + t - 160, a.h[2]); }
        else if (t >= 192 && t < 192 + PT_S3 && a.levels > 3)
        { rx3[t - 192] = (short)reflect101_closed(o3x + t - 192, a.w[3]); ry3[t - 192] = (short)reflect101_closed(o3y + t - 192, a.h[3]); }
    }
    __syncthreads();
    {
        // all of a thread's loads are issued before the first store (19 independent global loads in flight per thread:
        // the patch load is the only part of this kernel that waits for L2)
        constexpr int PER = (PT_S0 * PT_S0 + 255) / 256;
        uint8_t v[PER];
#pragma unroll
        for (int k = 0; k < PER; k++)
        {
            const int i = min(threadIdx.x + 256 * k, PT_S0 * PT_S0 - 1);
            const int py = i / PT_S0, px = i - py * PT_S0;
            v[k] = __ldg(a.det + (size_t)ry0[py] * a.det_pitch + rx0[px]);
        }
#pragma unroll
        for (int k = 0; k < PER; k++)
        {
            const int i = threadIdx.x + 256 * k;
            if (i < PT_S0 * PT_S0) p0[i] = v[k];
        }
    }
    __syncthreads();
    emit_level<PT_TILE>(a, 0, p0, PT_S0, o0x, o0y, PT_TILE * blockIdx.x, PT_TILE * blockIdx.y);
    if (a.levels < 2) return;
    for (int i = threadIdx.x; i < PT_S1 * PT_S1; i += blockDim.x)
    {
        const int py = i / PT_S1, px = i - py * PT_S1;
        p1[i] = (uint8_t)pyr_down_at(p0, PT_S0, o0x, o0y, rx1[px], ry1[py]);
    }
    __syncthreads();
    emit_level<(PT_TILE >> 1)>(a, 1, p1, PT_S1, o1x, o1y, (PT_TILE >> 1) * blockIdx.x, (PT_TILE >> 1) * blockIdx.y);
    if (a.levels < 3) return;
    if (threadIdx.x < PT_S2 * PT_S2)
    {
        const int py = threadIdx.x / PT_S2, px = threadIdx.x - py * PT_S2;
        p2[threadIdx.x] = (uint8_t)pyr_down_at(p1, PT_S1, o1x, o1y, rx2[px], ry2[py]);
    }
    __syncthreads();
    emit_level<(PT_TILE >> 2)>(a, 2, p2, PT_S2, o2x, o2y, (PT_TILE >> 2) * blockIdx.x, (PT_TILE >> 2) * blockIdx.y);
    if (a.levels < 4) return;
    if (threadIdx.x < PT_S3 * PT_S3)
    {
        const int py = threadIdx.x / PT_S3, px = threadIdx.x - py * PT_S3;
        p3[threadIdx.x] = (uint8_t)pyr_down_at(p2, PT_S2, o2x, o2y, rx3[px], ry3[py]);
    }
    __syncthreads();
    emit_level<(PT_TILE >> 3)>(a, 3, p3, PT_S3, o3x, o3y, (PT_TILE >> 3) * blockIdx.x, (PT_TILE >> 3) * blockIdx.y);
}

struct LkArg
{
    const uint8_t* prev[LK_MAX_LEVELS];
    const uint8_t* next[LK_MAX_LEVELS];
    const short2* deriv[LK_MAX_LEVELS];
    size_t img_pitch[LK_MAX_LEVELS];
    size_t deriv_pitch[LK_MAX_LEVELS];
    int w[LK_MAX_LEVELS], h[LK_MAX_LEVELS];
    int max_level;
};

__device__ __forceinline__ int descale(int v, int n) { return (v + (1 << (n - 1))) >> n; }

// One warp per point, four points per CTA.  The frame's parameters and points normally travel AS KERNEL ARGUMENTS
// (`pack`, up to LK_INLINE_POINTS points = 20.8 KB of the 32 KB parameter space): they ride the launch itself, so the
// chain needs neither a copy-engine transfer (which would queue behind the neighbouring frames' multi-megabyte
// uploads) nor SM reads of host memory over PCIe (measured: ~1 us per dependent read, serialised).  With
// pack.inline_points == 0 they are read from io.pts_in / io.prm_in (device memory) instead.  Device copies of the
// inputs are left for the kernels that follow; io.next_host / io.status_host (optional, mapped pinned host memory)
// receive a second copy of the results as posted PCIe writes.
static_assert(sizeof(TrackParams) == 24, "TrackParams is moved as six 32-bit words");
static_assert(sizeof(LkPack) <= 32764 - 512, "kernel parameter space");

__global__ void __launch_bounds__(128)
    k_lk_track(LkArg a, LkIo io, const __grid_constant__ LkPack pack)
{
    __shared__ uint32_t s_prm[6];
    __shared__ float s_pts[8];
    __shared__ float s_out[8];
    __shared__ uint32_t s_status;
    const int base = blockIdx.x * 4;
    if (pack.inline_points)
    {
        if (threadIdx.x < 6) s_prm[threadIdx.x] = reinterpret_cast<const uint32_t*>(&pack.prm)[threadIdx.x];
        else if (threadIdx.x >= 8 && threadIdx.x < 16 && base + 3 < LK_INLINE_POINTS)
            s_pts[threadIdx.x - 8] = reinterpret_cast<const float*>(pack.pts)[2 * base + (threadIdx.x - 8)];
    }
    else
    {
        if (threadIdx.x < 6) s_prm[threadIdx.x] = __ldcv(reinterpret_cast<const uint32_t*>(io.prm_in) + threadIdx.x);
        else if (threadIdx.x >= 8 && threadIdx.x < 16)
            s_pts[threadIdx.x - 8] = __ldcv(reinterpret_cast<const float*>(io.pts_in) + 2 * base + (threadIdx.x - 8));
    }
    if (threadIdx.x == 0) s_status = 0;
    __syncthreads();
    const TrackParams* prm = reinterpret_cast<const TrackParams*>(s_prm);
    const int n = prm->n;
    if (base >= n) return;  // whole CTA
    if (blockIdx.x == 0 && threadIdx.x < 6 && io.prm_copy) reinterpret_cast<uint32_t*>(io.prm_copy)[threadIdx.x] = s_prm[threadIdx.x];
    if (threadIdx.x < 8 && io.pts_copy) reinterpret_cast<float*>(io.pts_copy)[2 * base + threadIdx.x] = s_pts[threadIdx.x];

    const int wid = threadIdx.x >> 5;
    const int pt = base + wid;
    const double epsilon_sq = prm->lk_epsilon_sq;
    const int lane = threadIdx.x & 31;
    const float2 p0 = make_float2(s_pts[2 * wid], s_pts[2 * wid + 1]);
    const int max_level = (pt < n) ? a.max_level : -1;  // warps past the last point only take part in the barriers
    const float half = 5.0f;       // (winSize - 1) * 0.5
    const float FLT_SCALE = 1.0f / (float)(1 << 20);
    float2 out = make_float2(0.f, 0.f);
    bool ok = true;

    // window pixel assignment: idx = lane + 32*q, q < 4, idx < 121
    int wx[4], wy[4];
#pragma unroll
    for (int q = 0; q < 4; q++)
    {
        const int idx = lane + 32 * q;
        wy[q] = idx / WIN;
        wx[q] = idx - wy[q] * WIN;
    }

    for (int level = max_level; level >= 0; level--)
    {
        const int cols = a.w[level], rows = a.h[level];
        const float inv = (float)(1.0 / (double)(1 << level));
        float2 prevPt = make_float2(p0.x * inv, p0.y * inv);
        float2 nextPt = (level == a.max_level) ? prevPt : make_float2(out.x * 2.0f, out.y * 2.0f);
        out = nextPt;

        prevPt.x -= half;
        prevPt.y -= half;
        const int ipx = (int)floorf(prevPt.x), ipy = (int)floorf(prevPt.y);
        if (ipx < -WIN || ipx >= cols || ipy < -WIN || ipy >= rows)
        {
            if (level == 0) ok = false;
            continue;
        }
        float fa = prevPt.x - (float)ipx, fb = prevPt.y - (float)ipy;
        int iw00 = __float2int_rn((1.f - fa) * (1.f - fb) * 16384.f);
        int iw01 = __float2int_rn(fa * (1.f - fb) * 16384.f);
        int iw10 = __float2int_rn((1.f - fa) * fb * 16384.f);
        int iw11 = 16384 - iw00 - iw01 - iw10;

        const size_t ipitch = a.img_pitch[level], dpitch = a.deriv_pitch[level];
        const uint8_t* I = a.prev[level] + (size_t)(ipy + P) * ipitch + (ipx + P);
        const short2* D = a.deriv[level] + (size_t)(ipy + P) * dpitch + (ipx + P);

        int Ival[4], Ix[4], Iy[4];
        int sA11 = 0, sA12 = 0, sA22 = 0;
#pragma unroll
        for (int q = 0; q < 4; q++)
        {
            Ival[q] = 0; Ix[q] = 0; Iy[q] = 0;
            if (lane + 32 * q < WIN * WIN)
            {
                const uint8_t* s = I + (size_t)wy[q] * ipitch + wx[q];
                const short2* d = D + (size_t)wy[q] * dpitch + wx[q];
                Ival[q] = descale((int)s[0] * iw00 + (int)s[1] * iw01 + (int)s[ipitch] * iw10 + (int)s[ipitch + 1] * iw11, 9);
                const short2 d00 = d[0], d01 = d[1], d10 = d[dpitch], d11 = d[dpitch + 1];
                Ix[q] = descale((int)d00.x * iw00 + (int)d01.x * iw01 + (int)d10.x * iw10 + (int)d11.x * iw11, 14);
                Iy[q] = descale((int)d00.y * iw00 + (int)d01.y * iw01 + (int)d10.y * iw10 + (int)d11.y * iw11, 14);
                sA11 += Ix[q] * Ix[q];
                sA12 += Ix[q] * Iy[q];
                sA22 += Iy[q] * Iy[q];
            }
        }
        sA11 = __reduce_add_sync(0xffffffffu, sA11);
        sA12 = __reduce_add_sync(0xffffffffu, sA12);
        sA22 = __reduce_add_sync(0xffffffffu, sA22);

        const float A11 = (float)sA11 * FLT_SCALE, A12 = (float)sA12 * FLT_SCALE, A22 = (float)sA22 * FLT_SCALE;
        float Dt = A11 * A22 - A12 * A12;
        const float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (float)(2 * WIN * WIN);
        if ((double)minEig < 1e-4 || Dt < 1.1920928955078125e-07f)
        {
            if (level == 0) ok = false;
            continue;
        }
        Dt = 1.f / Dt;

        nextPt.x -= half;
        nextPt.y -= half;
        float2 prevDelta = make_float2(0.f, 0.f);
        const uint8_t* Jbase = a.next[level];
        // the lane's window offsets at this level, once: the iteration below is a latency-bound dependent chain and
        // its 64-bit address arithmetic was a tenth of the instructions on it
        int joff[4];
#pragma unroll
        for (int q = 0; q < 4; q++) joff[q] = wy[q] * (int)ipitch + wx[q];

        for (int j = 0; j < 5; j++)
        {
            const int inx = (int)floorf(nextPt.x), iny = (int)floorf(nextPt.y);
            if (inx < -WIN || inx >= cols || iny < -WIN || iny >= rows)
            {
                if (level == 0) ok = false;
                break;
            }
            fa = nextPt.x - (float)inx;
            fb = nextPt.y - (float)iny;
            iw00 = __float2int_rn((1.f - fa) * (1.f - fb) * 16384.f);
            iw01 = __float2int_rn(fa * (1.f - fb) * 16384.f);
            iw10 = __float2int_rn((1.f - fa) * fb * 16384.f);
            iw11 = 16384 - iw00 - iw01 - iw10;

            const uint8_t* J = Jbase + (size_t)(iny + P) * ipitch + (inx + P);
            // per-lane partial sums fit int32 (4 products of |diff| <= 8160 by |I'| <= 4080); the warp total does not,
            // so it is reduced exactly as two REDUX halves (hi = arithmetic >> 16, lo = low 16 bits) instead of a
            // 5-step 64-bit shuffle chain: the iteration is latency-bound and this is its longest dependent chain
            int pb1 = 0, pb2 = 0;
#pragma unroll
            for (int q = 0; q < 4; q++)
            {
                if (lane + 32 * q < WIN * WIN)
                {
                    const uint8_t* s = J + joff[q];
                    const int diff = descale((int)s[0] * iw00 + (int)s[1] * iw01 + (int)s[ipitch] * iw10 +
                                                 (int)s[ipitch + 1] * iw11, 9) - Ival[q];
                    pb1 += diff * Ix[q];
                    pb2 += diff * Iy[q];
                }
            }
            const long long sb1 = ((long long)__reduce_add_sync(0xffffffffu, pb1 >> 16) << 16) +
                                  (long long)__reduce_add_sync(0xffffffffu, pb1 & 0xffff);
            const long long sb2 = ((long long)__reduce_add_sync(0xffffffffu, pb2 >> 16) << 16) +
                                  (long long)__reduce_add_sync(0xffffffffu, pb2 & 0xffff);
            const float b1 = (float)sb1 * FLT_SCALE, b2 = (float)sb2 * FLT_SCALE;
            const float2 delta = make_float2((A12 * b2 - A22 * b1) * Dt, (A12 * b1 - A11 * b2) * Dt);
            nextPt.x += delta.x;
            nextPt.y += delta.y;
            out = make_float2(nextPt.x + half, nextPt.y + half);

            if ((double)delta.x * (double)delta.x + (double)delta.y * (double)delta.y <= epsilon_sq) break;
            if (j > 0 && fabs((double)(delta.x + prevDelta.x)) < 0.01 && fabs((double)(delta.y + prevDelta.y)) < 0.01)
            {
                out.x -= delta.x * 0.5f;
                out.y -= delta.y * 0.5f;
                break;
            }
            prevDelta = delta;
        }
    }
    if (lane == 0)
    {
        s_out[2 * wid] = out.x;
        s_out[2 * wid + 1] = out.y;
        atomicOr(&s_status, (ok ? 1u : 0u) << (8 * wid));
    }
    __syncthreads();
    // 4 points = one 32-byte row of coordinates and one 32-bit word of status bytes per CTA
    if (threadIdx.x < 8)
    {
        reinterpret_cast<float*>(io.next)[2 * base + threadIdx.x] = s_out[threadIdx.x];
        if (io.next_host) reinterpret_cast<float*>(io.next_host)[2 * base + threadIdx.x] = s_out[threadIdx.x];
    }
    else if (threadIdx.x == 8)
    {
        reinterpret_cast<uint32_t*>(io.status)[blockIdx.x] = s_status;
        if (io.status_host) reinterpret_cast<uint32_t*>(io.status_host)[blockIdx.x] = s_status;
    }
}

}  // namespace

lvkb200_status LkPyramid::prepare(int width, int height, cudaStream_t cs)
{
    if (width == w[0] && height == h[0] && levels > 0) return LVKB200_OK;
    // buildOpticalFlowPyramid stops when a level is not larger than the window
    LVKB_REQUIRE(width > WIN && height > WIN);
    int lw = width, lh = height;
    levels = 0;
    for (int l = 0; l <= LK_REF_MAX_LEVEL; l++)
    {
        if (l > 0)
        {
            lw = (lw + 1) / 2;
            lh = (lh + 1) / 2;
            if (lw <= WIN || lh <= WIN) break;
        }
        w[l] = lw;
        h[l] = lh;
        img_pitch[l] = (size_t)((lw + 2 * P + 15) / 16 * 16);
        deriv_pitch[l] = (size_t)((lw + 2 * P + 3) / 4 * 4);
        const size_t img_bytes = img_pitch[l] * (lh + 2 * P + 1) + 16;
        const size_t deriv_bytes = sizeof(short2) * (deriv_pitch[l] * (lh + 2 * P + 1) + 16);
        LVKB_CUDA(img[l].ensure(img_bytes));
        LVKB_CUDA(deriv[l].ensure(deriv_bytes));
        // the derivative planes' border is zero (BORDER_CONSTANT) and only the tile builder's interior is rewritten
        LVKB_CUDA(cudaMemsetAsync(img[l].ptr, 0, img_bytes, cs));
        LVKB_CUDA(cudaMemsetAsync(deriv[l].ptr, 0, deriv_bytes, cs));
        levels++;
    }
    // the memsets are ordered on the stream that builds and reads this pyramid (no legacy-stream work, no device-wide
    // synchronisation: another host thread may be capturing its own stream's graph at this moment)
    valid = false;
    return LVKB200_OK;
}

void LkPyramid::release()
{
    for (int l = 0; l < LK_MAX_LEVELS; l++)
    {
        img[l].release();
        deriv[l].release();
    }
    levels = 0;
    valid = false;
    w[0] = h[0] = 0;
}

lvkb200_status LkPyramid::build(cudaStream_t cs, const uint8_t* det, size_t det_pitch)
{
    // ---- default: one ordinary launch, one CTA per 32x32 tile across all levels (no grid-wide barriers)
    const char* forced = std::getenv("LVKB200_PYRAMID");  // "coop" / "levels": the older builders, kept for A/B tests
    if (!forced || !*forced || std::strcmp(forced, "tiles") == 0)
    {
        PyrArg pa{};
        pa.det = det;
        pa.det_pitch = det_pitch;
        pa.levels = levels;
        for (int l = 0; l < levels; l++)
        {
            pa.img[l] = img[l].as<uint8_t>();
            pa.deriv[l] = deriv[l].as<short2>();
            pa.img_pitch[l] = img_pitch[l];
            pa.deriv_pitch[l] = deriv_pitch[l];
            pa.w[l] = w[l];
            pa.h[l] = h[l];
        }
        k_pyramid_tiles<<<dim3(div_up(w[0], PT_TILE), div_up(h[0], PT_TILE)), 256, 0, cs>>>(pa);
        count_launches(1);
        LVKB_CUDA(cudaGetLastError());
        valid = true;
        return LVKB200_OK;
    }
    const bool allow_coop = std::strcmp(forced, "levels") != 0;

    // ---- one cooperative launch for the whole pyramid
    static int coop_blocks = -1;  // co-resident 256-thread CTAs of k_pyramid_fused on this device (0 = unsupported)
    if (coop_blocks < 0)
    {
        int dev = 0, coop = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pyramid_fused, 256, 0);
        coop_blocks = (coop && per_sm > 0) ? sms * std::min(per_sm, 2) : 0;
        cudaGetLastError();
    }
    if (coop_blocks > 0 && allow_coop)
    {
        PyrArg pa{};
        pa.det = det;
        pa.det_pitch = det_pitch;
        pa.levels = levels;
        for (int l = 0; l < levels; l++)
        {
            pa.img[l] = img[l].as<uint8_t>();
            pa.deriv[l] = deriv[l].as<short2>();
            pa.img_pitch[l] = img_pitch[l];
            pa.deriv_pitch[l] = deriv_pitch[l];
            pa.w[l] = w[l];
            pa.h[l] = h[l];
        }
        void* args[] = {&pa};
        LVKB_CUDA(cudaLaunchCooperativeKernel((const void*)k_pyramid_fused, dim3(coop_blocks), dim3(256), args, 0, cs));
        count_launches(1);
        valid = true;
        return LVKB200_OK;
    }

    // ---- fallback: one launch per level (devices without cooperative launch)
    const dim3 blk(32, 8);
    {
        const dim3 grid(div_up(w[0] + 2 * P, 32), div_up(h[0] + 2 * P, 8));
        k_pyr_pad0<<<grid, blk, 0, cs>>>(det, det_pitch, w[0], h[0], img[0].as<uint8_t>(), img_pitch[0]);
    }
    for (int l = 1; l < levels; l++)
    {
        const dim3 grid(div_up(w[l] + 2 * P, 32), div_up(h[l] + 2 * P, 8));
        k_pyr_down<<<grid, blk, 0, cs>>>(img[l - 1].as<uint8_t>(), img_pitch[l - 1], img[l].as<uint8_t>(),
                                         img_pitch[l], w[l], h[l]);
    }
    ScharrArg sa{};
    for (int l = 0; l < levels; l++)
    {
        sa.img[l] = img[l].as<uint8_t>();
        sa.deriv[l] = deriv[l].as<short2>();
        sa.img_pitch[l] = img_pitch[l];
        sa.deriv_pitch[l] = deriv_pitch[l];
        sa.w[l] = w[l];
        sa.h[l] = h[l];
    }
    const dim3 grid(div_up(w[0] + 2 * P, 32), div_up(h[0] + 2 * P, 8), levels);
    k_scharr<<<grid, blk, 0, cs>>>(sa);
    count_launches(levels + 1);
    LVKB_CUDA(cudaGetLastError());
    valid = true;
    return LVKB200_OK;
}

double lk_epsilon_for_call(int call_index)
{
    // SparsePyrLKOpticalFlowImpl::calc clamps and SQUARES its member TermCriteria::epsilon in place on every call
    // (upstream lkpyramid.cpp), and the reference keeps ONE tracker object for the life of the FrameTracker
    // (m_OpticalTracker, FrameTracker.cpp:41-48), so call n stops on |delta|^2 <= 0.01^(2^(n+1)):
    // 1e-4, 1e-8, 1e-16, ... and 0 from the 9th call on.  Reproduced, in double like the original.
    double eps = 0.01;
    for (int i = 0; i <= call_index && eps != 0.0; i++)
    {
        eps = std::min(std::max(eps, 0.0), 10.0);
        eps *= eps;
    }
    return eps;
}

lvkb200_status lk_track(cudaStream_t cs, const LkPyramid& prev, const LkPyramid& next, int max_points, const LkIo& io,
                        const LkPack* pack)
{
    // The grid covers max_points; CTAs beyond the frame's actual count exit at once.
    const int n = max_points;
    if (n <= 0) return LVKB200_OK;
    LVKB_REQUIRE((n & 3) == 0 && io.next && io.status);
    LVKB_REQUIRE(pack ? (pack->inline_points != 0 && n <= LK_INLINE_POINTS) : (io.pts_in && io.prm_in));
    static const LkPack no_pack{};  // inline_points == 0
    LVKB_REQUIRE(prev.levels == next.levels && prev.levels > 0 && prev.w[0] == next.w[0] && prev.h[0] == next.h[0]);
    LkArg a{};
    for (int l = 0; l < prev.levels; l++)
    {
        a.prev[l] = prev.img[l].as<uint8_t>();
        a.next[l] = next.img[l].as<uint8_t>();
        a.deriv[l] = prev.deriv[l].as<short2>();
        a.img_pitch[l] = prev.img_pitch[l];
        a.deriv_pitch[l] = prev.deriv_pitch[l];
        a.w[l] = prev.w[l];
        a.h[l] = prev.h[l];
    }
    a.max_level = prev.levels - 1;
    k_lk_track<<<div_up(n, 4), 128, 0, cs>>>(a, io, pack ? *pack : no_pack);
    count_launches(1);
    LVKB_CUDA(cudaGetLastError());
    return LVKB200_OK;
}

}  // namespace lvkb200
