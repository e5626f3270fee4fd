// K2 — FAST-9/16 corner score + 3x3 non-max suppression + ordered compaction for sm_100a.
//
// Replaces cv::FastFeatureDetector(threshold, nonmaxSuppression=true, TYPE_9_16)::detect(frame(bounds)) as called
// per detection region by FeatureDetector::detect (LiveVisionKit/Vision/FeatureDetector.cpp:130-134).
// OpenCV's arithmetic (upstream features2d fast.cpp / fast_score.cpp, not under /root/reference) restated:
//   * pixel p is a corner iff >= 9 contiguous pixels of the 16-pixel Bresenham circle (r = 3) are all > p+t or all < p-t;
//   * response = cornerScore<16> = max(t, A, B) - 1, A/B = max over the sixteen 9-arcs of the arc minimum of
//     (p - ring) / (ring - p);
//   * NMS keeps a corner iff its score is strictly greater than the scores of all 8 neighbours (non-corners score 0);
//   * rows/cols < 3 or >= dim-3 OF THE SUB-IMAGE never fire; keypoints are emitted in (y, x) lexicographic order.
// All regions of a frame are processed by one launch per kernel (blockIdx.z / .y = region), each with its own threshold.
//   k_fast_score   : 32x8 pixel tiles staged in shared memory with a 3-px halo; per-thread 16-bit brighter/darker ring
//                    masks, 9-contiguity by shift-and; arc-min/max score only for the few % of pixels that are corners.
//   k_fast_nms_row : one CTA per image row: NMS + warp-ballot ordered compaction into a per-row list.
//   k_fast_gather  : one CTA per region: exclusive scan of the row counts, then an ordered gather -> keypoint list.

#include "common.hpp"
#include "fast.hpp"

namespace lvkb200
{
namespace
{

constexpr int TW = 32, TH = 8, HALO = 3;
constexpr int SW = TW + 2 * HALO, SH = TH + 2 * HALO;

struct RegionsArg
{
    FastRegion r[FAST_MAX_REGIONS];
    int n;
};

// Bresenham circle, OpenCV order (dx, dy): (0,3),(1,3),(2,2),(3,1),(3,0),(3,-1),(2,-2),(1,-3),(0,-3),(-1,-3),(-2,-2),
// (-3,-1),(-3,0),(-3,1),(-2,2),(-1,3)
__device__ __constant__ int c_dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
__device__ __constant__ int c_dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

__device__ __forceinline__ bool has_arc9(unsigned m16)
{
    unsigned mm = m16 | (m16 << 16);
    unsigned r = mm & (mm >> 1);  // runs >= 2
    r &= r >> 2;                  // >= 4
    r &= r >> 4;                  // >= 8
    r &= mm >> 8;                 // >= 9
    return (r & 0xffffu) != 0;
}

__global__ void __launch_bounds__(TW* TH)
    k_fast_score(const uint8_t* __restrict__ img, size_t pitch, RegionsArg regs, uint8_t* __restrict__ score,
                 size_t score_pitch)
{
    const FastRegion rg = regs.r[blockIdx.z];
    const int tiles_x = (rg.w + TW - 1) / TW;
    const int bx = blockIdx.x % tiles_x, by = blockIdx.x / tiles_x;
    if (by * TH >= rg.h) return;

    __shared__ uint8_t tile[SH][SW + 2];
    const int ox = rg.x + bx * TW - HALO, oy = rg.y + by * TH - HALO;
    // stage (clamped to the REGION: pixels outside the sub-image are never used by a firing pixel)
    for (int idx = threadIdx.x; idx < SW * SH; idx += TW * TH)
    {
        const int r = idx / SW, c = idx - r * SW;
        const int gx = min(max(ox + c, rg.x), rg.x + rg.w - 1), gy = min(max(oy + r, rg.y), rg.y + rg.h - 1);
        tile[r][c] = __ldg(img + (size_t)gy * pitch + gx);
    }
    __syncthreads();

    const int tx = threadIdx.x % TW, ty = threadIdx.x / TW;
    const int lx = bx * TW + tx, ly = by * TH + ty;  // region-local
    if (lx >= rg.w || ly >= rg.h) return;

    int result = 0;
    if (lx >= 3 && ly >= 3 && lx < rg.w - 3 && ly < rg.h - 3)
    {
        const int v = tile[ty + HALO][tx + HALO];
        const int t = rg.threshold;
        int d[16];
        unsigned dark = 0, bright = 0;  // dark: ring < v - t  (d > t);  bright: ring > v + t (d < -t)
#pragma unroll
        for (int k = 0; k < 16; k++)
        {
            d[k] = v - (int)tile[ty + HALO + c_dy[k]][tx + HALO + c_dx[k]];
            dark |= (d[k] > t ? 1u : 0u) << k;
            bright |= (d[k] < -t ? 1u : 0u) << k;
        }
        if (has_arc9(dark) || has_arc9(bright))
        {
            // cornerScore<16> in OpenCV's own two-pass form (arcs [k..k+8] and [k+1..k+9] for even k, with its
            // early-outs).  NOTE: the shorter "max over 16 arcs of min/-max" formulation is miscompiled by
            // nvcc 12.9 for sm_100a (integer negation folded into VIMNMX3: tools/dbg/score_variants.cu shows 55 % wrong
            // results on a B200), so keep this form.
            int e[25];
#pragma unroll
            for (int k = 0; k < 25; k++) e[k] = d[k & 15];
#define D(i) e[(i)]
            int a0 = t;
#pragma unroll
            for (int k = 0; k < 16; k += 2)
            {
                int a = min(D(k + 1), D(k + 2));
                a = min(a, D(k + 3));
                if (a <= a0) continue;
                a = min(a, D(k + 4));
                a = min(a, D(k + 5));
                a = min(a, D(k + 6));
                a = min(a, D(k + 7));
                a = min(a, D(k + 8));
                a0 = max(a0, min(a, D(k)));
                a0 = max(a0, min(a, D(k + 9)));
            }
            int b0 = -a0;
#pragma unroll
            for (int k = 0; k < 16; k += 2)
            {
                int b = max(D(k + 1), D(k + 2));
                b = max(b, D(k + 3));
                b = max(b, D(k + 4));
                b = max(b, D(k + 5));
                if (b >= b0) continue;
                b = max(b, D(k + 6));
                b = max(b, D(k + 7));
                b = max(b, D(k + 8));
                b0 = min(b0, max(b, D(k)));
                b0 = min(b0, max(b, D(k + 9)));
            }
#undef D
            result = -b0 - 1;
        }
    }
    score[(size_t)(rg.y + ly) * score_pitch + rg.x + lx] = (uint8_t)result;
}

// One CTA per interior row of a region. Output: row_x[row_base + k] (region-local x), row_s[...] = score.
__global__ void __launch_bounds__(128)
    k_fast_nms_row(const uint8_t* __restrict__ score, size_t score_pitch, RegionsArg regs, int max_rows, int row_cap,
                   uint16_t* __restrict__ row_x, uint8_t* __restrict__ row_s, int* __restrict__ row_count)
{
    const FastRegion rg = regs.r[blockIdx.y];
    const int ly = blockIdx.x;  // region-local row
    if (ly >= rg.h) return;
    const size_t slot = (size_t)blockIdx.y * max_rows + ly;
    __shared__ int warp_cnt[4];
    __shared__ int base;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    if (ly < 3 || ly >= rg.h - 3)
    {
        if (threadIdx.x == 0) row_count[slot] = 0;
        return;
    }
    const uint8_t* r0 = score + (size_t)(rg.y + ly - 1) * score_pitch + rg.x;
    const uint8_t* r1 = r0 + score_pitch;
    const uint8_t* r2 = r1 + score_pitch;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int x0 = 3; x0 < rg.w - 3; x0 += 128)
    {
        const int lx = x0 + threadIdx.x;
        bool keep = false;
        int s = 0;
        if (lx < rg.w - 3)
        {
            s = r1[lx];
            if (s > 0)
                keep = s > r1[lx - 1] && s > r1[lx + 1] && s > r0[lx - 1] && s > r0[lx] && s > r0[lx + 1] &&
                       s > r2[lx - 1] && s > r2[lx] && s > r2[lx + 1];
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_cnt[wid] = __popc(bal);
        __syncthreads();
        int off = base;
        for (int w = 0; w < wid; w++) off += warp_cnt[w];
        if (keep)
        {
            const int k = off + __popc(bal & ((1u << lane) - 1u));
            if (k < row_cap)
            {
                row_x[slot * row_cap + k] = (uint16_t)lx;
                row_s[slot * row_cap + k] = (uint8_t)s;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) base += warp_cnt[0] + warp_cnt[1] + warp_cnt[2] + warp_cnt[3];
        __syncthreads();
    }
    if (threadIdx.x == 0) row_count[slot] = min(base, row_cap);
}

// One CTA per region: scan row counts, gather rows in order.
__global__ void __launch_bounds__(256)
    k_fast_gather(RegionsArg regs, int max_rows, int row_cap, const uint16_t* __restrict__ row_x,
                  const uint8_t* __restrict__ row_s, const int* __restrict__ row_count, int out_cap,
                  FastPoint* __restrict__ out, int* __restrict__ out_count)
{
    const FastRegion rg = regs.r[blockIdx.x];
    extern __shared__ int offs[];  // max_rows + 1
    const size_t rbase = (size_t)blockIdx.x * max_rows;
    // serial-in-chunks exclusive scan (rows <= a few thousand; one warp-synchronous pass is plenty)
    if (threadIdx.x == 0)
    {
        int acc = 0;
        for (int r = 0; r < rg.h; r++)
        {
            offs[r] = acc;
            acc += row_count[rbase + r];
        }
        offs[rg.h] = acc;
        out_count[blockIdx.x] = acc;
    }
    __syncthreads();
    FastPoint* o = out + (size_t)blockIdx.x * out_cap;
    for (int r = threadIdx.x / 32; r < rg.h; r += 8)
    {
        const int n = offs[r + 1] - offs[r];
        for (int k = threadIdx.x & 31; k < n; k += 32)
        {
            const int dst = offs[r] + k;
            if (dst < out_cap)
            {
                FastPoint p;
                p.x = (short)row_x[(rbase + r) * row_cap + k];
                p.y = (short)r;
                p.score = row_s[(rbase + r) * row_cap + k];
                o[dst] = p;
            }
        }
    }
}

}  // namespace

lvkb200_status FastDetector::prepare(int width, int height)
{
    if (width == w && height == h) return LVKB200_OK;
    LVKB_REQUIRE(width >= 7 && height >= 7 && width < 65536 && height < 65536);
    score_pitch = (size_t)((width + 15) / 16 * 16);
    row_cap = width / 2 + 2;  // strict 8-neighbour NMS: no two adjacent pixels survive
    max_rows = height;
    out_cap = ((width + 1) / 2) * ((height + 1) / 2) + 1;
    // No clearing needed (and none on the legacy stream, which would race with this stream's kernels):
    // k_fast_score writes every pixel of a region before k_fast_nms_row reads it, and nothing else is read.
    LVKB_CUDA(d_score.ensure(score_pitch * height));
    LVKB_CUDA(d_row_x.ensure(sizeof(uint16_t) * FAST_MAX_REGIONS * (size_t)max_rows * row_cap));
    LVKB_CUDA(d_row_s.ensure(sizeof(uint8_t) * FAST_MAX_REGIONS * (size_t)max_rows * row_cap));
    LVKB_CUDA(d_row_count.ensure(sizeof(int) * FAST_MAX_REGIONS * (size_t)max_rows));
    LVKB_CUDA(h_count.ensure(sizeof(int) * FAST_MAX_REGIONS));
    LVKB_CUDA(h_out.ensure(sizeof(FastPoint) * FAST_MAX_REGIONS * (size_t)out_cap));
    if (!done) LVKB_CUDA(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
    w = width;
    h = height;
    return LVKB200_OK;
}

void FastDetector::release()
{
    d_score.release(); d_row_x.release(); d_row_s.release(); d_row_count.release();
    h_count.release(); h_out.release();
    if (done) cudaEventDestroy(done);
    done = nullptr;
    w = h = 0;
}

lvkb200_status FastDetector::launch(cudaStream_t cs, const uint8_t* img, size_t pitch, const FastRegion* regions, int n)
{
    LVKB_REQUIRE(n >= 1 && n <= FAST_MAX_REGIONS);
    RegionsArg arg{};
    arg.n = n;
    int max_tiles = 0, max_h = 0;
    for (int i = 0; i < n; i++)
    {
        const FastRegion& r = regions[i];
        LVKB_REQUIRE(r.x >= 0 && r.y >= 0 && r.w > 0 && r.h > 0 && r.x + r.w <= w && r.y + r.h <= h);
        arg.r[i] = r;
        arg.r[i].threshold = std::min(std::max(r.threshold, 0), 255);  // cv::FAST clamps the threshold
        max_tiles = std::max(max_tiles, div_up(r.w, TW) * div_up(r.h, TH));
        max_h = std::max(max_h, r.h);
    }
    k_fast_score<<<dim3(max_tiles, 1, n), TW * TH, 0, cs>>>(img, pitch, arg, d_score.as<uint8_t>(), score_pitch);
    k_fast_nms_row<<<dim3(max_h, n), 128, 0, cs>>>(d_score.as<uint8_t>(), score_pitch, arg, max_rows, row_cap,
                                                   d_row_x.as<uint16_t>(), d_row_s.as<uint8_t>(),
                                                   d_row_count.as<int>());
    k_fast_gather<<<n, 256, sizeof(int) * (max_rows + 1), cs>>>(arg, max_rows, row_cap, d_row_x.as<uint16_t>(),
                                                                d_row_s.as<uint8_t>(), d_row_count.as<int>(), out_cap,
                                                                h_out.device_view<FastPoint>(),
                                                                h_count.device_view<int>());
    count_launches(3);
    LVKB_CUDA(cudaGetLastError());
    LVKB_CUDA(cudaEventRecord(done, cs));
    launched = n;
    return LVKB200_OK;
}

lvkb200_status FastDetector::fetch(std::vector<std::vector<FastPoint>>& out)
{
    const int n = launched;
    out.resize(n);
    LVKB_CUDA(cudaEventSynchronize(done));
    const int* cnt = h_count.as<int>();
    for (int i = 0; i < n; i++)
    {
        const int c = std::min(cnt[i], out_cap);
        const FastPoint* p = h_out.as<FastPoint>() + (size_t)i * out_cap;
        out[i].assign(p, p + c);
    }
    return LVKB200_OK;
}

}  // namespace lvkb200
