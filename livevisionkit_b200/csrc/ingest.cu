// K0 — fused colour->gray + cv::resize(INTER_AREA) for sm_100a.
//
// Replaces VideoFrame::viewAsFormat(GRAY) (LiveVisionKit/Data/VideoFrame.cpp:187-317: cvtColor / extractChannel)
// followed by cv::resize(gray, detection_resolution, INTER_AREA) (LiveVisionKit/Vision/FrameTracker.cpp:117).
// The full-resolution gray image is never materialised: the newest frame is read once (3 B/px, the only
// O(W*H) traffic of the tracking chain) and the <=130 KB detection image is written.
//
// Bit-exact restatement of OpenCV's arithmetic (upstream, not under /root/reference; verified against cv2 4.13):
//   * BGR/RGB(A)->GRAY: (B*3735 + G*19235 + R*9798 + 2^14) >> 15;   YUV->GRAY: channel 0.
//   * INTER_AREA, integer factor on both axes: block sum, then (sum+2)>>2 for 2x2, else rint(float(sum)*(1.f/area)).
//   * INTER_AREA, any fractional axis: per-axis (src index, float weight) tables; per source row
//     buf = ((0 + S0*a0) + S1*a1) + ... in float32 (no FMA), then sum = b0*buf0, sum += bj*bufj; result rint(sum).
// One CTA produces a 64x2 tile of the detection image: it stages the gray source footprint of the tile in shared
// memory with coalesced reads (lane = consecutive source pixel), then reduces horizontally and vertically.

#include <cmath>
#include <vector>

#include "common.hpp"
#include "ingest.hpp"

namespace lvkb200
{
namespace
{

constexpr int DT_W = 64;
constexpr int DT_H = 2;
constexpr int THREADS = 256;
constexpr int MAX_SW = 1024;  // staged source tile capacity
constexpr int MAX_SH = 24;
constexpr int MAX_YT = 14;  // vertical table entries per destination row

__global__ void __launch_bounds__(THREADS)
    k_ingest_gray_area(const uint8_t* __restrict__ src, size_t pitch, int px_stride, int c0, int c1, int c2, int dw,
                       int dh, const int2* __restrict__ xtab, const float* __restrict__ xw, int xt,
                       const int2* __restrict__ ytab, const float* __restrict__ yw, int yt, int fast, float fast_scale,
                       uint8_t* __restrict__ dst, size_t dst_pitch)
{
    __shared__ __align__(16) uint8_t gray[MAX_SH * MAX_SW];
    __shared__ float buf[DT_H * MAX_YT * DT_W];

    const int dx0 = blockIdx.x * DT_W, dy0 = blockIdx.y * DT_H;
    const int ndx = min(DT_W, dw - dx0), ndy = min(DT_H, dh - dy0);

    const int2 xf = __ldg(&xtab[dx0]), xl = __ldg(&xtab[dx0 + ndx - 1]);
    const int2 yf = __ldg(&ytab[dy0]), yl = __ldg(&ytab[dy0 + ndy - 1]);
    const int sx0 = xf.x, sw = xl.x + xl.y - xf.x;
    const int sy0 = yf.x, sh = yl.x + yl.y - yf.x;

    // ---- stage 1: gray source footprint -> smem (coalesced along x)
    // Fast path for packed 3-byte pixels on 4-byte aligned rows: 4 pixels = 12 bytes = three 32-bit loads per thread
    // (4x fewer load instructions than byte loads), one 32-bit shared store.
    const bool vec4 = (px_stride == 3) && ((pitch & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 3) == 0) &&
                      (((size_t)sx0 * 3) & 3) == 0 && !(c1 == 0 && c2 == 0);
    if (vec4)
    {
        const int groups = sw >> 2;  // whole groups of 4 pixels; the <= 3 pixel tail falls through to the byte loop
        for (int idx = threadIdx.x; idx < groups * sh; idx += THREADS)
        {
            const int r = idx / groups, gidx = idx - r * groups;
            const uint32_t* p = reinterpret_cast<const uint32_t*>(src + (size_t)(sy0 + r) * pitch + (size_t)sx0 * 3) + 3 * gidx;
            const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
            const int g0 = (c0 * (int)(w0 & 255) + c1 * (int)((w0 >> 8) & 255) + c2 * (int)((w0 >> 16) & 255) + (1 << 14)) >> 15;
            const int g1 = (c0 * (int)(w0 >> 24) + c1 * (int)(w1 & 255) + c2 * (int)((w1 >> 8) & 255) + (1 << 14)) >> 15;
            const int g2 = (c0 * (int)((w1 >> 16) & 255) + c1 * (int)(w1 >> 24) + c2 * (int)(w2 & 255) + (1 << 14)) >> 15;
            const int g3 = (c0 * (int)((w2 >> 8) & 255) + c1 * (int)((w2 >> 16) & 255) + c2 * (int)(w2 >> 24) + (1 << 14)) >> 15;
            *reinterpret_cast<uint32_t*>(&gray[r * MAX_SW + 4 * gidx]) = (uint32_t)g0 | ((uint32_t)g1 << 8) | ((uint32_t)g2 << 16) | ((uint32_t)g3 << 24);
        }
    }
    const int c_begin = vec4 ? (sw & ~3) : 0;
    const int tail = sw - c_begin;
    for (int idx = threadIdx.x; idx < tail * sh; idx += THREADS)
    {
        const int r = idx / tail, c = c_begin + (idx - r * tail);
        const uint8_t* p = src + (size_t)(sy0 + r) * pitch + (size_t)(sx0 + c) * px_stride;
        int g;
        if (c1 == 0 && c2 == 0)
            g = __ldg(p);  // YUV: extractChannel(0); GRAY: as is
        else
            g = (c0 * (int)__ldg(p) + c1 * (int)__ldg(p + 1) + c2 * (int)__ldg(p + 2) + (1 << 14)) >> 15;
        gray[r * MAX_SW + c] = (uint8_t)g;
    }
    __syncthreads();

    // ---- stage 2: horizontal pass per (dst row, table row, dst col)
    // flat work list: for dyl in [0,ndy): for j in [0, ycount(dy)): for dxl in [0, ndx)
    const int yc0 = yf.y, yc1 = (ndy > 1) ? yl.y : 0;
    const int total = (yc0 + yc1) * ndx;
    for (int idx = threadIdx.x; idx < total; idx += THREADS)
    {
        const int rowi = idx / ndx, dxl = idx - rowi * ndx;
        const int dyl = (rowi < yc0) ? 0 : 1;
        const int j = (dyl == 0) ? rowi : rowi - yc0;
        const int2 ye = (dyl == 0) ? yf : yl;
        const int r = ye.x + j - sy0;
        const int2 xe = __ldg(&xtab[dx0 + dxl]);
        const uint8_t* g = &gray[r * MAX_SW + (xe.x - sx0)];
        float v;
        if (fast)
        {
            int s = 0;
            for (int k = 0; k < xe.y; k++) s += g[k];
            v = __int_as_float(s);  // exact integer partial sum, carried bit-wise
        }
        else
        {
            const float* w = xw + (size_t)(dx0 + dxl) * xt;
            v = 0.0f;
            for (int k = 0; k < xe.y; k++) v = __fadd_rn(v, __fmul_rn((float)g[k], __ldg(&w[k])));
        }
        buf[(dyl * MAX_YT + j) * DT_W + dxl] = v;
    }
    __syncthreads();

    // ---- stage 3: vertical pass, one thread per destination pixel
    if (threadIdx.x < DT_W * DT_H)
    {
        const int dyl = threadIdx.x / DT_W, dxl = threadIdx.x - dyl * DT_W;
        if (dyl < ndy && dxl < ndx)
        {
            const int2 ye = (dyl == 0) ? yf : yl;
            int out;
            if (fast)
            {
                int s = 0;
                for (int j = 0; j < ye.y; j++) s += __float_as_int(buf[(dyl * MAX_YT + j) * DT_W + dxl]);
                if (fast == 2)
                    out = (s + 2) >> 2;  // OpenCV's 2x2 special case
                else
                    out = __float2int_rn(__fmul_rn((float)s, fast_scale));
            }
            else
            {
                const float* w = yw + (size_t)(dy0 + dyl) * yt;
                float s = __fmul_rn(__ldg(&w[0]), buf[(dyl * MAX_YT + 0) * DT_W + dxl]);
                for (int j = 1; j < ye.y; j++)
                    s = __fadd_rn(s, __fmul_rn(__ldg(&w[j]), buf[(dyl * MAX_YT + j) * DT_W + dxl]));
                out = __float2int_rn(s);
            }
            out = max(0, min(255, out));
            dst[(size_t)(dy0 + dyl) * dst_pitch + dx0 + dxl] = (uint8_t)out;
        }
    }
}

// Integer scale factors with packed 3-byte pixels and SX a multiple of 4 (1080p and 4K onto 480x270: SX = 4, 8): one
// thread per destination pixel, no shared memory and no tables.  A destination pixel's slice of a source row is
// 3*SX bytes = 3*SX/4 ALIGNED 32-bit words, consecutive lanes read consecutive slices, and all sy*3*SX/4 loads of a
// thread are independent.  Same arithmetic as the tiled kernel's integer path (exact integer sum, one rounding).
template <int SX>
__global__ void __launch_bounds__(128)
    k_ingest_gray_int(const uint8_t* __restrict__ src, size_t pitch, int c0, int c1, int c2, int dw, int dh, int sy,
                      int mode, float scale, uint8_t* __restrict__ dst, size_t dst_pitch)
{
    constexpr int W = 3 * SX / 4;
    const int dx = blockIdx.x * 32 + (threadIdx.x & 31), dy = blockIdx.y * 4 + (threadIdx.x >> 5);
    if (dx >= dw || dy >= dh) return;
    const uint8_t* row = src + (size_t)dy * sy * pitch + (size_t)dx * (3 * SX);
    const bool first_channel = (c1 == 0 && c2 == 0);  // YUV: extractChannel(0)
    int sum = 0;
#pragma unroll 4
    for (int r = 0; r < sy; r++, row += pitch)
    {
        uint32_t w[W];
#pragma unroll
        for (int k = 0; k < W; k++) w[k] = __ldg(reinterpret_cast<const uint32_t*>(row) + k);
#pragma unroll
        for (int g = 0; g < W; g += 3)
        {
            const uint32_t w0 = w[g], w1 = w[g + 1], w2 = w[g + 2];
            if (first_channel)
                sum += (int)(w0 & 255) + (int)(w0 >> 24) + (int)((w1 >> 16) & 255) + (int)((w2 >> 8) & 255);
            else
            {
                sum += (c0 * (int)(w0 & 255) + c1 * (int)((w0 >> 8) & 255) + c2 * (int)((w0 >> 16) & 255) + (1 << 14)) >> 15;
                sum += (c0 * (int)(w0 >> 24) + c1 * (int)(w1 & 255) + c2 * (int)((w1 >> 8) & 255) + (1 << 14)) >> 15;
                sum += (c0 * (int)((w1 >> 16) & 255) + c1 * (int)(w1 >> 24) + c2 * (int)(w2 & 255) + (1 << 14)) >> 15;
                sum += (c0 * (int)((w2 >> 8) & 255) + c1 * (int)((w2 >> 16) & 255) + c2 * (int)(w2 >> 24) + (1 << 14)) >> 15;
            }
        }
    }
    int out = (mode == 2) ? ((sum + 2) >> 2) : __float2int_rn(__fmul_rn((float)sum, scale));
    out = max(0, min(255, out));
    dst[(size_t)dy * dst_pitch + dx] = (uint8_t)out;
}

// OpenCV computeResizeAreaTab (upstream imgproc/resize.cpp), restated.
void build_axis(int ssize, int dsize, std::vector<int2>& tab, std::vector<float>& w, int& max_count)
{
    const double scale = (double)ssize / (double)dsize;
    std::vector<std::vector<std::pair<int, float>>> rows(dsize);
    max_count = 0;
    for (int d = 0; d < dsize; d++)
    {
        const double fsx1 = d * scale, fsx2 = fsx1 + scale;
        const double cell = std::min(scale, ssize - fsx1);
        int sx1 = (int)std::ceil(fsx1), sx2 = (int)std::floor(fsx2);
        sx2 = std::min(sx2, ssize - 1);
        sx1 = std::min(sx1, sx2);
        auto& e = rows[d];
        if (sx1 - fsx1 > 1e-3) e.emplace_back(sx1 - 1, (float)((sx1 - fsx1) / cell));
        for (int s = sx1; s < sx2; s++) e.emplace_back(s, (float)(1.0 / cell));
        if (fsx2 - sx2 > 1e-3) e.emplace_back(sx2, (float)(std::min(std::min(fsx2 - sx2, 1.0), cell) / cell));
        max_count = std::max(max_count, (int)e.size());
    }
    tab.resize(dsize);
    w.assign((size_t)dsize * max_count, 0.0f);
    for (int d = 0; d < dsize; d++)
    {
        tab[d] = make_int2(rows[d].empty() ? 0 : rows[d][0].first, (int)rows[d].size());
        for (size_t k = 0; k < rows[d].size(); k++) w[(size_t)d * max_count + k] = rows[d][k].second;
    }
}

}  // namespace

lvkb200_status IngestPlan::prepare(int src_w, int src_h, int dst_w, int dst_h, cudaStream_t cs)
{
    if (src_w == sw && src_h == sh && dst_w == dw && dst_h == dh) return LVKB200_OK;
    LVKB_REQUIRE(src_w >= dst_w && src_h >= dst_h);  // INTER_AREA up-scaling (bilinear-like) is not on the path
    std::vector<int2> xt, yt;
    std::vector<float> xwv, ywv;
    build_axis(src_w, dst_w, xt, xwv, xcount);
    build_axis(src_h, dst_h, yt, ywv, ycount);
    LVKB_REQUIRE(ycount <= MAX_YT);
    // staged footprint limits of one 64x2 tile
    const double scale_x = (double)src_w / dst_w, scale_y = (double)src_h / dst_h;
    LVKB_REQUIRE(scale_x * DT_W + 2 <= MAX_SW && scale_y * DT_H + 2 <= MAX_SH);

    isx = (int)std::lrint(scale_x);
    isy = (int)std::lrint(scale_y);
    fast = 0;
    fast_scale = 0.f;
    if (std::fabs(scale_x - isx) < 2.220446049250313e-16 && std::fabs(scale_y - isy) < 2.220446049250313e-16)
    {
        fast = (isx == 2 && isy == 2) ? 2 : 1;
        fast_scale = 1.0f / (float)(isx * isy);
        if (isx == 1 && isy == 1) fast = 1;  // same size: cv::resize copies; sum*1.0 is the identity
    }
    LVKB_CUDA(d_xtab.ensure(xt.size() * sizeof(int2)));
    LVKB_CUDA(d_ytab.ensure(yt.size() * sizeof(int2)));
    LVKB_CUDA(d_xw.ensure(xwv.size() * sizeof(float)));
    LVKB_CUDA(d_yw.ensure(ywv.size() * sizeof(float)));
    LVKB_CUDA(cudaStreamSynchronize(cs));  // tables of a previous geometry may still be in use
    // stream-ordered uploads (the stream is non-blocking: the legacy stream would not be ordered with its kernels)
    LVKB_CUDA(cudaMemcpyAsync(d_xtab.ptr, xt.data(), xt.size() * sizeof(int2), cudaMemcpyHostToDevice, cs));
    LVKB_CUDA(cudaMemcpyAsync(d_ytab.ptr, yt.data(), yt.size() * sizeof(int2), cudaMemcpyHostToDevice, cs));
    LVKB_CUDA(cudaMemcpyAsync(d_xw.ptr, xwv.data(), xwv.size() * sizeof(float), cudaMemcpyHostToDevice, cs));
    LVKB_CUDA(cudaMemcpyAsync(d_yw.ptr, ywv.data(), ywv.size() * sizeof(float), cudaMemcpyHostToDevice, cs));
    LVKB_CUDA(cudaStreamSynchronize(cs));  // the host vectors die at scope exit
    sw = src_w; sh = src_h; dw = dst_w; dh = dst_h;
    return LVKB200_OK;
}

void IngestPlan::release()
{
    d_xtab.release(); d_ytab.release(); d_xw.release(); d_yw.release();
    sw = sh = dw = dh = 0;
}

lvkb200_status IngestPlan::launch(cudaStream_t cs, const uint8_t* src, size_t pitch, lvkb200_format format,
                                  uint8_t* dst, size_t dst_pitch) const
{
    int stride = 3, c0 = 0, c1 = 0, c2 = 0;
    switch (format)
    {
        case LVKB200_BGR: c0 = 3735; c1 = 19235; c2 = 9798; break;             // cv::COLOR_BGR2GRAY
        case LVKB200_RGB: c0 = 9798; c1 = 19235; c2 = 3735; break;             // cv::COLOR_RGB2GRAY
        case LVKB200_BGRA: stride = 4; c0 = 3735; c1 = 19235; c2 = 9798; break;  // cv::COLOR_BGRA2GRAY
        case LVKB200_RGBA: stride = 4; c0 = 9798; c1 = 19235; c2 = 3735; break;  // cv::COLOR_RGBA2GRAY
        case LVKB200_YUV: c0 = 1; break;                                        // cv::extractChannel(0)
        case LVKB200_GRAY: stride = 1; c0 = 1; break;
        default: LVKB_REQUIRE(format != LVKB200_UNKNOWN);  // StabilizationFilter.cpp:71
    }
    const bool aligned = stride == 3 && (pitch & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 3) == 0;
    if (fast != 0 && aligned && (isx == 4 || isx == 8) && format != LVKB200_GRAY)
    {
        const dim3 g(div_up(dw, 32), div_up(dh, 4));
        if (isx == 4)
            k_ingest_gray_int<4><<<g, 128, 0, cs>>>(src, pitch, c0, c1, c2, dw, dh, isy, fast, fast_scale, dst, dst_pitch);
        else
            k_ingest_gray_int<8><<<g, 128, 0, cs>>>(src, pitch, c0, c1, c2, dw, dh, isy, fast, fast_scale, dst, dst_pitch);
        count_launches(1);
        LVKB_CUDA(cudaGetLastError());
        return LVKB200_OK;
    }
    const dim3 grid(div_up(dw, DT_W), div_up(dh, DT_H));
    k_ingest_gray_area<<<grid, THREADS, 0, cs>>>(src, pitch, stride, c0, c1, c2, dw, dh, d_xtab.as<int2>(),
                                                 d_xw.as<float>(), xcount, d_ytab.as<int2>(), d_yw.as<float>(), ycount,
                                                 fast, fast_scale, dst, dst_pitch);
    count_launches(1);
    LVKB_CUDA(cudaGetLastError());
    return LVKB200_OK;
}

}  // namespace lvkb200
