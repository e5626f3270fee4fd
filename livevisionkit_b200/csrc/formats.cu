// Frame ingest / egress kernels for sm_100a: OBS plane layouts <-> packed 8UC3 {Y, U, V}.
//
// Replaces the OpenCV/OpenCL sequences of Modules/OBS-Plugin/Interop/FrameIngest.cpp:
//   I4XXIngest::to_ocl (:479-522)  resize(U), resize(V) INTER_LINEAR + merge        -> k_planes_to_packed
//   NV12Ingest::to_ocl (:566-584)  resize(UV 8UC2) INTER_LINEAR + mixChannels       -> k_planes_to_packed (xstride 2)
//   P422Ingest::to_ocl (:618-645)  extractChannel + reshape + resize + mixChannels  -> k_planes_to_packed (xstride 4)
//   P444Ingest::to_ocl (:689-698)  mixChannels {1,0, 2,1, 3,2}                      -> k_planes_to_packed (no resample)
//   ...::to_obs (:526-560, :588-604, :649-678, :702-716)  split + resize INTER_AREA -> k_packed_to_planes
// Every layout is described by three PlaneRef {base, pitch, xstride}; one kernel per direction covers all of them.
//
// Arithmetic (bit-exact with OpenCV's CPU path, pinned by tests/test_formats_cpu.py against cv2):
//   upsample  : 11-bit fixed-point taps, horizontal pass kept at 22 bits, vertical pass
//               ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16), then (v + 2) >> 2       (resize.cpp VResizeLinear<uchar>)
//   2x2 area  : (a + b + c + d + 2) >> 2 on single-channel planes                             (ResizeAreaFastVec)
//               saturate_cast<uchar>(sum * 0.25f) = round half to even on NV12's 2-channel plane (generic loop: the
//               vector path only exists for 1, 3 and 4 channels)
//   2x1 area  : saturate_cast<uchar>((a + b) * 0.5f) = round half to even                    (resizeAreaFast_ generic)
// Roofline: HBM.  4:2:0 ingest reads 1.5 B/px and writes 3 B/px; egress the reverse.

#include <cmath>

#include "formats.hpp"

namespace lvkb200
{

std::vector<LinearTap> linear_taps(int src, int dst, bool horizontal)
{
    // resize.cpp: inv_scale = dsize / ssize (double); scale = 1 / inv_scale; f = (float)((d + 0.5) * scale - 0.5);
    // s = cvFloor(f); f -= s; weights = saturate_cast<short>((1 - f, f) * INTER_RESIZE_COEF_SCALE)  [cvRound: half even]
    std::vector<LinearTap> taps(static_cast<size_t>(dst));
    const double inv_scale = static_cast<double>(dst) / static_cast<double>(src);
    const double scale = 1.0 / inv_scale;
    for (int d = 0; d < dst; d++)
    {
        float f = static_cast<float>((d + 0.5) * scale - 0.5);
        int s = static_cast<int>(std::floor(f));
        f -= static_cast<float>(s);
        if (horizontal)
        {
            if (s < 0) { f = 0.f; s = 0; }
            if (s >= src - 1) { f = 0.f; s = src - 1; }
        }
        LinearTap t;
        t.ofs = s;
        t.w0 = static_cast<short>(std::nearbyintf((1.f - f) * 2048.f));
        t.w1 = static_cast<short>(std::nearbyintf(f * 2048.f));
        taps[static_cast<size_t>(d)] = t;
    }
    return taps;
}

cudaError_t FormatPlan::prepare(int w, int h, int cw, int ch, cudaStream_t cs)
{
    if (w == width && h == height && cw == chroma_w && ch == chroma_h) return cudaSuccess;
    width = height = 0;
    if (cw != w || ch != h)
    {
        const std::vector<LinearTap> xt = linear_taps(cw, w, true), yt = linear_taps(ch, h, false);
        cudaError_t e = xtab.ensure(xt.size() * sizeof(LinearTap));
        if (e != cudaSuccess) return e;
        e = ytab.ensure(yt.size() * sizeof(LinearTap));
        if (e != cudaSuccess) return e;
        // pageable source: the copy is staged by the runtime before the call returns, the vectors may die afterwards
        e = cudaMemcpyAsync(xtab.ptr, xt.data(), xt.size() * sizeof(LinearTap), cudaMemcpyHostToDevice, cs);
        if (e != cudaSuccess) return e;
        e = cudaMemcpyAsync(ytab.ptr, yt.data(), yt.size() * sizeof(LinearTap), cudaMemcpyHostToDevice, cs);
        if (e != cudaSuccess) return e;
    }
    width = w; height = h; chroma_w = cw; chroma_h = ch;
    return cudaSuccess;
}

namespace
{

struct ConstPlane
{
    const uint8_t* base;
    size_t pitch;
    int xstride;
};

// One chroma sample of the upsampled plane: the OpenCV two-pass fixed-point bilinear.
__device__ __forceinline__ unsigned upsample(const ConstPlane& p, int cw, int ch, LinearTap tx, LinearTap ty)
{
    const int x0 = tx.ofs, x1 = min(tx.ofs + 1, cw - 1);
    const int y0 = min(max(ty.ofs, 0), ch - 1), y1 = min(max(ty.ofs + 1, 0), ch - 1);
    const uint8_t* r0 = p.base + (size_t)y0 * p.pitch;
    const uint8_t* r1 = p.base + (size_t)y1 * p.pitch;
    const int s0 = (int)__ldg(r0 + x0 * p.xstride) * tx.w0 + (int)__ldg(r0 + x1 * p.xstride) * tx.w1;
    const int s1 = (int)__ldg(r1 + x0 * p.xstride) * tx.w0 + (int)__ldg(r1 + x1 * p.xstride) * tx.w1;
    const int v = ((ty.w0 * (s0 >> 4)) >> 16) + ((ty.w1 * (s1 >> 4)) >> 16);
    return (unsigned)((v + 2) >> 2);
}

// One thread = 4 horizontally adjacent destination pixels = 12 packed bytes = three 32-bit stores.
template <bool RESAMPLE>
__global__ void __launch_bounds__(256)
    k_planes_to_packed(ConstPlane Y, ConstPlane U, ConstPlane V, int W, int H, int cw, int ch,
                       const LinearTap* __restrict__ xtab, const LinearTap* __restrict__ ytab, uint8_t* __restrict__ dst,
                       size_t dst_pitch, int words_ok)
{
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y;
    if (x4 >= W) return;
    LinearTap ty{};
    if (RESAMPLE) ty = ytab[y];
    unsigned px[4][3];
#pragma unroll
    for (int i = 0; i < 4; i++)
    {
        const int x = min(x4 + i, W - 1);
        px[i][0] = __ldg(Y.base + (size_t)y * Y.pitch + (size_t)x * Y.xstride);
        if (RESAMPLE)
        {
            const LinearTap tx = xtab[x];
            px[i][1] = upsample(U, cw, ch, tx, ty);
            px[i][2] = upsample(V, cw, ch, tx, ty);
        }
        else
        {
            px[i][1] = __ldg(U.base + (size_t)y * U.pitch + (size_t)x * U.xstride);
            px[i][2] = __ldg(V.base + (size_t)y * V.pitch + (size_t)x * V.xstride);
        }
    }
    uint8_t* q = dst + (size_t)y * dst_pitch + 3 * (size_t)x4;
    if (words_ok && x4 + 3 < W)
    {
        uint32_t* q32 = reinterpret_cast<uint32_t*>(q);
        q32[0] = px[0][0] | (px[0][1] << 8) | (px[0][2] << 16) | (px[1][0] << 24);
        q32[1] = px[1][1] | (px[1][2] << 8) | (px[2][0] << 16) | (px[2][1] << 24);
        q32[2] = px[2][2] | (px[3][0] << 8) | (px[3][1] << 16) | (px[3][2] << 24);
    }
    else
    {
        for (int i = 0; i < 4 && x4 + i < W; i++)
        {
            q[3 * i] = (uint8_t)px[i][0]; q[3 * i + 1] = (uint8_t)px[i][1]; q[3 * i + 2] = (uint8_t)px[i][2];
        }
    }
}

// saturate_cast<uchar>((a + b) * 0.5f): cvRound = round half to even
__device__ __forceinline__ unsigned mean2_half_even(unsigned a, unsigned b)
{
    const unsigned s = a + b;
    return (s >> 1) + ((s & 1u) & ((s >> 1) & 1u));
}

// One thread = a block of 4 x SUB_Y source pixels: 4 (x SUB_Y) luma samples and 4 / SUB_X chroma samples per plane.
__device__ __forceinline__ unsigned mean4_half_even(unsigned s)
{
    const unsigned q = s >> 2, r = s & 3u;
    return q + ((r == 3u || (r == 2u && (q & 1u))) ? 1u : 0u);
}

template <int SUB_X, int SUB_Y>
__global__ void __launch_bounds__(256)
    k_packed_to_planes(const uint8_t* __restrict__ src, size_t src_pitch, int W, int H, PlaneRef Y, PlaneRef U, PlaneRef V,
                       PlaneRef A, int has_alpha, int half_even_2x2)
{
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y0 = blockIdx.y * SUB_Y;
    if (x4 >= W) return;
    unsigned c[SUB_Y][4][3];
#pragma unroll
    for (int r = 0; r < SUB_Y; r++)
    {
        const uint8_t* p = src + (size_t)min(y0 + r, H - 1) * src_pitch;
#pragma unroll
        for (int i = 0; i < 4; i++)
        {
            const int x = min(x4 + i, W - 1);
            c[r][i][0] = __ldg(p + 3 * x); c[r][i][1] = __ldg(p + 3 * x + 1); c[r][i][2] = __ldg(p + 3 * x + 2);
        }
    }
#pragma unroll
    for (int r = 0; r < SUB_Y; r++)
    {
        if (y0 + r >= H) break;
#pragma unroll
        for (int i = 0; i < 4; i++)
        {
            if (x4 + i >= W) break;
            Y.base[(size_t)(y0 + r) * Y.pitch + (size_t)(x4 + i) * Y.xstride] = (uint8_t)c[r][i][0];
            if (has_alpha) A.base[(size_t)(y0 + r) * A.pitch + (size_t)(x4 + i) * A.xstride] = 255;
        }
    }
    const int cy = blockIdx.y;
#pragma unroll
    for (int i = 0; i < 4; i += SUB_X)
    {
        if (x4 + i >= W) break;
        const int cx = (x4 + i) / SUB_X;
        unsigned u, v;
        if (SUB_X == 1 && SUB_Y == 1)
        {
            u = c[0][i][1]; v = c[0][i][2];
        }
        else if (SUB_X == 2 && SUB_Y == 1)
        {
            u = mean2_half_even(c[0][i][1], c[0][i + 1][1]);
            v = mean2_half_even(c[0][i][2], c[0][i + 1][2]);
        }
        else
        {
            const unsigned su = c[0][i][1] + c[0][i + 1][1] + c[SUB_Y - 1][i][1] + c[SUB_Y - 1][i + 1][1];
            const unsigned sv = c[0][i][2] + c[0][i + 1][2] + c[SUB_Y - 1][i][2] + c[SUB_Y - 1][i + 1][2];
            u = half_even_2x2 ? mean4_half_even(su) : (su + 2u) >> 2;
            v = half_even_2x2 ? mean4_half_even(sv) : (sv + 2u) >> 2;
        }
        U.base[(size_t)cy * U.pitch + (size_t)cx * U.xstride] = (uint8_t)u;
        V.base[(size_t)cy * V.pitch + (size_t)cx * V.xstride] = (uint8_t)v;
    }
}

}  // namespace

cudaError_t launch_planes_to_packed(cudaStream_t cs, const FormatPlan& plan, PlaneRef y, PlaneRef u, PlaneRef v,
                                    uint8_t* dst, size_t dst_pitch)
{
    const int W = plan.width, H = plan.height;
    const dim3 grid(div_up(div_up(W, 4), 256), H);
    const ConstPlane Y{y.base, y.pitch, y.xstride}, U{u.base, u.pitch, u.xstride}, V{v.base, v.pitch, v.xstride};
    const int words_ok = ((reinterpret_cast<uintptr_t>(dst) & 3u) == 0 && (dst_pitch & 3u) == 0) ? 1 : 0;
    if (plan.chroma_w != W || plan.chroma_h != H)
        k_planes_to_packed<true><<<grid, 256, 0, cs>>>(Y, U, V, W, H, plan.chroma_w, plan.chroma_h,
                                                       plan.xtab.as<LinearTap>(), plan.ytab.as<LinearTap>(), dst,
                                                       dst_pitch, words_ok);
    else
        k_planes_to_packed<false><<<grid, 256, 0, cs>>>(Y, U, V, W, H, W, H, nullptr, nullptr, dst, dst_pitch, words_ok);
    count_launches(1);
    return cudaGetLastError();
}

cudaError_t launch_packed_to_planes(cudaStream_t cs, const uint8_t* src, size_t src_pitch, int w, int h, int sub_x,
                                    int sub_y, PlaneRef y, PlaneRef u, PlaneRef v, const PlaneRef* alpha,
                                    bool interleaved_chroma)
{
    const int he = interleaved_chroma ? 1 : 0;
    const PlaneRef a = alpha ? *alpha : PlaneRef{nullptr, 0, 0};
    const dim3 grid(div_up(div_up(w, 4), 256), div_up(h, sub_y));
    if (sub_x == 1 && sub_y == 1)
        k_packed_to_planes<1, 1><<<grid, 256, 0, cs>>>(src, src_pitch, w, h, y, u, v, a, alpha ? 1 : 0, he);
    else if (sub_x == 2 && sub_y == 1)
        k_packed_to_planes<2, 1><<<grid, 256, 0, cs>>>(src, src_pitch, w, h, y, u, v, a, alpha ? 1 : 0, he);
    else if (sub_x == 2 && sub_y == 2)
        k_packed_to_planes<2, 2><<<grid, 256, 0, cs>>>(src, src_pitch, w, h, y, u, v, a, alpha ? 1 : 0, he);
    else
        return cudaErrorInvalidValue;
    count_launches(1);
    return cudaGetLastError();
}

}  // namespace lvkb200
