// K5 + K6a — device-side swap-erase compaction and robust homography estimation (batched RANSAC hypothesis scoring
// + sigma-weighted refit) for sm_100a.
//
// Replaces, without a host round trip between them,
//   * lvk::fast_filter(features, tracked, matched, status) (LiveVisionKit/Functions/Container.tpp:97-121, called at
//     Vision/FrameTracker.cpp:149): reverse-order swap-with-last erase of the unmatched points — the ORDER it
//     produces is part of the contract (it is the order the estimator and next frame's propagate() see);
//   * cv::findHomography(tracked, matched, mask, cv::UsacParams{MAGSAC, LO_SIGMA, 50 iterations, conf .99}) as
//     called by FrameTracker::estimate_global_motion (Vision/FrameTracker.cpp:337-359).
// OpenCV's USAC lives in calib3d/usac (third-party, not under /root/reference, no source in this image); its
// sampling stream and MAGSAC weighting tables cannot be reproduced bit-for-bit, so this is a GPU-native estimator
// with the same CONTRACT:
//   * deterministic (fixed hypothesis stream, like randomGeneratorState = 0);
//   * output H normalised to h33 = 1, double precision;
//   * mask[i] = 1  <=>  float32 forward reprojection error^2 of the RETURNED H < threshold^2 (exactly how cv2's
//     mask relates to its H);
//   * degenerate input (e.g. collinear points) -> no model.
// Parity vs cv2 is therefore: identical masks except points within epsilon of the threshold, H within the
// estimator's own input-order variance (stated and measured in tests/test_tracking_gpu.py).
//
// Kernels (one dependent chain on one stream; the point count lives in device memory so nothing waits for the host):
//   k_compact_swap_erase: single CTA; replays the reference's erase order on an index permutation, then gathers
//   k_ransac_hypotheses : 256 minimal 4-point models, one thread each, closed form (unit square -> quad, no solve)
//   k_ransac_score      : one CTA per hypothesis, all points scored with a cooperative-groups block reduction
//                         (truncated-quadratic / MSAC cost at the acceptance threshold)
//   k_ransac_refine     : single CTA: arg-min hypothesis, then 3 IRLS passes of a Hartley-normalised weighted DLT
//                         (sigma-consensus style weights), warp-parallel 8x8 Gauss-Jordan, final mask.

#include <cooperative_groups.h>
#include <cooperative_groups/reduce.h>

#include "common.hpp"
#include "ransac.hpp"

namespace cg = cooperative_groups;

namespace lvkb200
{
namespace
{

constexpr int HYP = RANSAC_HYPOTHESES;

// ---------------------------------------------------------------------------------------------------------------------
// The chain's tiny inputs/outputs (params, <= 2 601 points, status/mask/model) live in MAPPED pinned host memory that
// the kernels read/write directly (LK fetches its inputs, LK and the refine kernel deliver the results).  A
// cudaMemcpyAsync would queue on the copy engines behind the multi-megabyte frame upload/download of the neighbouring
// frames (pipelined operation) and stall the whole chain; separate transfer kernels cost ~9 us each.

// ---------------------------------------------------------------------------------------------------------------------
// fast_filter: for k = n-1 .. 0: if !keep[k]: data[k] = data.back(); data.pop_back()

constexpr int CT = 1024;

__global__ void __launch_bounds__(CT)
    k_compact_swap_erase(const float2* __restrict__ a, const float2* __restrict__ b, const uint8_t* __restrict__ keep,
                         const TrackParams* __restrict__ prm,
                         float2* __restrict__ a_out, float2* __restrict__ b_out, int* __restrict__ perm,
                         int* __restrict__ removed, int* __restrict__ n_out)
{
    __shared__ int warp_tot[CT / 32];
    __shared__ int s_base, s_size;
    const int n = prm->n;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    // ascending list of removed indices (ordered compaction)
    for (int i0 = 0; i0 < n; i0 += CT)
    {
        const int i = i0 + threadIdx.x;
        const bool rem = (i < n) && (keep[i] == 0);
        if (i < n) perm[i] = i;
        const unsigned bal = __ballot_sync(0xffffffffu, rem);
        if (lane == 0) warp_tot[wid] = __popc(bal);
        __syncthreads();
        int off = s_base;
        for (int w = 0; w < wid; w++) off += warp_tot[w];
        if (rem) removed[off + __popc(bal & ((1u << lane) - 1u))] = i;
        __syncthreads();
        if (threadIdx.x == 0)
        {
            int t = 0;
            for (int w = 0; w < CT / 32; w++) t += warp_tot[w];
            s_base += t;
        }
        __syncthreads();
    }
    // replay the erases from the highest index down (sequential by definition; usually a handful of points)
    if (threadIdx.x == 0)
    {
        int size = n;
        for (int r = s_base - 1; r >= 0; r--)
        {
            const int k = removed[r];
            perm[k] = perm[size - 1];
            size--;
        }
        s_size = size;
        *n_out = size;
    }
    __syncthreads();
    const int size = s_size;
    for (int i = threadIdx.x; i < size; i += CT)
    {
        const int src = perm[i];
        a_out[i] = a[src];
        b_out[i] = b[src];
    }
}

// ---------------------------------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

__device__ __forceinline__ float cross2(float2 a, float2 b, float2 c)
{
    return (b.x - a.x) * (c.y - a.y) - (b.y - a.y) * (c.x - a.x);
}

// Projective map of the unit square (0,0),(1,0),(1,1),(0,1) onto the quad q0..q3 (Heckbert 1989), row-major 3x3.
__device__ __forceinline__ bool square_to_quad(const double qx[4], const double qy[4], double m[9])
{
    const double dx1 = qx[1] - qx[2], dx2 = qx[3] - qx[2], sx = qx[0] - qx[1] + qx[2] - qx[3];
    const double dy1 = qy[1] - qy[2], dy2 = qy[3] - qy[2], sy = qy[0] - qy[1] + qy[2] - qy[3];
    const double den = dx1 * dy2 - dx2 * dy1;
    if (fabs(den) < 1e-12) return false;
    const double g = (sx * dy2 - dx2 * sy) / den, h = (dx1 * sy - sx * dy1) / den;
    m[0] = qx[1] - qx[0] + g * qx[1]; m[1] = qx[3] - qx[0] + h * qx[3]; m[2] = qx[0];
    m[3] = qy[1] - qy[0] + g * qy[1]; m[4] = qy[3] - qy[0] + h * qy[3]; m[5] = qy[0];
    m[6] = g; m[7] = h; m[8] = 1.0;
    return true;
}

__global__ void __launch_bounds__(128)
    k_ransac_hypotheses(const float2* __restrict__ src, const float2* __restrict__ dst, const int* __restrict__ n_ptr,
                        const TrackParams* __restrict__ prm, uint32_t seed, float* __restrict__ models)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= HYP) return;
    const int n = *n_ptr;
    float* out = models + (size_t)k * 9;
    out[8] = 0.0f;  // invalid until proven otherwise
    if (prm->model == 1)
    {
        // cv::estimateAffinePartial2D minimal model: 2 correspondences -> similarity [a -b tx; b a ty]
        if (n < 2) return;
        const int i0 = (int)(hash32(seed ^ hash32((uint32_t)k * 977u)) % (uint32_t)n);
        int i1 = (int)(hash32(seed ^ hash32((uint32_t)k * 977u + 1u)) % (uint32_t)(n - 1));
        if (i1 >= i0) i1++;
        const double x0 = src[i0].x, y0 = src[i0].y, x1 = src[i1].x, y1 = src[i1].y;
        const double u0 = dst[i0].x, v0 = dst[i0].y, u1 = dst[i1].x, v1 = dst[i1].y;
        const double dx = x1 - x0, dy = y1 - y0, du = u1 - u0, dv = v1 - v0;
        const double den = dx * dx + dy * dy;
        if (den < 1.0 || (du * du + dv * dv) < 1.0) return;  // coincident points: degenerate sample
        const double a = (dx * du + dy * dv) / den, b = (dx * dv - dy * du) / den;
        out[0] = (float)a; out[1] = (float)(-b); out[2] = (float)(u0 - (a * x0 - b * y0));
        out[3] = (float)b; out[4] = (float)a; out[5] = (float)(v0 - (b * x0 + a * y0));
        out[6] = 0.0f; out[7] = 0.0f; out[8] = 1.0f;
        return;
    }
    if (n < 4) return;

    int idx[4];
    uint32_t ctr = 0;
    for (int j = 0; j < 4; j++)
    {
        for (int tries = 0; tries < 64; tries++)
        {
            const int cand = (int)(hash32(seed ^ hash32((uint32_t)k * 977u + (ctr++))) % (uint32_t)n);
            bool dup = false;
            for (int q = 0; q < j; q++) dup |= (idx[q] == cand);
            if (!dup) { idx[j] = cand; break; }
            if (tries == 63) return;
        }
    }
    float2 p[4], q[4];
#pragma unroll
    for (int j = 0; j < 4; j++) { p[j] = src[idx[j]]; q[j] = dst[idx[j]]; }

    // degeneracy: no three (nearly) collinear points, and the sample must keep its orientation
    const int tri[4][3] = {{0, 1, 2}, {0, 1, 3}, {0, 2, 3}, {1, 2, 3}};
#pragma unroll
    for (int t = 0; t < 4; t++)
    {
        const float cs = cross2(p[tri[t][0]], p[tri[t][1]], p[tri[t][2]]);
        const float cd = cross2(q[tri[t][0]], q[tri[t][1]], q[tri[t][2]]);
        if (fabsf(cs) < 4.0f || fabsf(cd) < 4.0f) return;  // triangle area < 2 px^2
        if ((cs > 0.f) != (cd > 0.f)) return;
    }

    // H = S2Q(dst quad) * S2Q(src quad)^-1   (adjugate instead of the inverse; scale fixed by h33 = 1)
    double sx[4], sy[4], dx[4], dy[4], A[9], B[9];
#pragma unroll
    for (int j = 0; j < 4; j++) { sx[j] = p[j].x; sy[j] = p[j].y; dx[j] = q[j].x; dy[j] = q[j].y; }
    if (!square_to_quad(sx, sy, A) || !square_to_quad(dx, dy, B)) return;
    const double adj[9] = {A[4] * A[8] - A[5] * A[7], A[2] * A[7] - A[1] * A[8], A[1] * A[5] - A[2] * A[4],
                           A[5] * A[6] - A[3] * A[8], A[0] * A[8] - A[2] * A[6], A[2] * A[3] - A[0] * A[5],
                           A[3] * A[7] - A[4] * A[6], A[1] * A[6] - A[0] * A[7], A[0] * A[4] - A[1] * A[3]};
    double H[9];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) H[r * 3 + c] = B[r * 3] * adj[c] + B[r * 3 + 1] * adj[3 + c] + B[r * 3 + 2] * adj[6 + c];
    if (fabs(H[8]) < 1e-12) return;
    const double inv = 1.0 / H[8];
#pragma unroll
    for (int j = 0; j < 8; j++) out[j] = (float)(H[j] * inv);
    out[8] = 1.0f;
}

// OpenCV usac ReprojectionErrorForward::getError, float32
__device__ __forceinline__ float reproj_err2(const float m[9], float2 p, float2 q)
{
    const float z = 1.0f / (m[6] * p.x + m[7] * p.y + m[8]);
    const float dx = q.x - (m[0] * p.x + m[1] * p.y + m[2]) * z;
    const float dy = q.y - (m[3] * p.x + m[4] * p.y + m[5]) * z;
    return dx * dx + dy * dy;
}

__global__ void __launch_bounds__(256)
    k_ransac_score(const float2* __restrict__ src, const float2* __restrict__ dst, const int* __restrict__ n_ptr,
                   const float* __restrict__ models, const TrackParams* __restrict__ prm, float* __restrict__ scores)
{
    const float thr2 = prm->threshold_sq;
    cg::thread_block block = cg::this_thread_block();
    cg::thread_block_tile<32> warp = cg::tiled_partition<32>(block);
    __shared__ float m[9];
    __shared__ float partial[8];
    const int n = *n_ptr;
    if (threadIdx.x < 9) m[threadIdx.x] = models[(size_t)blockIdx.x * 9 + threadIdx.x];
    block.sync();
    if (m[8] == 0.0f || n < 4)
    {
        if (threadIdx.x == 0) scores[blockIdx.x] = 3.0e38f;
        return;
    }
    float cost = 0.0f;
    for (int i = threadIdx.x; i < n; i += blockDim.x)
    {
        const float e = reproj_err2(m, src[i], dst[i]);
        cost += (e == e) ? fminf(e, thr2) : thr2;  // NaN-safe truncated quadratic
    }
    cost = cg::reduce(warp, cost, cg::plus<float>());
    if (warp.thread_rank() == 0) partial[warp.meta_group_rank()] = cost;
    block.sync();
    if (threadIdx.x == 0)
    {
        float s = 0.0f;
        for (int w = 0; w < 8; w++) s += partial[w];
        scores[blockIdx.x] = s;
    }
}

constexpr int RT = 256;  // refine CTA size

// Entry (r, c) of the 8 x 9 augmented normal equations [A^T W A | A^T W b] of the weighted DLT (h33 = 1) from the 23
// moments S[0..5] = S0, S[6..11] = Su, S[12..17] = Sv, S[18..22] = Sq, each packed xx xy x yy y 1 (see refine_body).
__device__ __forceinline__ double gram_entry(const double* S, int r, int c)
{
    const int pack[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
    const int rb = r / 3, ri = r - 3 * rb;
    if (c == 8) return (rb == 0) ? S[6 + pack[ri][2]] : (rb == 1) ? S[12 + pack[ri][2]] : -S[18 + pack[ri][2]];
    const int cb = c / 3, ci = c - 3 * cb;
    if (rb == cb) return (rb < 2) ? S[pack[ri][ci]] : S[18 + pack[ri][ci]];
    if (rb + cb == 1) return 0.0;
    const int other = min(rb, cb), i = (rb < cb) ? ri : ci, j = (rb < cb) ? ci : ri;  // i: index in p, j: block-2 index
    return -S[(other == 0 ? 6 : 12) + pack[i][j]];
}

__device__ __forceinline__ void
    refine_body(const float2* __restrict__ src, const float2* __restrict__ dst, const int* __restrict__ n_ptr,
                const float* __restrict__ models, const float* __restrict__ scores,
                const TrackParams* __restrict__ prm, int iterations, RansacResult* __restrict__ result,
                uint8_t* __restrict__ mask)
{
    const float thr2 = prm->threshold_sq;
    cg::thread_block block = cg::this_thread_block();
    cg::thread_block_tile<32> warp = cg::tiled_partition<32>(block);
    __shared__ float s_best[RT / 32];
    __shared__ int s_besti[RT / 32];
    __shared__ double s_acc[RT / 32][8];
    __shared__ double s_part[RT / 32][32];
    __shared__ double s_mom[32];
    __shared__ double s_T[8];
    __shared__ float s_m[9];
    __shared__ int s_ok;
    const int tid = threadIdx.x, lane = warp.thread_rank(), wid = warp.meta_group_rank();
    const int n = *n_ptr;

    // ---- arg-min over hypotheses (lowest index wins ties)
    float best = 3.0e38f;
    int besti = -1;
    for (int k = tid; k < HYP; k += RT)
        if (scores[k] < best) { best = scores[k]; besti = k; }
    for (int o = 16; o > 0; o >>= 1)
    {
        const float ob = warp.shfl_xor(best, o);
        const int oi = warp.shfl_xor(besti, o);
        if (oi >= 0 && (besti < 0 || ob < best || (ob == best && oi < besti))) { best = ob; besti = oi; }
    }
    if (lane == 0) { s_best[wid] = best; s_besti[wid] = besti; }
    block.sync();
    if (tid == 0)
    {
        float b = 3.0e38f;
        int bi = -1;
        for (int w = 0; w < RT / 32; w++)
            if (s_besti[w] >= 0 && (bi < 0 || s_best[w] < b || (s_best[w] == b && s_besti[w] < bi))) { b = s_best[w]; bi = s_besti[w]; }
        s_ok = (n >= 4 && bi >= 0 && b < 2.9e38f) ? 1 : 0;
        if (s_ok)
            for (int j = 0; j < 9; j++) s_m[j] = models[(size_t)bi * 9 + j];
        result->n = n;
    }
    block.sync();
    if (!s_ok)
    {
        if (tid == 0) { result->found = 0; result->inliers = 0; }
        for (int i = tid; i < n; i += RT) mask[i] = 0;
        return;
    }

    if (prm->model == 1)
    {
        // cv::estimateAffinePartial2D: the mask is the inlier set of the best minimal model (RANSACPointSetRegistrator),
        // the returned transform is the least-squares similarity over those inliers (what its LM refinement converges
        // to for this linear model).  Centred closed form: a = S(x'u'+y'v')/S(x'^2+y'^2), b = S(x'v'-y'u')/S(..).
        float m[9];
        for (int j = 0; j < 9; j++) m[j] = s_m[j];
        double c[5] = {0, 0, 0, 0, 0};
        for (int i = tid; i < n; i += RT)
        {
            const float e = reproj_err2(m, src[i], dst[i]);
            const bool in = e < thr2;
            mask[i] = in ? 1 : 0;
            if (in) { c[0] += 1.0; c[1] += src[i].x; c[2] += src[i].y; c[3] += dst[i].x; c[4] += dst[i].y; }
        }
        for (int j = 0; j < 5; j++) c[j] = cg::reduce(warp, c[j], cg::plus<double>());
        if (lane == 0) for (int j = 0; j < 5; j++) s_acc[wid][j] = c[j];
        block.sync();
        if (tid < 5)
        {
            double s = 0;
            for (int w = 0; w < RT / 32; w++) s += s_acc[w][tid];
            s_T[tid] = s;
        }
        block.sync();
        const double cnt = s_T[0];
        const double cx = s_T[1] / fmax(cnt, 1.0), cy = s_T[2] / fmax(cnt, 1.0);
        const double cu = s_T[3] / fmax(cnt, 1.0), cv = s_T[4] / fmax(cnt, 1.0);
        double q[3] = {0, 0, 0};
        for (int i = tid; i < n; i += RT)
        {
            if (!mask[i]) continue;
            const double x = src[i].x - cx, y = src[i].y - cy, u = dst[i].x - cu, v = dst[i].y - cv;
            q[0] += x * x + y * y; q[1] += x * u + y * v; q[2] += x * v - y * u;
        }
        for (int j = 0; j < 3; j++) q[j] = cg::reduce(warp, q[j], cg::plus<double>());
        block.sync();
        if (lane == 0) for (int j = 0; j < 3; j++) s_acc[wid][j] = q[j];
        block.sync();
        if (tid == 0)
        {
            double s[3] = {0, 0, 0};
            for (int w = 0; w < RT / 32; w++) for (int j = 0; j < 3; j++) s[j] += s_acc[w][j];
            double a = m[0], b = m[3], tx = m[2], ty = m[5];
            if (cnt >= 2.0 && s[0] > 1e-9)
            {
                a = s[1] / s[0]; b = s[2] / s[0];
                tx = cu - (a * cx - b * cy); ty = cv - (b * cx + a * cy);
            }
            result->h[0] = a; result->h[1] = -b; result->h[2] = tx;
            result->h[3] = b; result->h[4] = a; result->h[5] = ty;
            result->h[6] = 0.0; result->h[7] = 0.0; result->h[8] = 1.0;
            result->inliers = (int)cnt;
            result->found = 1;
        }
        return;
    }

    // ---- Hartley normalisation of both point sets, once (it only conditions the normal equations)
    {
        double c[4] = {0, 0, 0, 0};
        for (int i = tid; i < n; i += RT) { c[0] += src[i].x; c[1] += src[i].y; c[2] += dst[i].x; c[3] += dst[i].y; }
        for (int j = 0; j < 4; j++) c[j] = cg::reduce(warp, c[j], cg::plus<double>());
        if (lane == 0) for (int j = 0; j < 4; j++) s_acc[wid][j] = c[j];
        block.sync();
        if (tid < 4)
        {
            double s = 0;
            for (int w = 0; w < RT / 32; w++) s += s_acc[w][tid];
            s_T[tid] = s / n;
        }
        block.sync();
        const double cx = s_T[0], cy = s_T[1], cu = s_T[2], cv = s_T[3];
        double d[2] = {0, 0};
        for (int i = tid; i < n; i += RT)
        {
            d[0] += sqrt((src[i].x - cx) * (src[i].x - cx) + (src[i].y - cy) * (src[i].y - cy));
            d[1] += sqrt((dst[i].x - cu) * (dst[i].x - cu) + (dst[i].y - cv) * (dst[i].y - cv));
        }
        for (int j = 0; j < 2; j++) d[j] = cg::reduce(warp, d[j], cg::plus<double>());
        block.sync();
        if (lane == 0) { s_acc[wid][0] = d[0]; s_acc[wid][1] = d[1]; }
        block.sync();
        if (tid < 2)
        {
            double s = 0;
            for (int w = 0; w < RT / 32; w++) s += s_acc[w][tid];
            s_T[4 + tid] = (s > 1e-9) ? 1.4142135623730951 * n / s : 1.0;
        }
        block.sync();
    }
    const double cx = s_T[0], cy = s_T[1], cu = s_T[2], cv = s_T[3], s1 = s_T[4], s2 = s_T[5];

    // ---- IRLS: weighted DLT with h33 = 1 in normalised coordinates.
    // weights: sigma-consensus style, smooth and compactly supported: w = (1 - e/c)^2 for e < c, c = 2.25 * thr^2
    // (support 1.5x the acceptance radius), so points just outside the threshold still pull a little, far ones not at all.
    //
    // The normal equations A^T W A h = A^T W b of the rows [p 0 -u p | u], [0 p -v p | v] (p = (x, y, 1)) have block
    // structure: every entry is +-one of the 23 moments  S0 = sum w p p^T, Su = sum w u p p^T, Sv = sum w v p p^T,
    // Sq = sum w (u^2+v^2) p p^T  (gram_entry), so a thread carries 23 accumulators instead of 36 + 8, and the warp
    // total is taken with a recursive-halving reduce-scatter (31 shuffles for all 23 values; lane l ends up owning
    // moment l) instead of 44 five-step butterflies: the shuffle unit was what this single-CTA kernel waited on.
    const float c_sup = 2.25f * thr2;
    for (int it = 0; it < iterations; it++)
    {
        float m[9];
        for (int j = 0; j < 9; j++) m[j] = s_m[j];
        double acc[32];
#pragma unroll
        for (int j = 0; j < 32; j++) acc[j] = 0.0;
        for (int i = tid; i < n; i += RT)
        {
            const float2 p = src[i], d = dst[i];
            const float e = reproj_err2(m, p, d);
            if (!(e == e) || e >= c_sup) continue;
            const float t = 1.0f - e / c_sup;
            const double w = (double)(t * t);
            const double x = (p.x - cx) * s1, y = (p.y - cy) * s1, u = (d.x - cu) * s2, v = (d.y - cv) * s2;
            const double wx = w * x, wy = w * y;
            const double pp[6] = {wx * x, wx * y, wx, wy * y, wy, w};  // w * p p^T, packed xx xy x yy y 1
            const double q = u * u + v * v;
#pragma unroll
            for (int j = 0; j < 6; j++)
            {
                acc[j] += pp[j];
                acc[6 + j] += u * pp[j];
                acc[12 + j] += v * pp[j];
                if (j < 5) acc[18 + j] += q * pp[j];
            }
        }
        // reduce-scatter over the warp: after the step with `half`, slot i holds moment i + (lane & ~(half - 1)) % 32
#pragma unroll
        for (int half = 16; half >= 1; half >>= 1)
        {
            const bool upper = (lane & half) != 0;
#pragma unroll
            for (int i = 0; i < half; i++)
            {
                const double keep = upper ? acc[i + half] : acc[i];
                const double send = upper ? acc[i] : acc[i + half];
                acc[i] = keep + warp.shfl_xor(send, half);
            }
        }
        s_part[wid][lane] = acc[0];  // this warp's total of moment `lane`
        block.sync();
        if (tid < 32)
        {
            double s = 0;
            for (int w = 0; w < RT / 32; w++) s += s_part[w][tid];
            s_mom[tid] = s;
        }
        block.sync();
        if (s_mom[5] < 4.0) break;  // sum of weights: support collapsed, keep the current model

        // warp 0: Gauss-Jordan on the 8x9 augmented SPD system, lane r owns row r
        if (wid == 0)
        {
            double row[9];
            const int r = lane & 7;
#pragma unroll
            for (int c = 0; c < 9; c++) row[c] = gram_entry(s_mom, r, c);
            bool ok = true;
            for (int c = 0; c < 8; c++)
            {
                double prow[9];
                for (int k = 0; k < 9; k++) prow[k] = warp.shfl(row[k], c);
                if (fabs(prow[c]) < 1e-14) { ok = false; break; }
                const double inv = 1.0 / prow[c];
                if (r == c) { for (int k = 0; k < 9; k++) row[k] = prow[k] * inv; }
                else { const double f = row[c] * inv; for (int k = 0; k < 9; k++) row[k] -= f * prow[k]; }
            }
            double h[8];
            for (int k = 0; k < 8; k++) h[k] = warp.shfl(row[8], k);
            if (lane == 0 && ok)
            {
                // denormalise: H = T2^-1 * Hn * T1,  T = [s 0 -s*c; 0 s -s*c; 0 0 1]
                const double Hn[9] = {h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], 1.0};
                double A[9];  // Hn * T1
                for (int rr = 0; rr < 3; rr++)
                {
                    A[rr * 3 + 0] = Hn[rr * 3 + 0] * s1;
                    A[rr * 3 + 1] = Hn[rr * 3 + 1] * s1;
                    A[rr * 3 + 2] = Hn[rr * 3 + 2] - Hn[rr * 3 + 0] * s1 * cx - Hn[rr * 3 + 1] * s1 * cy;
                }
                double Hd[9];  // T2^-1 = [1/s 0 c; 0 1/s c; 0 0 1]
                for (int k = 0; k < 3; k++)
                {
                    Hd[0 + k] = A[0 + k] / s2 + cu * A[6 + k];
                    Hd[3 + k] = A[3 + k] / s2 + cv * A[6 + k];
                    Hd[6 + k] = A[6 + k];
                }
                if (fabs(Hd[8]) > 1e-12)
                {
                    const double invh = 1.0 / Hd[8];
                    for (int k = 0; k < 9; k++) { result->h[k] = Hd[k] * invh; s_m[k] = (float)(Hd[k] * invh); }
                }
            }
        }
        block.sync();
    }

    // ---- result + final mask with the returned model
    float m[9];
    for (int j = 0; j < 9; j++) m[j] = s_m[j];
    int inl = 0;
    for (int i = tid; i < n; i += RT)
    {
        const float e = reproj_err2(m, src[i], dst[i]);
        const uint8_t in = (e < thr2) ? 1 : 0;
        mask[i] = in;
        inl += in;
    }
    inl = cg::reduce(warp, inl, cg::plus<int>());
    if (lane == 0) s_besti[wid] = inl;
    block.sync();
    if (tid == 0)
    {
        int total = 0;
        for (int w = 0; w < RT / 32; w++) total += s_besti[w];
        result->inliers = total;
        result->found = 1;
        if (iterations == 0 || result->h[8] == 0.0)
            for (int k = 0; k < 9; k++) result->h[k] = (double)s_m[k];
    }
}

static_assert(sizeof(RansacResult) % 4 == 0, "RansacResult is moved as 32-bit words");

__device__ __forceinline__ void copy16(const uint8_t* src, uint8_t* dst, int bytes)
{
    for (int i = threadIdx.x; i < (bytes + 15) / 16; i += RT)
        reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
}

// The estimator's last kernel also delivers the chain's results: it copies the used part of the result block (LK
// matches + status of the prm->n tracked points, inlier mask, model) into MAPPED PINNED HOST memory with 16-byte
// posted PCIe writes, so no transfer step follows on the frame's critical path and no copy engine is involved.
__global__ void __launch_bounds__(RT)
    k_ransac_refine(const float2* __restrict__ src, const float2* __restrict__ dst, const int* __restrict__ n_ptr,
                    const float* __restrict__ models, const float* __restrict__ scores,
                    const TrackParams* __restrict__ prm, int iterations, RansacResult* __restrict__ result,
                    uint8_t* __restrict__ mask, TrackOutCopy out)
{
    constexpr int RESULT_WORDS = sizeof(RansacResult) / 4;
    if (threadIdx.x < RESULT_WORDS) reinterpret_cast<uint32_t*>(result)[threadIdx.x] = 0u;
    __syncthreads();
    refine_body(src, dst, n_ptr, models, scores, prm, iterations, result, mask);
    if (!out.host) return;
    __syncthreads();  // the block's global writes (mask, result) are visible to all of its threads
    const int tracked = prm->n, estimated = *n_ptr;
    copy16(out.dev, out.host, tracked * (int)sizeof(float2));
    copy16(out.dev + out.off_status, out.host + out.off_status, tracked);
    copy16(out.dev + out.off_mask, out.host + out.off_mask, estimated);
    copy16(out.dev + out.off_result, out.host + out.off_result, (int)sizeof(RansacResult));
}

// Local-motion mode has no estimator kernel: one small CTA delivers the LK results (same 16-byte posted writes).
// Per-CTA result writes from the 200-600 LK CTAs themselves cost ~0.4 us EACH on the PCIe write path (LK: 21 -> 217 us).
__global__ void __launch_bounds__(RT) k_track_out_copy(const TrackParams* __restrict__ prm, TrackOutCopy out)
{
    const int tracked = prm->n;
    copy16(out.dev, out.host, tracked * (int)sizeof(float2));
    copy16(out.dev + out.off_status, out.host + out.off_status, tracked);
}

}  // namespace

lvkb200_status track_out_copy(cudaStream_t cs, const TrackParams* d_params, const TrackOutCopy& out)
{
    LVKB_REQUIRE(out.dev != nullptr && out.host != nullptr);
    k_track_out_copy<<<1, RT, 0, cs>>>(d_params, out);
    count_launches(1);
    LVKB_CUDA(cudaGetLastError());
    return LVKB200_OK;
}

lvkb200_status compact_swap_erase(cudaStream_t cs, const float2* d_a, const float2* d_b, const uint8_t* d_keep,
                                  const TrackParams* d_params, float2* d_a_out, float2* d_b_out, int* d_perm,
                                  int* d_removed, int* d_n_out)
{
    k_compact_swap_erase<<<1, CT, 0, cs>>>(d_a, d_b, d_keep, d_params, d_a_out, d_b_out, d_perm, d_removed, d_n_out);
    count_launches(1);
    LVKB_CUDA(cudaGetLastError());
    return LVKB200_OK;
}

lvkb200_status ransac_homography(cudaStream_t cs, const float2* d_src, const float2* d_dst, const int* d_n,
                                 const TrackParams* d_params, float* d_models, float* d_scores,
                                 RansacResult* d_result, uint8_t* d_mask, const TrackOutCopy& out)
{
    k_ransac_hypotheses<<<div_up(HYP, 128), 128, 0, cs>>>(d_src, d_dst, d_n, d_params, 0x9E3779B9u, d_models);
    k_ransac_score<<<HYP, 256, 0, cs>>>(d_src, d_dst, d_n, d_models, d_params, d_scores);
    k_ransac_refine<<<1, RT, 0, cs>>>(d_src, d_dst, d_n, d_models, d_scores, d_params, RANSAC_REFINE_ITERS, d_result,
                                      d_mask, out);
    count_launches(3);
    LVKB_CUDA(cudaGetLastError());
    return LVKB200_OK;
}

}  // namespace lvkb200
