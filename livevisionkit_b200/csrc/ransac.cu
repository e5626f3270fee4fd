// K6a — robust homography estimation (batched RANSAC hypothesis scoring + sigma-weighted refit) for sm_100a.
//
// Replaces cv::findHomography(tracked, matched, mask, cv::UsacParams{MAGSAC, LO_SIGMA, 50 iterations, conf .99}) as
// called by FrameTracker::estimate_global_motion (LiveVisionKit/Vision/FrameTracker.cpp:337-359).
// OpenCV's USAC lives in calib3d/usac (third-party, not under /root/reference, no source in this image); its
// sampling stream and MAGSAC weighting tables cannot be reproduced bit-for-bit, so this is a GPU-native estimator
// with the same CONTRACT:
//   * deterministic (fixed hypothesis stream, like randomGeneratorState = 0);
//   * output H normalised to h33 = 1, double precision;
//   * mask[i] = 1  <=>  float32 forward reprojection error^2 of the RETURNED H < threshold^2 (exactly how cv2's
//     mask relates to its H);
//   * degenerate input (e.g. collinear points) -> no model.
// Parity vs cv2 is therefore: identical masks except points within epsilon of the threshold, H within the
// estimator's own input-order variance (stated and measured in tests/test_ransac_gpu.py).
//
// Kernels (one dependent chain, no host round trip in between):
//   k_ransac_hypotheses : 256 minimal 4-point models, one thread each (8x8 solve in double, degeneracy tests)
//   k_ransac_score      : one CTA per hypothesis, all points scored with a cooperative-groups block reduction
//                         (truncated-quadratic / MSAC cost at the acceptance threshold)
//   k_ransac_refine     : single CTA: arg-min hypothesis, then 5 IRLS passes of a normalised weighted DLT
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

__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

__device__ __forceinline__ float cross2(float2 a, float2 b, float2 c)
{
    return (b.x - a.x) * (c.y - a.y) - (b.y - a.y) * (c.x - a.x);
}

// Gaussian elimination with partial pivoting, n = 8, in double.  Returns false when singular.
__device__ bool solve8(double A[8][9])
{
    for (int c = 0; c < 8; c++)
    {
        int piv = c;
        double best = fabs(A[c][c]);
        for (int r = c + 1; r < 8; r++)
            if (fabs(A[r][c]) > best) { best = fabs(A[r][c]); piv = r; }
        if (best < 1e-12) return false;
        if (piv != c)
            for (int k = 0; k < 9; k++) { const double t = A[c][k]; A[c][k] = A[piv][k]; A[piv][k] = t; }
        const double inv = 1.0 / A[c][c];
        for (int r = 0; r < 8; r++)
        {
            if (r == c) continue;
            const double f = A[r][c] * inv;
            for (int k = c; k < 9; k++) A[r][k] -= f * A[c][k];
        }
    }
    for (int r = 0; r < 8; r++) A[r][8] /= A[r][r];
    return true;
}

__global__ void __launch_bounds__(128)
    k_ransac_hypotheses(const float2* __restrict__ src, const float2* __restrict__ dst, int n, uint32_t seed,
                        float* __restrict__ models)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= HYP) return;
    float* out = models + (size_t)k * 9;
    out[8] = 0.0f;  // invalid until proven otherwise

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
    for (int j = 0; j < 4; j++) { p[j] = src[idx[j]]; q[j] = dst[idx[j]]; }

    // degeneracy: no three (nearly) collinear points, and the sample must keep its orientation
    const int tri[4][3] = {{0, 1, 2}, {0, 1, 3}, {0, 2, 3}, {1, 2, 3}};
    for (int t = 0; t < 4; t++)
    {
        const float cs = cross2(p[tri[t][0]], p[tri[t][1]], p[tri[t][2]]);
        const float cd = cross2(q[tri[t][0]], q[tri[t][1]], q[tri[t][2]]);
        if (fabsf(cs) < 4.0f || fabsf(cd) < 4.0f) return;  // triangle area < 2 px^2
        if ((cs > 0.f) != (cd > 0.f)) return;
    }

    double A[8][9];
    for (int j = 0; j < 4; j++)
    {
        const double x = p[j].x, y = p[j].y, u = q[j].x, v = q[j].y;
        double* r0 = A[2 * j];
        double* r1 = A[2 * j + 1];
        r0[0] = x; r0[1] = y; r0[2] = 1; r0[3] = 0; r0[4] = 0; r0[5] = 0; r0[6] = -u * x; r0[7] = -u * y; r0[8] = u;
        r1[0] = 0; r1[1] = 0; r1[2] = 0; r1[3] = x; r1[4] = y; r1[5] = 1; r1[6] = -v * x; r1[7] = -v * y; r1[8] = v;
    }
    if (!solve8(A)) return;
    for (int j = 0; j < 8; j++) out[j] = (float)A[j][8];
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
    k_ransac_score(const float2* __restrict__ src, const float2* __restrict__ dst, int n, const float* __restrict__ models,
                   float thr2, float* __restrict__ scores)
{
    cg::thread_block block = cg::this_thread_block();
    cg::thread_block_tile<32> warp = cg::tiled_partition<32>(block);
    __shared__ float m[9];
    __shared__ float partial[8];
    if (threadIdx.x < 9) m[threadIdx.x] = models[(size_t)blockIdx.x * 9 + threadIdx.x];
    block.sync();
    if (m[8] == 0.0f)
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

constexpr int RT = 256;   // refine CTA size (44 double accumulators per thread: keep the register budget)
constexpr int WC = 8;     // cached weights per thread (n <= RT*WC points never recompute)
constexpr int NACC = 44;  // 36 (upper triangle of A^T W A) + 8 (A^T W b)

__global__ void __launch_bounds__(RT)
    k_ransac_refine(const float2* __restrict__ src, const float2* __restrict__ dst, int n, const float* __restrict__ models,
                    const float* __restrict__ scores, float thr2, int iterations, RansacResult* __restrict__ result,
                    uint8_t* __restrict__ mask)
{
    cg::thread_block block = cg::this_thread_block();
    cg::thread_block_tile<32> warp = cg::tiled_partition<32>(block);
    __shared__ float s_best[RT / 32];
    __shared__ int s_besti[RT / 32];
    __shared__ double s_acc[RT / 32][NACC];
    __shared__ double s_norm[RT / 32][6];
    __shared__ double s_T[10];  // src: cx, cy, s ; dst: cx, cy, s ; total weight ...
    __shared__ float s_m[9];
    __shared__ int s_ok;
    const int tid = threadIdx.x, lane = warp.thread_rank(), wid = warp.meta_group_rank();

    // ---- arg-min over hypotheses (lowest index wins ties)
    float best = 3.0e38f;
    int besti = -1;
    for (int k = tid; k < HYP; k += RT)
        if (scores[k] < best) { best = scores[k]; besti = k; }
    for (int o = 16; o > 0; o >>= 1)
    {
        const float ob = warp.shfl_xor(best, o);
        const int oi = warp.shfl_xor(besti, o);
        if (ob < best || (ob == best && oi >= 0 && (besti < 0 || oi < besti))) { best = ob; besti = oi; }
    }
    if (lane == 0) { s_best[wid] = best; s_besti[wid] = besti; }
    block.sync();
    if (tid == 0)
    {
        float b = 3.0e38f;
        int bi = -1;
        for (int w = 0; w < RT / 32; w++)
            if (s_besti[w] >= 0 && (s_best[w] < b || (s_best[w] == b && s_besti[w] < bi))) { b = s_best[w]; bi = s_besti[w]; }
        s_ok = (bi >= 0 && b < 2.9e38f) ? 1 : 0;
        if (s_ok)
            for (int j = 0; j < 9; j++) s_m[j] = models[(size_t)bi * 9 + j];
    }
    block.sync();
    if (!s_ok)
    {
        if (tid == 0) result->found = 0;
        for (int i = tid; i < n; i += RT) mask[i] = 0;
        return;
    }

    // ---- IRLS: weighted, Hartley-normalised DLT with h33 = 1 in normalised coordinates
    // weights: sigma-consensus style, smooth and compactly supported: w = (1 - e/c)^2 for e < c, c = 2.25 * thr^2
    // (support 1.5x the acceptance radius), so points just outside the threshold still pull a little, far ones not at all.
    const float c_sup = 2.25f * thr2;
    for (int it = 0; it < iterations; it++)
    {
        float m[9];
        for (int j = 0; j < 9; j++) m[j] = s_m[j];

        // pass 1: weights + weighted centroids / scales
        double nsum[6] = {0, 0, 0, 0, 0, 0};  // w, w*x, w*y, w*u, w*v, (unused)
        float wloc[WC];
        for (int q = 0; q < WC; q++) wloc[q] = 0.f;
        for (int q = 0, i = tid; i < n; i += RT, q++)
        {
            const float2 p = src[i], d = dst[i];
            const float e = reproj_err2(m, p, d);
            float w = 0.0f;
            if (e == e && e < c_sup) { const float t = 1.0f - e / c_sup; w = t * t; }
            if (q < WC) wloc[q] = w;
            nsum[0] += w; nsum[1] += (double)w * p.x; nsum[2] += (double)w * p.y;
            nsum[3] += (double)w * d.x; nsum[4] += (double)w * d.y;
        }
        for (int j = 0; j < 5; j++) nsum[j] = cg::reduce(warp, nsum[j], cg::plus<double>());
        if (lane == 0) for (int j = 0; j < 5; j++) s_norm[wid][j] = nsum[j];
        block.sync();
        if (tid < 5)
        {
            double s = 0;
            for (int w = 0; w < RT / 32; w++) s += s_norm[w][tid];
            s_T[tid] = s;
        }
        block.sync();
        const double wsum = s_T[0];
        if (wsum < 4.0) break;  // support collapsed: keep the current model
        const double cx = s_T[1] / wsum, cy = s_T[2] / wsum, cu = s_T[3] / wsum, cv = s_T[4] / wsum;
        block.sync();
        // mean distances
        double dsum[2] = {0, 0};
        for (int q = 0, i = tid; i < n; i += RT, q++)
        {
            const float2 p = src[i], d = dst[i];
            float w;
            if (q < WC) w = wloc[q];
            else { const float e = reproj_err2(m, p, d); w = 0.f; if (e == e && e < c_sup) { const float t = 1.0f - e / c_sup; w = t * t; } }
            dsum[0] += (double)w * sqrt((p.x - cx) * (p.x - cx) + (p.y - cy) * (p.y - cy));
            dsum[1] += (double)w * sqrt((d.x - cu) * (d.x - cu) + (d.y - cv) * (d.y - cv));
        }
        for (int j = 0; j < 2; j++) dsum[j] = cg::reduce(warp, dsum[j], cg::plus<double>());
        if (lane == 0) { s_norm[wid][0] = dsum[0]; s_norm[wid][1] = dsum[1]; }
        block.sync();
        if (tid < 2)
        {
            double s = 0;
            for (int w = 0; w < RT / 32; w++) s += s_norm[w][tid];
            s_T[5 + tid] = s;
        }
        block.sync();
        const double s1 = (s_T[5] > 1e-9) ? 1.4142135623730951 * wsum / s_T[5] : 1.0;
        const double s2 = (s_T[6] > 1e-9) ? 1.4142135623730951 * wsum / s_T[6] : 1.0;

        // pass 2: normal equations in normalised coordinates
        double acc[NACC];
        for (int j = 0; j < NACC; j++) acc[j] = 0.0;
        for (int q = 0, i = tid; i < n; i += RT, q++)
        {
            const float2 p = src[i], d = dst[i];
            float w;
            if (q < WC) w = wloc[q];
            else { const float e = reproj_err2(m, p, d); w = 0.f; if (e == e && e < c_sup) { const float t = 1.0f - e / c_sup; w = t * t; } }
            if (w == 0.0f) continue;
            const double x = (p.x - cx) * s1, y = (p.y - cy) * s1, u = (d.x - cu) * s2, v = (d.y - cv) * s2;
            const double r0[8] = {x, y, 1, 0, 0, 0, -u * x, -u * y};
            const double r1[8] = {0, 0, 0, x, y, 1, -v * x, -v * y};
            int t = 0;
            for (int a = 0; a < 8; a++)
                for (int b = a; b < 8; b++) acc[t++] += w * (r0[a] * r0[b] + r1[a] * r1[b]);
            for (int a = 0; a < 8; a++) acc[36 + a] += w * (r0[a] * u + r1[a] * v);
        }
        for (int j = 0; j < NACC; j++) acc[j] = cg::reduce(warp, acc[j], cg::plus<double>());
        if (lane == 0) for (int j = 0; j < NACC; j++) s_acc[wid][j] = acc[j];
        block.sync();
        if (tid < NACC)
        {
            double s = 0;
            for (int w = 0; w < RT / 32; w++) s += s_acc[w][tid];
            s_acc[0][tid] = s;
        }
        block.sync();

        // warp 0: Gauss-Jordan on the 8x9 augmented SPD system, lane r owns row r
        if (wid == 0)
        {
            double row[9];
            const int r = lane & 7;
            {
                int t = 0;
                double full[8][8];
                for (int a = 0; a < 8; a++)
                    for (int b = a; b < 8; b++) { full[a][b] = s_acc[0][t]; full[b][a] = s_acc[0][t]; t++; }
                for (int k = 0; k < 8; k++) row[k] = full[r][k];
                row[8] = s_acc[0][36 + r];
            }
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

}  // namespace

lvkb200_status ransac_homography(cudaStream_t cs, const float2* d_src, const float2* d_dst, int n, float threshold,
                                 float* d_models, float* d_scores, RansacResult* d_result, uint8_t* d_mask)
{
    LVKB_REQUIRE(n >= 4);
    const float thr2 = threshold * threshold;
    LVKB_CUDA(cudaMemsetAsync(d_result, 0, sizeof(RansacResult), cs));
    k_ransac_hypotheses<<<div_up(HYP, 128), 128, 0, cs>>>(d_src, d_dst, n, 0x9E3779B9u, d_models);
    k_ransac_score<<<HYP, 256, 0, cs>>>(d_src, d_dst, n, d_models, thr2, d_scores);
    k_ransac_refine<<<1, RT, 0, cs>>>(d_src, d_dst, n, d_models, d_scores, thr2, RANSAC_REFINE_ITERS, d_result, d_mask);
    count_launches(3);
    LVKB_CUDA(cudaGetLastError());
    return LVKB200_OK;
}

}  // namespace lvkb200
