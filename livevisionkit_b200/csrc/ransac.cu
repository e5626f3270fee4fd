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
//   k_ransac_score      : one CTA per hypothesis: its minimal 4-point model in closed form (unit square -> quad, no
//                         solve; one thread), then all points scored with a cooperative-groups block reduction
//                         (truncated-quadratic / MSAC cost at the acceptance threshold) and an atomicMin of the
//                         packed (cost, index) key = the arg-min over the hypotheses
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

// Minimal model number k of the fixed hypothesis stream -> out[0..8] (out[8] == 0: degenerate sample, no model).
__device__ __forceinline__ void make_hypothesis(int k, int n, const float2* __restrict__ src, const float2* __restrict__ dst,
                                                const TrackParams* __restrict__ prm, uint32_t seed, float* out)
{
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
                   float* __restrict__ models, const TrackParams* __restrict__ prm, float* __restrict__ scores,
                   unsigned long long* __restrict__ best_key, uint32_t seed)
{
    const float thr2 = prm->threshold_sq;
    cg::thread_block block = cg::this_thread_block();
    cg::thread_block_tile<32> warp = cg::tiled_partition<32>(block);
    __shared__ float m[9];
    __shared__ float partial[8];
    const int n = *n_ptr;
    // the CTA's own hypothesis first (one thread, ~1 us of dependent FP64): a separate 256-thread kernel for the 256
    // closed-form models cost a launch and a drain on the frame's critical path; the other threads pull the first
    // correspondences they will score towards L1 meanwhile
    if (threadIdx.x == 0)
    {
        float h[9];
        make_hypothesis((int)blockIdx.x, n, src, dst, prm, seed, h);
        for (int j = 0; j < 9; j++) { m[j] = h[j]; models[(size_t)blockIdx.x * 9 + j] = h[j]; }
    }
    else if ((int)threadIdx.x < n)
    {
        asm volatile("prefetch.global.L1 [%0];" ::"l"(src + threadIdx.x));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(dst + threadIdx.x));
    }
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
        // arg-min over the hypotheses, lowest index wins ties: costs are >= 0, so their bit patterns order like the values
        if (s < 2.9e38f) atomicMin(best_key, ((unsigned long long)__float_as_uint(s) << 32) | (unsigned)blockIdx.x);
    }
}

constexpr int RT = 512;  // refine CTA size
constexpr int PPT = 6;   // correspondences a thread keeps in registers (RT * PPT = 3072 >= the tracker's point capacity)

// Entry (r, c) of the 8 x 9 augmented normal equations [A^T W A | A^T W b] of the weighted DLT (h33 = 1) is +- one of the
// 23 moments S[0..5] = S0, S[6..11] = Su, S[12..17] = Sv, S[18..22] = Sq, each packed xx xy x yy y 1 (see refine_body):
// -> index into S (-1: structural zero) and sign.
__device__ __forceinline__ void gram_index(int r, int c, int& idx, bool& neg)
{
    auto pack = [](int i, int j) { const int a = min(i, j), b = max(i, j); return a * (5 - a) / 2 + b; };  // symmetric 3x3
    const int rb = r / 3, ri = r - 3 * rb;
    neg = false;
    if (c == 8)
    {
        idx = (rb == 0 ? 6 : rb == 1 ? 12 : 18) + pack(ri, 2);
        neg = rb == 2;
        return;
    }
    const int cb = c / 3, ci = c - 3 * cb;
    if (rb == cb) { idx = (rb < 2 ? 0 : 18) + pack(ri, ci); return; }
    if (rb + cb == 1) { idx = -1; return; }
    const int other = min(rb, cb), i = (rb < cb) ? ri : ci, j = (rb < cb) ? ci : ri;  // i: index in p, j: block-2 index
    idx = (other == 0 ? 6 : 12) + pack(i, j);
    neg = true;
}

// 1 / d to full double precision without the IEEE division's range tests and slow path (d is a pivot / scale here):
// hardware seed (>= 20 bits) + two Newton steps.
__device__ __forceinline__ double fast_rcp(double d)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    r = fma(r, fma(-d, r, 1.0), r);
    r = fma(r, fma(-d, r, 1.0), r);
    return r;
}

// Block total of up to 32 per-thread double accumulators -> s_mom[0..31].  Within a warp a recursive-halving
// reduce-scatter (31 shuffles for all 32 values; lane l ends up owning slot l) instead of 32 five-step butterflies, then
// one lane per slot adds the warps' totals.  Two barriers.
__device__ __forceinline__ void block_sum32(double (&acc)[32], double (*s_part)[32], double* s_mom)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1)
    {
        const bool upper = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; i++)
        {
            const double keep = upper ? acc[i + half] : acc[i];
            const double send = upper ? acc[i] : acc[i + half];
            acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    s_part[wid][lane] = acc[0];  // this warp's total of slot `lane`
    __syncthreads();
    if (threadIdx.x < 32)
    {
        double s = 0;
#pragma unroll
        for (int w = 0; w < RT / 32; w++) s += s_part[w][threadIdx.x];
        s_mom[threadIdx.x] = s;
    }
    __syncthreads();
}

__device__ __forceinline__ void
    refine_body(const float2* __restrict__ src, const float2* __restrict__ dst, const int* __restrict__ n_ptr,
                const float* __restrict__ models, const float* __restrict__ scores,
                const TrackParams* __restrict__ prm, int iterations, RansacResult* __restrict__ result,
                uint8_t* __restrict__ mask)
{
    const float thr2 = prm->threshold_sq;
    __shared__ double s_part[RT / 32][32];
    __shared__ double s_mom[32];
    __shared__ float s_m[9];
    __shared__ int s_ok, s_count;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n = *n_ptr;

    // ---- the correspondences live in registers for the whole kernel (every pass below would re-read them from L2)
    float2 ps[PPT], pd[PPT];
#pragma unroll
    for (int k = 0; k < PPT; k++)
    {
        const int i = tid + k * RT;
        ps[k] = i < n ? src[i] : make_float2(0.0f, 0.0f);
        pd[k] = i < n ? dst[i] : make_float2(0.0f, 0.0f);
    }
    // BODY sees the index i and the correspondence (p, d); points beyond the register tile come from global memory
#define LVKB_FOR_POINTS(...)                                                                                           \
    {                                                                                                                 \
        _Pragma("unroll") for (int k_ = 0; k_ < PPT; k_++)                                                            \
        {                                                                                                             \
            const int i = tid + k_ * RT;                                                                              \
            if (i < n) { const float2 p = ps[k_], d = pd[k_]; __VA_ARGS__ }                                                  \
        }                                                                                                             \
        for (int i = tid + PPT * RT; i < n; i += RT) { const float2 p = src[i], d = dst[i]; __VA_ARGS__ }                    \
    }

    // ---- best hypothesis: the scoring pass left the arg-min as a packed (cost bits, index) key behind the scores
    if (tid == 0)
    {
        unsigned long long* const key_ptr = reinterpret_cast<unsigned long long*>(const_cast<float*>(scores + HYP));
        const unsigned long long key = *key_ptr;
        *key_ptr = ~0ull;  // armed for the next scoring pass (the buffer is created in this state)
        const int bi = (int)(unsigned)(key & 0xffffffffull);
        s_ok = (n >= 4 && key != ~0ull && bi < HYP) ? 1 : 0;
        if (s_ok)
            for (int j = 0; j < 9; j++) s_m[j] = models[(size_t)bi * 9 + j];
        s_count = 0;
        result->n = n;
    }
    __syncthreads();
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
        double acc[32];
#pragma unroll
        for (int j = 0; j < 32; j++) acc[j] = 0.0;
        LVKB_FOR_POINTS({
            const bool in = reproj_err2(m, p, d) < thr2;
            mask[i] = in ? 1 : 0;
            if (in) { acc[0] += 1.0; acc[1] += p.x; acc[2] += p.y; acc[3] += d.x; acc[4] += d.y; }
        })
        block_sum32(acc, s_part, s_mom);
        const double cnt = s_mom[0];
        const double cx = s_mom[1] / fmax(cnt, 1.0), cy = s_mom[2] / fmax(cnt, 1.0);
        const double cu = s_mom[3] / fmax(cnt, 1.0), cv = s_mom[4] / fmax(cnt, 1.0);
        __syncthreads();  // s_mom is reused below
#pragma unroll
        for (int j = 0; j < 32; j++) acc[j] = 0.0;
        LVKB_FOR_POINTS({
            if (reproj_err2(m, p, d) < thr2)
            {
                const double x = p.x - cx, y = p.y - cy, u = d.x - cu, v = d.y - cv;
                acc[0] += x * x + y * y; acc[1] += x * u + y * v; acc[2] += x * v - y * u;
            }
        })
        block_sum32(acc, s_part, s_mom);
        if (tid == 0)
        {
            double a = m[0], b = m[3], tx = m[2], ty = m[5];
            if (cnt >= 2.0 && s_mom[0] > 1e-9)
            {
                a = s_mom[1] / s_mom[0]; b = s_mom[2] / s_mom[0];
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

    // ---- isotropic normalisation of both point sets from ONE pass (it only conditions the normal equations): centroid
    // to the origin, RMS distance to sqrt(2)
    double cx, cy, cu, cv, s1, s2;
    {
        double acc[32];
#pragma unroll
        for (int j = 0; j < 32; j++) acc[j] = 0.0;
        LVKB_FOR_POINTS({
            (void)i;
            acc[0] += p.x; acc[1] += p.y; acc[2] += d.x; acc[3] += d.y;
            acc[4] += (double)p.x * p.x + (double)p.y * p.y;
            acc[5] += (double)d.x * d.x + (double)d.y * d.y;
        })
        block_sum32(acc, s_part, s_mom);
        const double inv_n = 1.0 / (double)n;
        cx = s_mom[0] * inv_n; cy = s_mom[1] * inv_n; cu = s_mom[2] * inv_n; cv = s_mom[3] * inv_n;
        const double v1 = s_mom[4] * inv_n - (cx * cx + cy * cy), v2 = s_mom[5] * inv_n - (cu * cu + cv * cv);
        s1 = v1 > 1e-12 ? sqrt(2.0 / v1) : 1.0;
        s2 = v2 > 1e-12 ? sqrt(2.0 / v2) : 1.0;
        __syncthreads();  // s_mom is reused by the first IRLS pass
    }

    // ---- IRLS: weighted DLT with h33 = 1 in normalised coordinates.
    // weights: sigma-consensus style, smooth and compactly supported: w = (1 - e/c)^2 for e < c, c = 2.25 * thr^2
    // (support 1.5x the acceptance radius), so points just outside the threshold still pull a little, far ones not at all.
    //
    // The normal equations A^T W A h = A^T W b of the rows [p 0 -u p | u], [0 p -v p | v] (p = (x, y, 1)) have block
    // structure: every entry is +-one of the 23 moments  S0 = sum w p p^T, Su = sum w u p p^T, Sv = sum w v p p^T,
    // Sq = sum w (u^2+v^2) p p^T  (gram_entry), so a thread carries 23 accumulators instead of 36 + 8.
    const float c_sup = 2.25f * thr2;
    const float inv_c_sup = 1.0f / c_sup;
    // Gram row of lane (lane & 7): moment index (-1: structural zero) and sign of entry (r, c), see gram_entry
    int gidx[9];
    bool gneg[9];
    {
        const int r = lane & 7;
#pragma unroll
        for (int c = 0; c < 9; c++) gram_index(r, c, gidx[c], gneg[c]);
    }
    for (int it = 0; it < iterations; it++)
    {
        float m[9];
        for (int j = 0; j < 9; j++) m[j] = s_m[j];
        double acc[32];
#pragma unroll
        for (int j = 0; j < 32; j++) acc[j] = 0.0;
        LVKB_FOR_POINTS({
            (void)i;
            const float e = reproj_err2(m, p, d);
            if (e == e && e < c_sup)
            {
                const float t = 1.0f - e * inv_c_sup;
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
        })
        block_sum32(acc, s_part, s_mom);
        if (s_mom[5] < 4.0) break;  // sum of weights: support collapsed, keep the current model

        // warp 0: Gauss-Jordan on the 8x9 augmented SPD system, lane r owns row r.  The other 15 warps wait at the
        // barrier below for this serial stretch, so it is kept short: table-free Gram assembly (index + sign per entry
        // computed once, before the loop), reciprocal pivots without the IEEE division, and the denormalisation spread
        // over nine lanes (one output coefficient each).
        if (wid == 0)
        {
            double row[9];
            const int r = lane & 7;
#pragma unroll
            for (int c = 0; c < 9; c++) row[c] = gidx[c] < 0 ? 0.0 : (gneg[c] ? -s_mom[gidx[c]] : s_mom[gidx[c]]);
            bool ok = true;
#pragma unroll
            for (int c = 0; c < 8; c++)
            {
                double prow[9];
#pragma unroll
                for (int k = 0; k < 9; k++) prow[k] = __shfl_sync(0xffffffffu, row[k], c);
                ok = ok && fabs(prow[c]) >= 1e-14;
                const double inv = fast_rcp(prow[c]);
                if (r == c)
                {
#pragma unroll
                    for (int k = 0; k < 9; k++) row[k] = prow[k] * inv;
                }
                else
                {
                    const double f = row[c] * inv;
#pragma unroll
                    for (int k = 0; k < 9; k++) row[k] = fma(-f, prow[k], row[k]);
                }
            }
            double h[9];
#pragma unroll
            for (int k = 0; k < 8; k++) h[k] = __shfl_sync(0xffffffffu, row[8], k);
            h[8] = 1.0;
            // denormalise: H = T2^-1 * Hn * T1,  T = [s 0 -s*c; 0 s -s*c; 0 0 1]; lane k < 9 computes coefficient k
            const int kk = lane < 9 ? lane : 0, orow = kk / 3, ocol = kk - 3 * orow;
            // A = Hn * T1: column ocol of rows `orow` and 2
            auto a_entry = [&](int rr) {
                const double h0 = h[rr * 3], h1 = h[rr * 3 + 1], h2 = h[rr * 3 + 2];
                return ocol == 0 ? h0 * s1 : ocol == 1 ? h1 * s1 : h2 - (h0 * cx + h1 * cy) * s1;
            };
            const double a2 = a_entry(2);
            double hd;  // T2^-1 = [1/s 0 c; 0 1/s c; 0 0 1]
            if (orow == 2) hd = a2;
            else
            {
                const double ar = orow == 0 ? a_entry(0) : a_entry(1);
                hd = ar * fast_rcp(s2) + (orow == 0 ? cu : cv) * a2;
            }
            const double hd8 = __shfl_sync(0xffffffffu, hd, 8);
            ok = ok && fabs(hd8) > 1e-12 && hd8 == hd8;
            if (ok && lane < 9)
            {
                const double v = hd * fast_rcp(hd8);
                result->h[lane] = v;
                s_m[lane] = (float)v;
            }
        }
        __syncthreads();
    }

    // ---- result + final mask with the returned model
    float m[9];
    for (int j = 0; j < 9; j++) m[j] = s_m[j];
    int inl = 0;
    LVKB_FOR_POINTS({
        const uint8_t in = (reproj_err2(m, p, d) < thr2) ? 1 : 0;
        mask[i] = in;
        inl += in;
    })
    inl = __reduce_add_sync(0xffffffffu, inl);
    if (lane == 0 && inl) atomicAdd(&s_count, inl);
    __syncthreads();
    if (tid == 0)
    {
        result->inliers = s_count;
        result->found = 1;
        if (iterations == 0 || result->h[8] == 0.0)
            for (int k = 0; k < 9; k++) result->h[k] = (double)s_m[k];
    }
#undef LVKB_FOR_POINTS
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
    unsigned long long* best_key = reinterpret_cast<unsigned long long*>(d_scores + HYP);  // behind the scores, 8-byte aligned
    k_ransac_score<<<HYP, 256, 0, cs>>>(d_src, d_dst, d_n, d_models, d_params, d_scores, best_key, 0x9E3779B9u);
    k_ransac_refine<<<1, RT, 0, cs>>>(d_src, d_dst, d_n, d_models, d_scores, d_params, RANSAC_REFINE_ITERS, d_result,
                                      d_mask, out);
    count_launches(2);
    LVKB_CUDA(cudaGetLastError());
    return LVKB200_OK;
}

}  // namespace lvkb200
