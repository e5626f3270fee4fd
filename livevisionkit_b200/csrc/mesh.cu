// K6c — FrameTracker::estimate_local_motions (LiveVisionKit/Vision/FrameTracker.cpp:200-321) on the device.
//
// The reference assembles a sparse system  A x = b  per frame and hands it to Eigen::LeastSquaresConjugateGradient
// (diagonal preconditioner, tolerance FLT_EPSILON, at most 2*cols iterations, warm start from the previous mesh):
//   rows 0 .. n-1          temporal:    ts * x[k]                         = ts * x_prev[k]      (FrameTracker.cpp:380-400)
//   rows n .. n+S-1        similarity:  four non-zeros per row            = 0                   (FrameTracker.cpp:402-457)
//   rows n+S+2i, n+S+2i+1  feature i:   bilinear weights of its mesh cell = matched point i     (FrameTracker.cpp:236-262)
// For the OBS "Vector Field" preset (16x16 vertices, n = 512 unknowns, ~800 features) that is ~135 CG iterations of two
// sparse products over ~3 000 rows: 2.7 ms in the host solver, i.e. 20x everything else on the frame's critical path.
//
// Here: ONE CTA keeps the whole system in shared memory and iterates with four barriers per CG step.
//   * A p   : thread per row unit (a feature = one unit = its x and y row, which share their four weights);
//   * A^T r : thread per unknown; the feature part is gathered CELL-wise (the features of the <= 4 cells around a
//             vertex, found through a per-frame counting sort by cell), so no atomics and a fixed summation order;
//   * dot products: warp shuffles + one shared round, every thread ends up with the total (no broadcast barrier).
// The kernel follows k_compact_swap_erase on the tracking stream, so the host never touches the system: it reads
// the mesh, the inlier mask and the LK results from mapped pinned memory after one stream synchronisation.
// Arithmetic: float32 as in the reference; summation ORDER differs from Eigen's (and from the sequential CPU
// restatement), so results agree to rounding (tests: <= 1e-3 px on the mesh, identical masks), not bit-wise.
// Compiled with --fmad=false: the weights / the acceptance test are the reference's unfused expressions.

#include <cfloat>
#include <cstdlib>
#include <cstring>

#include "mesh.hpp"

namespace lvkb200
{
namespace
{

constexpr int T = MESH_CGLS_THREADS;
constexpr int NW = T / 32;

struct MeshSys
{
    int cols, rows;  // vertices
    int n;           // unknowns = 2 * cols * rows
    int S;           // similarity rows
    int csc_nnz;     // non-zeros of the similarity rows (= 4 S), column-major copy
    int cap;         // feature capacity the shared-memory carve-up was sized for
    const uint16_t* sim_col;  // 4 per similarity row
    const float* sim_val;
    const int* csc_ptr;       // n + 1
    const uint16_t* csc_row;  // row index (absolute, >= n)
    const float* csc_val;
};

// Shared-memory carve-up (also evaluated on the host to size the launch).
struct Carve
{
    size_t x, p, s, invd, r, q, fw, sim_val, csc_val, csc_ptr, cell_start, cell_cnt, sim_col, csc_row, fcell, forder,
        red, total;
    __host__ __device__ Carve(int n, int S, int nnz, int cells, int cap)
    {
        size_t o = 0;
        auto take = [&o](size_t bytes) { const size_t at = o; o += (bytes + 15) & ~size_t(15); return at; };
        const size_t m = (size_t)n + S + 2 * (size_t)cap;
        x = take(4 * (size_t)n); p = take(4 * (size_t)n); s = take(4 * (size_t)n); invd = take(4 * (size_t)n);
        r = take(4 * m); q = take(4 * m);
        fw = take(16 * (size_t)cap);
        sim_val = take(16 * (size_t)S); csc_val = take(4 * (size_t)nnz);
        csc_ptr = take(4 * ((size_t)n + 1)); cell_start = take(4 * ((size_t)cells + 1)); cell_cnt = take(4 * (size_t)cells);
        sim_col = take(8 * (size_t)S); csc_row = take(2 * (size_t)nnz);
        fcell = take(2 * (size_t)cap + 8); forder = take(2 * (size_t)cap);
        red = take(2 * NW * sizeof(float2));
        total = o;
    }
};

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum of (a, b) over the CTA, returned to EVERY thread.  `slot` (NW float2) must not be rewritten before all threads
// have passed another barrier (the caller alternates two slots).
__device__ __forceinline__ float2 block_sum2(float a, float b, float2* slot)
{
    a = warp_sum(a);
    b = warp_sum(b);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) slot[wid] = make_float2(a, b);
    __syncthreads();
    static_assert(NW <= 32, "one shuffle round covers the warp totals");
    const float2 t = slot[lane < NW ? lane : 0];
    return make_float2(warp_sum(lane < NW ? t.x : 0.0f), warp_sum(lane < NW ? t.y : 0.0f));
}

struct Ctx
{
    const MeshSys& sys;
    float *x, *p, *s, *invd, *r, *q;
    float4* fw;
    const float *sim_val, *csc_val;
    const int *csc_ptr, *cell_start;
    const uint16_t *sim_col, *csc_row, *fcell, *forder;
    int N, n, S, units;
    float ts;
};

// out[rows] = A v for the row units this thread owns (unit u: temporal row u | similarity row | feature = two rows).
// ASSIGN(row, value) stores one row's product.
template <typename Assign>
__device__ __forceinline__ void spmv_rows(const Ctx& c, const float* __restrict__ v, Assign assign)
{
    const int cols = c.sys.cols;
    for (int u = threadIdx.x; u < c.units; u += T)
    {
        if (u < c.n)
            assign(u, c.ts * v[u]);
        else if (u < c.n + c.S)
        {
            const int k = u - c.n;
            const ushort4 ci = reinterpret_cast<const ushort4*>(c.sim_col)[k];
            const float4 cv = reinterpret_cast<const float4*>(c.sim_val)[k];
            assign(u, cv.x * v[ci.x] + cv.y * v[ci.y] + cv.z * v[ci.z] + cv.w * v[ci.w]);
        }
        else
        {
            const int f = u - c.n - c.S;
            const float4 w = c.fw[f];
            const int cell = c.fcell[f];
            const int cy = cell / (cols - 1), cx = cell - cy * (cols - 1);
            const int i00 = 2 * (cy * cols + cx), i10 = i00 + 2, i01 = i00 + 2 * cols, i11 = i01 + 2;
            const int row = c.n + c.S + 2 * f;
            assign(row, w.x * v[i00] + w.y * v[i01] + w.z * v[i11] + w.w * v[i10]);
            assign(row + 1, w.x * v[i00 + 1] + w.y * v[i01 + 1] + w.z * v[i11 + 1] + w.w * v[i10 + 1]);
        }
    }
}

// (A^T u)[col] for one unknown.  SQUARE: the column's squared norm instead (u ignored) -> the preconditioner.
template <bool SQUARE>
__device__ __forceinline__ float spmv_t_col(const Ctx& c, const float* __restrict__ u, int col)
{
    const int cols = c.sys.cols, rows = c.sys.rows, gw = cols - 1, gh = rows - 1;
    float acc = SQUARE ? c.ts * c.ts : c.ts * u[col];
    for (int k = c.csc_ptr[col]; k < c.csc_ptr[col + 1]; k++)
    {
        const float a = c.csc_val[k];
        acc += SQUARE ? a * a : a * u[c.csc_row[k]];
    }
    const int vtx = col >> 1, comp = col & 1;
    const int vy = vtx / cols, vx = vtx - vy * cols;
    const float* fwf = reinterpret_cast<const float*>(c.fw);
    const int base = c.n + c.S + comp;
#pragma unroll
    for (int q = 0; q < 4; q++)
    {
        // the vertex is corner i00 of cell (vx, vy) [w0], i10 of (vx-1, vy) [w3], i01 of (vx, vy-1) [w1],
        // i11 of (vx-1, vy-1) [w2]
        const int dx = (q & 1) ? -1 : 0, dy = (q & 2) ? -1 : 0;
        const int wsel = (q == 0) ? 0 : (q == 1) ? 3 : (q == 2) ? 1 : 2;
        const int cx = vx + dx, cy = vy + dy;
        if (cx < 0 || cy < 0 || cx >= gw || cy >= gh) continue;
        const int cell = cy * gw + cx;
        for (int k = c.cell_start[cell]; k < c.cell_start[cell + 1]; k++)
        {
            const int f = c.forder[k];
            const float a = fwf[4 * f + wsel];
            acc += SQUARE ? a * a : a * u[base + 2 * f];
        }
    }
    return acc;
}

// (A^T u)[col], the feature part gathered by the two lanes of a vertex TOGETHER: lane `comp` (= col & 1) walks the two
// cells 2*comp and 2*comp + 1 around the vertex and accumulates BOTH components (the x and the y row of a feature share
// their weight and sit next to each other in u), then the lanes swap the half the other one needs.  Half the
// sequential chain of spmv_t_col (the dependent shared-memory loads of the gather bound the CG step), one weight load
// instead of two per feature.  Must be called by all 32 lanes with col = lane-consecutive values (col >= n: idle lane).
__device__ __forceinline__ float spmv_t_col_pair(const Ctx& c, const float* __restrict__ u, int col)
{
    const int cols = c.sys.cols, rows = c.sys.rows, gw = cols - 1, gh = rows - 1;
    const bool live = col < c.n;
    const int comp = col & 1;
    float acc = 0.0f, ax = 0.0f, ay = 0.0f;
    if (live)
    {
        acc = c.ts * u[col];
        for (int k = c.csc_ptr[col]; k < c.csc_ptr[col + 1]; k++) acc += c.csc_val[k] * u[c.csc_row[k]];
        const int vtx = col >> 1;
        const int vy = vtx / cols, vx = vtx - vy * cols;
        const float* fwf = reinterpret_cast<const float*>(c.fw);
        const float2* uf = reinterpret_cast<const float2*>(u + c.n + c.S);  // (x row, y row) of feature f: n + S is even
#pragma unroll
        for (int h = 0; h < 2; h++)
        {
            // the vertex is corner i00 of cell (vx, vy) [w0], i10 of (vx-1, vy) [w3], i01 of (vx, vy-1) [w1],
            // i11 of (vx-1, vy-1) [w2]
            const int q = 2 * comp + h;
            const int dx = (q & 1) ? -1 : 0, dy = (q & 2) ? -1 : 0;
            const int wsel = (q == 0) ? 0 : (q == 1) ? 3 : (q == 2) ? 1 : 2;
            const int cx = vx + dx, cy = vy + dy;
            if (cx < 0 || cy < 0 || cx >= gw || cy >= gh) continue;
            const int cell = cy * gw + cx;
            for (int k = c.cell_start[cell]; k < c.cell_start[cell + 1]; k++)
            {
                const int f = c.forder[k];
                const float a = fwf[4 * f + wsel];
                const float2 uv = uf[f];
                ax += a * uv.x;
                ay += a * uv.y;
            }
        }
    }
    // lane comp = 0 owns the x column: it needs the partner's ax; lane comp = 1 owns the y column
    const float other = __shfl_xor_sync(0xffffffffu, comp ? ax : ay, 1);
    return acc + ((comp ? ay : ax) + other);
}

__device__ __forceinline__ void copy16(const uint8_t* src, uint8_t* dst, int bytes)
{
    for (int i = threadIdx.x; i < (bytes + 15) / 16; i += T)
        reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
}

__global__ void __launch_bounds__(T, 1)
    k_mesh_cgls(MeshSys sys, MeshSolveParams prm, const float2* __restrict__ src, const float2* __restrict__ dst,
                const int* __restrict__ n_ptr, const TrackParams* __restrict__ tp, float* __restrict__ state,
                uint8_t* __restrict__ mask, uint8_t* __restrict__ res_dev, uint8_t* __restrict__ res_host,
                TrackOutCopy out)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int cells = (sys.cols - 1) * (sys.rows - 1);
    const Carve cv(sys.n, sys.S, sys.csc_nnz, cells, sys.cap);
    float* const x = reinterpret_cast<float*>(smem + cv.x);
    float* const p = reinterpret_cast<float*>(smem + cv.p);
    float* const s = reinterpret_cast<float*>(smem + cv.s);
    float* const invd = reinterpret_cast<float*>(smem + cv.invd);
    float* const r = reinterpret_cast<float*>(smem + cv.r);
    float* const q = reinterpret_cast<float*>(smem + cv.q);
    float4* const fw = reinterpret_cast<float4*>(smem + cv.fw);
    float* const sim_val = reinterpret_cast<float*>(smem + cv.sim_val);
    float* const csc_val = reinterpret_cast<float*>(smem + cv.csc_val);
    int* const csc_ptr = reinterpret_cast<int*>(smem + cv.csc_ptr);
    int* const cell_start = reinterpret_cast<int*>(smem + cv.cell_start);
    int* const cell_cnt = reinterpret_cast<int*>(smem + cv.cell_cnt);
    uint16_t* const sim_col = reinterpret_cast<uint16_t*>(smem + cv.sim_col);
    uint16_t* const csc_row = reinterpret_cast<uint16_t*>(smem + cv.csc_row);
    uint16_t* const fcell = reinterpret_cast<uint16_t*>(smem + cv.fcell);
    uint16_t* const forder = reinterpret_cast<uint16_t*>(smem + cv.forder);
    float2* const red = reinterpret_cast<float2*>(smem + cv.red);

    const int tid = threadIdx.x;
    const int n = sys.n, S = sys.S, cols = sys.cols, rows = sys.rows, gw = cols - 1, gh = rows - 1;
    const int N = min(*n_ptr, sys.cap);
    MeshSolveResult* const hdr = reinterpret_cast<MeshSolveResult*>(res_dev);
    float* const mesh_out = reinterpret_cast<float*>(res_dev + sizeof(MeshSolveResult));
    int iterations = 0;
    const bool solve = N >= prm.min_samples;  // FrameTracker.cpp:152-156: too few samples -> no estimate at all

    if (solve)
    {
        // ---- stage the static system, the warm start and this frame's features
        for (int i = tid; i < n; i += T) x[i] = state[i];
        for (int i = tid; i <= n; i += T) csc_ptr[i] = sys.csc_ptr[i];
        for (int i = tid; i < 4 * S; i += T) { sim_col[i] = sys.sim_col[i]; sim_val[i] = sys.sim_val[i]; }
        for (int i = tid; i < sys.csc_nnz; i += T) { csc_row[i] = sys.csc_row[i]; csc_val[i] = sys.csc_val[i]; }
        for (int i = tid; i < cells; i += T) cell_cnt[i] = 0;
        __syncthreads();
        for (int i = tid; i < N; i += T)
        {
            // FrameTracker.cpp:236-262: mesh cell of the tracked point and its barycentric weights (Math.tpp:247-265)
            const float2 t = src[i];
            int kx = (int)max(min(__float2ll_rz(t.x / prm.key_w), (long long)INT_MAX), (long long)INT_MIN);
            int ky = (int)max(min(__float2ll_rz(t.y / prm.key_h), (long long)INT_MAX), (long long)INT_MIN);
            kx = min(max(kx, 0), gw - 1);
            ky = min(max(ky, 0), gh - 1);
            const float p0x = (float)kx * prm.key_w, p0y = (float)ky * prm.key_h;
            const float p1x = (float)(kx + 1) * prm.key_w, p1y = (float)(ky + 1) * prm.key_h;
            const float rx = fminf(p0x, p1x), ry = fminf(p0y, p1y);
            const float rw = fmaxf(p0x, p1x) - rx, rh = fmaxf(p0y, p1y) - ry;
            const float inv_area = 1.0f / (rw * rh);
            const float x2 = rx + rw, y2 = ry + rh;
            const float rx1 = x2 - t.x, ry1 = y2 - t.y, rx2 = t.x - rx, ry2 = t.y - ry;
            fw[i] = make_float4(rx1 * ry1 * inv_area, rx1 * ry2 * inv_area, rx2 * ry2 * inv_area, rx2 * ry1 * inv_area);
            const int cell = ky * gw + kx;
            fcell[i] = (uint16_t)cell;
            atomicAdd(&cell_cnt[cell], 1);
        }
        __syncthreads();
        if (n == 8)
        {
            // ---- the library-default 2x2 mesh (8 unknowns, ONE cell: every feature touches all four vertices).  The
            // sparse transpose product would walk all N features sequentially per unknown; instead the normal equations
            // are formed once - they are block structured: the x and the y rows of a feature share their weights, so
            // A^T A = [W 0; 0 W] interleaved with W = sum_f w w^T (10 sums) and A^T b from sum_f w m.x, sum_f w m.y
            // (8 sums) - and the SAME preconditioned CG iteration (oracle/lscg_ref.c) runs on the 8x8 system in one
            // thread.  Identical in exact arithmetic; the float32 summation order differs (tests: <= 5e-3 px).
            float acc[18];
#pragma unroll
            for (int k = 0; k < 18; k++) acc[k] = 0.0f;
            for (int i = tid; i < N; i += T)
            {
                const float4 w4 = fw[i];
                const float wv[4] = {w4.x, w4.w, w4.y, w4.z};  // vertices 0 (i00), 1 (i10), 2 (i01), 3 (i11)
                const float2 m = dst[i];
                int k = 0;
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int b = a; b < 4; b++) acc[k++] += wv[a] * wv[b];
#pragma unroll
                for (int a = 0; a < 4; a++) { acc[10 + a] += wv[a] * m.x; acc[14 + a] += wv[a] * m.y; }
            }
#pragma unroll
            for (int k = 0; k < 18; k++)
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_down_sync(0xffffffffu, acc[k], o);
            if ((tid & 31) == 0)
#pragma unroll
                for (int k = 0; k < 18; k++) q[(tid >> 5) * 18 + k] = acc[k];
            __syncthreads();
            if (tid == 0)
            {
                float sum[18];
                for (int k = 0; k < 18; k++)
                {
                    float t = 0.0f;
                    for (int w = 0; w < NW; w++) t += q[w * 18 + k];
                    sum[k] = t;
                }
                const float ts = prm.temporal_weight;
                float Nm[8][8], g[8], xs[8];
                for (int a = 0; a < 8; a++)
                {
                    for (int b = 0; b < 8; b++) Nm[a][b] = 0.0f;
                    xs[a] = x[a];
                    g[a] = ts * (ts * xs[a]);  // temporal rows: A = ts I, b = ts x_prev
                    Nm[a][a] = ts * ts;
                }
                int k = 0;
                for (int a = 0; a < 4; a++)
                    for (int b = a; b < 4; b++, k++)
                        for (int comp = 0; comp < 2; comp++)
                        {
                            Nm[2 * a + comp][2 * b + comp] += sum[k];
                            if (a != b) Nm[2 * b + comp][2 * a + comp] += sum[k];
                        }
                for (int a = 0; a < 4; a++) { g[2 * a] += sum[10 + a]; g[2 * a + 1] += sum[14 + a]; }
                for (int r0 = 0; r0 < S; r0++)  // similarity rows (right-hand side 0)
                    for (int a = 0; a < 4; a++)
                        for (int b = 0; b < 4; b++)
                            Nm[sim_col[4 * r0 + a]][sim_col[4 * r0 + b]] += sim_val[4 * r0 + a] * sim_val[4 * r0 + b];
                float invdg[8], nres[8], pp[8], Np[8];
                float rhs_norm2 = 0.0f, res_norm2 = 0.0f, abs_new = 0.0f;
                for (int a = 0; a < 8; a++)
                {
                    invdg[a] = Nm[a][a] > 0.0f ? 1.0f / Nm[a][a] : 1.0f;
                    float t = g[a];
                    for (int b = 0; b < 8; b++) t -= Nm[a][b] * xs[b];
                    nres[a] = t;
                    rhs_norm2 += g[a] * g[a];
                }
                if (rhs_norm2 == 0.0f)
                {
                    for (int a = 0; a < 8; a++) xs[a] = 0.0f;
                }
                else
                {
                    const float threshold = FLT_EPSILON * FLT_EPSILON * rhs_norm2;
                    for (int a = 0; a < 8; a++) res_norm2 += nres[a] * nres[a];
                    if (!(res_norm2 < threshold))
                    {
                        for (int a = 0; a < 8; a++) { pp[a] = invdg[a] * nres[a]; abs_new += nres[a] * pp[a]; }
                        while (iterations < 16)
                        {
                            float pNp = 0.0f;
                            for (int a = 0; a < 8; a++)
                            {
                                float t = 0.0f;
                                for (int b = 0; b < 8; b++) t += Nm[a][b] * pp[b];
                                Np[a] = t;
                                pNp += pp[a] * t;
                            }
                            const float alpha = abs_new / pNp;
                            res_norm2 = 0.0f;
                            for (int a = 0; a < 8; a++)
                            {
                                xs[a] += alpha * pp[a];
                                nres[a] -= alpha * Np[a];
                                res_norm2 += nres[a] * nres[a];
                            }
                            if (res_norm2 < threshold) break;
                            const float abs_old = abs_new;
                            abs_new = 0.0f;
                            for (int a = 0; a < 8; a++) abs_new += nres[a] * (invdg[a] * nres[a]);
                            const float beta = abs_new / abs_old;
                            for (int a = 0; a < 8; a++) pp[a] = invdg[a] * nres[a] + beta * pp[a];
                            iterations++;
                        }
                    }
                }
                for (int a = 0; a < 8; a++) x[a] = xs[a];
            }
            __syncthreads();
        }
        else
        {
        // exclusive scan of the cell populations (<= a few hundred cells: one warp, 32 cells per step)
        if (tid < 32)
        {
            int carry = 0;
            for (int c0 = 0; c0 < cells; c0 += 32)
            {
                const int c = c0 + tid;
                const int v = c < cells ? cell_cnt[c] : 0;
                int incl = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1)
                {
                    const int t2 = __shfl_up_sync(0xffffffffu, incl, o);
                    if (tid >= o) incl += t2;
                }
                if (c < cells) cell_start[c] = carry + incl - v;
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (tid == 0) cell_start[cells] = carry;
        }
        __syncthreads();
        // stable placement: feature i goes behind the features j < i of its cell (fixed summation order, no atomics)
        for (int i = tid; i < N; i += T)
        {
            const uint16_t me = fcell[i];
            int rank = 0;
            for (int j = 0; j < i; j++) rank += (fcell[j] == me);
            forder[cell_start[me] + rank] = (uint16_t)i;
        }
        __syncthreads();

        Ctx c{sys, x, p, s, invd, r, q, fw, sim_val, csc_val, csc_ptr, cell_start, sim_col, csc_row, fcell, forder,
              N, n, S, n + S + N, prm.temporal_weight};
        const float ts = prm.temporal_weight;

        // ---- LeastSquareDiagonalPreconditioner, residual = b - A x, z = A^T b
        // q <- b  (temporal rows: ts * x_prev, similarity rows: 0, feature rows: the matched point)
        for (int u = tid; u < c.units; u += T)
        {
            if (u < n) q[u] = ts * x[u];
            else if (u < n + S) q[u] = 0.0f;
            else
            {
                const float2 m = dst[u - n - S];
                q[n + S + 2 * (u - n - S)] = m.x;
                q[n + S + 2 * (u - n - S) + 1] = m.y;
            }
        }
        spmv_rows(c, x, [&](int row, float v) { r[row] = v; });  // A x (own rows, read back by the same thread below)
        __syncthreads();
        float zz = 0.0f;
        for (int col = tid; col < n; col += T)
        {
            const float d = spmv_t_col<true>(c, nullptr, col);
            invd[col] = d > 0.0f ? 1.0f / d : 1.0f;
            const float z = spmv_t_col<false>(c, q, col);  // A^T b
            zz += z * z;
        }
        const float rhs_norm2 = block_sum2(zz, 0.0f, red).x;
        // residual = b - A x
        for (int i = tid; i < n + S + 2 * N; i += T) r[i] = q[i] - r[i];
        __syncthreads();

        if (rhs_norm2 == 0.0f)
        {
            for (int i = tid; i < n; i += T) x[i] = 0.0f;
        }
        else
        {
            const float threshold = FLT_EPSILON * FLT_EPSILON * rhs_norm2;
            const int n_up = (n + 31) & ~31;  // whole warps enter the paired gather
            float ss = 0.0f, sz = 0.0f;
            for (int col = tid; col < n_up; col += T)
            {
                const float v = spmv_t_col_pair(c, r, col);
                if (col < n)
                {
                    s[col] = v;
                    const float z = invd[col] * v;
                    p[col] = z;
                    ss += v * v;
                    sz += v * z;
                }
            }
            const float2 t0 = block_sum2(ss, sz, red + NW);
            float abs_new = t0.y;
            if (!(t0.x < threshold))
            {
                __syncthreads();  // p complete
                const int max_iters = 2 * n;
                while (iterations < max_iters)
                {
                    // q = A p ; alpha = abs_new / |q|^2
                    float qq = 0.0f;
                    spmv_rows(c, p, [&](int row, float v) { q[row] = v; qq += v * v; });
                    const float alpha = abs_new / block_sum2(qq, 0.0f, red).x;
                    // x += alpha p ; residual -= alpha q  (q rows are this thread's own)
                    for (int i = tid; i < n; i += T) x[i] += alpha * p[i];
                    for (int u = tid; u < c.units; u += T)
                    {
                        if (u < n + S) r[u] -= alpha * q[u];
                        else
                        {
                            const int row = n + S + 2 * (u - n - S);
                            r[row] -= alpha * q[row];
                            r[row + 1] -= alpha * q[row + 1];
                        }
                    }
                    __syncthreads();
                    // s = A^T residual ; z = M^-1 s
                    ss = 0.0f; sz = 0.0f;
                    for (int col = tid; col < n_up; col += T)
                    {
                        const float v = spmv_t_col_pair(c, r, col);
                        if (col < n)
                        {
                            s[col] = v;
                            ss += v * v;
                            sz += v * (invd[col] * v);
                        }
                    }
                    const float2 t1 = block_sum2(ss, sz, red + NW);
                    if (t1.x < threshold) break;
                    const float beta = t1.y / abs_new;
                    abs_new = t1.y;
                    for (int col = tid; col < n; col += T) p[col] = invd[col] * s[col] + beta * p[col];
                    iterations++;
                    __syncthreads();
                }
            }
        }
        }  // sparse path (n != 8)
        __syncthreads();

        // ---- results: new state, mesh copy for the host, inlier mask (FrameTracker.cpp:279-300)
        for (int i = tid; i < n; i += T)
        {
            const float v = x[i];
            state[i] = v;
            mesh_out[i] = v;
        }
        for (int i = tid; i < N; i += T)
        {
            const float4 w = fw[i];
            const int cell = fcell[i];
            const int cy = cell / gw, cx = cell - cy * gw;
            const int i00 = 2 * (cy * cols + cx), i10 = i00 + 2, i01 = i00 + 2 * cols, i11 = i01 + 2;
            // row product in the reference's triplet order i00, i01, i11, i10
            const float ex = ((w.x * x[i00] + w.y * x[i01]) + w.z * x[i11]) + w.w * x[i10];
            const float ey = ((w.x * x[i00 + 1] + w.y * x[i01 + 1]) + w.z * x[i11 + 1]) + w.w * x[i10 + 1];
            const float2 m = dst[i];
            mask[i] = (fabsf(ex - m.x) + fabsf(ey - m.y)) < prm.acceptance ? 1 : 0;
        }
    }
    if (tid == 0)
    {
        hdr->solved = solve ? 1 : 0;
        hdr->iterations = iterations;
        hdr->n = N;
        hdr->pad = 0;
    }
    __syncthreads();  // the block's global writes (mask, header, mesh) are visible to all of its threads
    if (res_host) copy16(res_dev, res_host, (int)sizeof(MeshSolveResult) + (solve ? 4 * n : 0));
    if (out.host)
    {
        const int tracked = tp->n;
        copy16(out.dev, out.host, tracked * (int)sizeof(float2));
        copy16(out.dev + out.off_status, out.host + out.off_status, tracked);
        if (solve) copy16(out.dev + out.off_mask, out.host + out.off_mask, N);
    }
}


// ------------------------------------------------------------------------------------------------------------------
// Round 2: the sparse solve re-tiled for issue rate (k_mesh_cgls2).  The ncu source view of the kernel above on the
// "Vector Field" preset (profiles/r02_mesh_cgls_v1_lines.txt: 688 us, 136 iterations, 17 400 warp-instructions per
// iteration at 44 % issue-active) showed where an iteration went: 43 % in the transposed product (a lane per unknown
// walking the features of its four cells: the warp runs as long as its most crowded vertex, ~26 visits of 15
// instructions), 22 % in the row products (two integer divisions and a three-way branch per row unit), 9 % in block
// reductions.  Here:
//   * 1024 threads.  The features are renumbered in cell order once per frame (one warp, match.any on the cell index;
//     weights and vertex index stored at the sorted position), so every per-iteration access is contiguous and
//     division-free.
//   * A^T r is split: warps 16-31 sum CHAINS - per cell, the four corner-weighted sums of its features' (r_x, r_y),
//     two lanes per cell (even / odd features), the cells dealt to the warps in order of population so that no lane
//     waits for a crowded neighbour; warps 0-15 meanwhile walk the similarity columns (lists padded to groups of
//     four with zero weights: LDS.128); after one barrier a column adds its four chain sums.
//   * q = A p stays in registers between the product and the residual update (a thread owns the same rows in both);
//     the kernel is instantiated per (unknowns, similarity rows, features) a thread may own.
// 396 us, 12 400 warp-instructions per iteration (profiles/r02_mesh_cgls2_lines.txt).  Same iteration as before
// (Eigen's LSCG, oracle/lscg_ref.c); the float32 summation order of the dot products and of the per-vertex gather
// differs again (tests: same tolerances as the kernel above).  Meshes beyond the slot limits below, and the 2x2 mesh,
// stay on k_mesh_cgls.
namespace v2
{

constexpr int T2 = 1024, NW2 = T2 / 32, COLT = 512;
constexpr int MAX_SLOT_N = 2, MAX_SLOT_S = 2, MAX_SLOT_F = 4;  // unknowns / similarity rows / features a thread may own
static_assert(NW2 == 32, "the block sums read one warp total per lane");

struct Carve2
{
    size_t x, p, s, invd, r, sw, sim_val, csc_val, P, csc_ptr, cell_start, cell_cnt, colP, sim_col, csc_row, si00,
        fcell, forder, chain_cell, red, total;
    __host__ __device__ Carve2(int n, int S, int nnz, int cells, int cap)
    {
        size_t o = 0;
        auto take = [&o](size_t bytes) { const size_t at = o; o += (bytes + 15) & ~size_t(15); return at; };
        x = take(4 * (size_t)n); p = take(4 * (size_t)n); s = take(4 * (size_t)n); invd = take(4 * (size_t)n);
        r = take(4 * ((size_t)n + S + 2 * (size_t)cap));
        sw = take(16 * (size_t)cap);
        sim_val = take(16 * (size_t)S); csc_val = take(4 * (size_t)nnz);
        P = take(8 * (4 * (size_t)cells + 1));
        csc_ptr = take(4 * ((size_t)n + 1)); cell_start = take(4 * ((size_t)cells + 1)); cell_cnt = take(4 * (size_t)cells);
        colP = take(8 * (size_t)n); sim_col = take(8 * (size_t)S); csc_row = take(2 * (size_t)nnz);
        si00 = take(2 * (size_t)cap); fcell = take(2 * (size_t)cap + 8); forder = take(2 * (size_t)cap);
        chain_cell = take(2 * (size_t)cells);
        red = take(NW2 * sizeof(float) + NW2 * sizeof(float2));
        total = o;
    }
};

__device__ __forceinline__ float block_sum1(float a, float* slot)
{
    a = warp_sum(a);
    if ((threadIdx.x & 31) == 0) slot[threadIdx.x >> 5] = a;
    __syncthreads();
    return warp_sum(slot[threadIdx.x & 31]);
}

// Sum of (a, b) over the COLUMN warps (threads < COLT; the others hold nothing), returned to every thread.
__device__ __forceinline__ float2 block_sum2(float a, float b, float2* slot)
{
    if (threadIdx.x < COLT)
    {
        a = warp_sum(a);
        b = warp_sum(b);
        if ((threadIdx.x & 31) == 0) slot[threadIdx.x >> 5] = make_float2(a, b);
    }
    __syncthreads();
    static_assert(COLT / 32 == 16, "the second round is a 16-lane butterfly");
    float2 t = slot[threadIdx.x & 15];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1)
    {
        t.x += __shfl_xor_sync(0xffffffffu, t.x, o);
        t.y += __shfl_xor_sync(0xffffffffu, t.y, o);
    }
    return t;
}

struct Sh
{
    float *x, *p, *s, *invd, *r;
    float4* sw;
    float *sim_val, *csc_val;
    float2* P;
    int *csc_ptr, *cell_start, *cell_cnt;
    ushort4* colP;
    uint16_t *sim_col, *csc_row, *si00, *fcell, *forder, *chain_cell;
    int n, S, N, cells, cols;
    float ts;
};

// emit(col, (A^T r)[col]) for every unknown, r = the row vector in h.r.  SQUARE: the squared column norms instead.
// Contains one barrier; the caller separates two calls by another one (h.P and h.s are rewritten).
template <bool SQUARE, typename Emit>
__device__ __forceinline__ void transpose_product(const Sh& h, Emit emit)
{
    const int tid = threadIdx.x;
    if (tid >= COLT)
    {
        const float2* uf = reinterpret_cast<const float2*>(h.r + h.n + h.S);  // (x row, y row) of a feature: n + S is even
        // two lanes per cell: the even and the odd features of its (sorted) list; whole warps stay in the loop for the
        // exchange (2 * cells is even: a pair is live or idle together)
        const int lane = tid & 31;
        for (int j0 = tid - COLT - lane; j0 < 2 * h.cells; j0 += T2 - COLT)
        {
            const int j = j0 + lane;
            const bool live = j < 2 * h.cells;
            const int cell = live ? h.chain_cell[j >> 1] : 0;
            const int k1 = live ? h.cell_start[cell + 1] : 0;
            float2 c0 = make_float2(0.0f, 0.0f), c1 = c0, c2 = c0, c3 = c0;
#pragma unroll 2
            for (int k = live ? h.cell_start[cell] + (j & 1) : 0; k < k1; k += 2)
            {
                const float4 w = h.sw[k];
                if (SQUARE)
                {
                    c0.x += w.x * w.x; c1.x += w.y * w.y; c2.x += w.z * w.z; c3.x += w.w * w.w;
                }
                else
                {
                    const float2 uv = uf[k];
                    c0.x += w.x * uv.x; c0.y += w.x * uv.y;
                    c1.x += w.y * uv.x; c1.y += w.y * uv.y;
                    c2.x += w.z * uv.x; c2.y += w.z * uv.y;
                    c3.x += w.w * uv.x; c3.y += w.w * uv.y;
                }
            }
            // the even lane writes corners 0 and 1, the odd lane corners 2 and 3: each sends the partner's half
            const bool odd = (j & 1) != 0;
            if (SQUARE) { c0.y = c0.x; c1.y = c1.x; c2.y = c2.x; c3.y = c3.x; }
            float4 give = odd ? make_float4(c0.x, c0.y, c1.x, c1.y) : make_float4(c2.x, c2.y, c3.x, c3.y);
            const float4 keep = odd ? make_float4(c2.x, c2.y, c3.x, c3.y) : make_float4(c0.x, c0.y, c1.x, c1.y);
            give.x = __shfl_xor_sync(0xffffffffu, give.x, 1); give.y = __shfl_xor_sync(0xffffffffu, give.y, 1);
            give.z = __shfl_xor_sync(0xffffffffu, give.z, 1); give.w = __shfl_xor_sync(0xffffffffu, give.w, 1);
            if (live)
                reinterpret_cast<float4*>(h.P + 4 * cell)[odd ? 1 : 0] =
                    make_float4(keep.x + give.x, keep.y + give.y, keep.z + give.z, keep.w + give.w);
        }
    }
    else
    {
        for (int col = tid; col < h.n; col += COLT)
        {
            float acc = SQUARE ? h.ts * h.ts : h.ts * h.r[col];
            const int k1 = h.csc_ptr[col + 1];
#pragma unroll 1
            for (int k = h.csc_ptr[col]; k < k1; k += 4)  // padded to whole groups of four (zero weights) on the host
            {
                const float4 a = *reinterpret_cast<const float4*>(h.csc_val + k);
                const ushort4 ri = *reinterpret_cast<const ushort4*>(h.csc_row + k);
                if (SQUARE) { acc += a.x * a.x; acc += a.y * a.y; acc += a.z * a.z; acc += a.w * a.w; }
                else
                {
                    const float u0 = h.r[ri.x], u1 = h.r[ri.y], u2 = h.r[ri.z], u3 = h.r[ri.w];
                    acc += a.x * u0; acc += a.y * u1; acc += a.z * u2; acc += a.w * u3;
                }
            }
            h.s[col] = acc;
        }
    }
    __syncthreads();
    if (tid < COLT)
    {
        const float* Pf = reinterpret_cast<const float*>(h.P);
        for (int col = tid; col < h.n; col += COLT)
        {
            const ushort4 pi = h.colP[col];
            const int comp = col & 1;
            // corner i00 of cell (vx, vy) and i10 of (vx-1, vy); i01 of (vx, vy-1) and i11 of (vx-1, vy-1)
            const float f = (Pf[2 * pi.x + comp] + Pf[2 * pi.y + comp]) + (Pf[2 * pi.z + comp] + Pf[2 * pi.w + comp]);
            emit(col, h.s[col] + f);
        }
    }
}

__device__ __forceinline__ void copy16(const uint8_t* src, uint8_t* dst, int bytes)
{
    for (int i = threadIdx.x; i < (bytes + 15) / 16; i += T2)
        reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
}

template <int SN, int SS, int SF>
__global__ void __launch_bounds__(T2, 1)
    k_mesh_cgls2(MeshSys sys, MeshSolveParams prm, const float2* __restrict__ src, const float2* __restrict__ dst,
                 const int* __restrict__ n_ptr, const TrackParams* __restrict__ tp, float* __restrict__ state,
                 uint8_t* __restrict__ mask, uint8_t* __restrict__ res_dev, uint8_t* __restrict__ res_host,
                 TrackOutCopy out)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x;
    const int n = sys.n, S = sys.S, cols = sys.cols, rows = sys.rows, gw = cols - 1, gh = rows - 1;
    const int cells = gw * gh;
    const Carve2 cv(n, S, sys.csc_nnz, cells, sys.cap);
    Sh h;
    h.x = reinterpret_cast<float*>(smem + cv.x); h.p = reinterpret_cast<float*>(smem + cv.p);
    h.s = reinterpret_cast<float*>(smem + cv.s); h.invd = reinterpret_cast<float*>(smem + cv.invd);
    h.r = reinterpret_cast<float*>(smem + cv.r); h.sw = reinterpret_cast<float4*>(smem + cv.sw);
    h.sim_val = reinterpret_cast<float*>(smem + cv.sim_val); h.csc_val = reinterpret_cast<float*>(smem + cv.csc_val);
    h.P = reinterpret_cast<float2*>(smem + cv.P);
    h.csc_ptr = reinterpret_cast<int*>(smem + cv.csc_ptr); h.cell_start = reinterpret_cast<int*>(smem + cv.cell_start);
    h.cell_cnt = reinterpret_cast<int*>(smem + cv.cell_cnt); h.colP = reinterpret_cast<ushort4*>(smem + cv.colP);
    h.sim_col = reinterpret_cast<uint16_t*>(smem + cv.sim_col); h.csc_row = reinterpret_cast<uint16_t*>(smem + cv.csc_row);
    h.si00 = reinterpret_cast<uint16_t*>(smem + cv.si00); h.fcell = reinterpret_cast<uint16_t*>(smem + cv.fcell);
    h.forder = reinterpret_cast<uint16_t*>(smem + cv.forder); h.chain_cell = reinterpret_cast<uint16_t*>(smem + cv.chain_cell);
    float* const red1 = reinterpret_cast<float*>(smem + cv.red);
    float2* const red2 = reinterpret_cast<float2*>(smem + cv.red + NW2 * sizeof(float));
    float* const x = h.x; float* const p = h.p; float* const r = h.r;

    const int N = min(*n_ptr, sys.cap);
    h.n = n; h.S = S; h.N = N; h.cells = cells; h.cols = cols; h.ts = prm.temporal_weight;
    const float ts = prm.temporal_weight;
    const int NS = n + S;
    MeshSolveResult* const hdr = reinterpret_cast<MeshSolveResult*>(res_dev);
    float* const mesh_out = reinterpret_cast<float*>(res_dev + sizeof(MeshSolveResult));
    int iterations = 0;
    const bool solve = N >= prm.min_samples;  // FrameTracker.cpp:152-156: too few samples -> no estimate at all

    if (solve)
    {
        // ---- stage the static system and the warm start
        for (int i = tid; i < n; i += T2) x[i] = state[i];
        for (int i = tid; i <= n; i += T2) h.csc_ptr[i] = sys.csc_ptr[i];
        for (int i = tid; i < 4 * S; i += T2) { h.sim_col[i] = sys.sim_col[i]; h.sim_val[i] = sys.sim_val[i]; }
        for (int i = tid; i < sys.csc_nnz; i += T2) { h.csc_row[i] = sys.csc_row[i]; h.csc_val[i] = sys.csc_val[i]; }
        for (int i = tid; i < cells; i += T2) h.cell_cnt[i] = 0;
        if (tid == 0) h.P[4 * cells] = make_float2(0.0f, 0.0f);  // what a vertex on the border adds for a missing cell
        __syncthreads();
        // ---- mesh cell of every tracked point (FrameTracker.cpp:236-246), cell populations
        for (int i = tid; i < N; i += T2)
        {
            const float2 t = src[i];
            int kx = (int)max(min(__float2ll_rz(t.x / prm.key_w), (long long)INT_MAX), (long long)INT_MIN);
            int ky = (int)max(min(__float2ll_rz(t.y / prm.key_h), (long long)INT_MAX), (long long)INT_MIN);
            kx = min(max(kx, 0), gw - 1);
            ky = min(max(ky, 0), gh - 1);
            const int cell = ky * gw + kx;
            h.fcell[i] = (uint16_t)cell;
            atomicAdd(&h.cell_cnt[cell], 1);
        }
        __syncthreads();
        if (tid < 32)
        {
            // exclusive scan of the cell populations (one warp, 32 cells per step)
            int carry = 0;
            for (int c0 = 0; c0 < cells; c0 += 32)
            {
                const int c = c0 + tid;
                const int v = c < cells ? h.cell_cnt[c] : 0;
                int incl = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1)
                {
                    const int t2 = __shfl_up_sync(0xffffffffu, incl, o);
                    if (tid >= o) incl += t2;
                }
                if (c < cells) h.cell_start[c] = carry + incl - v;
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (tid == 0) h.cell_start[cells] = carry;
        }
        else
        {
            // the four chain sums an unknown adds up: indices into P (4 per cell; slot 4*cells holds zero)
            for (int col = tid - 32; col < n; col += T2 - 32)
            {
                const int vtx = col >> 1;
                const int vy = vtx / cols, vx = vtx - vy * cols;
                const int none = 4 * cells;
                const bool l = vx > 0, t = vy > 0, rr = vx < gw, b = vy < gh;
                ushort4 pi;
                pi.x = (uint16_t)((rr && b) ? 4 * (vy * gw + vx) + 0 : none);            // i00 of (vx, vy): w0
                pi.y = (uint16_t)((l && b) ? 4 * (vy * gw + vx - 1) + 3 : none);         // i10 of (vx-1, vy): w3
                pi.z = (uint16_t)((rr && t) ? 4 * ((vy - 1) * gw + vx) + 1 : none);      // i01 of (vx, vy-1): w1
                pi.w = (uint16_t)((l && t) ? 4 * ((vy - 1) * gw + vx - 1) + 2 : none);   // i11 of (vx-1, vy-1): w2
                h.colP[col] = pi;
            }
        }
        __syncthreads();
        // ---- features into cell order (stable: ascending index inside a cell).  One warp walks the features 32 at a
        // time: lanes of the same cell find each other with match.any and take consecutive places behind the cell's
        // running fill mark - O(N / 32) steps instead of an O(N) comparison loop per feature.
        if (tid < 32)
        {
            int* const fill = reinterpret_cast<int*>(h.P);  // (the chain sums are not in use yet)
            for (int c = tid; c < cells; c += 32) fill[c] = h.cell_start[c];
            __syncwarp();
            for (int base = 0; base < N; base += 32)
            {
                const int i = base + tid;
                const unsigned cell = i < N ? h.fcell[i] : 0xFFFFu;  // the lanes behind the last feature: a group of their own
                const unsigned peers = __match_any_sync(0xffffffffu, cell);
                const int rank = __popc(peers & ((1u << tid) - 1u));
                if (i < N) h.forder[fill[cell] + rank] = (uint16_t)i;
                __syncwarp();
                if (i < N && rank == 0) fill[cell] += __popc(peers);
                __syncwarp();
            }
        }
        else
        {
            // cells by falling population: the order the chains are dealt to the warps in
            for (int c = tid - 32; c < cells; c += T2 - 32)
            {
                const int mine = h.cell_cnt[c];
                int rank = 0;
                for (int o = 0; o < cells; o++)
                {
                    const int other = h.cell_cnt[o];
                    rank += (other > mine) || (other == mine && o < c);
                }
                h.chain_cell[rank] = (uint16_t)c;
            }
        }
        __syncthreads();
        // barycentric weights of every point in its cell (Math.tpp:247-265), at the sorted position
        for (int k = tid; k < N; k += T2)
        {
            const int i = h.forder[k];
            const int me = h.fcell[i];
            const int ky = me / gw, kx = me - ky * gw;
            const float2 t = src[i];
            const float p0x = (float)kx * prm.key_w, p0y = (float)ky * prm.key_h;
            const float p1x = (float)(kx + 1) * prm.key_w, p1y = (float)(ky + 1) * prm.key_h;
            const float rx = fminf(p0x, p1x), ry = fminf(p0y, p1y);
            const float rw = fmaxf(p0x, p1x) - rx, rh = fmaxf(p0y, p1y) - ry;
            const float inv_area = 1.0f / (rw * rh);
            const float x2 = rx + rw, y2 = ry + rh;
            const float rx1 = x2 - t.x, ry1 = y2 - t.y, rx2 = t.x - rx, ry2 = t.y - ry;
            h.sw[k] = make_float4(rx1 * ry1 * inv_area, rx1 * ry2 * inv_area, rx2 * ry2 * inv_area, rx2 * ry1 * inv_area);
            h.si00[k] = (uint16_t)(2 * (ky * cols + kx));
        }
        __syncthreads();

        // ---- r <- b  (temporal rows: ts * x_prev, similarity rows: 0, feature rows: the matched point)
#pragma unroll
        for (int sl = 0; sl < SN; sl++)
        {
            const int u = tid + sl * T2;
            if (u < n) r[u] = ts * x[u];
        }
#pragma unroll
        for (int sl = 0; sl < SS; sl++)
        {
            const int k = tid + sl * T2;
            if (k < S) r[n + k] = 0.0f;
        }
        float2* const rf = reinterpret_cast<float2*>(r + NS);
        const float2* const p2 = reinterpret_cast<const float2*>(p);
        const float2* const x2v = reinterpret_cast<const float2*>(x);
#pragma unroll
        for (int sl = 0; sl < SF; sl++)
        {
            const int k = tid + sl * T2;
            if (k < N) rf[k] = dst[h.forder[k]];
        }
        __syncthreads();
        // ---- LeastSquareDiagonalPreconditioner, |A^T b|^2
        transpose_product<true>(h, [&](int col, float d) { h.invd[col] = d > 0.0f ? 1.0f / d : 1.0f; });
        __syncthreads();
        float zz = 0.0f;
        transpose_product<false>(h, [&](int, float z) { zz += z * z; });
        const float rhs_norm2 = block_sum1(zz, red1);
        // ---- residual = b - A x  (every thread its own rows)
#pragma unroll
        for (int sl = 0; sl < SN; sl++)
        {
            const int u = tid + sl * T2;
            if (u < n) r[u] = r[u] - ts * x[u];
        }
#pragma unroll
        for (int sl = 0; sl < SS; sl++)
        {
            const int k = tid + sl * T2;
            if (k < S)
            {
                const ushort4 ci = reinterpret_cast<const ushort4*>(h.sim_col)[k];
                const float4 cw = reinterpret_cast<const float4*>(h.sim_val)[k];
                r[n + k] = r[n + k] - (cw.x * x[ci.x] + cw.y * x[ci.y] + cw.z * x[ci.z] + cw.w * x[ci.w]);
            }
        }
#pragma unroll
        for (int sl = 0; sl < SF; sl++)
        {
            const int k = tid + sl * T2;
            if (k < N)
            {
                const float4 w = h.sw[k];
                const int a00 = h.si00[k] >> 1, a01 = a00 + cols;
                const float2 v00 = x2v[a00], v10 = x2v[a00 + 1], v01 = x2v[a01], v11 = x2v[a01 + 1];
                float2 b = rf[k];
                b.x = b.x - (w.x * v00.x + w.y * v01.x + w.z * v11.x + w.w * v10.x);
                b.y = b.y - (w.x * v00.y + w.y * v01.y + w.z * v11.y + w.w * v10.y);
                rf[k] = b;
            }
        }
        __syncthreads();

        if (rhs_norm2 == 0.0f)
        {
            for (int i = tid; i < n; i += T2) x[i] = 0.0f;
        }
        else
        {
            const float threshold = FLT_EPSILON * FLT_EPSILON * rhs_norm2;
            float ss = 0.0f, sz = 0.0f;
            transpose_product<false>(h, [&](int col, float v) {
                h.s[col] = v;
                const float z = h.invd[col] * v;
                p[col] = z;
                ss += v * v;
                sz += v * z;
            });
            const float2 t0 = block_sum2(ss, sz, red2);  // (its barrier also completes p)
            float abs_new = t0.y;
            if (!(t0.x < threshold))
            {
                const int max_iters = 2 * n;
                while (iterations < max_iters)
                {
                    // q = A p (kept in registers: the thread owns the same rows in the residual update)
                    float qn[SN], qs[SS], qfx[SF], qfy[SF];
                    float qq = 0.0f;
#pragma unroll
                    for (int sl = 0; sl < SN; sl++)
                    {
                        const int u = tid + sl * T2;
                        qn[sl] = 0.0f;
                        if (u < n)
                        {
                            const float v = ts * p[u];
                            qn[sl] = v;
                            qq += v * v;
                        }
                    }
#pragma unroll
                    for (int sl = 0; sl < SS; sl++)
                    {
                        const int k = tid + sl * T2;
                        qs[sl] = 0.0f;
                        if (k < S)
                        {
                            const ushort4 ci = reinterpret_cast<const ushort4*>(h.sim_col)[k];
                            const float4 cw = reinterpret_cast<const float4*>(h.sim_val)[k];
                            const float v = cw.x * p[ci.x] + cw.y * p[ci.y] + cw.z * p[ci.z] + cw.w * p[ci.w];
                            qs[sl] = v;
                            qq += v * v;
                        }
                    }
#pragma unroll
                    for (int sl = 0; sl < SF; sl++)
                    {
                        const int k = tid + sl * T2;
                        qfx[sl] = 0.0f; qfy[sl] = 0.0f;
                        if (k < N)
                        {
                            const float4 w = h.sw[k];
                            const int a00 = h.si00[k] >> 1, a01 = a00 + cols;
                            const float2 v00 = p2[a00], v10 = p2[a00 + 1], v01 = p2[a01], v11 = p2[a01 + 1];
                            const float vx = w.x * v00.x + w.y * v01.x + w.z * v11.x + w.w * v10.x;
                            const float vy = w.x * v00.y + w.y * v01.y + w.z * v11.y + w.w * v10.y;
                            qfx[sl] = vx; qfy[sl] = vy;
                            qq += vx * vx;
                            qq += vy * vy;
                        }
                    }
                    const float alpha = abs_new / block_sum1(qq, red1);
                    // x += alpha p ; residual -= alpha q
#pragma unroll
                    for (int sl = 0; sl < SN; sl++)
                    {
                        const int u = tid + sl * T2;
                        if (u < n)
                        {
                            x[u] += alpha * p[u];
                            r[u] -= alpha * qn[sl];
                        }
                    }
#pragma unroll
                    for (int sl = 0; sl < SS; sl++)
                    {
                        const int k = tid + sl * T2;
                        if (k < S) r[n + k] -= alpha * qs[sl];
                    }
#pragma unroll
                    for (int sl = 0; sl < SF; sl++)
                    {
                        const int k = tid + sl * T2;
                        if (k < N)
                        {
                            float2 b = rf[k];
                            b.x -= alpha * qfx[sl];
                            b.y -= alpha * qfy[sl];
                            rf[k] = b;
                        }
                    }
                    __syncthreads();
                    // s = A^T residual ; z = M^-1 s
                    ss = 0.0f; sz = 0.0f;
                    transpose_product<false>(h, [&](int col, float v) {
                        h.s[col] = v;
                        ss += v * v;
                        sz += v * (h.invd[col] * v);
                    });
                    const float2 t1 = block_sum2(ss, sz, red2);
                    if (t1.x < threshold) break;
                    const float beta = t1.y / abs_new;
                    abs_new = t1.y;
                    if (tid < COLT)
                        for (int col = tid; col < n; col += COLT) p[col] = h.invd[col] * h.s[col] + beta * p[col];
                    iterations++;
                    __syncthreads();
                }
            }
        }
        __syncthreads();

        // ---- results: new state, mesh copy for the host, inlier mask (FrameTracker.cpp:279-300)
        for (int i = tid; i < n; i += T2)
        {
            const float v = x[i];
            state[i] = v;
            mesh_out[i] = v;
        }
        for (int k = tid; k < N; k += T2)
        {
            const float4 w = h.sw[k];
            const int i = h.forder[k];
            const int i00 = h.si00[k], i10 = i00 + 2, i01 = i00 + 2 * cols, i11 = i01 + 2;
            // row product in the reference's triplet order i00, i01, i11, i10
            const float ex = ((w.x * x[i00] + w.y * x[i01]) + w.z * x[i11]) + w.w * x[i10];
            const float ey = ((w.x * x[i00 + 1] + w.y * x[i01 + 1]) + w.z * x[i11 + 1]) + w.w * x[i10 + 1];
            const float2 m = dst[i];
            mask[i] = (fabsf(ex - m.x) + fabsf(ey - m.y)) < prm.acceptance ? 1 : 0;
        }
    }
    if (tid == 0)
    {
        hdr->solved = solve ? 1 : 0;
        hdr->iterations = iterations;
        hdr->n = N;
        hdr->pad = 0;
    }
    __syncthreads();  // the block's global writes (mask, header, mesh) are visible to all of its threads
    if (res_host) copy16(res_dev, res_host, (int)sizeof(MeshSolveResult) + (solve ? 4 * n : 0));
    if (out.host)
    {
        const int tracked = tp->n;
        copy16(out.dev, out.host, tracked * (int)sizeof(float2));
        copy16(out.dev + out.off_status, out.host + out.off_status, tracked);
        if (solve) copy16(out.dev + out.off_mask, out.host + out.off_mask, N);
    }
}


// The slot triple (unknowns, similarity rows, features per thread) the kernel is instantiated for.
inline int slot_variant(int n, int S, int cap)
{
    if (n <= T2 && S <= T2 && cap <= T2) return 0;
    if (n <= T2 && S <= T2 && cap <= 3 * T2) return 1;
    return 2;
}

inline cudaError_t set_smem(int variant, size_t bytes)
{
    const int b = static_cast<int>(bytes);
    switch (variant)
    {
    case 0: return cudaFuncSetAttribute(k_mesh_cgls2<1, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, b);
    case 1: return cudaFuncSetAttribute(k_mesh_cgls2<1, 1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, b);
    default: return cudaFuncSetAttribute(k_mesh_cgls2<MAX_SLOT_N, MAX_SLOT_S, MAX_SLOT_F>, cudaFuncAttributeMaxDynamicSharedMemorySize, b);
    }
}

template <typename... Args>
inline void launch(int variant, size_t smem, cudaStream_t cs, Args... args)
{
    switch (variant)
    {
    case 0: k_mesh_cgls2<1, 1, 1><<<1, T2, smem, cs>>>(args...); break;
    case 1: k_mesh_cgls2<1, 1, 3><<<1, T2, smem, cs>>>(args...); break;
    default: k_mesh_cgls2<MAX_SLOT_N, MAX_SLOT_S, MAX_SLOT_F><<<1, T2, smem, cs>>>(args...); break;
    }
}

}  // namespace v2

}  // namespace

bool MeshCgls::configure(const MeshStaticRows& sys, int feature_capacity, cudaStream_t cs, cudaError_t* err)
{
    *err = cudaSuccess;
    const int n = 2 * sys.mesh_cols * sys.mesh_rows, S = sys.rows(), nnz = 4 * S;
    const int cells = (sys.mesh_cols - 1) * (sys.mesh_rows - 1);
    n_unknowns = 0;
    if (sys.mesh_cols < 2 || sys.mesh_rows < 2 || n + S > 65535 || feature_capacity > 65535 || cells > 65535) return false;
    int dev = 0, max_smem = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);

    // column-major copy of the similarity rows (stable: ascending row inside every column).  For the re-tiled kernel
    // every column's list is padded to whole groups of four with zero weights on the column's own (temporal) row.
    std::vector<int> count(n, 0);
    for (int k = 0; k < nnz; k++) count[sys.col[k]]++;
    const char* force_v1 = std::getenv("LVKB200_MESH_V1");
    int padded = 0;
    for (int c = 0; c < n; c++) padded += (count[c] + 3) / 4 * 4;
    // the re-tiled kernel for every sparse system within its per-thread slot limits (LVKB200_MESH_V1=1: the round-1 kernel)
    const v2::Carve2 cv2(n, S, padded, cells, feature_capacity);
    const bool use_v2 = n != 8 && n <= v2::MAX_SLOT_N * v2::T2 && S <= v2::MAX_SLOT_S * v2::T2 &&
                        feature_capacity <= v2::MAX_SLOT_F * v2::T2 && 4 * cells + 1 <= 65535 && (n + S) % 2 == 0 &&
                        cv2.total <= static_cast<size_t>(max_smem) && !(force_v1 && force_v1[0] == '1');
    const int cn = use_v2 ? padded : nnz;  // entries of the column-major copy
    const Carve cv1(n, S, nnz, cells, feature_capacity);
    const size_t smem_needed = use_v2 ? cv2.total : cv1.total;
    if (smem_needed > static_cast<size_t>(max_smem)) return false;
    const int slots_wanted = use_v2 ? v2::slot_variant(n, S, feature_capacity) : -1;
    *err = use_v2 ? v2::set_smem(slots_wanted, smem_needed)
                  : cudaFuncSetAttribute(k_mesh_cgls, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_needed);
    if (*err != cudaSuccess) return false;

    std::vector<int> ptr(n + 1, 0);
    for (int c = 0; c < n; c++) ptr[c + 1] = ptr[c] + (use_v2 ? (count[c] + 3) / 4 * 4 : count[c]);
    std::vector<uint16_t> crow(cn), scol(nnz);
    std::vector<float> cval(cn, 0.0f);
    for (int c = 0; c < n; c++)
        for (int k = ptr[c]; k < ptr[c + 1]; k++) crow[k] = static_cast<uint16_t>(c);  // padding: 0 * r[c]
    std::vector<int> fill(ptr.begin(), ptr.end() - 1);
    for (int r = 0; r < S; r++)
        for (int k = 4 * r; k < 4 * r + 4; k++)
        {
            const int at = fill[sys.col[k]]++;
            crow[at] = static_cast<uint16_t>(n + r);
            cval[at] = sys.val[k];
            scol[k] = static_cast<uint16_t>(sys.col[k]);
        }

    // one blob: [sim_val f32 x nnz | csc_val f32 x cn | csc_ptr i32 x (n+1) | sim_col u16 x nnz | csc_row u16 x cn]
    const size_t o_simval = 0, o_cscval = o_simval + 4 * (size_t)nnz, o_ptr = o_cscval + 4 * (size_t)cn,
                 o_simcol = o_ptr + 4 * ((size_t)n + 1), o_cscrow = o_simcol + 2 * (size_t)nnz,
                 bytes = o_cscrow + 2 * (size_t)cn;
    std::vector<uint8_t> blob(bytes);
    std::memcpy(blob.data() + o_simval, sys.val.data(), 4 * (size_t)nnz);
    std::memcpy(blob.data() + o_cscval, cval.data(), 4 * (size_t)cn);
    std::memcpy(blob.data() + o_ptr, ptr.data(), 4 * ((size_t)n + 1));
    std::memcpy(blob.data() + o_simcol, scol.data(), 2 * (size_t)nnz);
    std::memcpy(blob.data() + o_cscrow, crow.data(), 2 * (size_t)cn);
    if ((*err = d_static.ensure(bytes + 16)) != cudaSuccess) return false;
    if ((*err = d_state.ensure(4 * (size_t)n)) != cudaSuccess) return false;
    const size_t out_bytes = sizeof(MeshSolveResult) + 4 * (size_t)n + 16;
    if ((*err = d_out.ensure(out_bytes)) != cudaSuccess) return false;
    if ((*err = h_out.ensure(out_bytes)) != cudaSuccess) return false;
    std::memset(h_out.ptr, 0, out_bytes);
    // pageable source, completed before `blob` dies; on the stream's own queue (no legacy-stream work: another host thread
    // may be capturing its stream's graph right now).  configure is not on the per-frame path.
    if ((*err = cudaMemcpyAsync(d_static.ptr, blob.data(), bytes, cudaMemcpyHostToDevice, cs)) != cudaSuccess) return false;
    if ((*err = cudaStreamSynchronize(cs)) != cudaSuccess) return false;
    const bool resized = mesh_cols != sys.mesh_cols || mesh_rows != sys.mesh_rows;
    mesh_cols = sys.mesh_cols; mesh_rows = sys.mesh_rows;
    n_sim = S; csc_nnz = cn; capacity = feature_capacity; smem_bytes = smem_needed; slots = slots_wanted;
    n_unknowns = n;
    if (resized && (*err = cudaMemsetAsync(d_state.ptr, 0, 4 * (size_t)n, cs)) != cudaSuccess) { n_unknowns = 0; return false; }
    return true;
}

cudaError_t MeshCgls::reset_state(cudaStream_t cs)
{
    if (!ready()) return cudaSuccess;
    return cudaMemsetAsync(d_state.ptr, 0, 4 * (size_t)n_unknowns, cs);
}

cudaError_t MeshCgls::set_state(cudaStream_t cs, const float* mesh)
{
    cudaError_t e = cudaMemcpyAsync(d_state.ptr, mesh, 4 * (size_t)n_unknowns, cudaMemcpyHostToDevice, cs);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(cs);
}

cudaError_t MeshCgls::launch(cudaStream_t cs, const MeshSolveParams& prm, const float2* d_src, const float2* d_dst,
                             const int* d_n, const TrackParams* d_params, uint8_t* d_mask, const TrackOutCopy& out)
{
    const uint8_t* b = d_static.as<uint8_t>();
    const size_t nnz = 4 * static_cast<size_t>(n_sim), cn = static_cast<size_t>(csc_nnz), n = static_cast<size_t>(n_unknowns);
    MeshSys sys{};
    sys.cols = mesh_cols; sys.rows = mesh_rows; sys.n = n_unknowns; sys.S = n_sim; sys.csc_nnz = csc_nnz; sys.cap = capacity;
    sys.sim_val = reinterpret_cast<const float*>(b);
    sys.csc_val = reinterpret_cast<const float*>(b + 4 * nnz);
    sys.csc_ptr = reinterpret_cast<const int*>(b + 4 * nnz + 4 * cn);
    sys.sim_col = reinterpret_cast<const uint16_t*>(b + 4 * nnz + 4 * cn + 4 * (n + 1));
    sys.csc_row = reinterpret_cast<const uint16_t*>(b + 4 * nnz + 4 * cn + 4 * (n + 1) + 2 * nnz);
    if (slots >= 0)
        v2::launch(slots, smem_bytes, cs, sys, prm, d_src, d_dst, d_n, d_params, d_state.as<float>(), d_mask,
                   d_out.as<uint8_t>(), h_out.device_view<uint8_t>(), out);
    else
        k_mesh_cgls<<<1, T, smem_bytes, cs>>>(sys, prm, d_src, d_dst, d_n, d_params, d_state.as<float>(), d_mask,
                                              d_out.as<uint8_t>(), h_out.device_view<uint8_t>(), out);
    count_launches(1);
    return cudaGetLastError();
}

void MeshCgls::release()
{
    d_static.release(); d_state.release(); d_out.release(); h_out.release();
    n_unknowns = 0; mesh_cols = mesh_rows = 0;
}

}  // namespace lvkb200
