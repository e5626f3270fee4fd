// FrameTracker::estimate_local_motions (LiveVisionKit/Vision/FrameTracker.cpp:200-321) and
// generate_mesh_constraints (:380-457): sparse least-squares motion mesh.
// The system is tiny for the library default (2x2 mesh: 8 unknowns, 8 + 2N rows with <= 4 non-zeros each), so the
// preconditioned CG on the normal equations (Eigen::LeastSquaresConjugateGradient semantics: diagonal
// preconditioner, tolerance FLT_EPSILON, max 2*cols iterations, warm start) runs on the host in float32.
// Larger meshes (the 16x16 "Vector Field" preset) are solved on the device by k_mesh_cgls (mesh.cu), which takes its
// static rows and scalars from this class (export_static / device_params) and hands the solution back (adopt_state).
#pragma once

#include <cfloat>
#include <cmath>
#include <vector>

#include "host_logic.hpp"

namespace lvkb200
{

struct SparseRows
{
    std::vector<int> row_ptr{0};
    std::vector<int> col;
    std::vector<float> val;
    int rows() const { return static_cast<int>(row_ptr.size()) - 1; }
    void clear() { row_ptr.assign(1, 0); col.clear(); val.clear(); }
    void begin_row() {}
    void add(int c, float v) { col.push_back(c); val.push_back(v); }
    void end_row() { row_ptr.push_back(static_cast<int>(col.size())); }
    void truncate_rows(int n)
    {
        row_ptr.resize(n + 1);
        col.resize(row_ptr.back());
        val.resize(row_ptr.back());
    }
};

class MeshSolver
{
public:
    void configure(const lvkb200_settings& s)
    {
        const int mw = s.motion_resolution_width, mh = s.motion_resolution_height;
        region_w = static_cast<float>(s.detection_resolution_width);
        region_h = static_cast<float>(s.detection_resolution_height);
        const bool rebuild = mw != cols || mh != rows || A.rows() == 0 || ts != s.temporal_smoothing ||
                             ls != s.local_smoothing || grid_w != region_w || grid_h != region_h;
        ts = s.temporal_smoothing;
        ls = s.local_smoothing;
        acceptance = s.acceptance_threshold;
        if (mw != cols || mh != rows) mesh.assign(static_cast<size_t>(2) * mw * mh, 0.f);
        cols = mw; rows = mh;
        grid_w = region_w; grid_h = region_h;
        // mesh_grid: VirtualGrid(mesh_size, Rect2f(region.tl, (Size2f(mesh)/Size2f(grid)) * region.size))
        const int gw = cols - 1, gh = rows - 1;
        vg.set(cols, rows, 0.f, 0.f, (static_cast<float>(cols) / static_cast<float>(gw)) * region_w,
               (static_cast<float>(rows) / static_cast<float>(gh)) * region_h);
        if (rebuild) generate_constraints();
    }

    void restart() { std::fill(mesh.begin(), mesh.end(), 0.f); }

    std::vector<float>& state() { return mesh; }

    // Returns the motion mesh offsets (rows*cols*2) and the inlier mask.
    void estimate(const std::vector<float>& tracked, const std::vector<float>& matched, Mesh& offsets,
                  std::vector<uint8_t>& inliers, int* iterations = nullptr)
    {
        const int n = static_cast<int>(tracked.size() / 2);
        const int gw = cols - 1, gh = rows - 1;
        const int ncols = 2 * cols * rows;
        b.assign(static_cast<size_t>(static_count) + 2 * n, 0.f);
        for (int k = 0; k < ncols; k++) b[k] = ts * mesh[k];
        A.truncate_rows(static_count);
        for (int i = 0; i < n; i++)
        {
            const float sx = tracked[2 * i], sy = tracked[2 * i + 1];
            size_t kx, ky;
            vg.key_of(sx, sy, kx, ky);
            int k00x = std::min(std::max(static_cast<int>(kx), 0), gw);
            int k00y = std::min(std::max(static_cast<int>(ky), 0), gh);
            const int k11x = k00x + 1, k11y = k00y + 1;
            const int i00 = 2 * (k00y * cols + k00x), i11 = 2 * (k11y * cols + k11x);
            const int i10 = i00 + 2, i01 = i11 - 2;
            // barycentric_rect(Rect2f(p0, p1), src) — Functions/Math.tpp:247-265
            const float p0x = static_cast<float>(k00x) * vg.kw, p0y = static_cast<float>(k00y) * vg.kh;
            const float p1x = static_cast<float>(k11x) * vg.kw, p1y = static_cast<float>(k11y) * vg.kh;
            const float rx = std::min(p0x, p1x), ry = std::min(p0y, p1y);
            const float rw = std::max(p0x, p1x) - rx, rh = std::max(p0y, p1y) - ry;
            const float inv_area = 1.0f / (rw * rh);
            const float x1 = rx, x2 = rx + rw, y1 = ry, y2 = ry + rh;
            const float rx1 = x2 - sx, ry1 = y2 - sy, rx2 = sx - x1, ry2 = sy - y1;
            const float w0 = rx1 * ry1 * inv_area, w1 = rx1 * ry2 * inv_area, w2 = rx2 * ry2 * inv_area,
                        w3 = rx2 * ry1 * inv_area;
            add_row4(i00, w0, i01, w1, i11, w2, i10, w3);
            add_row4(i00 + 1, w0, i01 + 1, w1, i11 + 1, w2, i10 + 1, w3);
            b[static_count + 2 * i] = matched[2 * i];
            b[static_count + 2 * i + 1] = matched[2 * i + 1];
        }
        const int iters = lscg(ncols);
        if (iterations) *iterations = iters;

        inliers.resize(n);
        for (int i = 0; i < n; i++)
        {
            const int rxi = static_count + 2 * i, ryi = rxi + 1;
            float x = 0.f, y = 0.f;
            // triplet order of the reference: i00, i01, i11, i10 (row storage here is column-sorted; re-evaluate in
            // the reference order from the stored weights)
            eval_row(rxi, x);
            eval_row(ryi, y);
            const float err = std::fabs(x - b[rxi]) + std::fabs(y - b[ryi]);
            inliers[i] = err < acceptance ? 1 : 0;
        }
        offsets_from_state(offsets);
    }

    // The motion mesh as normalised offsets (FrameTracker.cpp:303-318) from the current solution.
    void offsets_from_state(Mesh& offsets) const
    {
        offsets.resize(static_cast<size_t>(2) * cols * rows);
        for (int r = 0; r < rows; r++)
            for (int c = 0; c < cols; c++)
            {
                const size_t k = (static_cast<size_t>(r) * cols + c) * 2;
                const float ax = static_cast<float>(c) * vg.kw, ay = static_cast<float>(r) * vg.kh;
                offsets[k] = (ax - mesh[k]) / region_w;
                offsets[k + 1] = (ay - mesh[k + 1]) / region_h;
            }
    }

    // ---- hand-over to the device solver (mesh.cu)
    int unknowns() const { return 2 * cols * rows; }
    int mesh_cols() const { return cols; }
    int mesh_rows() const { return rows; }
    float temporal_weight() const { return ts; }
    float acceptance_threshold() const { return acceptance; }
    float key_w() const { return vg.kw; }
    float key_h() const { return vg.kh; }
    // Bumped whenever generate_constraints() ran: the device copy of the static rows is stale.
    unsigned generation() const { return static_generation; }
    // The similarity rows (everything static except the diagonal temporal rows), four non-zeros each.
    void export_static(std::vector<int>& col4, std::vector<float>& val4) const
    {
        col4.clear(); val4.clear();
        for (int r = 2 * cols * rows; r < static_count; r++)
            for (int k = A.row_ptr[r]; k < A.row_ptr[r + 1]; k++)
            {
                col4.push_back(A.col[k]);
                val4.push_back(A.val[k]);
            }
    }
    void adopt_state(const float* solution) { std::copy(solution, solution + mesh.size(), mesh.begin()); }

private:
    void add_row4(int c0, float v0, int c1, float v1, int c2, float v2, int c3, float v3)
    {
        // Eigen setFromTriplets sums duplicates; the four vertices of a cell are distinct, keep insertion order
        A.add(c0, v0); A.add(c1, v1); A.add(c2, v2); A.add(c3, v3);
        A.end_row();
    }

    void eval_row(int r, float& out) const
    {
        float s = 0.f;
        bool first = true;
        for (int k = A.row_ptr[r]; k < A.row_ptr[r + 1]; k++)
        {
            const float t = A.val[k] * mesh[A.col[k]];
            s = first ? t : s + t;
            first = false;
        }
        out = s;
    }

    void generate_constraints()
    {
        A.clear();
        int index = 0;
        for (int r = 0; r < rows; r++)
            for (int c = 0; c < cols; c++, index++)
            {
                A.add(2 * index, ts); A.end_row();
                A.add(2 * index + 1, ts); A.end_row();
            }
        const double v1 = -(static_cast<double>(vg.kw) / static_cast<double>(vg.kh));  // -key_size().aspectRatio()
        const double v2 = -1.0 / v1;
        index = 0;
        for (int r = 0; r < rows; r++)
            for (int c = 0; c < cols; c++, index++)
            {
                int quad = 1;
                if (c % 4 == 0 && r % 4 == 0) quad = 3;
                else if ((c + r) % 2 != 1 && c != 0 && r != 0 && c != cols - 2 && r != rows - 2) continue;
                if (c >= cols - quad || r >= rows - quad) continue;
                const int i00 = 2 * index, i10 = i00 + 2 * quad;
                const int i01 = 2 * (index + quad * cols), i11 = i01 + 2 * quad;
                const float weight = ls;
                const float w1 = static_cast<float>(v1 * weight), w2 = static_cast<float>(v2 * weight);
                A.add(i00, -weight); A.add(i01, weight); A.add(i01 + 1, -w2); A.add(i11 + 1, w2); A.end_row();
                A.add(i00 + 1, -weight); A.add(i01, w2); A.add(i01 + 1, weight); A.add(i11, -w2); A.end_row();
                A.add(i00, -weight); A.add(i10, weight); A.add(i10 + 1, -w1); A.add(i11 + 1, w1); A.end_row();
                A.add(i00 + 1, -weight); A.add(i10, w1); A.add(i10 + 1, weight); A.add(i11, -w1); A.end_row();
            }
        static_count = A.rows();
        static_generation++;
    }

    void spmv(const std::vector<float>& x, std::vector<float>& y) const
    {
        const int m = A.rows();
        y.resize(m);
        for (int r = 0; r < m; r++)
        {
            float s = 0.f;
            for (int k = A.row_ptr[r]; k < A.row_ptr[r + 1]; k++) s += A.val[k] * x[A.col[k]];
            y[r] = s;
        }
    }
    void spmv_t(const std::vector<float>& x, std::vector<float>& y, int n) const
    {
        y.assign(n, 0.f);
        const int m = A.rows();
        for (int r = 0; r < m; r++)
            for (int k = A.row_ptr[r]; k < A.row_ptr[r + 1]; k++) y[A.col[k]] += A.val[k] * x[r];
    }
    static float dot(const std::vector<float>& a, const std::vector<float>& c)
    {
        float s = 0.f;
        for (size_t i = 0; i < a.size(); i++) s += a[i] * c[i];
        return s;
    }

    // Eigen least_square_conjugate_gradient with LeastSquareDiagonalPreconditioner, solveWithGuess(b, mesh)
    int lscg(int n)
    {
        const int m = A.rows();
        const int max_iters = 2 * n;
        const float tol = FLT_EPSILON;
        std::vector<float> invdiag(n, 0.f);
        for (int r = 0; r < m; r++)
            for (int k = A.row_ptr[r]; k < A.row_ptr[r + 1]; k++) invdiag[A.col[k]] += A.val[k] * A.val[k];
        for (int c = 0; c < n; c++) invdiag[c] = invdiag[c] > 0.f ? 1.0f / invdiag[c] : 1.0f;

        std::vector<float>& x = mesh;
        spmv(x, tmp);
        residual.resize(m);
        for (int r = 0; r < m; r++) residual[r] = b[r] - tmp[r];
        spmv_t(residual, nres, n);
        spmv_t(b, z, n);
        const float rhs_norm2 = dot(z, z);
        if (rhs_norm2 == 0.f)
        {
            std::fill(x.begin(), x.end(), 0.f);
            return 0;
        }
        const float threshold = tol * tol * rhs_norm2;
        float res_norm2 = dot(nres, nres);
        if (res_norm2 < threshold) return 0;
        p.resize(n);
        for (int c = 0; c < n; c++) p[c] = invdiag[c] * nres[c];
        float abs_new = dot(nres, p);
        int i = 0;
        while (i < max_iters)
        {
            spmv(p, tmp);
            const float alpha = abs_new / dot(tmp, tmp);
            for (int c = 0; c < n; c++) x[c] += alpha * p[c];
            for (int r = 0; r < m; r++) residual[r] -= alpha * tmp[r];
            spmv_t(residual, nres, n);
            res_norm2 = dot(nres, nres);
            if (res_norm2 < threshold) break;
            z.resize(n);
            for (int c = 0; c < n; c++) z[c] = invdiag[c] * nres[c];
            const float abs_old = abs_new;
            abs_new = dot(nres, z);
            const float beta = abs_new / abs_old;
            for (int c = 0; c < n; c++) p[c] = z[c] + beta * p[c];
            i++;
        }
        return i;
    }

    int cols = 0, rows = 0, static_count = 0;
    unsigned static_generation = 0;
    float ts = 1.f, ls = 20.f, acceptance = 8.f, region_w = 0, region_h = 0, grid_w = -1, grid_h = -1;
    VGrid vg;
    SparseRows A;
    std::vector<float> mesh, b, tmp, residual, nres, z, p;
};

}  // namespace lvkb200
