// Small host-side numerics on the path (double precision, a few dozen flops per frame).
#pragma once

#include <cmath>
#include <cstring>
#include <utility>

namespace lvkb200
{

// Solves the n x n system A x = b (row-major A, n <= 9) by Gaussian elimination with partial pivoting.
// Returns false when singular.
inline bool solve_dense(int n, double* A, double* b, double* x)
{
    for (int c = 0; c < n; c++)
    {
        int piv = c;
        double best = std::fabs(A[c * n + c]);
        for (int r = c + 1; r < n; r++)
        {
            const double v = std::fabs(A[r * n + c]);
            if (v > best) { best = v; piv = r; }
        }
        if (best < 1e-300) return false;
        if (piv != c)
        {
            for (int k = 0; k < n; k++) std::swap(A[c * n + k], A[piv * n + k]);
            std::swap(b[c], b[piv]);
        }
        const double inv = 1.0 / A[c * n + c];
        for (int r = c + 1; r < n; r++)
        {
            const double f = A[r * n + c] * inv;
            if (f == 0.0) continue;
            for (int k = c; k < n; k++) A[r * n + k] -= f * A[c * n + k];
            b[r] -= f * b[c];
        }
    }
    for (int r = n - 1; r >= 0; r--)
    {
        double s = b[r];
        for (int k = r + 1; k < n; k++) s -= A[r * n + k] * x[k];
        x[r] = s / A[r * n + r];
    }
    return true;
}

// cv::getPerspectiveTransform(src[4], dst[4]) as used by WarpMesh::apply (Math/WarpMesh.cpp:214):
// the 3x3 M (row-major, m[8] = 1) with dst_i ~ M * src_i.
inline bool perspective_transform_4pt(const float src[4][2], const float dst[4][2], double m[9])
{
    double A[64], b[8], x[8];
    std::memset(A, 0, sizeof(A));
    for (int i = 0; i < 4; i++)
    {
        const double sx = src[i][0], sy = src[i][1], dx = dst[i][0], dy = dst[i][1];
        double* r0 = &A[i * 8];
        double* r1 = &A[(i + 4) * 8];
        r0[0] = sx; r0[1] = sy; r0[2] = 1.0; r0[6] = -sx * dx; r0[7] = -sy * dx;
        r1[3] = sx; r1[4] = sy; r1[5] = 1.0; r1[6] = -sx * dy; r1[7] = -sy * dy;
        b[i] = dx;
        b[i + 4] = dy;
    }
    if (!solve_dense(8, A, b, x)) return false;
    for (int k = 0; k < 8; k++) m[k] = x[k];
    m[8] = 1.0;
    return true;
}

// 2x2 branch of WarpMesh::apply (Math/WarpMesh.cpp:196-217): mesh corner offsets (normalized, row-major
// [r][c][xy]) -> dst->src transform handed to the remap kernel.
inline bool mesh2x2_to_transform(const float* offsets, int width, int height, double t[9])
{
    const float w = static_cast<float>(width), h = static_cast<float>(height);
    const float dstc[4][2] = {{0.f, 0.f}, {w, 0.f}, {0.f, h}, {w, h}};
    float srcc[4][2];
    for (int k = 0; k < 4; k++)
    {
        // Point2f * cv::Scalar -> double product narrowed to float (Functions/Extensions.cpp:122-125)
        const float ox = static_cast<float>(static_cast<double>(offsets[2 * k]) * static_cast<double>(width));
        const float oy = static_cast<float>(static_cast<double>(offsets[2 * k + 1]) * static_cast<double>(height));
        srcc[k][0] = dstc[k][0] + ox;
        srcc[k][1] = dstc[k][1] + oy;
    }
    return perspective_transform_4pt(dstc, srcc, t);
}

}  // namespace lvkb200
