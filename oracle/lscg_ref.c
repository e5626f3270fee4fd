/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/easu_ref.c header for the usage rules).
 *
 * CPU restatement of Eigen 3.4's LeastSquaresConjugateGradient<SparseMatrix<float>>::solveWithGuess
 * with the default LeastSquareDiagonalPreconditioner, as called by
 *   LiveVisionKit/Vision/FrameTracker.cpp:274-276 (estimate_local_motions).
 * Eigen is a third-party dependency that is NOT under /root/reference (pinned 3.4,
 * Scripts/setup_deb.sh:133) and is absent from this image, so the published algorithm
 * (Eigen/src/IterativeLinearSolvers/LeastSquareConjugateGradient.h, least_square_conjugate_gradient)
 * is restated:  preconditioned CG on the normal equations, float32 throughout,
 * tolerance = FLT_EPSILON, maxIterations = 2*cols, warm start x0.
 * Parity status: UNPINNED (no Eigen here to compare against); summation order is plain sequential,
 * Eigen's is SIMD-blocked, so agreement with real Eigen is to float rounding, not bit-exact.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>

/* y = A x  (CSR, m rows) */
static void spmv(int m, const int* rp, const int* ci, const float* v, const float* x, float* y)
{
    for (int r = 0; r < m; r++)
    {
        float s = 0.0f;
        for (int k = rp[r]; k < rp[r + 1]; k++) s += v[k] * x[ci[k]];
        y[r] = s;
    }
}

/* y = A^T x  (n cols) */
static void spmv_t(int m, int n, const int* rp, const int* ci, const float* v, const float* x, float* y)
{
    for (int c = 0; c < n; c++) y[c] = 0.0f;
    for (int r = 0; r < m; r++)
        for (int k = rp[r]; k < rp[r + 1]; k++) y[ci[k]] += v[k] * x[r];
}

static float dot(int n, const float* a, const float* b)
{
    float s = 0.0f;
    for (int i = 0; i < n; i++) s += a[i] * b[i];
    return s;
}

/* Returns the iteration count; x holds the guess on entry and the solution on exit. */
int oracle_lscg_solve(int m, int n, const int* row_ptr, const int* col_idx, const float* vals, const float* b,
                      float* x, int max_iters, float tol)
{
    if (max_iters < 0) max_iters = 2 * n;
    if (tol < 0.0f) tol = FLT_EPSILON;

    float* residual = (float*)malloc(sizeof(float) * (size_t)m);
    float* tmp = (float*)malloc(sizeof(float) * (size_t)m);
    float* nres = (float*)malloc(sizeof(float) * (size_t)n);
    float* p = (float*)malloc(sizeof(float) * (size_t)n);
    float* z = (float*)malloc(sizeof(float) * (size_t)n);
    float* invdiag = (float*)malloc(sizeof(float) * (size_t)n);
    int iters = 0;

    /* LeastSquareDiagonalPreconditioner: 1 / ||A_col||^2 (1 when the column is empty). */
    for (int c = 0; c < n; c++) invdiag[c] = 0.0f;
    for (int r = 0; r < m; r++)
        for (int k = row_ptr[r]; k < row_ptr[r + 1]; k++) invdiag[col_idx[k]] += vals[k] * vals[k];
    for (int c = 0; c < n; c++) invdiag[c] = (invdiag[c] > 0.0f) ? 1.0f / invdiag[c] : 1.0f;

    spmv(m, row_ptr, col_idx, vals, x, tmp);
    for (int r = 0; r < m; r++) residual[r] = b[r] - tmp[r];
    spmv_t(m, n, row_ptr, col_idx, vals, residual, nres);

    spmv_t(m, n, row_ptr, col_idx, vals, b, z);
    float rhs_norm2 = dot(n, z, z);
    if (rhs_norm2 == 0.0f)
    {
        for (int c = 0; c < n; c++) x[c] = 0.0f;
        goto done;
    }
    {
        const float threshold = tol * tol * rhs_norm2;
        float res_norm2 = dot(n, nres, nres);
        if (res_norm2 < threshold) goto done;

        for (int c = 0; c < n; c++) p[c] = invdiag[c] * nres[c];
        float abs_new = dot(n, nres, p);
        while (iters < max_iters)
        {
            spmv(m, row_ptr, col_idx, vals, p, tmp);
            float alpha = abs_new / dot(m, tmp, tmp);
            for (int c = 0; c < n; c++) x[c] += alpha * p[c];
            for (int r = 0; r < m; r++) residual[r] -= alpha * tmp[r];
            spmv_t(m, n, row_ptr, col_idx, vals, residual, nres);

            res_norm2 = dot(n, nres, nres);
            if (res_norm2 < threshold) break;

            for (int c = 0; c < n; c++) z[c] = invdiag[c] * nres[c];
            float abs_old = abs_new;
            abs_new = dot(n, nres, z);
            float beta = abs_new / abs_old;
            for (int c = 0; c < n; c++) p[c] = z[c] + beta * p[c];
            iters++;
        }
    }
done:
    free(residual); free(tmp); free(nres); free(p); free(z); free(invdiag);
    return iters;
}
