/* CPU restatement of the operator closure on the reference's hot path.
 *
 * TEST INFRASTRUCTURE ONLY (checker + CPU baseline); never linked into or
 * called by the product library.
 *
 * The reference never sees a matrix: its solvers call `op * x`
 * (pykrylov/linop/linop.py:362-369 -> :356-360 -> :271-298) and the closure
 * supplied by the user does the arithmetic (linop.py:697-717 for Pysparse).
 * Pysparse is absent and un-pinned, so the stand-in closure is SciPy's
 * csr_matrix @ v: a sequential, column-sorted, non-fused
 *     sum = 0;  for jj in row:  sum += val[jj] * x[col[jj]]
 * per row.  These loops restate exactly that order; compile with
 * -ffp-contract=off so that gcc cannot fuse the multiply-add.
 * tests/test_oracle.py pins them bit-for-bit against scipy.sparse.
 */
#include <stdint.h>
#include <stddef.h>

/* y = A x, A in CSR (int32 indices). */
void csr_matvec_ref(int64_t nrows, const int32_t *rowptr, const int32_t *col,
                    const double *val, const double *x, double *y)
{
    for (int64_t i = 0; i < nrows; ++i) {
        double sum = 0.0;
        for (int32_t jj = rowptr[i]; jj < rowptr[i + 1]; ++jj)
            sum += val[jj] * x[col[jj]];
        y[i] = sum;
    }
}

/* y = A^T x with A in CSR == CSC product of the transpose: scatter in row
 * order, which is the order scipy's csc_matvec uses for (A.T) @ x. */
void csr_matvec_transp_ref(int64_t nrows, int64_t ncols, const int32_t *rowptr,
                           const int32_t *col, const double *val,
                           const double *x, double *y)
{
    for (int64_t j = 0; j < ncols; ++j) y[j] = 0.0;
    for (int64_t i = 0; i < nrows; ++i) {
        const double xi = x[i];
        for (int32_t jj = rowptr[i]; jj < rowptr[i + 1]; ++jj)
            y[col[jj]] += val[jj] * xi;
    }
}

/* Sequential left-to-right inner product (a fixed-order yardstick for the
 * device reductions; np.dot is the parity target, this is a cross-check). */
double dot_seq_ref(int64_t n, const double *a, const double *b)
{
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}
