/* OpenMP sparse products for the CPU oracle (test infrastructure / CPU baseline only).
 *
 * The reference executes every product of the hot path through PETSc's MatMult /
 * MatMultAdd on the CPU (preconditioner/preconditioner.py:406-432; control/control.py:2021
 * ff. through Firedrake's assemble(action(...))), one MPI rank per core.  The numpy/scipy
 * oracle is single threaded; these two loops let its sparse products use all host cores
 * so that the CPU baseline of bench.py is not a strawman.  Results are identical to
 * scipy's row-wise accumulation order (each row is summed sequentially by one thread). */
#include <stdint.h>

void oracle_csr_matvec(int32_t n, const int32_t *indptr, const int32_t *indices, const double *data,
                       const double *x, double *y)
{
#pragma omp parallel for schedule(static)
    for (int32_t r = 0; r < n; ++r) {
        double acc = 0.0;
        for (int32_t k = indptr[r]; k < indptr[r + 1]; ++k) acc += data[k] * x[indices[k]];
        y[r] = acc;
    }
}

/* Y[n x m] = A X[n_cols x m], row-major dense blocks */
void oracle_csr_matmat(int32_t n, int32_t m, const int32_t *indptr, const int32_t *indices,
                       const double *data, const double *X, double *Y)
{
#pragma omp parallel for schedule(static)
    for (int32_t r = 0; r < n; ++r) {
        double *yr = Y + (int64_t)r * m;
        for (int32_t j = 0; j < m; ++j) yr[j] = 0.0;
        for (int32_t k = indptr[r]; k < indptr[r + 1]; ++k) {
            const double a = data[k];
            const double *xr = X + (int64_t)indices[k] * m;
            for (int32_t j = 0; j < m; ++j) yr[j] += a * xr[j];
        }
    }
}
