// Dense inverse of the coarsest AMG level on the device (amg_setup.cpp leaves it here: AmgParams::device_inverse).
//
// The time sweeps want ONE dense product on the coarsest level (2205 rows at C2) instead of two more levels of
// dependent launches, so the set-up has to invert a matrix of a few thousand rows per hierarchy.  Gauss-Jordan on
// [A | I] touches all of it once per column: 2 n^3 flops over 16 n^2 bytes n times -- 172 GB of traffic at n = 2205,
// 5 s on one host core (measured, round 2: 4.96 s of a 6.1 s set-up).  On the device the 78 MB stay in L2 and a
// column step is two small kernels: 2205 x ~20 us.
//
// The arithmetic is the host algorithm's, operation for operation (amg_setup.cpp::dense_inverse, which stays as the
// checked CPU statement of it): partial pivoting with the first maximum, the pivot row scaled by the rounded
// reciprocal, every other row updated as a - f * p with the product and the difference rounded separately (no fused
// multiply-add: the host code is compiled without one), columns left of the pivot skipped.  The result is therefore
// bit-identical to the host's.
#include <cstdint>

#include "amg.cuh"

namespace {

constexpr int DI_T = 256;

// column c: first row >= c with the largest |a[r][c]|; the scaled pivot row goes to prow, the old row c to crow
// (both [2 lda]: the a part, then the inverse part), so that the elimination can run in place
__global__ void __launch_bounds__(DI_T) di_pivot_kernel(const double *__restrict__ a, const double *__restrict__ inv, int n,
                                                       int lda, int c, double *__restrict__ prow, double *__restrict__ crow,
                                                       int *__restrict__ piv_out, int *__restrict__ err)
{
    __shared__ double s_val[DI_T];
    __shared__ int s_idx[DI_T];
    double best = -1.0;
    int bi = n;
    for (int r = c + threadIdx.x; r < n; r += DI_T) {
        const double v = fabs(a[(size_t)r * lda + c]);
        if (v > best) {
            best = v;
            bi = r;
        }
    }
    s_val[threadIdx.x] = best;
    s_idx[threadIdx.x] = bi;
    __syncthreads();
    for (int o = DI_T / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const double v2 = s_val[threadIdx.x + o];
            const int i2 = s_idx[threadIdx.x + o];
            if (v2 > s_val[threadIdx.x] || (v2 == s_val[threadIdx.x] && i2 < s_idx[threadIdx.x])) {
                s_val[threadIdx.x] = v2;
                s_idx[threadIdx.x] = i2;
            }
        }
        __syncthreads();
    }
    const int piv = s_idx[0];
    if (!(s_val[0] > 0.0)) {      // singular (or NaN): report, leave the matrices alone
        if (threadIdx.x == 0) {
            *err = 1;
            *piv_out = -1;
        }
        return;
    }
    if (threadIdx.x == 0) *piv_out = piv;
    const double d = 1.0 / a[(size_t)piv * lda + c];
    const double *ap = a + (size_t)piv * lda, *ip = inv + (size_t)piv * lda;
    const double *ac = a + (size_t)c * lda, *ic = inv + (size_t)c * lda;
    for (int j = threadIdx.x; j < n; j += DI_T) {
        if (j >= c) {
            prow[j] = __dmul_rn(ap[j], d);
            crow[j] = ac[j];
        }
        prow[lda + j] = __dmul_rn(ip[j], d);
        crow[lda + j] = ic[j];
    }
}

// one CTA per row r: row c becomes the scaled pivot row, row piv continues as the old row c (the swap), every row
// but c loses f times the pivot row, f = its entry in column c
__global__ void __launch_bounds__(DI_T) di_eliminate_kernel(double *__restrict__ a, double *__restrict__ inv, int n, int lda,
                                                           int c, const double *__restrict__ prow,
                                                           const double *__restrict__ crow, const int *__restrict__ piv_p)
{
    const int piv = *piv_p;
    if (piv < 0) return;
    const int r = blockIdx.x;
    double *ar = a + (size_t)r * lda, *ir = inv + (size_t)r * lda;
    if (r == c) {
        for (int j = threadIdx.x; j < n; j += DI_T) {
            if (j >= c) ar[j] = prow[j];
            ir[j] = prow[lda + j];
        }
        return;
    }
    const bool from_c = (r == piv);
    const double *sa = from_c ? crow : ar, *si = from_c ? crow + lda : ir;
    const double f = sa[c];
    __syncthreads();      // f is read before anybody overwrites column c of this row
    if (f == 0.0) {      // nothing to eliminate (the host code skips the row); the swapped row still has to arrive
        if (from_c)
            for (int j = threadIdx.x; j < n; j += DI_T) {
                if (j >= c) ar[j] = sa[j];
                ir[j] = si[j];
            }
        return;
    }
    for (int j = threadIdx.x; j < n; j += DI_T) {
        if (j >= c) ar[j] = __dsub_rn(sa[j], __dmul_rn(f, prow[j]));
        ir[j] = __dsub_rn(si[j], __dmul_rn(f, prow[lda + j]));
    }
}

// inv -= e e^T (pseudo-inverse of a singular symmetric matrix with kernel vector e)
__global__ void di_subtract_outer_kernel(double *__restrict__ inv, const double *__restrict__ e, int n, int lda)
{
    const int i = blockIdx.x;
    for (int j = threadIdx.x; j < n; j += blockDim.x) inv[(size_t)i * lda + j] -= e[i] * e[j];
}

}  // namespace

int dense_inverse_device(ctl_handle_s *h, const HostCSR &A, const std::vector<double> &shift, double **Ainv_out, int *lda_out)
{
    const int n = A.n_rows;
    CTL_CHECK(n > 0 && A.n_cols == n && n <= 4096, CTL_ERR_ARG, "dense_inverse_device: square matrix of at most 4096 rows expected");
    CTL_CHECK(shift.empty() || (int)shift.size() == n, CTL_ERR_ARG, "dense_inverse_device: kernel vector of the wrong length");
    const int lda = (n + 1) & ~1;      // even: rows start 16-byte aligned (dense_gemv_kernel)
    std::vector<double> a((size_t)n * lda, 0.0), id((size_t)n * lda, 0.0);
    for (int i = 0; i < n; ++i) {
        id[(size_t)i * lda + i] = 1.0;
        for (int k = A.indptr[i]; k < A.indptr[i + 1]; ++k) a[(size_t)i * lda + A.indices[k]] += A.values[k];
    }
    if (!shift.empty())
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) a[(size_t)i * lda + j] += shift[i] * shift[j];
    double *d_a = nullptr, *d_inv = nullptr, *d_rows = nullptr, *d_e = nullptr;
    int *d_flags = nullptr;      // [0] pivot row of the current column, [1] error
    const size_t bytes = (size_t)n * lda * sizeof(double);
    int rc = CTL_OK;
    auto body = [&]() -> int {
        CTL_CUDA(cudaMalloc((void **)&d_a, bytes));
        CTL_CUDA(cudaMalloc((void **)&d_inv, bytes));
        CTL_CUDA(cudaMalloc((void **)&d_rows, (size_t)4 * lda * sizeof(double)));
        CTL_CUDA(cudaMalloc((void **)&d_flags, 2 * sizeof(int)));
        CTL_CUDA(cudaMemcpyAsync(d_a, a.data(), bytes, cudaMemcpyHostToDevice, h->stream));
        CTL_CUDA(cudaMemcpyAsync(d_inv, id.data(), bytes, cudaMemcpyHostToDevice, h->stream));
        CTL_CUDA(cudaMemsetAsync(d_flags, 0, 2 * sizeof(int), h->stream));
        CTL_CUDA(cudaMemsetAsync(d_rows, 0, (size_t)4 * lda * sizeof(double), h->stream));
        double *prow = d_rows, *crow = d_rows + 2 * lda;
        for (int c = 0; c < n; ++c) {
            di_pivot_kernel<<<1, DI_T, 0, h->stream>>>(d_a, d_inv, n, lda, c, prow, crow, d_flags, d_flags + 1);
            di_eliminate_kernel<<<n, DI_T, 0, h->stream>>>(d_a, d_inv, n, lda, c, prow, crow, d_flags);
        }
        h->launches += 2 * (int64_t)n;
        if (!shift.empty()) {
            CTL_CUDA(cudaMalloc((void **)&d_e, (size_t)n * sizeof(double)));
            CTL_CUDA(cudaMemcpyAsync(d_e, shift.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            di_subtract_outer_kernel<<<n, DI_T, 0, h->stream>>>(d_inv, d_e, n, lda);
            h->launches++;
        }
        int flags[2] = {0, 0};
        CTL_CUDA(cudaMemcpyAsync(flags, d_flags, sizeof(flags), cudaMemcpyDeviceToHost, h->stream));
        CTL_CUDA(cudaStreamSynchronize(h->stream));
        CTL_CUDA(cudaGetLastError());
        CTL_CHECK(flags[1] == 0, CTL_ERR_STATE, "AMG setup: singular coarse matrix");
        return CTL_OK;
    };
    rc = body();
    cudaFree(d_a);
    cudaFree(d_rows);
    cudaFree(d_flags);
    cudaFree(d_e);
    if (rc != CTL_OK) {
        cudaFree(d_inv);
        return rc;
    }
    *Ainv_out = d_inv;
    *lda_out = lda;
    return CTL_OK;
}

int dense_inverse_upload(ctl_handle_s *h, const std::vector<double> &Ainv, int n, double **Ainv_out, int *lda_out)
{
    CTL_CHECK(n > 0 && Ainv.size() == (size_t)n * n, CTL_ERR_ARG, "dense_inverse_upload: n x n values expected");
    const int lda = (n + 1) & ~1;
    std::vector<double> padded((size_t)n * lda, 0.0);
    for (int i = 0; i < n; ++i) std::copy(Ainv.begin() + (size_t)i * n, Ainv.begin() + (size_t)(i + 1) * n, padded.begin() + (size_t)i * lda);
    CTL_TRY(ctl_upload(h, Ainv_out, padded.data(), padded.size()));
    *lda_out = lda;
    return CTL_OK;
}
