// SELL-32 construction and the single-column SpMV kernel family (see sell.cuh).
// All kernels: one thread per row, 128 threads per CTA, HBM/L2-bound streaming of
// (4 + 8) bytes per stored entry plus the gathered x.
#include <algorithm>
#include <cstdlib>

#include "sell.cuh"

SellPattern::~SellPattern()
{
    cudaFree(slice_ptr);
    cudaFree(cols);
}

void sell_free(SellMat &m)
{
    cudaFree(m.vals);
    cudaFree(m.csr_ptr);
    cudaFree(m.csr_cols);
    cudaFree(m.csr_vals);
    m.vals = m.csr_vals = nullptr;
    m.csr_ptr = m.csr_cols = nullptr;
    m.lanes = 0;
    m.pat.reset();
}

int sell_from_csr(ctl_handle_s *h, const HostCSR &A, SellMat &out, bool force_csr)
{
    const double mean = A.n_rows ? (double)A.nnz() / A.n_rows : 0.0;
    // SELL-32 (one row per thread, coalesced) up to a mean row length of 24: measured at C2 against the
    // CSR-vector kernels, inner solve 0.842 ms (threshold 10) -> 0.792 (14) -> 0.774 (20-24) -> 0.779 (40)
    double sell_max_mean = 24.0;
    if (const char *e = getenv("CTL_SELL_MAX_MEAN")) sell_max_mean = atof(e);      // experiment
    // ... but only for levels with enough rows to fill the GPU with one row per thread: below ~50 k rows
    // the CSR-vector kernels (several threads per row) win (19 k-row level as SELL: 0.833 ms per solve)
    int sell_min_rows = 50000;
    if (const char *e = getenv("CTL_SELL_MIN_ROWS")) sell_min_rows = atoi(e);      // experiment
    const bool mesh_like = mean <= 10.0;          // fine-mesh stencils (short rows): always SELL
    if ((mesh_like || (mean <= sell_max_mean && A.n_rows >= sell_min_rows)) && !force_csr) {
        std::shared_ptr<SellPattern> pat;
        CTL_TRY(sell_build_pattern(h, A, pat));
        return sell_set_values(h, pat, A.values.data(), out);
    }
    auto pat = std::make_shared<SellPattern>();     // sizes only; no SELL arrays
    pat->n_rows = A.n_rows;
    pat->n_cols = A.n_cols;
    pat->nnz = A.nnz();
    out.pat = pat;
    // lanes per row: about a quarter of the mean row length.  Measured at C2 (inner solve, warm): 8 lanes
    // for the 13-entry rows of level 1 = 0.918 ms, 16 lanes = 1.189 ms, 4 lanes = 0.843 ms: the narrow
    // groups keep four times as many rows in flight, which is what these latency-bound kernels need.
    out.lanes = mean <= 16.0 ? 4 : (mean <= 48.0 ? 8 : 16);
    if (const char *e = getenv("CTL_CSR_LANES_SHIFT")) {          // experiment: wider / narrower row groups
        const int sh = atoi(e);
        out.lanes = std::max(2, std::min(32, sh >= 0 ? out.lanes << sh : out.lanes >> (-sh)));
    }
    CTL_TRY(ctl_upload(h, &out.csr_ptr, A.indptr.data(), A.indptr.size()));
    CTL_TRY(ctl_upload(h, &out.csr_cols, A.indices.data(), A.indices.size()));
    CTL_TRY(ctl_upload(h, &out.csr_vals, A.values.data(), A.values.size()));
    return CTL_OK;
}

int sell_build_pattern(ctl_handle_s *h, const HostCSR &A, std::shared_ptr<SellPattern> &out)
{
    auto p = std::make_shared<SellPattern>();
    p->n_rows = A.n_rows;
    p->n_cols = A.n_cols;
    p->n_slices = ceil_div(A.n_rows, 32);
    p->nnz = A.nnz();
    std::vector<int> sptr(p->n_slices + 1, 0);
    for (int s = 0; s < p->n_slices; ++s) {
        int w = 0;
        for (int r = 32 * s; r < std::min(A.n_rows, 32 * s + 32); ++r) w = std::max(w, A.indptr[r + 1] - A.indptr[r]);
        sptr[s + 1] = sptr[s] + 32 * w;
    }
    p->n_stored = sptr[p->n_slices];
    std::vector<int> cols((size_t)p->n_stored);
    p->csr_to_sell.resize(A.nnz());
    for (int s = 0; s < p->n_slices; ++s) {
        const int w = (sptr[s + 1] - sptr[s]) / 32;
        for (int lane = 0; lane < 32; ++lane) {
            const int r = 32 * s + lane;
            const int len = r < A.n_rows ? A.indptr[r + 1] - A.indptr[r] : 0;
            for (int k = 0; k < w; ++k) {
                const int64_t pos = (int64_t)sptr[s] + 32 * k + lane;
                if (k < len) {
                    cols[pos] = A.indices[A.indptr[r] + k];
                    p->csr_to_sell[A.indptr[r] + k] = pos;
                } else {
                    cols[pos] = r < A.n_rows ? std::min(r, A.n_cols - 1) : 0;
                }
            }
        }
    }
    CTL_TRY(ctl_upload(h, &p->slice_ptr, sptr.data(), sptr.size()));
    CTL_TRY(ctl_upload(h, &p->cols, cols.data(), cols.size()));
    out = p;
    return CTL_OK;
}

int sell_set_values(ctl_handle_s *h, const std::shared_ptr<SellPattern> &pat, const double *csr_values,
                    SellMat &out)
{
    std::vector<double> v((size_t)pat->n_stored, 0.0);
    for (size_t k = 0; k < pat->csr_to_sell.size(); ++k) v[pat->csr_to_sell[k]] = csr_values[k];
    out.pat = pat;
    CTL_TRY(ctl_upload(h, &out.vals, v.data(), v.size()));
    return CTL_OK;
}

namespace {

constexpr int ST = 128;

// One row of a SELL-32 slice.  The slice width is uniform across the warp, so the loop is
// divergence free; entries are fetched in chunks of SC with all index/value loads issued
// before the dependent gathers of x (memory-level parallelism instead of a serial
// load -> gather -> fma chain per entry).
constexpr int SC = 8;
__device__ __forceinline__ double sell_row_dot(const int *__restrict__ slice_ptr, const int *__restrict__ cols,
                                               const double *__restrict__ vals, const double *__restrict__ x,
                                               int row)
{
    const int s = row >> 5, lane = row & 31;
    const int beg = __ldg(slice_ptr + s) + lane, end = __ldg(slice_ptr + s + 1);
    double acc = 0.0;
    for (int p0 = beg; p0 < end; p0 += 32 * SC) {
        int c[SC];
        double v[SC];
#pragma unroll
        for (int j = 0; j < SC; ++j) {
            const int p = p0 + 32 * j;
            const bool ok = p < end;
            // streamed once per pass: evict-first, so that the 88 MB fine matrix does not flush the
            // coarse levels and the vectors out of L2 between two uses
            c[j] = ok ? __ldcs(cols + p) : 0;
            v[j] = ok ? __ldcs(vals + p) : 0.0;
        }
        double xv[SC];
#pragma unroll
        for (int j = 0; j < SC; ++j) xv[j] = __ldg(x + c[j]);
#pragma unroll
        for (int j = 0; j < SC; ++j) acc = fma(v[j], xv[j], acc);
    }
    return acc;
}

template <int MODE>
__global__ void __launch_bounds__(ST) sell_spmv_kernel(const int *__restrict__ slice_ptr, const int *__restrict__ cols,
                                                      const double *__restrict__ vals, const double *__restrict__ x,
                                                      const double *b, double *y, int n_rows)
{
    pdl_sync();
    const int row = blockIdx.x * ST + threadIdx.x;
    if (row >= n_rows) return;
    const double ax = sell_row_dot(slice_ptr, cols, vals, x, row);
    if (MODE == SELL_ASSIGN) y[row] = ax;
    else if (MODE == SELL_RESIDUAL) y[row] = b[row] - ax;
    else if (MODE == SELL_ADD) y[row] += ax;
    else y[row] -= ax;
}

__global__ void __launch_bounds__(ST) sell_cheb_kernel(const int *__restrict__ slice_ptr, const int *__restrict__ cols,
                                                      const double *__restrict__ vals, const double *__restrict__ dinv,
                                                      const double *__restrict__ b, const double *p_prev,
                                                      const double *__restrict__ p_cur, double *out, double a,
                                                      double bq, double c, int n_rows)
{
    pdl_sync();
    const int row = blockIdx.x * ST + threadIdx.x;
    if (row >= n_rows) return;
    const double ax = sell_row_dot(slice_ptr, cols, vals, p_cur, row);
    double r = bq * p_cur[row] + c * dinv[row] * (b[row] - ax);
    if (a != 0.0) r = fma(a, p_prev[row], r);
    out[row] = r;
}

__global__ void __launch_bounds__(ST) dinv_scale_kernel(const double *__restrict__ dinv, const double *__restrict__ b,
                                                       double *__restrict__ out, double c, int n)
{
    pdl_sync();
    const int i = blockIdx.x * ST + threadIdx.x;
    if (i < n) out[i] = c * dinv[i] * b[i];
}

// one warp per row of a small dense matrix
__global__ void __launch_bounds__(ST) dense_gemv_kernel(const double *__restrict__ A, const double *__restrict__ b,
                                                       double *__restrict__ y, int n)
{
    pdl_sync();
    const int row = blockIdx.x * (ST / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    double acc = 0.0;
    for (int j = lane; j < n; j += 32) acc = fma(A[(size_t)row * n + j], b[j], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) y[row] = acc;
}

__global__ void __launch_bounds__(ST) sell_spmv2_kernel(const int *__restrict__ slice_ptr, const int *__restrict__ cols,
                                                       const double *__restrict__ v1, const double *__restrict__ v2,
                                                       const double *__restrict__ x1, const double *__restrict__ x2,
                                                       const double *__restrict__ x3, double *__restrict__ y,
                                                       double alpha, double beta, int n_rows)
{
    pdl_sync();
    const int row = blockIdx.x * ST + threadIdx.x;
    if (row >= n_rows) return;
    const int s = row >> 5, lane = row & 31;
    const int beg = __ldg(slice_ptr + s), end = __ldg(slice_ptr + s + 1);
    double acc1 = 0.0, acc2 = 0.0;
    for (int p = beg + lane; p < end; p += 32) {
        const int c = __ldcs(cols + p);
        double xa = __ldg(x1 + c);
        if (x2) xa += __ldg(x2 + c);
        acc1 = fma(__ldcs(v1 + p), xa, acc1);
        if (x3) acc2 = fma(__ldcs(v2 + p), __ldg(x3 + c), acc2);
    }
    y[row] = alpha * acc1 + beta * acc2;
}

// ---- CSR-vector variants: T lanes per row, shuffle reduction
// STREAM: the matrix is large and read once per pass (fine-level restriction): evict-first loads
template <int T, bool STREAM>
__device__ __forceinline__ double csr_row_dot(const int *__restrict__ ptr, const int *__restrict__ cols,
                                              const double *__restrict__ vals, const double *__restrict__ x,
                                              int row, int lane)
{
    const int beg = __ldg(ptr + row), end = __ldg(ptr + row + 1);
    double acc = 0.0;
    for (int p = beg + lane; p < end; p += T) {
        const int c = STREAM ? __ldcs(cols + p) : __ldg(cols + p);
        const double v = STREAM ? __ldcs(vals + p) : __ldg(vals + p);
        acc = fma(v, __ldg(x + c), acc);
    }
#pragma unroll
    for (int o = T / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o, T);
    return acc;
}

template <int MODE, int T, bool STREAM>
__global__ void __launch_bounds__(ST) csrv_spmv_kernel(const int *__restrict__ ptr, const int *__restrict__ cols,
                                                      const double *__restrict__ vals, const double *__restrict__ x,
                                                      const double *b, double *y, int n_rows)
{
    pdl_sync();
    const int t = blockIdx.x * ST + threadIdx.x;
    const int row = t / T, lane = t % T;
    const int r = row < n_rows ? row : n_rows - 1;       // whole warp takes part in the shuffles
    const double ax = csr_row_dot<T, STREAM>(ptr, cols, vals, x, r, lane);
    if (row >= n_rows || lane != 0) return;
    if (MODE == SELL_ASSIGN) y[row] = ax;
    else if (MODE == SELL_RESIDUAL) y[row] = b[row] - ax;
    else if (MODE == SELL_ADD) y[row] += ax;
    else y[row] -= ax;
}

template <int T>
__global__ void __launch_bounds__(ST) csrv_cheb_kernel(const int *__restrict__ ptr, const int *__restrict__ cols,
                                                      const double *__restrict__ vals, const double *__restrict__ dinv,
                                                      const double *__restrict__ b, const double *p_prev,
                                                      const double *__restrict__ p_cur, double *out, double a,
                                                      double bq, double c, int n_rows)
{
    pdl_sync();
    const int t = blockIdx.x * ST + threadIdx.x;
    const int row = t / T, lane = t % T;
    const int r = row < n_rows ? row : n_rows - 1;
    const double ax = csr_row_dot<T, false>(ptr, cols, vals, p_cur, r, lane);
    if (row >= n_rows || lane != 0) return;
    double v = bq * p_cur[row] + c * dinv[row] * (b[row] - ax);
    if (a != 0.0) v = fma(a, p_prev[row], v);
    out[row] = v;
}

template <int T, bool STREAM>
void launch_csrv_spmv_s(ctl_handle_s *h, const SellMat &A, const double *x, double *y, const double *b, int mode)
{
    const int n = A.pat->n_rows, blocks = ceil_div((int64_t)n * T, ST);
    switch (mode) {
    case SELL_ASSIGN: pdl_launch(h, blocks, ST, csrv_spmv_kernel<SELL_ASSIGN, T, STREAM>, A.csr_ptr, A.csr_cols, A.csr_vals, x, b, y, n); break;
    case SELL_RESIDUAL: pdl_launch(h, blocks, ST, csrv_spmv_kernel<SELL_RESIDUAL, T, STREAM>, A.csr_ptr, A.csr_cols, A.csr_vals, x, b, y, n); break;
    case SELL_ADD: pdl_launch(h, blocks, ST, csrv_spmv_kernel<SELL_ADD, T, STREAM>, A.csr_ptr, A.csr_cols, A.csr_vals, x, b, y, n); break;
    default: pdl_launch(h, blocks, ST, csrv_spmv_kernel<SELL_SUB, T, STREAM>, A.csr_ptr, A.csr_cols, A.csr_vals, x, b, y, n); break;
    }
}

template <int T>
void launch_csrv_spmv(ctl_handle_s *h, const SellMat &A, const double *x, double *y, const double *b, int mode)
{
    // matrices above 32 MB (fine-level restriction) stream through L2 with evict-first
    if (A.pat->nnz * 12 > (32ll << 20)) launch_csrv_spmv_s<T, true>(h, A, x, y, b, mode);
    else launch_csrv_spmv_s<T, false>(h, A, x, y, b, mode);
}

}  // namespace

int sell_spmv(ctl_handle_s *h, const SellMat &A, const double *x, double *y, const double *b, int mode)
{
    if (h->recorder) {
        CTL_CHECK(A.lanes > 0, CTL_ERR_STATE, "fused tail: matrix is not in CSR form");
        FusedOp op{};
        op.type = FOP_SPMV; op.n = A.pat->n_rows; op.lanes = A.lanes; op.mode = mode;
        op.ptr = A.csr_ptr; op.cols = A.csr_cols; op.vals = A.csr_vals;
        op.cur = x; op.b = b; op.out = y;
        h->recorder->host.push_back(op);
        return CTL_OK;
    }
    const SellPattern &p = *A.pat;
    const int blocks = ceil_div(p.n_rows, ST);
    if (blocks == 0) return CTL_OK;
    if (A.lanes) {
        if (A.lanes == 2) launch_csrv_spmv<2>(h, A, x, y, b, mode);
        else if (A.lanes == 4) launch_csrv_spmv<4>(h, A, x, y, b, mode);
        else if (A.lanes == 8) launch_csrv_spmv<8>(h, A, x, y, b, mode);
        else if (A.lanes == 16) launch_csrv_spmv<16>(h, A, x, y, b, mode);
        else launch_csrv_spmv<32>(h, A, x, y, b, mode);
        h->launches++;
        CTL_CUDA(cudaGetLastError());
        return CTL_OK;
    }
    switch (mode) {
    case SELL_ASSIGN: pdl_launch(h, blocks, ST, sell_spmv_kernel<SELL_ASSIGN>, p.slice_ptr, p.cols, A.vals, x, b, y, p.n_rows); break;
    case SELL_RESIDUAL: pdl_launch(h, blocks, ST, sell_spmv_kernel<SELL_RESIDUAL>, p.slice_ptr, p.cols, A.vals, x, b, y, p.n_rows); break;
    case SELL_ADD: pdl_launch(h, blocks, ST, sell_spmv_kernel<SELL_ADD>, p.slice_ptr, p.cols, A.vals, x, b, y, p.n_rows); break;
    default: pdl_launch(h, blocks, ST, sell_spmv_kernel<SELL_SUB>, p.slice_ptr, p.cols, A.vals, x, b, y, p.n_rows); break;
    }
    h->launches++;
    CTL_CUDA(cudaGetLastError());
    return CTL_OK;
}

int sell_cheb_step(ctl_handle_s *h, const SellMat &A, const double *dinv, const double *b,
                   const double *p_prev, const double *p_cur, double *out, double a, double bq, double c)
{
    if (h->recorder) {
        CTL_CHECK(A.lanes > 0, CTL_ERR_STATE, "fused tail: matrix is not in CSR form");
        FusedOp op{};
        op.type = FOP_CHEB; op.n = A.pat->n_rows; op.lanes = A.lanes;
        op.ptr = A.csr_ptr; op.cols = A.csr_cols; op.vals = A.csr_vals;
        op.dinv = dinv; op.b = b; op.prev = p_prev; op.cur = p_cur; op.out = out;
        op.a = a; op.bq = bq; op.c = c;
        h->recorder->host.push_back(op);
        return CTL_OK;
    }
    const SellPattern &p = *A.pat;
    const int blocks = ceil_div(p.n_rows, ST);
    if (blocks == 0) return CTL_OK;
    if (A.lanes) {
        const int n = p.n_rows;
        if (A.lanes == 2)
            pdl_launch(h, ceil_div((int64_t)n * 2, ST), ST, csrv_cheb_kernel<2>, A.csr_ptr, A.csr_cols, A.csr_vals, dinv, b, p_prev, p_cur, out, a, bq, c, n);
        else if (A.lanes == 4)
            pdl_launch(h, ceil_div((int64_t)n * 4, ST), ST, csrv_cheb_kernel<4>, A.csr_ptr, A.csr_cols, A.csr_vals, dinv, b, p_prev, p_cur, out, a, bq, c, n);
        else if (A.lanes == 8)
            pdl_launch(h, ceil_div((int64_t)n * 8, ST), ST, csrv_cheb_kernel<8>, A.csr_ptr, A.csr_cols, A.csr_vals, dinv, b, p_prev, p_cur, out, a, bq, c, n);
        else if (A.lanes == 16)
            pdl_launch(h, ceil_div((int64_t)n * 16, ST), ST, csrv_cheb_kernel<16>, A.csr_ptr, A.csr_cols, A.csr_vals, dinv, b, p_prev, p_cur, out, a, bq, c, n);
        else
            pdl_launch(h, ceil_div((int64_t)n * 32, ST), ST, csrv_cheb_kernel<32>, A.csr_ptr, A.csr_cols, A.csr_vals, dinv, b, p_prev, p_cur, out, a, bq, c, n);
        h->launches++;
        CTL_CUDA(cudaGetLastError());
        return CTL_OK;
    }
    pdl_launch(h, blocks, ST, sell_cheb_kernel, p.slice_ptr, p.cols, A.vals, dinv, b, p_prev, p_cur, out, a,
                                                   bq, c, p.n_rows);
    h->launches++;
    CTL_CUDA(cudaGetLastError());
    return CTL_OK;
}

int vec_dinv_scale(ctl_handle_s *h, const double *dinv, const double *b, double *out, double c, int n)
{
    if (h->recorder) {
        FusedOp op{};
        op.type = FOP_DINV_SCALE; op.n = n; op.dinv = dinv; op.b = b; op.out = out; op.c = c;
        h->recorder->host.push_back(op);
        return CTL_OK;
    }
    if (n == 0) return CTL_OK;
    pdl_launch(h, ceil_div(n, ST), ST, dinv_scale_kernel, dinv, b, out, c, n);
    h->launches++;
    CTL_CUDA(cudaGetLastError());
    return CTL_OK;
}

int dense_gemv(ctl_handle_s *h, const double *Ainv, const double *b, double *y, int n)
{
    if (h->recorder) {
        FusedOp op{};
        op.type = FOP_GEMV; op.n = n; op.vals = Ainv; op.b = b; op.out = y;
        h->recorder->host.push_back(op);
        return CTL_OK;
    }
    if (n == 0) return CTL_OK;
    pdl_launch(h, ceil_div(n, ST / 32), ST, dense_gemv_kernel, Ainv, b, y, n);
    h->launches++;
    CTL_CUDA(cudaGetLastError());
    return CTL_OK;
}

int sell_spmv2(ctl_handle_s *h, const SellMat &A1, const SellMat &A2, const double *x1, const double *x2,
               const double *x3, double *y, double alpha, double beta)
{
    const SellPattern &p = *A1.pat;
    const int blocks = ceil_div(p.n_rows, ST);
    if (blocks == 0) return CTL_OK;
    pdl_launch(h, blocks, ST, sell_spmv2_kernel, p.slice_ptr, p.cols, A1.vals, A2.vals, x1, x2, x3, y, alpha,
                                                    beta, p.n_rows);
    h->launches++;
    CTL_CUDA(cudaGetLastError());
    return CTL_OK;
}

// ---------------------------------------------------------------------------------------
// Fused coarse tail.  Below the second AMG level every operation is a few microseconds of
// dependent L2 round trips on a few thousand rows, and a V-cycle issues about ten of them
// per level: as separate kernels (even inside a CUDA graph) each costs about 5 us.  The
// recorded program of the sub-cycle (same primitives, same order, same arithmetic) runs as
// ONE cooperative kernel with a grid barrier between operations.
// ---------------------------------------------------------------------------------------
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace {

constexpr int FT = 1024;    // threads per CTA of the fused kernel

template <int T>
__device__ __forceinline__ void fused_rows(const FusedOp &op, int gtid, int gthreads)
{
    const int lane = gtid % T;
    const int groups = gthreads / T;
    const int n_pad = (op.n + groups - 1) / groups * groups;      // every thread joins the shuffles
    for (int row = gtid / T; row < n_pad; row += groups) {
        const bool live = row < op.n;
        const int r = live ? row : op.n - 1;
        const int beg = __ldg(op.ptr + r), end = __ldg(op.ptr + r + 1);
        double acc = 0.0;
        for (int p = beg + lane; p < end; p += T) acc = fma(__ldg(op.vals + p), op.cur[__ldg(op.cols + p)], acc);
#pragma unroll
        for (int o = T / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o, T);
        if (!live || lane != 0) continue;
        if (op.type == FOP_CHEB) {
            double v = op.bq * op.cur[r] + op.c * op.dinv[r] * (op.b[r] - acc);
            if (op.a != 0.0) v = fma(op.a, op.prev[r], v);
            op.out[r] = v;
        } else if (op.mode == SELL_ASSIGN) op.out[r] = acc;
        else if (op.mode == SELL_RESIDUAL) op.out[r] = op.b[r] - acc;
        else if (op.mode == SELL_ADD) op.out[r] += acc;
        else op.out[r] -= acc;
    }
}

__device__ __forceinline__ void fused_execute(const FusedOp &op, int gtid, int gthreads)
{
    if (op.type == FOP_DINV_SCALE) {
        for (int r = gtid; r < op.n; r += gthreads) op.out[r] = op.c * op.dinv[r] * op.b[r];
    } else if (op.type == FOP_COPY) {
        for (int r = gtid; r < op.n; r += gthreads) op.out[r] = op.cur[r];
    } else if (op.type == FOP_GEMV) {
        const int lane = gtid & 31, warps = gthreads >> 5;
        for (int row = gtid >> 5; row < op.n; row += warps) {
            double acc = 0.0;
            for (int j = lane; j < op.n; j += 32) acc = fma(__ldg(op.vals + (size_t)row * op.n + j), op.b[j], acc);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) op.out[row] = acc;
        }
    } else {
        switch (op.lanes) {
        case 2: fused_rows<2>(op, gtid, gthreads); break;
        case 4: fused_rows<4>(op, gtid, gthreads); break;
        case 8: fused_rows<8>(op, gtid, gthreads); break;
        case 16: fused_rows<16>(op, gtid, gthreads); break;
        default: fused_rows<32>(op, gtid, gthreads); break;
        }
    }
}

__global__ void __launch_bounds__(FT) fused_tail_kernel(const FusedOp *__restrict__ ops, int n_ops)
{
    cg::grid_group grid = cg::this_grid();
    const int gtid = blockIdx.x * FT + threadIdx.x;
    const int gthreads = gridDim.x * FT;
    for (int i = 0; i < n_ops; ++i) {
        const FusedOp op = ops[i];
        fused_execute(op, gtid, gthreads);
        grid.sync();
    }
}

// The same program on ONE thread-block cluster: the operations of the small levels (<= 20 k rows,
// matrices resident in L2) need a few thousand threads, and a cluster barrier costs a fraction of a
// grid barrier or of a kernel boundary.
__global__ void __launch_bounds__(FT) fused_cluster_kernel(const FusedOp *__restrict__ ops, int n_ops)
{
    cg::cluster_group cl = cg::this_cluster();
    const int gtid = cl.block_rank() * FT + threadIdx.x;
    const int gthreads = cl.num_blocks() * FT;
    for (int i = 0; i < n_ops; ++i) {
        const FusedOp op = ops[i];
        fused_execute(op, gtid, gthreads);
        __threadfence();
        cl.sync();
    }
}

}  // namespace

int vec_copy_n(ctl_handle_s *h, double *dst, const double *src, int n)
{
    if (h->recorder) {
        FusedOp op{};
        op.type = FOP_COPY; op.n = n; op.cur = src; op.out = dst;
        h->recorder->host.push_back(op);
        return CTL_OK;
    }
    CTL_CUDA(cudaMemcpyAsync(dst, src, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    return CTL_OK;
}

int fused_upload(ctl_handle_s *h, FusedProgram &p)
{
    p.n_ops = (int)p.host.size();
    if (p.n_ops == 0) return CTL_OK;
    CTL_CUDA(cudaMalloc((void **)&p.dev, p.host.size() * sizeof(FusedOp)));
    CTL_CUDA(cudaMemcpyAsync(p.dev, p.host.data(), p.host.size() * sizeof(FusedOp), cudaMemcpyHostToDevice, h->stream));
    CTL_CUDA(cudaStreamSynchronize(h->stream));
    return CTL_OK;
}

void fused_free(FusedProgram &p)
{
    cudaFree(p.dev);
    p.dev = nullptr;
    p.n_ops = 0;
    p.host.clear();
}

int fused_run(ctl_handle_s *h, const FusedProgram &p)
{
    if (p.n_ops == 0) return CTL_OK;
    static int n_sm = 0;
    if (!n_sm) cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, h->cfg.device);
    const FusedOp *ops = p.dev;
    int n_ops = p.n_ops;
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(FT);
    cfg.stream = h->stream;
    cudaLaunchAttribute at;
    cfg.attrs = &at;
    cfg.numAttrs = 1;
    if (p.cluster > 0) {
        cfg.gridDim = dim3(p.cluster);
        at.id = cudaLaunchAttributeClusterDimension;
        at.val.clusterDim.x = p.cluster;
        at.val.clusterDim.y = 1;
        at.val.clusterDim.z = 1;
        CTL_CUDA(cudaLaunchKernelEx(&cfg, fused_cluster_kernel, ops, n_ops));
    } else {
        static int n_cta = 0;
        if (!n_cta) {
            n_cta = 32;
            if (const char *e = getenv("CTL_FUSED_CTAS")) n_cta = std::max(1, std::min(n_sm, atoi(e)));
        }
        cfg.gridDim = dim3(n_cta);
        at.id = cudaLaunchAttributeCooperative;
        at.val.cooperative = 1;
        CTL_CUDA(cudaLaunchKernelEx(&cfg, fused_tail_kernel, ops, n_ops));
    }
    h->launches++;
    return CTL_OK;
}

// largest cluster (16, else 8) of FT-thread CTAs the device can co-schedule for the fused kernel
int fused_cluster_size(ctl_handle_s *h)
{
    for (int want : {16, 8}) {
        if (want > 8 &&
            cudaFuncSetAttribute(fused_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
            cudaGetLastError();
            continue;
        }
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(want);
        cfg.blockDim = dim3(FT);
        cudaLaunchAttribute at;
        at.id = cudaLaunchAttributeClusterDimension;
        at.val.clusterDim.x = want;
        at.val.clusterDim.y = 1;
        at.val.clusterDim.z = 1;
        cfg.attrs = &at;
        cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, fused_cluster_kernel, &cfg) == cudaSuccess && n > 0) return want;
        cudaGetLastError();
    }
    return 0;
}
