// SELL-32 matrices and the single-column SpMV kernel family (sell.cu).
//
// Every product inside the time sweeps and the AMG V-cycles acts on ONE spatial vector
// (SURVEY.md rows K8, K9), so rows cannot be spread over a warp the way the time-batched
// kernels do.  Sliced ELLPACK with slice height 32 gives one row per thread with fully
// coalesced index/value loads: entry k of row r is stored at slice_ptr[r / 32] + 32 k +
// r % 32.  Padding entries have value 0 and the row's own index as column.
#pragma once
#include <memory>

#include "common.cuh"

struct SellPattern {
    int n_rows = 0, n_cols = 0, n_slices = 0;
    int64_t n_stored = 0, nnz = 0;
    int *slice_ptr = nullptr;          // device, n_slices + 1
    int *cols = nullptr;               // device, n_stored
    std::vector<int64_t> csr_to_sell;  // host: CSR entry -> stored position
    ~SellPattern();
};

struct SellMat {
    std::shared_ptr<SellPattern> pat;
    double *vals = nullptr;            // device, n_stored (owned)
    // Rows much longer than a fine-mesh stencil (coarse AMG operators, restrictions) are
    // latency-bound with one thread per row: they are stored as plain CSR instead and
    // processed by `lanes` threads per row (lanes = 0: SELL, one thread per row).
    int lanes = 0;
    int *csr_ptr = nullptr, *csr_cols = nullptr;
    double *csr_vals = nullptr;
    int n_rows() const { return pat->n_rows; }
    bool valid() const { return pat && (vals || csr_vals); }
};

// build the pattern from a host CSR (column indices need not be sorted)
int sell_build_pattern(ctl_handle_s *h, const HostCSR &A, std::shared_ptr<SellPattern> &out);
// lay one value set (CSR order, host) out on a pattern
int sell_set_values(ctl_handle_s *h, const std::shared_ptr<SellPattern> &pat, const double *csr_values,
                    SellMat &out);
// one call for matrices that own their pattern: picks SELL or CSR-vector by mean row length
// (force_csr: always CSR, the only format the fused coarse-tail kernel reads)
int sell_from_csr(ctl_handle_s *h, const HostCSR &A, SellMat &out, bool force_csr = false);

// ---- fused coarse tail: while a recorder is installed on the handle, the primitives below
// append their operation to it instead of launching a kernel; the recorded program is later
// executed by ONE cooperative kernel with grid-wide barriers between operations (sell.cu).
enum FusedType { FOP_DINV_SCALE = 0, FOP_CHEB = 1, FOP_SPMV = 2, FOP_GEMV = 3, FOP_COPY = 4 };
struct FusedOp {
    int type, n, lanes, mode;
    const int *ptr, *cols;
    const double *vals;
    const double *dinv, *b, *prev, *cur;
    double *out;
    double a, bq, c;
};
struct FusedProgram {
    std::vector<FusedOp> host;
    FusedOp *dev = nullptr;
    int n_ops = 0;
    int cluster = 0;     // > 0: run on one thread-block cluster of this many CTAs (else cooperative grid)
};
int fused_upload(ctl_handle_s *h, FusedProgram &p);
int fused_run(ctl_handle_s *h, const FusedProgram &p);
int fused_cluster_size(ctl_handle_s *h);
void fused_free(FusedProgram &p);
int vec_copy_n(ctl_handle_s *h, double *dst, const double *src, int n);   // recordable device copy
void sell_free(SellMat &m);

enum SellMode {
    SELL_ASSIGN = 0,      // y  = A x
    SELL_RESIDUAL = 1,    // y  = b - A x
    SELL_ADD = 2,         // y += A x
    SELL_SUB = 3          // y -= A x
};
int sell_spmv(ctl_handle_s *h, const SellMat &A, const double *x, double *y, const double *b, int mode);

// Chebyshev / Jacobi step: out = a * p_prev + bq * p_cur + c * dinv .* (b - A p_cur)
// (p_prev may be null when a == 0; out may alias p_prev)
int sell_cheb_step(ctl_handle_s *h, const SellMat &A, const double *dinv, const double *b,
                   const double *p_prev, const double *p_cur, double *out, double a, double bq, double c);
// out = c * dinv .* b   (first step from a zero guess)
int vec_dinv_scale(ctl_handle_s *h, const double *dinv, const double *b, double *out, double c, int n);
// y = Ainv b, dense row-major n x n
int dense_gemv(ctl_handle_s *h, const double *Ainv, const double *b, double *y, int n);
// two-matrix product on one shared pattern (backward-sweep right-hand side, pc.cu):
//   y = alpha * A1 (x1 + x2) + beta * A2 x3      (x2, x3 may be null)
int sell_spmv2(ctl_handle_s *h, const SellMat &A1, const SellMat &A2, const double *x1, const double *x2,
               const double *x3, double *y, double alpha, double beta);
