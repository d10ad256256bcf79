// Aggregation-AMG V-cycle on the device: the inner solver of the
// Schur-complement time sweeps, standing in for "preonly + hypre boomeramg, max_iter 2"
// (control/control.py:2056-2067 and the nine other sites).  The hierarchy is set up on
// the host once per distinct matrix (amg_setup.cpp); a cycle is a fixed sequence of SELL
// SpMV-family kernels (sell.cu), so the time sweeps can be captured in a CUDA graph.
#include "amg.cuh"

#include <cstdlib>
#include <stdexcept>

#include "cheb_coefficients.h"

// levels >= FUSED_FROM of a V-cycle run as one cooperative kernel (sell.cu: fused coarse tail)
static bool fusion_enabled()
{
    const char *e = getenv("CTL_FUSED");
    return e && (e[0] == '1' || e[0] == '2');
}

static int fused_from_level()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("CTL_FUSED_FROM");
        v = e ? std::max(1, atoi(e)) : 2;
    }
    return v;
}
#define FUSED_FROM fused_from_level()
// the fused kernel reads CSR only; without it every matrix takes the format sell_from_csr picks
#define FORCE_CSR(l) (fusion_enabled() && (l) >= FUSED_FROM)

static int dev_alloc(ctl_handle_s *h, double **p, size_t n)
{
    CTL_CUDA(cudaMalloc((void **)p, std::max<size_t>(n, 1) * sizeof(double)));
    CTL_CUDA(cudaMemsetAsync(*p, 0, std::max<size_t>(n, 1) * sizeof(double), h->stream));
    return CTL_OK;
}

static int vcycle(ctl_handle_s *h, AmgHierarchyDev &H, int l, const double *b, double *x, bool zero_guess);

int amg_build(ctl_handle_s *h, const HostCSR &A0, const AmgParams &p,
              const std::shared_ptr<SellPattern> &fine_pattern, AmgHierarchyDev &H)
{
    H.params = p;
    try {
        if (H.host.empty()) amg_setup_host(A0, p, H.host);      // else: set up earlier (pc.cu builds many in parallel)
    } catch (const std::exception &e) {
        ctl_set_error(h, e.what());
        return CTL_ERR_STATE;
    }
    const int nl = (int)H.host.size();
    H.dev.assign(nl, AmgLevelDev());
    H.bytes_per_cycle = 0;
    for (int l = 0; l < nl; ++l) {
        AmgLevelHost &Lh = H.host[l];
        AmgLevelDev &Ld = H.dev[l];
        // level 0 is distributed by rows (this rank's block, ghost columns appended); the
        // coarse levels are replicated on every rank
        const bool dist0 = (l == 0);
        const int rb = dist0 ? h->row_begin : 0;
        Ld.n = dist0 ? h->n_loc : Lh.A.n_rows;
        const int ghosts = dist0 ? h->n_halo : 0;
        Ld.rho = Lh.rho;
        if (dist0) {
            CTL_CHECK(fine_pattern != nullptr, CTL_ERR_ARG, "amg_build: level 0 needs the mesh pattern");
            std::vector<double> lv(h->loc_entry.size());
            for (size_t q = 0; q < lv.size(); ++q) lv[q] = Lh.A.values[h->loc_entry[q]];
            CTL_TRY(sell_set_values(h, fine_pattern, lv.data(), Ld.A));
        } else {
            CTL_TRY(sell_from_csr(h, Lh.A, Ld.A, FORCE_CSR(l)));
        }
        CTL_TRY(ctl_upload(h, &Ld.dinv, Lh.dinv.data() + rb, (size_t)Ld.n));
        CTL_TRY(dev_alloc(h, &Ld.r, Ld.n + ghosts));
        CTL_TRY(dev_alloc(h, &Ld.t0, Ld.n + ghosts));
        CTL_TRY(dev_alloc(h, &Ld.t1, Ld.n + ghosts));
        if (l > 0) {
            CTL_TRY(dev_alloc(h, &Ld.x, Ld.n));
            CTL_TRY(dev_alloc(h, &Ld.b, Ld.n));
        }
        const int64_t spmv = 12 * Lh.A.nnz() + 4ll * (Ld.n + 1) + 16ll * Ld.n;
        if (l + 1 < nl) {
            if (dist0 && h->cfg.world > 1) {
                HostCSR Pl, Rl;               // rows of P owned by this rank, and their transpose
                Pl.n_rows = Ld.n;
                Pl.n_cols = Lh.P.n_cols;
                Pl.indptr.assign(Ld.n + 1, 0);
                const int k0 = Lh.P.indptr[rb];
                for (int r = 0; r <= Ld.n; ++r) Pl.indptr[r] = Lh.P.indptr[rb + r] - k0;
                Pl.indices.assign(Lh.P.indices.begin() + k0, Lh.P.indices.begin() + Lh.P.indptr[rb + Ld.n]);
                Pl.values.assign(Lh.P.values.begin() + k0, Lh.P.values.begin() + Lh.P.indptr[rb + Ld.n]);
                csr_transpose(Pl, Rl);
                CTL_TRY(sell_from_csr(h, Pl, Ld.P));
                CTL_TRY(sell_from_csr(h, Rl, Ld.R));
            } else {
                CTL_TRY(sell_from_csr(h, Lh.P, Ld.P, FORCE_CSR(l)));
                CTL_TRY(sell_from_csr(h, Lh.R, Ld.R, FORCE_CSR(l)));
            }
            // per cycle: (2 nu - 1 or 2 nu) smoother products + 1 residual, restriction, prolongation
            const int nu_l = (l == 0 && p.nu_fine > 0) ? p.nu_fine : p.nu;
            H.bytes_per_cycle += spmv * (2 * nu_l) + 2 * (12 * Lh.P.nnz() + 16ll * Ld.n);
        } else if (!Lh.Ainv.empty()) {
            CTL_TRY(ctl_upload(h, &Ld.Ainv, Lh.Ainv.data(), Lh.Ainv.size()));
            H.bytes_per_cycle += 8ll * Ld.n * Ld.n;
        } else {
            H.bytes_per_cycle += spmv * p.nu;
        }
    }
    H.fused_from = 0;
    if (h->cfg.world > 1)
        CTL_CHECK(nl > 1, CTL_ERR_ARG, "amg_build: the problem is too small for a multi-rank hierarchy (single level)");
    // record the sub-cycle below level FUSED_FROM - 1 once (zero guess at entry, exactly what vcycle() issues)
    // (opt-in, CTL_FUSED=1: round-1 measurement showed the cooperative kernel slower than the separate
    // launches, 1.56 ms against 1.27 ms per inner solve: the operations are bound by dependent memory
    // round trips, not by launch gaps)
    const char *fe = getenv("CTL_FUSED");
    if (nl > FUSED_FROM && fe && (fe[0] == '1' || fe[0] == '2')) {
        H.fused.cluster = fe[0] == '2' ? fused_cluster_size(h) : 0;      // 2: one thread-block cluster
        h->recorder = &H.fused;
        const int rc = vcycle(h, H, FUSED_FROM, H.dev[FUSED_FROM].b, H.dev[FUSED_FROM].x, true);
        h->recorder = nullptr;
        CTL_TRY(rc);
        CTL_TRY(fused_upload(h, H.fused));
        H.fused_from = FUSED_FROM;
    }
    if (p.acc_lo > 0.0) {
        CTL_TRY(dev_alloc(h, &H.acc_r, H.dev[0].n + h->n_halo));
        CTL_TRY(dev_alloc(h, &H.acc_z, H.dev[0].n + h->n_halo));
        CTL_TRY(dev_alloc(h, &H.acc_p, H.dev[0].n + h->n_halo));
    }
    return CTL_OK;
}

void amg_free(AmgHierarchyDev &H)
{
    fused_free(H.fused);
    H.fused_from = 0;
    cudaFree(H.acc_r);
    cudaFree(H.acc_z);
    cudaFree(H.acc_p);
    H.acc_r = H.acc_z = H.acc_p = nullptr;
    for (AmgLevelDev &L : H.dev) {
        sell_free(L.A);
        sell_free(L.P);
        sell_free(L.R);
        cudaFree(L.dinv);
        cudaFree(L.x);
        cudaFree(L.b);
        cudaFree(L.r);
        cudaFree(L.t0);
        cudaFree(L.t1);
        cudaFree(L.Ainv);
    }
    H.dev.clear();
    H.host.clear();
}

// nu Chebyshev steps on D^-1 A over [lo rho, hi rho] (oracle/cheb.py::chebyshev), result in x.
// Iterates alternate between x and t0; for a non-zero guess and odd nu one copy moves the
// result back into x.
// ghost entries of a level-0 vector must be current before a product with the distributed A
static inline int halo0(ctl_handle_s *h, int l, double *v) { return l == 0 ? ctl_halo_exchange_vec(h, v) : CTL_OK; }

static int smooth(ctl_handle_s *h, const AmgParams &p, AmgLevelDev &L, int l, const double *b, double *x, bool zero_guess)
{
    double scale;
    std::vector<double> om;
    const int nu = (l == 0 && p.nu_fine > 0) ? p.nu_fine : p.nu;
    cheb_coefficients(p.lo * L.rho, p.hi * L.rho, nu, &scale, om);
    // Where iterate p_k lives.  A step reads p_{k-1} through the gather (must not be the buffer
    // being written) and p_{k-2} element-wise (may be).  Zero guess: alternate x / t0 so that
    // p_nu lands in x.  Non-zero guess (p_0 = x): rotate x / t0 / t1 backwards from p_nu = x;
    // this never overwrites a buffer that is still gathered unless nu % 3 == 1, in which case
    // the iterates alternate t0 / x and one copy moves an odd-nu result back into x.
    double *buf3[3] = {x, L.t0, L.t1};
    const bool rotate3 = !zero_guess && (nu % 3 != 1);
    auto where = [&](int k) -> double * {
        if (k == 0) return x;
        if (zero_guess) return ((nu - k) & 1) ? L.t0 : x;
        if (rotate3) return buf3[(nu - k) % 3];
        return (k & 1) ? L.t0 : x;
    };
    if (zero_guess) {
        CTL_TRY(vec_dinv_scale(h, L.dinv, b, where(1), scale, L.n));
    } else {
        // p_1 = x + scale D^-1 (b - A x)
        CTL_TRY(halo0(h, l, x));
        CTL_TRY(sell_cheb_step(h, L.A, L.dinv, b, nullptr, x, where(1), 0.0, 1.0, scale));
    }
    for (int k = 2; k <= nu; ++k) {
        const double w = om[k - 2];
        const bool no_prev = (k == 2 && zero_guess);
        CTL_TRY(halo0(h, l, where(k - 1)));
        CTL_TRY(sell_cheb_step(h, L.A, L.dinv, b, no_prev ? nullptr : where(k - 2), where(k - 1), where(k),
                               no_prev ? 0.0 : (1.0 - w), w, w * scale));
    }
    if (where(nu) != x) CTL_TRY(vec_copy_n(h, x, where(nu), L.n));
    return CTL_OK;
}

static int vcycle(ctl_handle_s *h, AmgHierarchyDev &H, int l, const double *b, double *x, bool zero_guess)
{
    AmgLevelDev &L = H.dev[l];
    const int last = (int)H.dev.size() - 1;
    if (l == last) {
        if (L.Ainv) return dense_gemv(h, L.Ainv, b, x, L.n);
        return smooth(h, H.params, L, l, b, x, zero_guess);
    }
    AmgLevelDev &C = H.dev[l + 1];
    CTL_TRY(smooth(h, H.params, L, l, b, x, zero_guess));
    CTL_TRY(halo0(h, l, x));
    CTL_TRY(sell_spmv(h, L.A, x, L.r, b, SELL_RESIDUAL));
    CTL_TRY(sell_spmv(h, L.R, L.r, C.b, nullptr, SELL_ASSIGN));
    if (l == 0) CTL_TRY(ctl_allreduce_sum(h, C.b, C.n));      // partial restrictions of the row blocks
    if (H.fused_from == l + 1 && !h->recorder) CTL_TRY(fused_run(h, H.fused));
    else CTL_TRY(vcycle(h, H, l + 1, C.b, C.x, true));
    CTL_TRY(sell_spmv(h, L.P, C.x, x, nullptr, SELL_ADD));
    CTL_TRY(smooth(h, H.params, L, l, b, x, false));
    return CTL_OK;
}

namespace {
// out = a x + b y + c z on n-vectors (x may be null when a == 0; out may alias x)
__global__ void lincomb3_kernel(double *out, double a, const double *x, double b, const double *__restrict__ y,
                                double c, const double *__restrict__ z, int n)
{
    pdl_sync();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double r = c * z[i];
    if (y) r = fma(b, y[i], r);
    if (x) r = fma(a, x[i], r);
    out[i] = r;
}
}  // namespace

int amg_solve(ctl_handle_s *h, AmgHierarchyDev &H, const double *b, double *x)
{
    const AmgParams &p = H.params;
    if (p.acc_lo <= 0.0) {
        for (int c = 0; c < p.cycles; ++c) CTL_TRY(vcycle(h, H, 0, b, x, c == 0));
        return CTL_OK;
    }
    // Chebyshev semi-iteration on (V-cycle * A) over [acc_lo, acc_hi]: oracle/amg.py::solve
    const int n = H.dev[0].n;
    double scale;
    std::vector<double> om;
    cheb_coefficients(p.acc_lo, p.acc_hi, p.cycles, &scale, om);
    double *buf[2] = {x, H.acc_p};
    auto slot = [&](int k) { return (p.cycles - k) & 1; };      // p_cycles lands in x
    const int blocks = ceil_div(n, 256);
    CTL_TRY(vcycle(h, H, 0, b, H.acc_z, true));
    pdl_launch(h, blocks, 256, lincomb3_kernel, buf[slot(1)], 0.0, nullptr, 0.0, nullptr, scale, H.acc_z, n);
    h->launches++;
    for (int k = 2; k <= p.cycles; ++k) {
        const double w = om[k - 2];
        CTL_TRY(halo0(h, 0, buf[slot(k - 1)]));
        CTL_TRY(sell_spmv(h, H.dev[0].A, buf[slot(k - 1)], H.acc_r, b, SELL_RESIDUAL));
        CTL_TRY(vcycle(h, H, 0, H.acc_r, H.acc_z, true));
        pdl_launch(h, blocks, 256, lincomb3_kernel, buf[slot(k)], 1.0 - w, k == 2 ? nullptr : buf[slot(k - 2)], w,
                                                      buf[slot(k - 1)], w * scale, H.acc_z, n);
        h->launches++;
    }
    CTL_CUDA(cudaGetLastError());
    return CTL_OK;
}
