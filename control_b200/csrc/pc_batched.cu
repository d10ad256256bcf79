// Time-batched pieces of the in-built block preconditioner:
// everything of Instationary.construct_pc's pc_linear (control/control.py:1995-2048 CN,
// 2191-2237 BE) that treats the N time blocks independently runs as ONE kernel over the
// [n x ld] time-fastest panel instead of N Firedrake/PETSc calls:
//
//   u0_first   T_1^-1 (alternating-sign suffix scan along time, in registers), Dirichlet
//              mask, first Chebyshev/Jacobi step                     (1997, 2003, 1984-1991)
//   cheb_step  fused SpMM + three-term Chebyshev update on D^-1 M     (solver_0, 1970-1982)
//   u0_final   scaling 2/tau | 1/tau | 1/(tau eps) and T_2^-1         (2008-2014, 2202-2206)
//   schur_rhs  b = T_2 (L u_0) - b_1, masked, then T_2^-1             (2017-2053, 2209-2237)
//
// Thread mapping as in kkt_apply.cu: G lanes per row, every lane owns CPL consecutive time columns
// (ld = G * CPL: CPL = 2 up to 64 columns, 4 or 8 for 128 / 256 columns).
#include <algorithm>

#include "common.cuh"
#include "panel.cuh"
#include "pc.cuh"

namespace {

// ------------------------------------------------------------------ u0_first
template <bool CN, int G, int CPL>
__global__ void __launch_bounds__(256) u0_first_kernel(const double *__restrict__ b0, const uint8_t *__restrict__ bc,
                                                      const double *__restrict__ dinv, double *__restrict__ btil,
                                                      double *__restrict__ p1, double c, int n_rows, int N, int ld)
{
    const RowMap m = row_map<G, CPL>(n_rows);
    const size_t off = (size_t)m.r * ld + m.c0;
    double v[CPL];
    load_cols<CPL>(b0 + off, v);
    if (bc[m.r]) {
#pragma unroll
        for (int k = 0; k < CPL; ++k) v[k] = 0.0;
    }
    if (CN) t1_inv<G, CPL>(v, m.l);
    if (!m.live) return;
    store_cols<CPL>(btil + off, v);
    const double s = c * dinv[m.r];
#pragma unroll
    for (int k = 0; k < CPL; ++k) v[k] *= s;
    store_cols<CPL>(p1 + off, v);
}

// ------------------------------------------------------------------ cheb_step
// out = a p_prev + bq p_cur + c dinv (btil - M_bc p_cur); M_bc = M with constrained
// rows/columns replaced by the identity (assemble(M, bcs), control/control.py:1971-1972)
template <int G, int CPL>
__global__ void __launch_bounds__(256) cheb_step_kernel(const int *__restrict__ indptr, const int *__restrict__ indices,
                                                       const double *__restrict__ Mv, const uint8_t *__restrict__ bc,
                                                       const double *__restrict__ dinv, const double *__restrict__ btil,
                                                       const double *p_prev, const double *__restrict__ p_cur,
                                                       const double *__restrict__ halo, double *out, double a,
                                                       double bq, double c, int n_rows, int ld)
{
    const RowMap m = row_map<G, CPL>(n_rows);
    if (!m.live) return;
    const int r = m.r;
    const size_t off = (size_t)r * ld + m.c0;
    double ax[CPL], pc[CPL];
    load_cols<CPL>(p_cur + off, pc);
    if (bc[r]) {
#pragma unroll
        for (int k = 0; k < CPL; ++k) ax[k] = pc[k];
    } else {
#pragma unroll
        for (int k = 0; k < CPL; ++k) ax[k] = 0.0;
        const int kb = __ldg(indptr + r), ke = __ldg(indptr + r + 1);
#pragma unroll 2
        for (int k = kb; k < ke; ++k) {
            const int col = __ldg(indices + k);
            const double mval = __ldg(Mv + k);
            const double *src = col < n_rows ? p_cur + (size_t)col * ld : halo + (size_t)(col - n_rows) * ld;
            double x[CPL];
            load_cols<CPL>(src + m.c0, x);
#pragma unroll
            for (int q = 0; q < CPL; ++q) ax[q] = fma(mval, x[q], ax[q]);
        }
    }
    double bt[CPL], o[CPL];
    load_cols<CPL>(btil + off, bt);
    const double s = c * dinv[r];
#pragma unroll
    for (int k = 0; k < CPL; ++k) o[k] = bq * pc[k] + s * (bt[k] - ax[k]);
    if (a != 0.0) {
#pragma unroll
        for (int q = 0; q < CPL / 2; ++q) {
            const double2 pp = *reinterpret_cast<const double2 *>(p_prev + off + 2 * q);
            o[2 * q] = fma(a, pp.x, o[2 * q]);
            o[2 * q + 1] = fma(a, pp.y, o[2 * q + 1]);
        }
    }
    store_cols<CPL>(out + off, o);
}

// ------------------------------------------------------------------ u0_final
template <bool CN, int G, int CPL>
__global__ void __launch_bounds__(256) u0_final_kernel(const double *__restrict__ p, const uint8_t *__restrict__ bc,
                                                      const double *wrap_b, double *__restrict__ u0, double sc,
                                                      double sc_last, int n_rows, int N, int ld)
{
    const RowMap m = row_map<G, CPL>(n_rows);
    const size_t off = (size_t)m.r * ld + m.c0;
    double v[CPL];
    load_cols<CPL>(p + off, v);
    if (CN) {
#pragma unroll
        for (int k = 0; k < CPL; ++k) v[k] *= sc;
        t2_inv<G, CPL>(v, m.l);
    } else {
#pragma unroll
        for (int k = 0; k < CPL; ++k) v[k] *= (m.c0 + k == N - 1) ? sc_last : sc;
    }
#pragma unroll
    for (int k = 0; k < CPL; ++k)
        if (m.c0 + k >= N) v[k] = 0.0;
    if (!m.live) return;
    if (bc[m.r]) {
        if (wrap_b) load_cols<CPL>(wrap_b + off, v);
        else {
#pragma unroll
            for (int k = 0; k < CPL; ++k) v[k] = 0.0;
        }
    }
    store_cols<CPL>(u0 + off, v);
}

// ------------------------------------------------------------------ schur_rhs
// triangular mode: out = T_2^-1 mask(T_2 mask(L u_0) - b_1) (CN) / mask(L u_0 - b_1) (BE)
// diagonal mode (u0 == nullptr): out = mask(b_1)
template <bool CN, bool PER_LEVEL, int G, int CPL>
__global__ void __launch_bounds__(256) schur_rhs_kernel(const int *__restrict__ indptr, const int *__restrict__ indices,
                                                       const double *__restrict__ Mv, const double *__restrict__ Kv,
                                                       const uint8_t *__restrict__ bc, const double *__restrict__ u0,
                                                       const double *__restrict__ halo, const double *__restrict__ b1,
                                                       double *__restrict__ out, double tau, int n_rows, int N, int ld)
{
    const RowMap m = row_map<G, CPL>(n_rows);
    const int r = m.r;
    const size_t off = (size_t)r * ld + m.c0;
    double o[CPL], bb[CPL];
    const bool masked = bc[r] != 0;
    load_cols<CPL>(b1 + off, bb);
    if (u0) {
        double mv[CPL], kv[CPL], mvp[CPL], kvp[CPL];
#pragma unroll
        for (int k = 0; k < CPL; ++k) mv[k] = kv[k] = 0.0;
        const int kb = __ldg(indptr + r), ke = __ldg(indptr + r + 1);
        for (int k = kb; k < ke; ++k) {
            const int col = __ldg(indices + k);
            const double mval = __ldg(Mv + k);
            const double *src = col < n_rows ? u0 + (size_t)col * ld : halo + (size_t)(col - n_rows) * ld;
            double x[CPL], kk[CPL];
            load_cols<CPL>(src + m.c0, x);
            if (PER_LEVEL) load_cols<CPL>(Kv + (size_t)k * ld + m.c0, kk);
            else {
                const double ks = __ldg(Kv + k);
#pragma unroll
                for (int q = 0; q < CPL; ++q) kk[q] = ks;
            }
#pragma unroll
            for (int q = 0; q < CPL; ++q) {
                mv[q] = fma(mval, x[q], mv[q]);
                kv[q] = fma(kk[q], x[q], kv[q]);
            }
        }
        time_prev<G, CPL>(mv, mvp, m.l);
        time_prev<G, CPL>(kv, kvp, m.l);
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
            // CN: (L u)_i = (h K_{i+1} + M) u_i + (h K_i - M) u_{i-1}   (block_10, control.py:2942-2947)
            // BE: (L u)_i = (tau K_i + M) u_i - M u_{i-1}                (block_10, control.py:2912-2917)
            const double w = CN ? 0.5 * tau : tau;
            o[q] = (w * kv[q] + mv[q]) + (CN ? (w * kvp[q] - mvp[q]) : -mvp[q]);
            if (m.c0 + q >= N || masked) o[q] = 0.0;
        }
        if (CN) {   // T_2: add the previous block
            double op[CPL];
            time_prev<G, CPL>(o, op, m.l);
#pragma unroll
            for (int q = 0; q < CPL; ++q) {
                o[q] += op[q];
                if (m.c0 + q >= N) o[q] = 0.0;
            }
        }
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
            o[q] -= bb[q];
            if (masked) o[q] = 0.0;
        }
        if (CN) {
            t2_inv<G, CPL>(o, m.l);
#pragma unroll
            for (int q = 0; q < CPL; ++q)
                if (m.c0 + q >= N) o[q] = 0.0;
        }
    } else {
#pragma unroll
        for (int q = 0; q < CPL; ++q) o[q] = masked ? 0.0 : bb[q];
    }
    if (!m.live) return;
    store_cols<CPL>(out + off, o);
}

// rows listed in bc_rows: u[row, :] = wrap_b ? wrap_b[row, :] : 0
__global__ void bc_fixup_kernel(const int *__restrict__ bc_rows, int n_bc, const double *wrap_b, double *u, int ld)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_bc * ld) return;
    const size_t off = (size_t)bc_rows[i / ld] * ld + (i % ld);
    u[off] = wrap_b ? wrap_b[off] : 0.0;
}

// single-panel transposes between the time-fastest panel [n][ld] and a time-slowest array
// [N][stride] (stride >= n: ghost entries may follow the owned rows)
constexpr int TILE = 32;
__global__ void panel_tf_to_ts_kernel(const double *__restrict__ src, double *__restrict__ dst, int n, int N, int ld,
                                      size_t stride)
{
    __shared__ double tile[TILE][TILE + 1];
    const int r0 = blockIdx.x * TILE, j0 = blockIdx.y * TILE;
    for (int rr = threadIdx.y; rr < TILE; rr += blockDim.y) {
        const int r = r0 + rr, j = j0 + threadIdx.x;
        tile[rr][threadIdx.x] = (r < n && j < ld) ? src[(size_t)r * ld + j] : 0.0;
    }
    __syncthreads();
    for (int jj = threadIdx.y; jj < TILE; jj += blockDim.y) {
        const int j = j0 + jj, r = r0 + threadIdx.x;
        if (j < N && r < n) dst[(size_t)j * stride + r] = tile[threadIdx.x][jj];
    }
}

__global__ void panel_ts_to_tf_kernel(const double *__restrict__ src, double *__restrict__ dst, int n, int N, int ld,
                                      size_t stride)
{
    __shared__ double tile[TILE][TILE + 1];
    const int r0 = blockIdx.x * TILE, j0 = blockIdx.y * TILE;
    for (int jj = threadIdx.y; jj < TILE; jj += blockDim.y) {
        const int j = j0 + jj, r = r0 + threadIdx.x;
        tile[jj][threadIdx.x] = (j < N && r < n) ? src[(size_t)j * stride + r] : 0.0;
    }
    __syncthreads();
    for (int rr = threadIdx.y; rr < TILE; rr += blockDim.y) {
        const int r = r0 + rr, j = j0 + threadIdx.x;
        if (r < n && j < ld) dst[(size_t)r * ld + j] = tile[threadIdx.x][rr];
    }
}

}  // namespace

int pcb_u0_first(ctl_handle_s *h, const double *b0, const double *dinv, double *btil, double *p1, double c)
{
    const int G = h->ld, nb = blocks_for(h->n_loc, h->ld);
    if (h->cfg.CN) {
        DISPATCH_G(G, (u0_first_kernel<true, GG, CC><<<nb, 256, 0, h->stream>>>(b0, h->d_bcmask, dinv, btil, p1, c, h->n_loc, h->N, h->ld)));
    } else {
        DISPATCH_G(G, (u0_first_kernel<false, GG, CC><<<nb, 256, 0, h->stream>>>(b0, h->d_bcmask, dinv, btil, p1, c, h->n_loc, h->N, h->ld)));
    }
    h->launches++;
    CTL_CUDA(cudaGetLastError());
    return CTL_OK;
}

int pcb_cheb_step(ctl_handle_s *h, const double *dinv, const double *btil, const double *p_prev, const double *p_cur,
                  double *out, double a, double bq, double c)
{
    const int G = h->ld, nb = blocks_for(h->n_loc, h->ld);
    DISPATCH_G(G, (cheb_step_kernel<GG, CC><<<nb, 256, 0, h->stream>>>(h->d_indptr, h->d_indices, h->d_M, h->d_bcmask, dinv, btil, p_prev,
                                                                 p_cur, h->d_halo, out, a, bq, c, h->n_loc, h->ld)));
    h->launches++;
    CTL_CUDA(cudaGetLastError());
    return CTL_OK;
}

int pcb_u0_final(ctl_handle_s *h, const double *p, const double *wrap_b, double *u0, double sc, double sc_last)
{
    const int G = h->ld, nb = blocks_for(h->n_loc, h->ld);
    if (h->cfg.CN) {
        DISPATCH_G(G, (u0_final_kernel<true, GG, CC><<<nb, 256, 0, h->stream>>>(p, h->d_bcmask, wrap_b, u0, sc, sc_last, h->n_loc, h->N, h->ld)));
    } else {
        DISPATCH_G(G, (u0_final_kernel<false, GG, CC><<<nb, 256, 0, h->stream>>>(p, h->d_bcmask, wrap_b, u0, sc, sc_last, h->n_loc, h->N, h->ld)));
    }
    h->launches++;
    CTL_CUDA(cudaGetLastError());
    return CTL_OK;
}

int pcb_schur_rhs(ctl_handle_s *h, const double *u0, const double *b1, double *out)
{
    const int G = h->ld, nb = blocks_for(h->n_loc, h->ld);
    const bool cn = h->cfg.CN != 0, pl = h->per_level;
#define SR(CNV, PLV)                                                                                                   \
    DISPATCH_G(G, (schur_rhs_kernel<CNV, PLV, GG, CC><<<nb, 256, 0, h->stream>>>(h->d_indptr, h->d_indices, h->d_M, h->d_K, h->d_bcmask, \
                                                                           u0, h->d_halo, b1, out, h->cfg.tau, h->n_loc, h->N, h->ld)))
    if (cn && pl) { SR(true, true); }
    else if (cn) { SR(true, false); }
    else if (pl) { SR(false, true); }
    else { SR(false, false); }
#undef SR
    h->launches++;
    CTL_CUDA(cudaGetLastError());
    return CTL_OK;
}

int pcb_bc_fixup(ctl_handle_s *h, const int *bc_rows, int n_bc, const double *wrap_b, double *u)
{
    if (n_bc == 0) return CTL_OK;
    const int total = n_bc * h->ld;
    bc_fixup_kernel<<<ceil_div(total, 256), 256, 0, h->stream>>>(bc_rows, n_bc, wrap_b, u, h->ld);
    h->launches++;
    CTL_CUDA(cudaGetLastError());
    return CTL_OK;
}

int pcb_panel_to_ts(ctl_handle_s *h, const double *panel_tf, double *ts, size_t stride)
{
    dim3 block(TILE, 8), grid(ceil_div(h->n_loc, TILE), ceil_div(h->ld, TILE));
    panel_tf_to_ts_kernel<<<grid, block, 0, h->stream>>>(panel_tf, ts, h->n_loc, h->N, h->ld, stride);
    h->launches++;
    CTL_CUDA(cudaGetLastError());
    return CTL_OK;
}

int pcb_ts_to_panel(ctl_handle_s *h, const double *ts, double *panel_tf, size_t stride)
{
    dim3 block(TILE, 8), grid(ceil_div(h->n_loc, TILE), ceil_div(h->ld, TILE));
    panel_ts_to_tf_kernel<<<grid, block, 0, h->stream>>>(ts, panel_tf, h->n_loc, h->N, h->ld, stride);
    h->launches++;
    CTL_CUDA(cudaGetLastError());
    return CTL_OK;
}
