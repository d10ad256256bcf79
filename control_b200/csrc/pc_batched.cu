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
// Thread mapping as in kkt_apply.cu: G = ld/2 lanes per row, lane l owns columns 2l, 2l+1.
#include <algorithm>

#include "common.cuh"
#include "pc.cuh"

namespace {

__device__ __forceinline__ double2 ldg2(const double *p)
{
    return __ldg(reinterpret_cast<const double2 *>(p));
}

// inclusive scans over the G lanes of a row group
template <int G>
__device__ __forceinline__ double group_prefix(double v, int l)
{
#pragma unroll
    for (int d = 1; d < G; d <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, v, d, G);
        if (l >= d) v += t;
    }
    return v;
}

template <int G>
__device__ __forceinline__ double group_suffix(double v, int l)
{
#pragma unroll
    for (int d = 1; d < G; d <<= 1) {
        const double t = __shfl_down_sync(0xffffffffu, v, d, G);
        if (l + d < G) v += t;
    }
    return v;
}

// T_1^-1: x_i <- sum_{k >= i} (-1)^(k-i) x_k   (control/control.py:63-78)
template <int G>
__device__ __forceinline__ void t1_inv(double &x0, double &x1, int l)
{
    const double z0 = x0, z1 = -x1;                 // column 2l is even
    const double s = group_suffix<G>(z0 + z1, l);   // sum over columns >= 2l
    x0 = s;
    x1 = -(s - z0);
}

// T_2^-1: x_i <- sum_{k <= i} (-1)^(i-k) x_k   (control/control.py:81-96)
template <int G>
__device__ __forceinline__ void t2_inv(double &x0, double &x1, int l)
{
    const double z0 = x0, z1 = -x1;
    const double s = group_prefix<G>(z0 + z1, l);   // sum over columns <= 2l+1
    x0 = s - z1;
    x1 = -s;
}

struct RowMap {
    int row, r, l, c0;
    bool live;
};

template <int G>
__device__ __forceinline__ RowMap row_map(int n_rows)
{
    RowMap m;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * (blockDim.x >> 5)) + (threadIdx.x >> 5);
    m.row = warp * (32 / G) + lane / G;
    m.live = m.row < n_rows;
    m.r = m.live ? m.row : n_rows - 1;
    m.l = lane % G;
    m.c0 = 2 * m.l;
    return m;
}

// ------------------------------------------------------------------ u0_first
template <bool CN, int G>
__global__ void __launch_bounds__(256) u0_first_kernel(const double *__restrict__ b0, const uint8_t *__restrict__ bc,
                                                      const double *__restrict__ dinv, double *__restrict__ btil,
                                                      double *__restrict__ p1, double c, int n_rows, int N, int ld)
{
    const RowMap m = row_map<G>(n_rows);
    const size_t off = (size_t)m.r * ld + m.c0;
    double2 v = ldg2(b0 + off);
    if (bc[m.r]) v = make_double2(0.0, 0.0);
    if (CN) t1_inv<G>(v.x, v.y, m.l);
    if (!m.live) return;
    *reinterpret_cast<double2 *>(btil + off) = v;
    const double s = c * dinv[m.r];
    *reinterpret_cast<double2 *>(p1 + off) = make_double2(s * v.x, s * v.y);
}

// ------------------------------------------------------------------ cheb_step
// out = a p_prev + bq p_cur + c dinv (btil - M_bc p_cur); M_bc = M with constrained
// rows/columns replaced by the identity (assemble(M, bcs), control/control.py:1971-1972)
template <int G>
__global__ void __launch_bounds__(256) cheb_step_kernel(const int *__restrict__ indptr, const int *__restrict__ indices,
                                                       const double *__restrict__ Mv, const uint8_t *__restrict__ bc,
                                                       const double *__restrict__ dinv, const double *__restrict__ btil,
                                                       const double *p_prev, const double *__restrict__ p_cur,
                                                       const double *__restrict__ halo, double *out, double a,
                                                       double bq, double c, int n_rows, int ld)
{
    const RowMap m = row_map<G>(n_rows);
    if (!m.live) return;
    const int r = m.r;
    const size_t off = (size_t)r * ld + m.c0;
    double ax0 = 0.0, ax1 = 0.0;
    const double2 pc = ldg2(p_cur + off);
    if (bc[r]) {
        ax0 = pc.x;
        ax1 = pc.y;
    } else {
        const int kb = __ldg(indptr + r), ke = __ldg(indptr + r + 1);
#pragma unroll 4
        for (int k = kb; k < ke; ++k) {
            const int col = __ldg(indices + k);
            const double mval = __ldg(Mv + k);
            const double *src = col < n_rows ? p_cur + (size_t)col * ld : halo + (size_t)(col - n_rows) * ld;
            const double2 x = ldg2(src + m.c0);
            ax0 = fma(mval, x.x, ax0);
            ax1 = fma(mval, x.y, ax1);
        }
    }
    const double2 bt = ldg2(btil + off);
    const double s = c * dinv[r];
    double o0 = bq * pc.x + s * (bt.x - ax0);
    double o1 = bq * pc.y + s * (bt.y - ax1);
    if (a != 0.0) {
        const double2 pp = *reinterpret_cast<const double2 *>(p_prev + off);
        o0 = fma(a, pp.x, o0);
        o1 = fma(a, pp.y, o1);
    }
    *reinterpret_cast<double2 *>(out + off) = make_double2(o0, o1);
}

// ------------------------------------------------------------------ u0_final
template <bool CN, int G>
__global__ void __launch_bounds__(256) u0_final_kernel(const double *__restrict__ p, const uint8_t *__restrict__ bc,
                                                      const double *wrap_b, double *__restrict__ u0, double sc,
                                                      double sc_last, int n_rows, int N, int ld)
{
    const RowMap m = row_map<G>(n_rows);
    const size_t off = (size_t)m.r * ld + m.c0;
    double2 v = ldg2(p + off);
    if (CN) {
        v.x *= sc;
        v.y *= sc;
        t2_inv<G>(v.x, v.y, m.l);
    } else {
        v.x *= (m.c0 == N - 1) ? sc_last : sc;
        v.y *= (m.c0 + 1 == N - 1) ? sc_last : sc;
    }
    if (m.c0 >= N) v.x = 0.0;
    if (m.c0 + 1 >= N) v.y = 0.0;
    if (!m.live) return;
    if (bc[m.r]) v = wrap_b ? ldg2(wrap_b + off) : make_double2(0.0, 0.0);
    *reinterpret_cast<double2 *>(u0 + off) = v;
}

// ------------------------------------------------------------------ schur_rhs
// triangular mode: out = T_2^-1 mask(T_2 mask(L u_0) - b_1) (CN) / mask(L u_0 - b_1) (BE)
// diagonal mode (u0 == nullptr): out = mask(b_1)
template <bool CN, bool PER_LEVEL, int G>
__global__ void __launch_bounds__(256) schur_rhs_kernel(const int *__restrict__ indptr, const int *__restrict__ indices,
                                                       const double *__restrict__ Mv, const double *__restrict__ Kv,
                                                       const uint8_t *__restrict__ bc, const double *__restrict__ u0,
                                                       const double *__restrict__ halo, const double *__restrict__ b1,
                                                       double *__restrict__ out, double tau, int n_rows, int N, int ld)
{
    const RowMap m = row_map<G>(n_rows);
    const int r = m.r;
    const size_t off = (size_t)r * ld + m.c0;
    const bool first = m.l == 0;
    double o0 = 0.0, o1 = 0.0;
    const bool masked = bc[r] != 0;
    if (u0) {
        double mv0 = 0, mv1 = 0, kv0 = 0, kv1 = 0;
        const int kb = __ldg(indptr + r), ke = __ldg(indptr + r + 1);
#pragma unroll 2
        for (int k = kb; k < ke; ++k) {
            const int col = __ldg(indices + k);
            const double mval = __ldg(Mv + k);
            const double *src = col < n_rows ? u0 + (size_t)col * ld : halo + (size_t)(col - n_rows) * ld;
            const double2 x = ldg2(src + m.c0);
            mv0 = fma(mval, x.x, mv0);
            mv1 = fma(mval, x.y, mv1);
            if (PER_LEVEL) {
                const double2 kk = ldg2(Kv + (size_t)k * ld + m.c0);
                kv0 = fma(kk.x, x.x, kv0);
                kv1 = fma(kk.y, x.y, kv1);
            } else {
                const double kk = __ldg(Kv + k);
                kv0 = fma(kk, x.x, kv0);
                kv1 = fma(kk, x.y, kv1);
            }
        }
        double t;
        t = __shfl_up_sync(0xffffffffu, mv1, 1, G); const double mvp0 = first ? 0.0 : t;
        if (CN) {
            const double h = 0.5 * tau;
            t = __shfl_up_sync(0xffffffffu, kv1, 1, G); const double kvp0 = first ? 0.0 : t;
            // (L u)_i = (h K_{i+1} + M) u_i + (h K_i - M) u_{i-1}   (block_10, control.py:2942-2947)
            o0 = (h * kv0 + mv0) + (h * kvp0 - mvp0);
            o1 = (h * kv1 + mv1) + (h * kv0 - mv0);
        } else {
            // (L u)_i = (tau K_i + M) u_i - M u_{i-1}                (block_10, control.py:2912-2917)
            o0 = (tau * kv0 + mv0) - mvp0;
            o1 = (tau * kv1 + mv1) - mv0;
        }
        if (m.c0 >= N) o0 = 0.0;
        if (m.c0 + 1 >= N) o1 = 0.0;
        if (masked) { o0 = 0.0; o1 = 0.0; }
        if (CN) {   // T_2: add the previous block
            t = __shfl_up_sync(0xffffffffu, o1, 1, G); const double op0 = first ? 0.0 : t;
            const double n1 = o1 + o0;
            o0 = o0 + op0;
            o1 = n1;
            if (m.c0 >= N) o0 = 0.0;
            if (m.c0 + 1 >= N) o1 = 0.0;
        }
        const double2 bb = ldg2(b1 + off);
        o0 -= bb.x;
        o1 -= bb.y;
        if (masked) { o0 = 0.0; o1 = 0.0; }
        if (CN) {
            t2_inv<G>(o0, o1, m.l);
            if (m.c0 >= N) o0 = 0.0;
            if (m.c0 + 1 >= N) o1 = 0.0;
        }
    } else {
        const double2 bb = ldg2(b1 + off);
        o0 = masked ? 0.0 : bb.x;
        o1 = masked ? 0.0 : bb.y;
    }
    if (!m.live) return;
    *reinterpret_cast<double2 *>(out + off) = make_double2(o0, o1);
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

inline int blocks_for(int n_rows, int G) { return ceil_div(n_rows, 8 * (32 / G)); }

}  // namespace

#define DISPATCH_G(G, CALL)                                                                \
    switch (G) {                                                                           \
    case 4: { constexpr int GG = 4; CALL; } break;                                         \
    case 8: { constexpr int GG = 8; CALL; } break;                                         \
    case 16: { constexpr int GG = 16; CALL; } break;                                       \
    default: { constexpr int GG = 32; CALL; } break;                                       \
    }

int pcb_u0_first(ctl_handle_s *h, const double *b0, const double *dinv, double *btil, double *p1, double c)
{
    const int G = h->ld / 2, nb = blocks_for(h->n_loc, G);
    if (h->cfg.CN) {
        DISPATCH_G(G, (u0_first_kernel<true, GG><<<nb, 256, 0, h->stream>>>(b0, h->d_bcmask, dinv, btil, p1, c, h->n_loc, h->N, h->ld)));
    } else {
        DISPATCH_G(G, (u0_first_kernel<false, GG><<<nb, 256, 0, h->stream>>>(b0, h->d_bcmask, dinv, btil, p1, c, h->n_loc, h->N, h->ld)));
    }
    h->launches++;
    CTL_CUDA(cudaGetLastError());
    return CTL_OK;
}

int pcb_cheb_step(ctl_handle_s *h, const double *dinv, const double *btil, const double *p_prev, const double *p_cur,
                  double *out, double a, double bq, double c)
{
    const int G = h->ld / 2, nb = blocks_for(h->n_loc, G);
    DISPATCH_G(G, (cheb_step_kernel<GG><<<nb, 256, 0, h->stream>>>(h->d_indptr, h->d_indices, h->d_M, h->d_bcmask, dinv, btil, p_prev,
                                                                 p_cur, h->d_halo, out, a, bq, c, h->n_loc, h->ld)));
    h->launches++;
    CTL_CUDA(cudaGetLastError());
    return CTL_OK;
}

int pcb_u0_final(ctl_handle_s *h, const double *p, const double *wrap_b, double *u0, double sc, double sc_last)
{
    const int G = h->ld / 2, nb = blocks_for(h->n_loc, G);
    if (h->cfg.CN) {
        DISPATCH_G(G, (u0_final_kernel<true, GG><<<nb, 256, 0, h->stream>>>(p, h->d_bcmask, wrap_b, u0, sc, sc_last, h->n_loc, h->N, h->ld)));
    } else {
        DISPATCH_G(G, (u0_final_kernel<false, GG><<<nb, 256, 0, h->stream>>>(p, h->d_bcmask, wrap_b, u0, sc, sc_last, h->n_loc, h->N, h->ld)));
    }
    h->launches++;
    CTL_CUDA(cudaGetLastError());
    return CTL_OK;
}

int pcb_schur_rhs(ctl_handle_s *h, const double *u0, const double *b1, double *out)
{
    const int G = h->ld / 2, nb = blocks_for(h->n_loc, G);
    const bool cn = h->cfg.CN != 0, pl = h->per_level;
#define SR(CNV, PLV)                                                                                                   \
    DISPATCH_G(G, (schur_rhs_kernel<CNV, PLV, GG><<<nb, 256, 0, h->stream>>>(h->d_indptr, h->d_indices, h->d_M, h->d_K, h->d_bcmask, \
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
