// Fused all-at-once KKT operator apply (SURVEY.md rows K1 + K2 + K3).
//
// Replaces MultiBlockSystemMatrix.mult (preconditioner/preconditioner.py:375-543): the
// 8N-4 (CN) / 6N-4 (BE) MatMultAdd calls over separately assembled blocks
// (control/control.py:2889-2978), the T_1 / T_2 time-coupling transforms (437-470) and the
// DirichletBCNullspace pre/post corrections (384-393, 527-537) become ONE kernel:
//
//   * vectors live in the time-fastest layout X[row][ld]: the N time columns of a spatial
//     dof are contiguous, so a (sub-)warp reads the whole N-column row segment of a
//     gathered dof with coalesced 16-byte loads (lane l owns columns 2l, 2l+1);
//   * every time step shares the pattern and the values of M and K, so the matrix is read
//     once per row, not once per block (per-level values: [nnz][ld] panels, same access);
//   * the four products MV, KV, MZ, KZ of a row stay in registers; the block stencil and
//     T_1 / T_2 are neighbour exchanges along the time axis = warp shuffles;
//   * Dirichlet columns are eliminated from the value arrays at setup (x_c = P x) and
//     Dirichlet rows return x (y = P A P x + (I - P) x).
//
// HBM-bound: algorithmic bytes per apply = 32 n N + 20 nnz + 4 (n+1).
#include <algorithm>

#include "common.cuh"

namespace {

constexpr int STAGE_CAP = 1024;   // CSR entries a CTA stages in shared memory
constexpr int CHUNK = 8;
#ifndef SCHUNK
#define SCHUNK 4
#endif
#ifndef SMINB
#define SMINB 4
#endif          // matrix entries staged per pass (covers a P1 2-D row)

struct KktArgs {
    int n_rows;                   // owned rows
    int n_own_cols;               // columns < n_own_cols come from x, the rest from halo
    int N, ld;
    const int *indptr, *indices;
    const double *Mv, *Kv, *KTv;  // value sets (BC columns zeroed)
    const uint8_t *bcmask;
    const double *xv, *xz;        // input panels  [n_rows x ld]
    const double *hv, *hz;        // ghost rows    [n_halo x ld]
    double *y0, *y1;              // output panels
    double tau, beta;
    const int *tile_uptr, *tile_ucols;   // TMA tile plan (see common.cuh)
    int tile_count_off;
    const uint8_t *tile_slot;
    const uint8_t *rec;           // record stream (CTL_KKT_TMA=3), see common.cuh
    const int *rec_off;           // per row block, units of 16 bytes
    int rec_max;
};

__device__ __forceinline__ double2 ldg2(const double *p)
{
    return __ldg(reinterpret_cast<const double2 *>(p));
}

// G = lanes per row (ld = 2 G columns); 32 / G rows per warp.
template <bool CN, bool PER_LEVEL, int G>
__global__ void __launch_bounds__(256) kkt_apply_kernel(const KktArgs a)
{
    constexpr int ROWS_PER_WARP = 32 / G;
    const int lane = threadIdx.x & 31;
    const int sub = lane / G;
    const int l = lane % G;
    const int warp = (blockIdx.x * (blockDim.x >> 5)) + (threadIdx.x >> 5);
    const int row = warp * ROWS_PER_WARP + sub;
    const bool live = row < a.n_rows;
    const int r = live ? row : a.n_rows - 1;        // keep the whole warp in the shuffles
    const int c0 = 2 * l;
    const int ld = a.ld;

    double mv0 = 0, mv1 = 0, kv0 = 0, kv1 = 0, mz0 = 0, mz1 = 0, kz0 = 0, kz1 = 0;

    const int kbeg = a.indptr[r], kend = a.indptr[r + 1];
    for (int k0 = kbeg; k0 < kend; k0 += CHUNK) {
        int col[CHUNK];
        double m[CHUNK];
#pragma unroll
        for (int j = 0; j < CHUNK; ++j) {
            const int k = k0 + j;
            const bool ok = k < kend;
            col[j] = ok ? __ldg(a.indices + k) : r;
            m[j] = ok ? __ldg(a.Mv + k) : 0.0;
        }
        double2 xv[CHUNK], xz[CHUNK];
#pragma unroll
        for (int j = 0; j < CHUNK; ++j) {
            const int c = col[j];
            const bool own = c < a.n_own_cols;
            const double *pv = own ? a.xv + (size_t)c * ld : a.hv + (size_t)(c - a.n_own_cols) * ld;
            const double *pz = own ? a.xz + (size_t)c * ld : a.hz + (size_t)(c - a.n_own_cols) * ld;
            xv[j] = ldg2(pv + c0);
            xz[j] = ldg2(pz + c0);
        }
        if (!PER_LEVEL) {
            double kk[CHUNK], kt[CHUNK];
#pragma unroll
            for (int j = 0; j < CHUNK; ++j) {
                const int k = k0 + j;
                const bool ok = k < kend;
                kk[j] = ok ? __ldg(a.Kv + k) : 0.0;
                kt[j] = ok ? __ldg(a.KTv + k) : 0.0;
            }
#pragma unroll
            for (int j = 0; j < CHUNK; ++j) {
                mv0 = fma(m[j], xv[j].x, mv0);
                mv1 = fma(m[j], xv[j].y, mv1);
                mz0 = fma(m[j], xz[j].x, mz0);
                mz1 = fma(m[j], xz[j].y, mz1);
                kv0 = fma(kk[j], xv[j].x, kv0);
                kv1 = fma(kk[j], xv[j].y, kv1);
                kz0 = fma(kt[j], xz[j].x, kz0);
                kz1 = fma(kt[j], xz[j].y, kz1);
            }
        } else {
#pragma unroll
            for (int j = 0; j < CHUNK; ++j) {
                const int k = k0 + j;
                if (k < kend) {
                    const double2 kk = ldg2(a.Kv + (size_t)k * ld + c0);
                    const double2 kt = ldg2(a.KTv + (size_t)k * ld + c0);
                    mv0 = fma(m[j], xv[j].x, mv0);
                    mv1 = fma(m[j], xv[j].y, mv1);
                    mz0 = fma(m[j], xz[j].x, mz0);
                    mz1 = fma(m[j], xz[j].y, mz1);
                    kv0 = fma(kk.x, xv[j].x, kv0);
                    kv1 = fma(kk.y, xv[j].y, kv1);
                    kz0 = fma(kt.x, xz[j].x, kz0);
                    kz1 = fma(kt.y, xz[j].y, kz1);
                }
            }
        }
    }

    // ---- block stencil + T_1 / T_2 along the time axis (columns c0, c0+1 of this lane)
    const unsigned full = 0xffffffffu;
    const bool first = (l == 0), last = (l == G - 1);
    const int N = a.N;
    const bool in0 = c0 < N, in1 = c0 + 1 < N;
    double y00, y01, y10, y11;
    if (CN) {
        const double h = 0.5 * a.tau, hb = h / a.beta;
        double t;
        t = __shfl_up_sync(full, mv1, 1, G);   const double mvp0 = first ? 0.0 : t;
        t = __shfl_up_sync(full, kv1, 1, G);   const double kvp0 = first ? 0.0 : t;
        t = __shfl_down_sync(full, kz0, 1, G); const double kzn1 = last ? 0.0 : t;
        t = __shfl_down_sync(full, mz0, 1, G); const double mzn1 = last ? 0.0 : t;
        // rows of the untransformed block system (control/control.py:2938-2958)
        double r00 = h * (mvp0 + mv0) + h * (kz0 + kz1) + mz0 - mz1;
        double r01 = h * (mv0 + mv1) + h * (kz1 + kzn1) + mz1 - mzn1;
        double r10 = h * (kvp0 + kv0) - mvp0 + mv0 - hb * (mz0 + mz1);
        double r11 = h * (kv0 + kv1) - mv0 + mv1 - hb * (mz1 + mzn1);
        if (!in0) { r00 = 0.0; r10 = 0.0; }
        if (!in1) { r01 = 0.0; r11 = 0.0; }
        // T_1: add the next block row; T_2: add the previous one (preconditioner.py:33-60)
        t = __shfl_down_sync(full, r00, 1, G); const double r0n = last ? 0.0 : t;
        t = __shfl_up_sync(full, r11, 1, G);   const double r1p = first ? 0.0 : t;
        y00 = r00 + r01;
        y01 = r01 + r0n;
        y10 = r10 + r1p;
        y11 = r11 + r10;
    } else {
        const double tau = a.tau, tb = tau / a.beta;
        double t;
        t = __shfl_up_sync(full, mv1, 1, G);   const double mvp0 = first ? 0.0 : t;
        t = __shfl_down_sync(full, mz0, 1, G); const double mzn1 = last ? 0.0 : t;
        // control/control.py:2907-2928, 2960-2978: block_00 last row None, block_11 row 0 None
        y00 = tau * kz0 + mz0 + ((c0 < N - 1) ? (tau * mv0 - mz1) : 0.0);
        y01 = tau * kz1 + mz1 + ((c0 + 1 < N - 1) ? (tau * mv1 - mzn1) : 0.0);
        y10 = tau * kv0 + mv0 + ((c0 >= 1) ? (-mvp0 - tb * mz0) : 0.0);
        y11 = tau * kv1 + mv1 + (-mv0 - tb * mz1);
    }
    if (!in0) { y00 = 0.0; y10 = 0.0; }
    if (!in1) { y01 = 0.0; y11 = 0.0; }
    if (!live) return;
    if (a.bcmask[r]) {           // y = (I - P) x on constrained rows
        const double2 xv = ldg2(a.xv + (size_t)r * ld + c0);
        const double2 xz = ldg2(a.xz + (size_t)r * ld + c0);
        y00 = xv.x; y01 = xv.y; y10 = xz.x; y11 = xz.y;
    }
    *reinterpret_cast<double2 *>(a.y0 + (size_t)r * ld + c0) = make_double2(y00, y01);
    *reinterpret_cast<double2 *>(a.y1 + (size_t)r * ld + c0) = make_double2(y10, y11);
}


// ---------------------------------------------------------------------------------------
// Staged variant (the one normally launched): a CTA owns `rows_per_cta` consecutive rows.
// Their CSR entries are contiguous, so the CTA first copies them into shared memory with
// fully coalesced loads, already converted to what the inner loop needs: the BYTE offset of
// the gathered X row (32 bit) and the (M, K) value pair as one 16-byte word.  Afterwards the
// only global latency a warp sees per row is the gather of the X row segments themselves.
// History (profiles/): v1 unstaged = latency bound (1.51 ms); v2 staged = 0.74 ms but issue
// bound (67% issue slots, 574 M warp instructions of which 11% DFMA: address IMADs, constant
// reloads, predication selects); v3 below cuts the per-entry overhead to two LDS, one
// integer add, two LDG.128 and eight DFMA.
// ---------------------------------------------------------------------------------------
template <bool CN, bool PER_LEVEL, bool SYM, bool HALO, int G, int SC>
__global__ void __launch_bounds__(256, SMINB) kkt_apply_staged_kernel(const KktArgs a, const int rows_per_cta,
                                                                     const int cap)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: [cap+1] double2 (m, k) | [cap+1] double kt (only !SYM) | [cap+1] unsigned off | [rows+1] int ptr
    double2 *s_mk = reinterpret_cast<double2 *>(smem_raw);
    double *s_kt = reinterpret_cast<double *>(s_mk + (cap + 1));
    unsigned *s_off = reinterpret_cast<unsigned *>(SYM || PER_LEVEL ? s_kt : s_kt + (cap + 1));
    int *s_ptr = reinterpret_cast<int *>(s_off + (cap + 1));

    constexpr int RPW = 32 / G;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int sub = lane / G, l = lane % G;
    const int c0 = 2 * l;
    const int ld = a.ld;
    const int r0 = blockIdx.x * rows_per_cta;
    const int nrows = min(rows_per_cta, a.n_rows - r0);
    const int n_own = a.n_own_cols;
    const char *__restrict__ xv_b = reinterpret_cast<const char *>(a.xv);
    const char *__restrict__ xz_b = reinterpret_cast<const char *>(a.xz);
    const char *__restrict__ hv_b = reinterpret_cast<const char *>(a.hv);
    const char *__restrict__ hz_b = reinterpret_cast<const char *>(a.hz);
    const unsigned lane_b = (unsigned)c0 * 8u;
    const unsigned row_b = (unsigned)ld * 8u;

    for (int i = threadIdx.x; i <= nrows; i += blockDim.x) s_ptr[i] = __ldg(a.indptr + r0 + i);
    __syncthreads();
    const int kb = s_ptr[0];
    const int cnt = s_ptr[nrows] - kb;
    for (int k = threadIdx.x; k < cnt; k += blockDim.x) {
        const int c = __ldg(a.indices + kb + k);
        // ghost rows (HALO) are addressed relative to the halo buffers: flag them in bit 31
        s_off[k] = (HALO && c >= n_own) ? (0x80000000u | ((unsigned)(c - n_own) * row_b)) : (unsigned)c * row_b;
        s_mk[k] = make_double2(__ldg(a.Mv + kb + k), PER_LEVEL ? 0.0 : __ldg(a.Kv + kb + k));
        if (!PER_LEVEL && !SYM) s_kt[k] = __ldg(a.KTv + kb + k);
    }
    if (threadIdx.x == 0) {          // sentinel entry: zero values, a valid row to gather
        s_off[cap] = 0u;
        s_mk[cap] = make_double2(0.0, 0.0);
        if (!PER_LEVEL && !SYM) s_kt[cap] = 0.0;
    }
    __syncthreads();

    const unsigned full = 0xffffffffu;
    const bool first = (l == 0), last = (l == G - 1);
    const int N = a.N;
    const bool in0 = c0 < N, in1 = c0 + 1 < N;
    const double tau = a.tau, beta = a.beta;
    const int nwarps = blockDim.x >> 5;

    for (int base = wid * RPW; base < nrows; base += nwarps * RPW) {
        const int lr_raw = base + sub;
        const bool live = lr_raw < nrows;
        const int lr = live ? lr_raw : nrows - 1;
        const int r = r0 + lr;
        const int kbeg = s_ptr[lr] - kb, kend = s_ptr[lr + 1] - kb;
        double mv0 = 0, mv1 = 0, kv0 = 0, kv1 = 0, mz0 = 0, mz1 = 0, kz0 = 0, kz1 = 0;
        for (int k0 = kbeg; k0 < kend; k0 += SC) {
            unsigned off[SC];
            int kk[SC];
#pragma unroll
            for (int j = 0; j < SC; ++j) {
                kk[j] = (k0 + j < kend) ? k0 + j : cap;       // padding slots read the sentinel
                off[j] = s_off[kk[j]];
            }
            double2 xv[SC], xz[SC];
#pragma unroll
            for (int j = 0; j < SC; ++j) {
                if (HALO && (off[j] & 0x80000000u)) {
                    const unsigned o = (off[j] & 0x7fffffffu) + lane_b;
                    xv[j] = __ldg(reinterpret_cast<const double2 *>(hv_b + o));
                    xz[j] = __ldg(reinterpret_cast<const double2 *>(hz_b + o));
                } else {
                    const unsigned o = off[j] + lane_b;
                    xv[j] = __ldg(reinterpret_cast<const double2 *>(xv_b + o));
                    xz[j] = __ldg(reinterpret_cast<const double2 *>(xz_b + o));
                }
            }
#pragma unroll
            for (int j = 0; j < SC; ++j) {
                const double2 mk = s_mk[kk[j]];
                mv0 = fma(mk.x, xv[j].x, mv0);
                mv1 = fma(mk.x, xv[j].y, mv1);
                mz0 = fma(mk.x, xz[j].x, mz0);
                mz1 = fma(mk.x, xz[j].y, mz1);
                if (!PER_LEVEL) {
                    const double kt = SYM ? mk.y : s_kt[kk[j]];
                    kv0 = fma(mk.y, xv[j].x, kv0);
                    kv1 = fma(mk.y, xv[j].y, kv1);
                    kz0 = fma(kt, xz[j].x, kz0);
                    kz1 = fma(kt, xz[j].y, kz1);
                } else if (kk[j] != cap) {
                    const double2 kp = ldg2(a.Kv + (size_t)(kb + kk[j]) * ld + c0);
                    const double2 kt = ldg2(a.KTv + (size_t)(kb + kk[j]) * ld + c0);
                    kv0 = fma(kp.x, xv[j].x, kv0);
                    kv1 = fma(kp.y, xv[j].y, kv1);
                    kz0 = fma(kt.x, xz[j].x, kz0);
                    kz1 = fma(kt.y, xz[j].y, kz1);
                }
            }
        }
        double y00, y01, y10, y11;
        if (CN) {
            const double h = 0.5 * tau, hb = h / beta;
            double t;
            t = __shfl_up_sync(full, mv1, 1, G);   const double mvp0 = first ? 0.0 : t;
            t = __shfl_up_sync(full, kv1, 1, G);   const double kvp0 = first ? 0.0 : t;
            t = __shfl_down_sync(full, kz0, 1, G); const double kzn1 = last ? 0.0 : t;
            t = __shfl_down_sync(full, mz0, 1, G); const double mzn1 = last ? 0.0 : t;
            double r00 = h * (mvp0 + mv0) + h * (kz0 + kz1) + mz0 - mz1;
            double r01 = h * (mv0 + mv1) + h * (kz1 + kzn1) + mz1 - mzn1;
            double r10 = h * (kvp0 + kv0) - mvp0 + mv0 - hb * (mz0 + mz1);
            double r11 = h * (kv0 + kv1) - mv0 + mv1 - hb * (mz1 + mzn1);
            if (!in0) { r00 = 0.0; r10 = 0.0; }
            if (!in1) { r01 = 0.0; r11 = 0.0; }
            t = __shfl_down_sync(full, r00, 1, G); const double r0n = last ? 0.0 : t;
            t = __shfl_up_sync(full, r11, 1, G);   const double r1p = first ? 0.0 : t;
            y00 = r00 + r01;
            y01 = r01 + r0n;
            y10 = r10 + r1p;
            y11 = r11 + r10;
        } else {
            const double tb = tau / beta;
            double t;
            t = __shfl_up_sync(full, mv1, 1, G);   const double mvp0 = first ? 0.0 : t;
            t = __shfl_down_sync(full, mz0, 1, G); const double mzn1 = last ? 0.0 : t;
            y00 = tau * kz0 + mz0 + ((c0 < N - 1) ? (tau * mv0 - mz1) : 0.0);
            y01 = tau * kz1 + mz1 + ((c0 + 1 < N - 1) ? (tau * mv1 - mzn1) : 0.0);
            y10 = tau * kv0 + mv0 + ((c0 >= 1) ? (-mvp0 - tb * mz0) : 0.0);
            y11 = tau * kv1 + mv1 + (-mv0 - tb * mz1);
        }
        if (!in0) { y00 = 0.0; y10 = 0.0; }
        if (!in1) { y01 = 0.0; y11 = 0.0; }
        if (live) {
            const size_t ro = (size_t)r * ld + c0;
            if (a.bcmask[r]) {
                const double2 xv = ldg2(a.xv + ro);
                const double2 xz = ldg2(a.xz + ro);
                y00 = xv.x; y01 = xv.y; y10 = xz.x; y11 = xz.y;
            }
            // streaming stores: the 1 GB result must not push the X rows still to be gathered out of L2
            __stcs(reinterpret_cast<double2 *>(a.y0 + ro), make_double2(y00, y01));
            __stcs(reinterpret_cast<double2 *>(a.y1 + ro), make_double2(y10, y11));
        }
    }
}

template <bool CN, bool PER_LEVEL, bool SYM, bool HALO>
void launch_staged_h(const KktArgs &a, int G, int rows_per_cta, int cap, int chunk, cudaStream_t s)
{
    const int blocks = ceil_div(a.n_rows, rows_per_cta);
    const size_t smem = (size_t)(cap + 1) * (16 + ((SYM || PER_LEVEL) ? 0 : 8) + 4) + (size_t)(rows_per_cta + 1) * 4;
#define LS(GG, CC) kkt_apply_staged_kernel<CN, PER_LEVEL, SYM, HALO, GG, CC><<<blocks, 256, smem, s>>>(a, rows_per_cta, cap)
    if (G == 32 && !PER_LEVEL) {
        // chunk = entries gathered per pass; matched to the row length so that no padding
        // slots are processed (7 for a P1 triangle mesh, 5 for the 15-point P1 tetrahedra stencil)
        switch (chunk) {
        case 5: LS(32, 5); break;
        case 7: LS(32, 7); break;
        case 8: LS(32, 8); break;
        default: LS(32, 4); break;
        }
        return;
    }
    switch (G) {
    case 4: LS(4, 4); break;
    case 8: LS(8, 4); break;
    case 16: LS(16, 4); break;
    default: LS(32, 4); break;
    }
#undef LS
}

template <bool CN, bool PER_LEVEL, bool SYM>
void launch_staged(const KktArgs &a, int G, int rows_per_cta, int cap, int chunk, bool halo, cudaStream_t s)
{
    if (halo) launch_staged_h<CN, PER_LEVEL, SYM, true>(a, G, rows_per_cta, cap, chunk, s);
    else launch_staged_h<CN, PER_LEVEL, SYM, false>(a, G, rows_per_cta, cap, chunk, s);
}

// ---------------------------------------------------------------------------------------
// TMA-staged variant (opt-in, CTL_KKT_TMA=1; round-1 result: correct, but 0.73 ms against 0.66 ms for
// the LDG-gather kernel above): v3 is bound by the L1 LDG data pipe (7 gathered 1 KB row segments per
// row, profiles/r01_kkt_apply_v3.txt).  Here a CTA owns 32 consecutive rows; the UNIQUE X
// rows they reference (host-built tile plan, about 3x34 for a 7-point mesh stencil) are
// copied once into shared memory by the TMA engine -- one cp.async.bulk of ld*8 bytes per
// row and panel, completion on an mbarrier -- while the other warps stage the CSR entries.
// All gathers of the inner loop are then LDS.128 from the tile (shared memory delivers
// 128 B/clk/SM against about 64 B/clk for LDG hits), and HBM -> SM traffic no longer passes
// through registers.  Two CTAs per SM (about 108 KB each) overlap one CTA's copy with the
// other's arithmetic.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p)
{
    return (unsigned)__cvta_generic_to_shared(p);
}

template <bool CN, bool SYM, bool HALO, int G, int TR>
__global__ void __launch_bounds__(TR * 8, 512 / (TR * 8) * 2) kkt_apply_tma_kernel(const KktArgs a, const int cap, const int umax)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int ld = a.ld;
    const unsigned row_b = (unsigned)ld * 8u;
    // layout: tile_v [umax*row_b] | tile_z [umax*row_b] | mk [cap+1] double2 | kt [cap+1] double (!SYM)
    //         | off [cap+1] unsigned | ptr [TR+1] int | mbarrier (8 B)
    unsigned char *tile_v = smem_raw;
    unsigned char *tile_z = tile_v + (size_t)umax * row_b;
    double2 *s_mk = reinterpret_cast<double2 *>(tile_z + (size_t)umax * row_b);
    double *s_kt = reinterpret_cast<double *>(s_mk + (cap + 1));
    unsigned *s_off = reinterpret_cast<unsigned *>(SYM ? s_kt : s_kt + (cap + 1));
    int *s_ptr = reinterpret_cast<int *>(s_off + (cap + 1));
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(
        (reinterpret_cast<uintptr_t>(s_ptr + (TR + 1)) + 7) & ~(uintptr_t)7);

    constexpr int RPW = 32 / G;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int sub = lane / G, l = lane % G;
    const int c0 = 2 * l;
    const int r0 = blockIdx.x * TR;
    const int nrows = min(TR, a.n_rows - r0);
    const int ub = __ldg(a.tile_uptr + blockIdx.x);
    const int n_runs = __ldg(a.tile_uptr + blockIdx.x + 1) - ub;
    const int U = __ldg(a.tile_ucols + a.tile_count_off + blockIdx.x);
    const unsigned mbar_s = smem_u32(mbar);

    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_s));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (wid == 0) {
        // producer warp: arm the barrier with the byte count, then one bulk copy per row and panel
        if (lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_s), "r"(2u * (unsigned)U * row_b)
                         : "memory");
        __syncwarp();
        for (int u = lane; u < n_runs; u += 32) {
            const int c = __ldg(a.tile_ucols + 3 * (ub + u));
            const unsigned bytes = (unsigned)__ldg(a.tile_ucols + 3 * (ub + u) + 1) * row_b;
            const size_t dst = (size_t)__ldg(a.tile_ucols + 3 * (ub + u) + 2) * row_b;
            const double *sv, *sz;
            if (HALO && c >= a.n_own_cols) {
                sv = a.hv + (size_t)(c - a.n_own_cols) * ld;
                sz = a.hz + (size_t)(c - a.n_own_cols) * ld;
            } else {
                sv = a.xv + (size_t)c * ld;
                sz = a.xz + (size_t)c * ld;
            }
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(tile_v + dst)),
                         "l"(sv), "r"(bytes), "r"(mbar_s)
                         : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(tile_z + dst)),
                         "l"(sz), "r"(bytes), "r"(mbar_s)
                         : "memory");
        }
    }
    // everybody: stage the CSR entries of the block (slot -> byte offset inside the tile)
    for (int i = threadIdx.x; i <= nrows; i += blockDim.x) s_ptr[i] = __ldg(a.indptr + r0 + i);
    __syncthreads();
    const int kb = s_ptr[0];
    const int cnt = s_ptr[nrows] - kb;
    for (int k = threadIdx.x; k < cnt; k += blockDim.x) {
        s_off[k] = (unsigned)__ldg(a.tile_slot + kb + k) * row_b;
        s_mk[k] = make_double2(__ldg(a.Mv + kb + k), __ldg(a.Kv + kb + k));
        if (!SYM) s_kt[k] = __ldg(a.KTv + kb + k);
    }
    if (threadIdx.x == 0) {
        s_off[cap] = 0u;
        s_mk[cap] = make_double2(0.0, 0.0);
        if (!SYM) s_kt[cap] = 0.0;
    }
    __syncthreads();
    // wait for the tile (phase 0 of the barrier)
    {
        unsigned done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done)
                         : "r"(mbar_s), "r"(0u)
                         : "memory");
        }
    }

    const unsigned full = 0xffffffffu;
    const bool first = (l == 0), last = (l == G - 1);
    const int N = a.N;
    const bool in0 = c0 < N, in1 = c0 + 1 < N;
    const double tau = a.tau, beta = a.beta;
    const unsigned lane_b = (unsigned)c0 * 8u;
    const int nwarps = blockDim.x >> 5;

    for (int base = wid * RPW; base < nrows; base += nwarps * RPW) {
        const int lr_raw = base + sub;
        const bool live = lr_raw < nrows;
        const int lr = live ? lr_raw : nrows - 1;
        const int r = r0 + lr;
        const int kbeg = s_ptr[lr] - kb, kend = s_ptr[lr + 1] - kb;
        double mv0 = 0, mv1 = 0, kv0 = 0, kv1 = 0, mz0 = 0, mz1 = 0, kz0 = 0, kz1 = 0;
        for (int k0 = kbeg; k0 < kend; k0 += SCHUNK) {
#pragma unroll
            for (int j = 0; j < SCHUNK; ++j) {
                const int kk = (k0 + j < kend) ? k0 + j : cap;
                const unsigned o = s_off[kk] + lane_b;
                const double2 xv = *reinterpret_cast<const double2 *>(tile_v + o);
                const double2 xz = *reinterpret_cast<const double2 *>(tile_z + o);
                const double2 mk = s_mk[kk];
                const double kt = SYM ? mk.y : s_kt[kk];
                mv0 = fma(mk.x, xv.x, mv0);
                mv1 = fma(mk.x, xv.y, mv1);
                mz0 = fma(mk.x, xz.x, mz0);
                mz1 = fma(mk.x, xz.y, mz1);
                kv0 = fma(mk.y, xv.x, kv0);
                kv1 = fma(mk.y, xv.y, kv1);
                kz0 = fma(kt, xz.x, kz0);
                kz1 = fma(kt, xz.y, kz1);
            }
        }
        double y00, y01, y10, y11;
        if (CN) {
            const double h = 0.5 * tau, hb = h / beta;
            double t;
            t = __shfl_up_sync(full, mv1, 1, G);   const double mvp0 = first ? 0.0 : t;
            t = __shfl_up_sync(full, kv1, 1, G);   const double kvp0 = first ? 0.0 : t;
            t = __shfl_down_sync(full, kz0, 1, G); const double kzn1 = last ? 0.0 : t;
            t = __shfl_down_sync(full, mz0, 1, G); const double mzn1 = last ? 0.0 : t;
            double r00 = h * (mvp0 + mv0) + h * (kz0 + kz1) + mz0 - mz1;
            double r01 = h * (mv0 + mv1) + h * (kz1 + kzn1) + mz1 - mzn1;
            double r10 = h * (kvp0 + kv0) - mvp0 + mv0 - hb * (mz0 + mz1);
            double r11 = h * (kv0 + kv1) - mv0 + mv1 - hb * (mz1 + mzn1);
            if (!in0) { r00 = 0.0; r10 = 0.0; }
            if (!in1) { r01 = 0.0; r11 = 0.0; }
            t = __shfl_down_sync(full, r00, 1, G); const double r0n = last ? 0.0 : t;
            t = __shfl_up_sync(full, r11, 1, G);   const double r1p = first ? 0.0 : t;
            y00 = r00 + r01;
            y01 = r01 + r0n;
            y10 = r10 + r1p;
            y11 = r11 + r10;
        } else {
            const double tb = tau / beta;
            double t;
            t = __shfl_up_sync(full, mv1, 1, G);   const double mvp0 = first ? 0.0 : t;
            t = __shfl_down_sync(full, mz0, 1, G); const double mzn1 = last ? 0.0 : t;
            y00 = tau * kz0 + mz0 + ((c0 < N - 1) ? (tau * mv0 - mz1) : 0.0);
            y01 = tau * kz1 + mz1 + ((c0 + 1 < N - 1) ? (tau * mv1 - mzn1) : 0.0);
            y10 = tau * kv0 + mv0 + ((c0 >= 1) ? (-mvp0 - tb * mz0) : 0.0);
            y11 = tau * kv1 + mv1 + (-mv0 - tb * mz1);
        }
        if (!in0) { y00 = 0.0; y10 = 0.0; }
        if (!in1) { y01 = 0.0; y11 = 0.0; }
        if (live) {
            const size_t ro = (size_t)r * ld + c0;
            if (a.bcmask[r]) {
                const double2 xv = ldg2(a.xv + ro);
                const double2 xz = ldg2(a.xz + ro);
                y00 = xv.x; y01 = xv.y; y10 = xz.x; y11 = xz.y;
            }
            *reinterpret_cast<double2 *>(a.y0 + ro) = make_double2(y00, y01);
            *reinterpret_cast<double2 *>(a.y1 + ro) = make_double2(y10, y11);
        }
    }
}

// ---------------------------------------------------------------------------------------
// Persistent two-stage variant of the TMA-staged kernel (opt-in, CTL_KKT_TMA=2; end of round 1: bit-identical
// to the staged kernel on a B200 but 2.2x slower in this first form -- the CSR slice of a block is still staged
// with two dependent global round trips inside the single resident CTA).  The
// one-shot kernel above copies a tile and then consumes it, relying on a second resident CTA for overlap;
// here every CTA loops over row blocks and the bulk copies of block i+1 are issued into the other stage
// BEFORE block i is consumed, so a whole tile (55-104 KB per SM, no registers) is in flight while the SM
// computes -- the "independent gathers in flight" that bound the LDG kernel.
// Stage reuse is ordered by the __syncthreads() that ends an iteration; each stage has its own mbarrier,
// whose k-th use completes phase k & 1.
// ---------------------------------------------------------------------------------------
// REC (CTL_KKT_TMA=3): the CSR slice of a row block is not staged by the CTA (two dependent global round trips
// per block, what made the first form slow) but arrives as ONE more bulk copy of the block's record (api.cu) into
// the stage, counted on the same mbarrier: an iteration is then issue-next / wait / consume / barrier.
template <bool CN, bool SYM, bool HALO, int G, int TR, bool REC>
__global__ void __launch_bounds__(TR * 8) kkt_apply_tma_pipe_kernel(const KktArgs a, const int cap, const int umax,
                                                                   const int n_blocks)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int ld = a.ld;
    const unsigned row_b = (unsigned)ld * 8u;
    const size_t tile_b = (size_t)umax * row_b;
    // layout !REC: stage 0 [tile_v | tile_z] | stage 1 [tile_v | tile_z] | mk [cap+1] double2 | kt [cap+1] double
    //              (!SYM) | off [cap+1] unsigned | ptr [TR+1] int | 2 mbarriers
    // layout  REC: stage 0 [tile_v | tile_z | record] | stage 1 [...] | 2 mbarriers
    const size_t stage_b = 2 * tile_b + (REC ? (size_t)a.rec_max : 0);
    unsigned char *tiles = smem_raw;
    double2 *s_mk = reinterpret_cast<double2 *>(tiles + 2 * stage_b);
    double *s_kt = reinterpret_cast<double *>(s_mk + (cap + 1));
    unsigned *s_off = reinterpret_cast<unsigned *>(SYM ? s_kt : s_kt + (cap + 1));
    int *s_ptr = reinterpret_cast<int *>(s_off + (cap + 1));
    unsigned long long *mbar = REC ? reinterpret_cast<unsigned long long *>(tiles + 2 * stage_b)
                                   : reinterpret_cast<unsigned long long *>(
                                         (reinterpret_cast<uintptr_t>(s_ptr + (TR + 1)) + 7) & ~(uintptr_t)7);
    constexpr unsigned HDR = ((TR + 1) * 4 + 15) & ~15;

    constexpr int RPW = 32 / G;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int sub = lane / G, l = lane % G;
    const int c0 = 2 * l;
    const unsigned mbar_s0 = smem_u32(mbar), mbar_s1 = smem_u32(mbar + 1);

    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_s0));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_s1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // producer (warp 0): arm the stage's barrier with the byte count, then one bulk copy per run and panel
    auto issue_tile = [&](const int blk, const int st) {
        const unsigned mb = st ? mbar_s1 : mbar_s0;
        unsigned char *tv = tiles + (size_t)st * stage_b;
        unsigned char *tz = tv + tile_b;
        const int ub = __ldg(a.tile_uptr + blk);
        const int n_runs = __ldg(a.tile_uptr + blk + 1) - ub;
        const int U = __ldg(a.tile_ucols + a.tile_count_off + blk);
        unsigned rec_bytes = 0;
        size_t rec_src = 0;
        if (REC) {
            const int o0 = __ldg(a.rec_off + blk), o1 = __ldg(a.rec_off + blk + 1);
            rec_src = (size_t)o0 * 16;
            rec_bytes = (unsigned)(o1 - o0) * 16u;
        }
        if (lane == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb),
                         "r"(2u * (unsigned)U * row_b + rec_bytes)
                         : "memory");
            if (REC)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32(tz + tile_b)),
                             "l"(a.rec + rec_src), "r"(rec_bytes), "r"(mb)
                             : "memory");
        }
        __syncwarp();
        for (int u = lane; u < n_runs; u += 32) {
            const int c = __ldg(a.tile_ucols + 3 * (ub + u));
            const unsigned bytes = (unsigned)__ldg(a.tile_ucols + 3 * (ub + u) + 1) * row_b;
            const size_t dst = (size_t)__ldg(a.tile_ucols + 3 * (ub + u) + 2) * row_b;
            const double *sv, *sz;
            if (HALO && c >= a.n_own_cols) {
                sv = a.hv + (size_t)(c - a.n_own_cols) * ld;
                sz = a.hz + (size_t)(c - a.n_own_cols) * ld;
            } else {
                sv = a.xv + (size_t)c * ld;
                sz = a.xz + (size_t)c * ld;
            }
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(tv + dst)),
                         "l"(sv), "r"(bytes), "r"(mb)
                         : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(tz + dst)),
                         "l"(sz), "r"(bytes), "r"(mb)
                         : "memory");
        }
    };

    const unsigned full = 0xffffffffu;
    const bool first = (l == 0), last = (l == G - 1);
    const int N = a.N;
    const bool in0 = c0 < N, in1 = c0 + 1 < N;
    const double tau = a.tau, beta = a.beta;
    const unsigned lane_b = (unsigned)c0 * 8u;
    const int nwarps = blockDim.x >> 5;

    if (wid == 0 && (int)blockIdx.x < n_blocks) issue_tile(blockIdx.x, 0);
    int it = 0;
    for (int blk = blockIdx.x; blk < n_blocks; blk += gridDim.x, ++it) {
        const int st = it & 1;
        const unsigned phase = (unsigned)(it >> 1) & 1u;
        const int nb = blk + gridDim.x;
        // stage st^1 was last read in iteration it-1, which ended with a __syncthreads(): safe to overwrite
        if (wid == 0 && nb < n_blocks) issue_tile(nb, st ^ 1);
        const int r0 = blk * TR;
        const int nrows = min(TR, a.n_rows - r0);
        int kb = 0, sentinel = cap;
        if (!REC) {
            for (int i = threadIdx.x; i <= nrows; i += blockDim.x) s_ptr[i] = __ldg(a.indptr + r0 + i);
            __syncthreads();
            kb = s_ptr[0];
            const int cnt = s_ptr[nrows] - kb;
            for (int k = threadIdx.x; k < cnt; k += blockDim.x) {
                s_off[k] = (unsigned)__ldg(a.tile_slot + kb + k) * row_b;
                s_mk[k] = make_double2(__ldg(a.Mv + kb + k), __ldg(a.Kv + kb + k));
                if (!SYM) s_kt[k] = __ldg(a.KTv + kb + k);
            }
            if (threadIdx.x == 0) {
                s_off[cap] = 0u;
                s_mk[cap] = make_double2(0.0, 0.0);
                if (!SYM) s_kt[cap] = 0.0;
            }
            __syncthreads();
        }
        {   // wait for this stage's tile
            const unsigned mb = st ? mbar_s1 : mbar_s0;
            unsigned done = 0;
            while (!done) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done)
                             : "r"(mb), "r"(phase)
                             : "memory");
            }
        }
        const unsigned char *tile_v = tiles + (size_t)st * stage_b;
        const unsigned char *tile_z = tile_v + tile_b;
        // where this block's CSR slice lives: the CTA-staged arrays, or the record that came with the tile
        const int *b_ptr = s_ptr;
        const double2 *b_mk = s_mk;
        const double *b_kt = s_kt;
        const unsigned *b_off = s_off;
        if (REC) {
            const unsigned char *rec = tile_z + tile_b;
            b_ptr = reinterpret_cast<const int *>(rec);
            const int cnt = b_ptr[TR];
            sentinel = cnt;                                   // entry cnt of a record is the zero sentinel
            b_mk = reinterpret_cast<const double2 *>(rec + HDR);
            b_kt = reinterpret_cast<const double *>(rec + HDR + (size_t)(cnt + 1) * 16);
            b_off = reinterpret_cast<const unsigned *>(
                rec + HDR + (size_t)(cnt + 1) * 16 + (SYM ? 0 : (((size_t)(cnt + 1) * 8 + 15) & ~(size_t)15)));
        }
        for (int base = wid * RPW; base < nrows; base += nwarps * RPW) {
            const int lr_raw = base + sub;
            const bool live = lr_raw < nrows;
            const int lr = live ? lr_raw : nrows - 1;
            const int r = r0 + lr;
            const int kbeg = b_ptr[lr] - kb, kend = b_ptr[lr + 1] - kb;
            double mv0 = 0, mv1 = 0, kv0 = 0, kv1 = 0, mz0 = 0, mz1 = 0, kz0 = 0, kz1 = 0;
            for (int k0 = kbeg; k0 < kend; k0 += SCHUNK) {
#pragma unroll
                for (int j = 0; j < SCHUNK; ++j) {
                    const int kk = (k0 + j < kend) ? k0 + j : sentinel;
                    const unsigned o = b_off[kk] + lane_b;
                    const double2 xv = *reinterpret_cast<const double2 *>(tile_v + o);
                    const double2 xz = *reinterpret_cast<const double2 *>(tile_z + o);
                    const double2 mk = b_mk[kk];
                    const double kt = SYM ? mk.y : b_kt[kk];
                    mv0 = fma(mk.x, xv.x, mv0);
                    mv1 = fma(mk.x, xv.y, mv1);
                    mz0 = fma(mk.x, xz.x, mz0);
                    mz1 = fma(mk.x, xz.y, mz1);
                    kv0 = fma(mk.y, xv.x, kv0);
                    kv1 = fma(mk.y, xv.y, kv1);
                    kz0 = fma(kt, xz.x, kz0);
                    kz1 = fma(kt, xz.y, kz1);
                }
            }
            double y00, y01, y10, y11;
            if (CN) {
                const double h = 0.5 * tau, hb = h / beta;
                double t;
                t = __shfl_up_sync(full, mv1, 1, G);   const double mvp0 = first ? 0.0 : t;
                t = __shfl_up_sync(full, kv1, 1, G);   const double kvp0 = first ? 0.0 : t;
                t = __shfl_down_sync(full, kz0, 1, G); const double kzn1 = last ? 0.0 : t;
                t = __shfl_down_sync(full, mz0, 1, G); const double mzn1 = last ? 0.0 : t;
                double r00 = h * (mvp0 + mv0) + h * (kz0 + kz1) + mz0 - mz1;
                double r01 = h * (mv0 + mv1) + h * (kz1 + kzn1) + mz1 - mzn1;
                double r10 = h * (kvp0 + kv0) - mvp0 + mv0 - hb * (mz0 + mz1);
                double r11 = h * (kv0 + kv1) - mv0 + mv1 - hb * (mz1 + mzn1);
                if (!in0) { r00 = 0.0; r10 = 0.0; }
                if (!in1) { r01 = 0.0; r11 = 0.0; }
                t = __shfl_down_sync(full, r00, 1, G); const double r0n = last ? 0.0 : t;
                t = __shfl_up_sync(full, r11, 1, G);   const double r1p = first ? 0.0 : t;
                y00 = r00 + r01;
                y01 = r01 + r0n;
                y10 = r10 + r1p;
                y11 = r11 + r10;
            } else {
                const double tb = tau / beta;
                double t;
                t = __shfl_up_sync(full, mv1, 1, G);   const double mvp0 = first ? 0.0 : t;
                t = __shfl_down_sync(full, mz0, 1, G); const double mzn1 = last ? 0.0 : t;
                y00 = tau * kz0 + mz0 + ((c0 < N - 1) ? (tau * mv0 - mz1) : 0.0);
                y01 = tau * kz1 + mz1 + ((c0 + 1 < N - 1) ? (tau * mv1 - mzn1) : 0.0);
                y10 = tau * kv0 + mv0 + ((c0 >= 1) ? (-mvp0 - tb * mz0) : 0.0);
                y11 = tau * kv1 + mv1 + (-mv0 - tb * mz1);
            }
            if (!in0) { y00 = 0.0; y10 = 0.0; }
            if (!in1) { y01 = 0.0; y11 = 0.0; }
            if (live) {
                const size_t ro = (size_t)r * ld + c0;
                if (a.bcmask[r]) {
                    const double2 xv = ldg2(a.xv + ro);
                    const double2 xz = ldg2(a.xz + ro);
                    y00 = xv.x; y01 = xv.y; y10 = xz.x; y11 = xz.y;
                }
                __stcs(reinterpret_cast<double2 *>(a.y0 + ro), make_double2(y00, y01));
                __stcs(reinterpret_cast<double2 *>(a.y1 + ro), make_double2(y10, y11));
            }
        }
        __syncthreads();      // every read of stage st and of the CSR staging is done
    }
}

template <bool CN, bool SYM, bool HALO, int G, int TR, bool REC>
cudaError_t launch_tma_pipe_gt(const KktArgs &a, int cap, int umax, cudaStream_t s)
{
    const int n_blocks = ceil_div(a.n_rows, TR);
    const size_t smem = REC ? (size_t)4 * umax * a.ld * 8 + (size_t)2 * a.rec_max + 32
                            : (size_t)4 * umax * a.ld * 8 + (size_t)(cap + 1) * (16 + (SYM ? 0 : 8) + 4) + (TR + 1) * 4 + 32;
    auto kern = kkt_apply_tma_pipe_kernel<CN, SYM, HALO, G, TR, REC>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        if (n_sm <= 0) n_sm = 148;
    }
    const int per_sm = std::max(1, (int)((227 * 1024) / (smem + 1024)));
    const int grid = std::min(n_blocks, n_sm * per_sm);
    kern<<<grid, TR * 8, smem, s>>>(a, cap, umax, n_blocks);
    return cudaSuccess;
}

template <bool CN, bool SYM, bool HALO>
cudaError_t launch_tma_pipe(const KktArgs &a, int G, int cap, int umax, int tile_rows, cudaStream_t s)
{
    // ld = 64 (G = 32) only: the configuration the pipeline is meant for
    if (G != 32) return cudaErrorInvalidValue;
    if (a.rec) {
        if (tile_rows == 16) return launch_tma_pipe_gt<CN, SYM, HALO, 32, 16, true>(a, cap, umax, s);
        return launch_tma_pipe_gt<CN, SYM, HALO, 32, 32, true>(a, cap, umax, s);
    }
    if (tile_rows == 16) return launch_tma_pipe_gt<CN, SYM, HALO, 32, 16, false>(a, cap, umax, s);
    return launch_tma_pipe_gt<CN, SYM, HALO, 32, 32, false>(a, cap, umax, s);
}

template <bool CN, bool SYM, bool HALO, int G, int TR>
cudaError_t launch_tma_gt(const KktArgs &a, int cap, int umax, cudaStream_t s)
{
    const int blocks = ceil_div(a.n_rows, TR);
    const size_t smem = (size_t)2 * umax * a.ld * 8 + (size_t)(cap + 1) * (16 + (SYM ? 0 : 8) + 4) + (TR + 1) * 4 + 16;
    auto kern = kkt_apply_tma_kernel<CN, SYM, HALO, G, TR>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<blocks, TR * 8, smem, s>>>(a, cap, umax);
    return cudaSuccess;
}

template <bool CN, bool SYM, bool HALO, int G>
cudaError_t launch_tma_g(const KktArgs &a, int cap, int umax, int tile_rows, cudaStream_t s)
{
    if (tile_rows == 16) return launch_tma_gt<CN, SYM, HALO, G, 16>(a, cap, umax, s);
    return launch_tma_gt<CN, SYM, HALO, G, 32>(a, cap, umax, s);
}

template <bool CN, bool SYM, bool HALO>
cudaError_t launch_tma(const KktArgs &a, int G, int cap, int umax, int tile_rows, cudaStream_t s)
{
    switch (G) {
    case 4: return launch_tma_g<CN, SYM, HALO, 4>(a, cap, umax, tile_rows, s);
    case 8: return launch_tma_g<CN, SYM, HALO, 8>(a, cap, umax, tile_rows, s);
    case 16: return launch_tma_g<CN, SYM, HALO, 16>(a, cap, umax, tile_rows, s);
    default: return launch_tma_g<CN, SYM, HALO, 32>(a, cap, umax, tile_rows, s);
    }
}

// ---------------------------------------------------------------------------------------
// Wide variant for more than 64 time blocks (ld = 128 or 256): one warp per row, every lane
// owns CPL = ld / 32 CONSECUTIVE time columns, so the block stencil and T_1 / T_2 only cross
// lanes at the ends of a lane's column group.  Same staging as the kernel above.
// ---------------------------------------------------------------------------------------
template <int CPL>
__device__ __forceinline__ void time_prev(const double (&a)[CPL], double (&out)[CPL], int lane)
{
    const double t = __shfl_up_sync(0xffffffffu, a[CPL - 1], 1);
    out[0] = lane == 0 ? 0.0 : t;
#pragma unroll
    for (int c = 1; c < CPL; ++c) out[c] = a[c - 1];
}

template <int CPL>
__device__ __forceinline__ void time_next(const double (&a)[CPL], double (&out)[CPL], int lane)
{
    const double t = __shfl_down_sync(0xffffffffu, a[0], 1);
#pragma unroll
    for (int c = 0; c < CPL - 1; ++c) out[c] = a[c + 1];
    out[CPL - 1] = lane == 31 ? 0.0 : t;
}

template <bool CN, bool PER_LEVEL, bool SYM, bool HALO, int CPL>
__global__ void __launch_bounds__(256) kkt_apply_wide_kernel(const KktArgs a, const int rows_per_cta, const int cap)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2 *s_mk = reinterpret_cast<double2 *>(smem_raw);
    double *s_kt = reinterpret_cast<double *>(s_mk + (cap + 1));
    int *s_col = reinterpret_cast<int *>(SYM || PER_LEVEL ? s_kt : s_kt + (cap + 1));
    int *s_ptr = s_col + (cap + 1);

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int c0 = CPL * lane;
    const int ld = a.ld;
    const int r0 = blockIdx.x * rows_per_cta;
    const int nrows = min(rows_per_cta, a.n_rows - r0);

    for (int i = threadIdx.x; i <= nrows; i += blockDim.x) s_ptr[i] = __ldg(a.indptr + r0 + i);
    __syncthreads();
    const int kb = s_ptr[0];
    const int cnt = s_ptr[nrows] - kb;
    for (int k = threadIdx.x; k < cnt; k += blockDim.x) {
        s_col[k] = __ldg(a.indices + kb + k);
        s_mk[k] = make_double2(__ldg(a.Mv + kb + k), PER_LEVEL ? 0.0 : __ldg(a.Kv + kb + k));
        if (!PER_LEVEL && !SYM) s_kt[k] = __ldg(a.KTv + kb + k);
    }
    __syncthreads();

    const int N = a.N;
    const double tau = a.tau, beta = a.beta;
    for (int lr = wid; lr < nrows; lr += (blockDim.x >> 5)) {
        const int r = r0 + lr;
        const int kbeg = s_ptr[lr] - kb, kend = s_ptr[lr + 1] - kb;
        double mv[CPL], kv[CPL], mz[CPL], kz[CPL];
#pragma unroll
        for (int c = 0; c < CPL; ++c) mv[c] = kv[c] = mz[c] = kz[c] = 0.0;
        for (int k = kbeg; k < kend; ++k) {
            const int col = s_col[k];
            const bool own = !HALO || col < a.n_own_cols;
            const double *pv = own ? a.xv + (size_t)col * ld : a.hv + (size_t)(col - a.n_own_cols) * ld;
            const double *pz = own ? a.xz + (size_t)col * ld : a.hz + (size_t)(col - a.n_own_cols) * ld;
            const double2 mk = s_mk[k];
            const double kt = (SYM || PER_LEVEL) ? mk.y : s_kt[k];
#pragma unroll
            for (int q = 0; q < CPL / 2; ++q) {
                const double2 xv = ldg2(pv + c0 + 2 * q);
                const double2 xz = ldg2(pz + c0 + 2 * q);
                double2 kk = make_double2(mk.y, mk.y), kkt = make_double2(kt, kt);
                if (PER_LEVEL) {
                    kk = ldg2(a.Kv + (size_t)(kb + k) * ld + c0 + 2 * q);
                    kkt = ldg2(a.KTv + (size_t)(kb + k) * ld + c0 + 2 * q);
                }
                mv[2 * q] = fma(mk.x, xv.x, mv[2 * q]);
                mv[2 * q + 1] = fma(mk.x, xv.y, mv[2 * q + 1]);
                mz[2 * q] = fma(mk.x, xz.x, mz[2 * q]);
                mz[2 * q + 1] = fma(mk.x, xz.y, mz[2 * q + 1]);
                kv[2 * q] = fma(kk.x, xv.x, kv[2 * q]);
                kv[2 * q + 1] = fma(kk.y, xv.y, kv[2 * q + 1]);
                kz[2 * q] = fma(kkt.x, xz.x, kz[2 * q]);
                kz[2 * q + 1] = fma(kkt.y, xz.y, kz[2 * q + 1]);
            }
        }
        double y0[CPL], y1[CPL];
        if (CN) {
            const double h = 0.5 * tau, hb = h / beta;
            double mvp[CPL], kvp[CPL], kzn[CPL], mzn[CPL], r0v[CPL], r1v[CPL], r0n[CPL], r1p[CPL];
            time_prev<CPL>(mv, mvp, lane);
            time_prev<CPL>(kv, kvp, lane);
            time_next<CPL>(kz, kzn, lane);
            time_next<CPL>(mz, mzn, lane);
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                const bool in = c0 + c < N;
                r0v[c] = in ? h * (mvp[c] + mv[c]) + h * (kz[c] + kzn[c]) + mz[c] - mzn[c] : 0.0;
                r1v[c] = in ? h * (kvp[c] + kv[c]) - mvp[c] + mv[c] - hb * (mz[c] + mzn[c]) : 0.0;
            }
            time_next<CPL>(r0v, r0n, lane);
            time_prev<CPL>(r1v, r1p, lane);
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                y0[c] = r0v[c] + r0n[c];
                y1[c] = r1v[c] + r1p[c];
            }
        } else {
            const double tb = tau / beta;
            double mvp[CPL], mzn[CPL];
            time_prev<CPL>(mv, mvp, lane);
            time_next<CPL>(mz, mzn, lane);
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                const int col = c0 + c;
                y0[c] = tau * kz[c] + mz[c] + ((col < N - 1) ? (tau * mv[c] - mzn[c]) : 0.0);
                y1[c] = tau * kv[c] + mv[c] + ((col >= 1) ? (-mvp[c] - tb * mz[c]) : 0.0);
            }
        }
        const size_t ro = (size_t)r * ld + c0;
        const bool bc = a.bcmask[r] != 0;
#pragma unroll
        for (int q = 0; q < CPL / 2; ++q) {
            double2 o0 = make_double2(c0 + 2 * q < N ? y0[2 * q] : 0.0, c0 + 2 * q + 1 < N ? y0[2 * q + 1] : 0.0);
            double2 o1 = make_double2(c0 + 2 * q < N ? y1[2 * q] : 0.0, c0 + 2 * q + 1 < N ? y1[2 * q + 1] : 0.0);
            if (bc) {
                o0 = ldg2(a.xv + ro + 2 * q);
                o1 = ldg2(a.xz + ro + 2 * q);
            }
            __stcs(reinterpret_cast<double2 *>(a.y0 + ro + 2 * q), o0);
            __stcs(reinterpret_cast<double2 *>(a.y1 + ro + 2 * q), o1);
        }
    }
}

template <bool CN, bool PER_LEVEL, bool SYM, bool HALO>
void launch_wide_h(const KktArgs &a, int rows_per_cta, int cap, cudaStream_t s)
{
    const int blocks = ceil_div(a.n_rows, rows_per_cta);
    const size_t smem = (size_t)(cap + 1) * (16 + ((SYM || PER_LEVEL) ? 0 : 8) + 4) + (size_t)(rows_per_cta + 1) * 4;
    if (a.ld == 128) kkt_apply_wide_kernel<CN, PER_LEVEL, SYM, HALO, 4><<<blocks, 256, smem, s>>>(a, rows_per_cta, cap);
    else kkt_apply_wide_kernel<CN, PER_LEVEL, SYM, HALO, 8><<<blocks, 256, smem, s>>>(a, rows_per_cta, cap);
}

template <bool CN>
void launch_wide(const KktArgs &a, bool per_level, bool sym, bool halo, int rows_per_cta, int cap, cudaStream_t s)
{
    if (per_level) {
        if (halo) launch_wide_h<CN, true, false, true>(a, rows_per_cta, cap, s);
        else launch_wide_h<CN, true, false, false>(a, rows_per_cta, cap, s);
    } else if (sym) {
        if (halo) launch_wide_h<CN, false, true, true>(a, rows_per_cta, cap, s);
        else launch_wide_h<CN, false, true, false>(a, rows_per_cta, cap, s);
    } else {
        if (halo) launch_wide_h<CN, false, false, true>(a, rows_per_cta, cap, s);
        else launch_wide_h<CN, false, false, false>(a, rows_per_cta, cap, s);
    }
}

// ---------------------------------------------------------------------------------------
// Grouped-rows variant (opt-in, CTL_KKT_GROUP=2|4; ld = 64, one rank, time-independent symmetric K).
// The staged kernel above is bound by the L1 data pipe: every row gathers its own 7 X row segments
// (2 panels x 512 B each) although consecutive rows of a mesh matrix share most of their columns.
// Here a warp owns R consecutive rows and gathers the UNION of their columns once (host-built plan,
// api.cu: 10 columns instead of 14 for R = 2 on the P1 triangle stencil, 16 instead of 28 for R = 4);
// every union entry carries R (m, k) pairs -- zero where a row lacks the column -- so the gathered
// segment feeds R rows from registers.  On-chip gather traffic drops by 29 % / 43 %, matrix bytes
// grow (36 B per union entry against 20 B per CSR entry: +2 % / +6 % of the HBM traffic of an apply),
// DFMA count grows by the zero entries.  Same time stencil / T_1, T_2 / Dirichlet epilogue per row.
// ---------------------------------------------------------------------------------------
template <bool CN>
__device__ __forceinline__ void kkt_row_epilogue(const KktArgs &a, const int r, const int lane, const double mv0,
                                                 const double mv1, const double kv0, const double kv1, const double mz0,
                                                 const double mz1, const double kz0, const double kz1)
{
    constexpr int G = 32;
    const unsigned full = 0xffffffffu;
    const int c0 = 2 * lane;
    const bool first = (lane == 0), last = (lane == G - 1);
    const int N = a.N;
    const bool in0 = c0 < N, in1 = c0 + 1 < N;
    const double tau = a.tau, beta = a.beta;
    double y00, y01, y10, y11;
    if (CN) {
        const double h = 0.5 * tau, hb = h / beta;
        double t;
        t = __shfl_up_sync(full, mv1, 1, G);   const double mvp0 = first ? 0.0 : t;
        t = __shfl_up_sync(full, kv1, 1, G);   const double kvp0 = first ? 0.0 : t;
        t = __shfl_down_sync(full, kz0, 1, G); const double kzn1 = last ? 0.0 : t;
        t = __shfl_down_sync(full, mz0, 1, G); const double mzn1 = last ? 0.0 : t;
        double r00 = h * (mvp0 + mv0) + h * (kz0 + kz1) + mz0 - mz1;
        double r01 = h * (mv0 + mv1) + h * (kz1 + kzn1) + mz1 - mzn1;
        double r10 = h * (kvp0 + kv0) - mvp0 + mv0 - hb * (mz0 + mz1);
        double r11 = h * (kv0 + kv1) - mv0 + mv1 - hb * (mz1 + mzn1);
        if (!in0) { r00 = 0.0; r10 = 0.0; }
        if (!in1) { r01 = 0.0; r11 = 0.0; }
        t = __shfl_down_sync(full, r00, 1, G); const double r0n = last ? 0.0 : t;
        t = __shfl_up_sync(full, r11, 1, G);   const double r1p = first ? 0.0 : t;
        y00 = r00 + r01;
        y01 = r01 + r0n;
        y10 = r10 + r1p;
        y11 = r11 + r10;
    } else {
        const double tb = tau / beta;
        double t;
        t = __shfl_up_sync(full, mv1, 1, G);   const double mvp0 = first ? 0.0 : t;
        t = __shfl_down_sync(full, mz0, 1, G); const double mzn1 = last ? 0.0 : t;
        y00 = tau * kz0 + mz0 + ((c0 < N - 1) ? (tau * mv0 - mz1) : 0.0);
        y01 = tau * kz1 + mz1 + ((c0 + 1 < N - 1) ? (tau * mv1 - mzn1) : 0.0);
        y10 = tau * kv0 + mv0 + ((c0 >= 1) ? (-mvp0 - tb * mz0) : 0.0);
        y11 = tau * kv1 + mv1 + (-mv0 - tb * mz1);
    }
    if (!in0) { y00 = 0.0; y10 = 0.0; }
    if (!in1) { y01 = 0.0; y11 = 0.0; }
    const size_t ro = (size_t)r * a.ld + c0;
    if (a.bcmask[r]) {
        const double2 xv = ldg2(a.xv + ro);
        const double2 xz = ldg2(a.xz + ro);
        y00 = xv.x; y01 = xv.y; y10 = xz.x; y11 = xz.y;
    }
    __stcs(reinterpret_cast<double2 *>(a.y0 + ro), make_double2(y00, y01));
    __stcs(reinterpret_cast<double2 *>(a.y1 + ro), make_double2(y10, y11));
}

template <bool CN, int R, int SC>
__global__ void __launch_bounds__(256) kkt_apply_group_kernel(const KktArgs a, const int *__restrict__ gptr,
                                                             const int *__restrict__ gcols,
                                                             const double2 *__restrict__ gvals,
                                                             const int groups_per_cta, const int cap)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: [(cap+1) * R] double2 (m, k) | [cap+1] unsigned byte offset of the gathered X row | [groups+1] int
    double2 *s_val = reinterpret_cast<double2 *>(smem_raw);
    unsigned *s_off = reinterpret_cast<unsigned *>(s_val + (size_t)(cap + 1) * R);
    int *s_ptr = reinterpret_cast<int *>(s_off + (cap + 1));

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    const int n_groups = (a.n_rows + R - 1) / R;
    const int g0 = blockIdx.x * groups_per_cta;
    const int ng = min(groups_per_cta, n_groups - g0);
    const char *__restrict__ xv_b = reinterpret_cast<const char *>(a.xv);
    const char *__restrict__ xz_b = reinterpret_cast<const char *>(a.xz);
    const unsigned lane_b = (unsigned)lane * 16u;
    const unsigned row_b = (unsigned)a.ld * 8u;

    for (int i = threadIdx.x; i <= ng; i += blockDim.x) s_ptr[i] = __ldg(gptr + g0 + i);
    __syncthreads();
    const int kb = s_ptr[0];
    const int cnt = s_ptr[ng] - kb;
    for (int k = threadIdx.x; k < cnt; k += blockDim.x) s_off[k] = (unsigned)__ldg(gcols + kb + k) * row_b;
    for (int k = threadIdx.x; k < cnt * R; k += blockDim.x) s_val[k] = __ldg(gvals + (size_t)kb * R + k);
    if (threadIdx.x == 0) {          // sentinel entry: zero values, a valid row to gather
        s_off[cap] = 0u;
#pragma unroll
        for (int q = 0; q < R; ++q) s_val[(size_t)cap * R + q] = make_double2(0.0, 0.0);
    }
    __syncthreads();

    for (int g = wid; g < ng; g += nwarps) {
        const int kbeg = s_ptr[g] - kb, kend = s_ptr[g + 1] - kb;
        double mv0[R], mv1[R], kv0[R], kv1[R], mz0[R], mz1[R], kz0[R], kz1[R];
#pragma unroll
        for (int q = 0; q < R; ++q) mv0[q] = mv1[q] = kv0[q] = kv1[q] = mz0[q] = mz1[q] = kz0[q] = kz1[q] = 0.0;
        for (int k0 = kbeg; k0 < kend; k0 += SC) {
            unsigned off[SC];
            int kk[SC];
#pragma unroll
            for (int j = 0; j < SC; ++j) {
                kk[j] = (k0 + j < kend) ? k0 + j : cap;
                off[j] = s_off[kk[j]];
            }
            double2 xv[SC], xz[SC];
#pragma unroll
            for (int j = 0; j < SC; ++j) {
                const unsigned o = off[j] + lane_b;
                xv[j] = __ldg(reinterpret_cast<const double2 *>(xv_b + o));
                xz[j] = __ldg(reinterpret_cast<const double2 *>(xz_b + o));
            }
#pragma unroll
            for (int j = 0; j < SC; ++j) {
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    const double2 mk = s_val[(size_t)kk[j] * R + q];
                    mv0[q] = fma(mk.x, xv[j].x, mv0[q]);
                    mv1[q] = fma(mk.x, xv[j].y, mv1[q]);
                    mz0[q] = fma(mk.x, xz[j].x, mz0[q]);
                    mz1[q] = fma(mk.x, xz[j].y, mz1[q]);
                    kv0[q] = fma(mk.y, xv[j].x, kv0[q]);
                    kv1[q] = fma(mk.y, xv[j].y, kv1[q]);
                    kz0[q] = fma(mk.y, xz[j].x, kz0[q]);
                    kz1[q] = fma(mk.y, xz[j].y, kz1[q]);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const int r = (g0 + g) * R + q;
            if (r < a.n_rows)        // warp-uniform
                kkt_row_epilogue<CN>(a, r, lane, mv0[q], mv1[q], kv0[q], kv1[q], mz0[q], mz1[q], kz0[q], kz1[q]);
        }
    }
}

template <bool CN, int R>
void launch_group_r(const KktArgs &a, const ctl_handle_s *h, cudaStream_t s)
{
    const int gpc = 64 / R;                                  // 64 rows per CTA
    const int cap = gpc * h->group_umax;
    const int n_groups = ceil_div(a.n_rows, R);
    const int blocks = ceil_div(n_groups, gpc);
    const size_t smem = (size_t)(cap + 1) * (16 * R + 4) + (size_t)(gpc + 1) * 4;
    const double2 *gv = reinterpret_cast<const double2 *>(h->d_gvals);
    // gather chunk with the least padding for the longest union list (ties: the larger chunk)
    const int u = h->group_umax;
    const int p4 = ceil_div(u, 4) * 4, p5 = ceil_div(u, 5) * 5, p8 = ceil_div(u, 8) * 8;
    if (p5 < p4 && p5 <= p8) kkt_apply_group_kernel<CN, R, 5><<<blocks, 256, smem, s>>>(a, h->d_gptr, h->d_gcols, gv, gpc, cap);
    else if (p8 <= p4 && R == 2) kkt_apply_group_kernel<CN, R, 8><<<blocks, 256, smem, s>>>(a, h->d_gptr, h->d_gcols, gv, gpc, cap);
    else kkt_apply_group_kernel<CN, R, 4><<<blocks, 256, smem, s>>>(a, h->d_gptr, h->d_gcols, gv, gpc, cap);
}

// ---------------------------------------------------------------------------------------
// Warp-specialised form of the record-fed TMA pipeline (opt-in, CTL_KKT_TMA=4; ld = 64; compiled at the end of
// round 1, NOT yet run).  Warp 0 only produces: for every row block of this CTA it waits until the stage is
// free (`empty` mbarrier, one arrival per consumer warp), then issues the bulk copies of the X tile and of the
// block's CSR record (`full` mbarrier, transaction count).  TR / 4 consumer warps wait on `full`, consume the
// stage from shared memory and release it.  No __syncthreads() inside the loop: the producer's dependent
// global loads (run table, record offsets) and the copies of block i+1 overlap the consumption of block i.
// ---------------------------------------------------------------------------------------
template <bool CN, bool SYM, bool HALO, int TR>
__global__ void __launch_bounds__(TR * 8 + 32) kkt_apply_tma_ws_kernel(const KktArgs a, const int umax, const int n_blocks)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int ld = a.ld;
    const unsigned row_b = (unsigned)ld * 8u;
    const size_t tile_b = (size_t)umax * row_b;
    const size_t stage_b = 2 * tile_b + (size_t)a.rec_max;         // [tile_v | tile_z | record]
    unsigned char *tiles = smem_raw;
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(tiles + 2 * stage_b);   // full[2], empty[2]
    constexpr unsigned HDR = ((TR + 1) * 4 + 15) & ~15;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int n_cons = (blockDim.x >> 5) - 1;

    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar + 0)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar + 1)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar + 2)), "r"(n_cons));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar + 3)), "r"(n_cons));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto wait_parity = [](const unsigned mb, const unsigned parity) {
        unsigned done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done)
                         : "r"(mb), "r"(parity)
                         : "memory");
        }
    };

    if (wid == 0) {
        // ------------------------------------------------------------------ producer
        int it = 0;
        for (int blk = blockIdx.x; blk < n_blocks; blk += gridDim.x, ++it) {
            const int st = it & 1;
            const unsigned full_mb = smem_u32(mbar + st), empty_mb = smem_u32(mbar + 2 + st);
            if (it >= 2) wait_parity(empty_mb, (unsigned)((it >> 1) + 1) & 1u);     // use (it/2 - 1) of the stage released
            unsigned char *tv = tiles + (size_t)st * stage_b;
            unsigned char *tz = tv + tile_b;
            const int ub = __ldg(a.tile_uptr + blk);
            const int n_runs = __ldg(a.tile_uptr + blk + 1) - ub;
            const int U = __ldg(a.tile_ucols + a.tile_count_off + blk);
            const int o0 = __ldg(a.rec_off + blk), o1 = __ldg(a.rec_off + blk + 1);
            const unsigned rec_bytes = (unsigned)(o1 - o0) * 16u;
            if (lane == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full_mb),
                             "r"(2u * (unsigned)U * row_b + rec_bytes)
                             : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32(tz + tile_b)),
                             "l"(a.rec + (size_t)o0 * 16), "r"(rec_bytes), "r"(full_mb)
                             : "memory");
            }
            __syncwarp();
            for (int u = lane; u < n_runs; u += 32) {
                const int c = __ldg(a.tile_ucols + 3 * (ub + u));
                const unsigned bytes = (unsigned)__ldg(a.tile_ucols + 3 * (ub + u) + 1) * row_b;
                const size_t dst = (size_t)__ldg(a.tile_ucols + 3 * (ub + u) + 2) * row_b;
                const double *sv, *sz;
                if (HALO && c >= a.n_own_cols) {
                    sv = a.hv + (size_t)(c - a.n_own_cols) * ld;
                    sz = a.hz + (size_t)(c - a.n_own_cols) * ld;
                } else {
                    sv = a.xv + (size_t)c * ld;
                    sz = a.xz + (size_t)c * ld;
                }
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32(tv + dst)),
                             "l"(sv), "r"(bytes), "r"(full_mb)
                             : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32(tz + dst)),
                             "l"(sz), "r"(bytes), "r"(full_mb)
                             : "memory");
            }
            __syncwarp();
        }
        return;
    }

    // ---------------------------------------------------------------------- consumers
    const unsigned lane_b = (unsigned)lane * 16u;
    const int cw = wid - 1;
    int it = 0;
    for (int blk = blockIdx.x; blk < n_blocks; blk += gridDim.x, ++it) {
        const int st = it & 1;
        wait_parity(smem_u32(mbar + st), (unsigned)(it >> 1) & 1u);
        const int r0 = blk * TR;
        const int nrows = min(TR, a.n_rows - r0);
        const unsigned char *tile_v = tiles + (size_t)st * stage_b;
        const unsigned char *tile_z = tile_v + tile_b;
        const unsigned char *rec = tile_z + tile_b;
        const int *b_ptr = reinterpret_cast<const int *>(rec);
        const int cnt = b_ptr[TR];                                    // entry cnt of a record is the zero sentinel
        const double2 *b_mk = reinterpret_cast<const double2 *>(rec + HDR);
        const double *b_kt = reinterpret_cast<const double *>(rec + HDR + (size_t)(cnt + 1) * 16);
        const unsigned *b_off = reinterpret_cast<const unsigned *>(
            rec + HDR + (size_t)(cnt + 1) * 16 + (SYM ? 0 : (((size_t)(cnt + 1) * 8 + 15) & ~(size_t)15)));
        for (int lr = cw; lr < nrows; lr += n_cons) {
            const int kbeg = b_ptr[lr], kend = b_ptr[lr + 1];
            double mv0 = 0, mv1 = 0, kv0 = 0, kv1 = 0, mz0 = 0, mz1 = 0, kz0 = 0, kz1 = 0;
            for (int k0 = kbeg; k0 < kend; k0 += SCHUNK) {
#pragma unroll
                for (int j = 0; j < SCHUNK; ++j) {
                    const int kk = (k0 + j < kend) ? k0 + j : cnt;
                    const unsigned o = b_off[kk] + lane_b;
                    const double2 xv = *reinterpret_cast<const double2 *>(tile_v + o);
                    const double2 xz = *reinterpret_cast<const double2 *>(tile_z + o);
                    const double2 mk = b_mk[kk];
                    const double kt = SYM ? mk.y : b_kt[kk];
                    mv0 = fma(mk.x, xv.x, mv0);
                    mv1 = fma(mk.x, xv.y, mv1);
                    mz0 = fma(mk.x, xz.x, mz0);
                    mz1 = fma(mk.x, xz.y, mz1);
                    kv0 = fma(mk.y, xv.x, kv0);
                    kv1 = fma(mk.y, xv.y, kv1);
                    kz0 = fma(kt, xz.x, kz0);
                    kz1 = fma(kt, xz.y, kz1);
                }
            }
            kkt_row_epilogue<CN>(a, r0 + lr, lane, mv0, mv1, kv0, kv1, mz0, mz1, kz0, kz1);
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(mbar + 2 + st)) : "memory");
    }
}

template <bool CN, bool SYM, bool HALO, int TR>
cudaError_t launch_tma_ws_t(const KktArgs &a, int umax, cudaStream_t s)
{
    const int n_blocks = ceil_div(a.n_rows, TR);
    const size_t smem = (size_t)4 * umax * a.ld * 8 + (size_t)2 * a.rec_max + 64;
    auto kern = kkt_apply_tma_ws_kernel<CN, SYM, HALO, TR>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int dev = 0, n_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (n_sm <= 0) n_sm = 148;
    const int per_sm = std::max(1, (int)((227 * 1024) / (smem + 1024)));
    kern<<<std::min(n_blocks, n_sm * per_sm), TR * 8 + 32, smem, s>>>(a, umax, n_blocks);
    return cudaSuccess;
}

template <bool CN, bool SYM, bool HALO>
cudaError_t launch_tma_ws(const KktArgs &a, int umax, int tile_rows, cudaStream_t s)
{
    if (tile_rows == 16) return launch_tma_ws_t<CN, SYM, HALO, 16>(a, umax, s);
    return launch_tma_ws_t<CN, SYM, HALO, 32>(a, umax, s);
}

template <bool CN, bool PER_LEVEL>
void launch_g(const KktArgs &a, int G, cudaStream_t s)
{
    const int threads = 256;
    const int rows_per_block = (threads / 32) * (32 / G);
    const int blocks = ceil_div(a.n_rows, rows_per_block);
    switch (G) {
    case 4: kkt_apply_kernel<CN, PER_LEVEL, 4><<<blocks, threads, 0, s>>>(a); break;
    case 8: kkt_apply_kernel<CN, PER_LEVEL, 8><<<blocks, threads, 0, s>>>(a); break;
    case 16: kkt_apply_kernel<CN, PER_LEVEL, 16><<<blocks, threads, 0, s>>>(a); break;
    default: kkt_apply_kernel<CN, PER_LEVEL, 32><<<blocks, threads, 0, s>>>(a); break;
    }
}

}  // namespace

int ctl_kkt_apply_tf(ctl_handle_s *h, const double *x_tf, double *y_tf)
{
    CTL_CHECK(h->assembled, CTL_ERR_STATE, "ctl_kkt_apply: ctl_assemble has not been called");
    CTL_CHECK(h->cfg.world == 1 || h->comm, CTL_ERR_STATE, "ctl_kkt_apply: call ctl_comm_init first (world > 1)");
    if (h->n_halo > 0) CTL_TRY(ctl_halo_exchange(h, x_tf));
    KktArgs a;
    a.n_rows = h->n_loc;
    a.n_own_cols = h->n_loc;
    a.N = h->N;
    a.ld = h->ld;
    a.indptr = h->d_indptr;
    a.indices = h->d_indices;
    a.Mv = h->d_M;
    a.Kv = h->d_K;
    a.KTv = h->d_KT;
    a.bcmask = h->d_bcmask;
    const size_t panel = (size_t)h->n_loc * h->ld;
    a.xv = x_tf;
    a.xz = x_tf + panel;
    a.hv = h->d_halo;
    a.hz = h->d_halo ? h->d_halo + (size_t)h->n_halo * h->ld : nullptr;
    a.y0 = y_tf;
    a.y1 = y_tf + panel;
    a.tau = h->cfg.tau;
    a.beta = h->cfg.beta;
    const int G = h->ld / 2;
    // rows per CTA: as many consecutive rows as fit the shared-memory entry budget
    const int max_len = std::max(1, h->max_row_len);
    int rows_per_cta = std::min(64, (STAGE_CAP / max_len) / 8 * 8);
    // the staged kernel addresses gathered rows with 31-bit byte offsets
    const bool fits = (size_t)std::max(h->n_loc, h->n_halo) * h->ld * 8 < 0x7fffffffull;
    const bool staged = rows_per_cta >= 8 && !h->force_unstaged && fits;
    const bool halo = h->n_halo > 0;
    if (h->ld > 64) {
        CTL_CHECK(rows_per_cta >= 8, CTL_ERR_ARG, "ctl_kkt_apply: rows too long for the wide (N > 64) kernel");
        const bool sym = h->d_KT == h->d_K;
        if (h->cfg.CN) launch_wide<true>(a, h->per_level, sym, halo, rows_per_cta, rows_per_cta * max_len, h->stream);
        else launch_wide<false>(a, h->per_level, sym, halo, rows_per_cta, rows_per_cta * max_len, h->stream);
        h->launches++;
        CTL_CUDA(cudaGetLastError());
        return CTL_OK;
    }
    if (h->group_ready && !h->force_unstaged && !halo && !h->per_level && h->ld == 64 && h->d_KT == h->d_K && fits &&
        (size_t)(64 / h->group_R * h->group_umax + 1) * (16 * h->group_R + 4) <= 40 * 1024) {
        if (h->cfg.CN) {
            if (h->group_R == 4) launch_group_r<true, 4>(a, h, h->stream);
            else launch_group_r<true, 2>(a, h, h->stream);
        } else {
            if (h->group_R == 4) launch_group_r<false, 4>(a, h, h->stream);
            else launch_group_r<false, 2>(a, h, h->stream);
        }
        h->launches++;
        CTL_CUDA(cudaGetLastError());
        return CTL_OK;
    }
    a.tile_uptr = h->d_tile_uptr;
    a.tile_ucols = h->d_tile_ucols;
    a.tile_slot = h->d_tile_slot;
    a.tile_count_off = h->tile_count_off;
    // TMA-staged kernel: time-independent K, tile plan available, tile + entries fit in shared memory
    const int tcap = h->tile_rows * max_len;
    const size_t tma_smem = (size_t)2 * h->tile_umax * h->ld * 8 + (size_t)(tcap + 1) * 28 + 33 * 4 + 16;
    const bool use_tma = h->tile_rows > 0 && !h->per_level && !h->force_unstaged && !h->no_tma && tma_smem <= 113 * 1024;
    const bool use_rec = h->tma_rec && h->d_rec && h->rec_max > 0;
    a.rec = use_rec ? h->d_rec : nullptr;
    a.rec_off = h->d_rec_off;
    a.rec_max = h->rec_max;
    const size_t pipe_smem = use_rec ? (size_t)4 * h->tile_umax * h->ld * 8 + (size_t)2 * h->rec_max + 32
                                     : (size_t)4 * h->tile_umax * h->ld * 8 + (size_t)(tcap + 1) * 28 + 33 * 4 + 32;
    if (h->tma_ws && use_rec && h->tile_rows > 0 && !h->per_level && !h->force_unstaged && G == 32 &&
        pipe_smem + 32 <= 227 * 1024) {
        const bool sym = h->d_KT == h->d_K;
        cudaError_t e;
        if (h->cfg.CN) {
            if (sym) e = halo ? launch_tma_ws<true, true, true>(a, h->tile_umax, h->tile_rows, h->stream) : launch_tma_ws<true, true, false>(a, h->tile_umax, h->tile_rows, h->stream);
            else e = halo ? launch_tma_ws<true, false, true>(a, h->tile_umax, h->tile_rows, h->stream) : launch_tma_ws<true, false, false>(a, h->tile_umax, h->tile_rows, h->stream);
        } else {
            if (sym) e = halo ? launch_tma_ws<false, true, true>(a, h->tile_umax, h->tile_rows, h->stream) : launch_tma_ws<false, true, false>(a, h->tile_umax, h->tile_rows, h->stream);
            else e = halo ? launch_tma_ws<false, false, true>(a, h->tile_umax, h->tile_rows, h->stream) : launch_tma_ws<false, false, false>(a, h->tile_umax, h->tile_rows, h->stream);
        }
        CTL_CUDA(e);
    } else if (h->tma_pipe && h->tile_rows > 0 && !h->per_level && !h->force_unstaged && G == 32 && pipe_smem <= 227 * 1024) {
        const bool sym = h->d_KT == h->d_K;
        cudaError_t e;
        if (h->cfg.CN) {
            if (sym) e = halo ? launch_tma_pipe<true, true, true>(a, G, tcap, h->tile_umax, h->tile_rows, h->stream) : launch_tma_pipe<true, true, false>(a, G, tcap, h->tile_umax, h->tile_rows, h->stream);
            else e = halo ? launch_tma_pipe<true, false, true>(a, G, tcap, h->tile_umax, h->tile_rows, h->stream) : launch_tma_pipe<true, false, false>(a, G, tcap, h->tile_umax, h->tile_rows, h->stream);
        } else {
            if (sym) e = halo ? launch_tma_pipe<false, true, true>(a, G, tcap, h->tile_umax, h->tile_rows, h->stream) : launch_tma_pipe<false, true, false>(a, G, tcap, h->tile_umax, h->tile_rows, h->stream);
            else e = halo ? launch_tma_pipe<false, false, true>(a, G, tcap, h->tile_umax, h->tile_rows, h->stream) : launch_tma_pipe<false, false, false>(a, G, tcap, h->tile_umax, h->tile_rows, h->stream);
        }
        CTL_CUDA(e);
    } else if (use_tma) {
        const bool sym = h->d_KT == h->d_K;
        cudaError_t e;
        if (h->cfg.CN) {
            if (sym) e = halo ? launch_tma<true, true, true>(a, G, tcap, h->tile_umax, h->tile_rows, h->stream) : launch_tma<true, true, false>(a, G, tcap, h->tile_umax, h->tile_rows, h->stream);
            else e = halo ? launch_tma<true, false, true>(a, G, tcap, h->tile_umax, h->tile_rows, h->stream) : launch_tma<true, false, false>(a, G, tcap, h->tile_umax, h->tile_rows, h->stream);
        } else {
            if (sym) e = halo ? launch_tma<false, true, true>(a, G, tcap, h->tile_umax, h->tile_rows, h->stream) : launch_tma<false, true, false>(a, G, tcap, h->tile_umax, h->tile_rows, h->stream);
            else e = halo ? launch_tma<false, false, true>(a, G, tcap, h->tile_umax, h->tile_rows, h->stream) : launch_tma<false, false, false>(a, G, tcap, h->tile_umax, h->tile_rows, h->stream);
        }
        CTL_CUDA(e);
    } else if (staged) {
        const int cap = rows_per_cta * max_len;
        const bool sym = h->d_KT == h->d_K;
        if (h->cfg.CN) {
            if (h->per_level) launch_staged<true, true, false>(a, G, rows_per_cta, cap, h->gather_chunk, halo, h->stream);
            else if (sym) launch_staged<true, false, true>(a, G, rows_per_cta, cap, h->gather_chunk, halo, h->stream);
            else launch_staged<true, false, false>(a, G, rows_per_cta, cap, h->gather_chunk, halo, h->stream);
        } else {
            if (h->per_level) launch_staged<false, true, false>(a, G, rows_per_cta, cap, h->gather_chunk, halo, h->stream);
            else if (sym) launch_staged<false, false, true>(a, G, rows_per_cta, cap, h->gather_chunk, halo, h->stream);
            else launch_staged<false, false, false>(a, G, rows_per_cta, cap, h->gather_chunk, halo, h->stream);
        }
    } else if (h->cfg.CN) {
        if (h->per_level) launch_g<true, true>(a, G, h->stream);
        else launch_g<true, false>(a, G, h->stream);
    } else {
        if (h->per_level) launch_g<false, true>(a, G, h->stream);
        else launch_g<false, false>(a, G, h->stream);
    }
    h->launches++;
    CTL_CUDA(cudaGetLastError());
    return CTL_OK;
}
