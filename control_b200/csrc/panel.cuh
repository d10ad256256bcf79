// Device helpers of the time-fastest [n x ld] panels, shared by the batched preconditioner
// kernels (pc_batched.cu) and the Stokes couplings (stokes.cu): G lanes per row, every lane
// owns CPL consecutive time columns (ld = G * CPL: CPL = 2 up to 64 columns, 4 or 8 for
// 128 / 256 columns); the time-coupling transforms are shuffles and register scans.
#pragma once
#include "common.cuh"

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

// A lane owns CPL consecutive time columns c0 .. c0+CPL-1 (c0 = CPL * l, CPL even).
// T_1^-1: x_i <- sum_{k >= i} (-1)^(k-i) x_k   (control/control.py:63-78)
template <int G, int CPL>
__device__ __forceinline__ void t1_inv(double (&x)[CPL], int l)
{
    double z[CPL], tot = 0.0;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
        z[c] = (c & 1) ? -x[c] : x[c];             // (-1)^column, c0 is even
        tot += z[c];
    }
    double s = group_suffix<G>(tot, l);             // sum over columns >= c0
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
        x[c] = (c & 1) ? -s : s;
        s -= z[c];
    }
}

// T_2^-1: x_i <- sum_{k <= i} (-1)^(i-k) x_k   (control/control.py:81-96)
template <int G, int CPL>
__device__ __forceinline__ void t2_inv(double (&x)[CPL], int l)
{
    double z[CPL], tot = 0.0;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
        z[c] = (c & 1) ? -x[c] : x[c];
        tot += z[c];
    }
    double s = group_prefix<G>(tot, l) - tot;       // sum over columns < c0
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
        s += z[c];
        x[c] = (c & 1) ? -s : s;
    }
}

// value of the previous time column (0 before the first)
template <int G, int CPL>
__device__ __forceinline__ void time_prev(const double (&a)[CPL], double (&out)[CPL], int l)
{
    const double t = __shfl_up_sync(0xffffffffu, a[CPL - 1], 1, G);
    out[0] = l == 0 ? 0.0 : t;
#pragma unroll
    for (int c = 1; c < CPL; ++c) out[c] = a[c - 1];
}

struct RowMap {
    int row, r, l, c0;
    bool live;
};

template <int G, int CPL>
__device__ __forceinline__ RowMap row_map(int n_rows)
{
    RowMap m;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * (blockDim.x >> 5)) + (threadIdx.x >> 5);
    m.row = warp * (32 / G) + lane / G;
    m.live = m.row < n_rows;
    m.r = m.live ? m.row : n_rows - 1;
    m.l = lane % G;
    m.c0 = CPL * m.l;
    return m;
}

template <int CPL>
__device__ __forceinline__ void load_cols(const double *p, double (&x)[CPL])
{
#pragma unroll
    for (int q = 0; q < CPL / 2; ++q) {
        const double2 v = ldg2(p + 2 * q);
        x[2 * q] = v.x;
        x[2 * q + 1] = v.y;
    }
}

template <int CPL>
__device__ __forceinline__ void store_cols(double *p, const double (&x)[CPL])
{
#pragma unroll
    for (int q = 0; q < CPL / 2; ++q) *reinterpret_cast<double2 *>(p + 2 * q) = make_double2(x[2 * q], x[2 * q + 1]);
}

// T_1: x_i <- x_i + x_{i+1}; T_2: x_i <- x_i + x_{i-1}   (preconditioner/preconditioner.py:33-60)
template <int G, int CPL>
__device__ __forceinline__ void time_next(const double (&a)[CPL], double (&out)[CPL], int l)
{
    const double t = __shfl_down_sync(0xffffffffu, a[0], 1, G);
#pragma unroll
    for (int c = 0; c + 1 < CPL; ++c) out[c] = a[c + 1];
    out[CPL - 1] = l == G - 1 ? 0.0 : t;
}

inline int blocks_for(int n_rows, int ld) { const int G = ld >= 64 ? 32 : ld / 2; return ceil_div(n_rows, 8 * (32 / G)); }

}  // namespace

// (lanes per row, columns per lane) for a padded row length ld = G * CPL
#define DISPATCH_G(LD, CALL)                                                               \
    switch (LD) {                                                                          \
    case 8: { constexpr int GG = 4, CC = 2; CALL; } break;                                 \
    case 16: { constexpr int GG = 8, CC = 2; CALL; } break;                                \
    case 32: { constexpr int GG = 16, CC = 2; CALL; } break;                               \
    case 64: { constexpr int GG = 32, CC = 2; CALL; } break;                               \
    case 128: { constexpr int GG = 32, CC = 4; CALL; } break;                              \
    default: { constexpr int GG = 32, CC = 8; CALL; } break;                               \
    }
