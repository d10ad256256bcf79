// Conversion between the reference's block-major mixed-vector layout
// (preconditioner/preconditioner.py:276-287: 2N contiguous n-blocks) and the library's
// time-fastest panels (two [n x ld] row-major panels).  One tiled transpose through
// shared memory per direction; only used on entry / exit of a library call, never inside
// the Krylov loop (SURVEY.md section 8b "Vector layout at the boundary").
#include "common.cuh"

namespace {

constexpr int TILE = 32;

// src: [N][n] (row j = time block j), dst: [n][ld]; columns >= N are zero-filled.
__global__ void bm_to_tf_kernel(const double *__restrict__ src, double *__restrict__ dst,
                                int n, int N, int ld)
{
    __shared__ double tile[TILE][TILE + 1];
    const size_t panel_src = (size_t)N * n, panel_dst = (size_t)n * ld;
    src += blockIdx.z * panel_src;
    dst += blockIdx.z * panel_dst;
    const int r0 = blockIdx.x * TILE, j0 = blockIdx.y * TILE;
    for (int jj = threadIdx.y; jj < TILE; jj += blockDim.y) {
        const int j = j0 + jj, r = r0 + threadIdx.x;
        tile[jj][threadIdx.x] = (j < N && r < n) ? src[(size_t)j * n + r] : 0.0;
    }
    __syncthreads();
    for (int rr = threadIdx.y; rr < TILE; rr += blockDim.y) {
        const int r = r0 + rr, j = j0 + threadIdx.x;
        if (r < n && j < ld) dst[(size_t)r * ld + j] = tile[threadIdx.x][rr];
    }
}

__global__ void tf_to_bm_kernel(const double *__restrict__ src, double *__restrict__ dst,
                                int n, int N, int ld)
{
    __shared__ double tile[TILE][TILE + 1];
    const size_t panel_src = (size_t)n * ld, panel_dst = (size_t)N * n;
    src += blockIdx.z * panel_src;
    dst += blockIdx.z * panel_dst;
    const int r0 = blockIdx.x * TILE, j0 = blockIdx.y * TILE;
    for (int rr = threadIdx.y; rr < TILE; rr += blockDim.y) {
        const int r = r0 + rr, j = j0 + threadIdx.x;
        tile[rr][threadIdx.x] = (r < n && j < ld) ? src[(size_t)r * ld + j] : 0.0;
    }
    __syncthreads();
    for (int jj = threadIdx.y; jj < TILE; jj += blockDim.y) {
        const int j = j0 + jj, r = r0 + threadIdx.x;
        if (j < N && r < n) dst[(size_t)j * n + r] = tile[threadIdx.x][jj];
    }
}

}  // namespace

int ctl_to_tf(ctl_handle_s *h, const double *src_bm, double *dst_tf)
{
    dim3 block(TILE, 8);
    dim3 grid(ceil_div(h->n_loc, TILE), ceil_div(h->ld, TILE), 2);
    bm_to_tf_kernel<<<grid, block, 0, h->stream>>>(src_bm, dst_tf, h->n_loc, h->N, h->ld);
    h->launches++;
    CTL_CUDA(cudaGetLastError());
    return CTL_OK;
}

int ctl_to_bm(ctl_handle_s *h, const double *src_tf, double *dst_bm)
{
    dim3 block(TILE, 8);
    dim3 grid(ceil_div(h->n_loc, TILE), ceil_div(h->ld, TILE), 2);
    tf_to_bm_kernel<<<grid, block, 0, h->stream>>>(src_tf, dst_bm, h->n_loc, h->N, h->ld);
    h->launches++;
    CTL_CUDA(cudaGetLastError());
    return CTL_OK;
}
