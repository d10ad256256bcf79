// Discrete objective J_h on the device ("next" row f3):
//   J_h = 1/2 sum_i w_i (v_i - vhat_i)^T M (v_i - vhat_i) + 1/(2 beta) sum_i w_i zeta_i^T M zeta_i
// over the n_t time levels, trapezoid weights for CN, tau for BE (the row weights of block_00 /
// block_11, control/control.py:2915-2923, 2940-2953; the reference itself never evaluates J).
// One kernel over (rows x levels) with the FULL mass matrix (no Dirichlet elimination: vhat need not
// vanish on the boundary), fixed-order two-stage reduction, 2 n_t doubles read back.
#include "common.cuh"

namespace {

constexpr int QT = 256;

__global__ void __launch_bounds__(QT) quadform_partial_kernel(const int *__restrict__ ptr, const int *__restrict__ cols,
                                                             const double *__restrict__ Mv, const double *__restrict__ v,
                                                             const double *__restrict__ zeta,
                                                             const double *__restrict__ vhat, double *__restrict__ partial,
                                                             int n)
{
    __shared__ double sh[2][QT / 32];
    const size_t base = (size_t)blockIdx.y * n;
    const int row = blockIdx.x * QT + threadIdx.x;
    double qd = 0.0, qz = 0.0;
    if (row < n) {
        double ad = 0.0, az = 0.0;
        for (int k = ptr[row]; k < ptr[row + 1]; ++k) {
            const int c = cols[k];
            const double m = Mv[k];
            ad = fma(m, v[base + c] - vhat[base + c], ad);
            az = fma(m, zeta[base + c], az);
        }
        qd = (v[base + row] - vhat[base + row]) * ad;
        qz = zeta[base + row] * az;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        qd += __shfl_xor_sync(0xffffffffu, qd, o);
        qz += __shfl_xor_sync(0xffffffffu, qz, o);
    }
    if ((threadIdx.x & 31) == 0) {
        sh[0][threadIdx.x >> 5] = qd;
        sh[1][threadIdx.x >> 5] = qz;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < QT / 32; ++w) {
            a += sh[0][w];
            b += sh[1][w];
        }
        double *p = partial + 2 * ((size_t)blockIdx.y * gridDim.x + blockIdx.x);
        p[0] = a;
        p[1] = b;
    }
}

// out[level][0..1] = sums over the row blocks, in block order
__global__ void quadform_finish_kernel(const double *__restrict__ partial, double *__restrict__ out, int blocks)
{
    if (threadIdx.x >= 2) return;
    const double *p = partial + 2 * (size_t)blockIdx.x * blocks + threadIdx.x;
    double s = 0.0;
    for (int b = 0; b < blocks; ++b) s += p[2 * (size_t)b];
    out[2 * blockIdx.x + threadIdx.x] = s;
}

}  // namespace

extern "C" int ctl_objective(ctl_handle h, const double *v, const double *zeta, const double *v_hat, double *out)
{
    CTL_CHECK(h && v && zeta && v_hat && out, CTL_ERR_ARG, "ctl_objective: null argument");
    CTL_CHECK(h->assembled, CTL_ERR_STATE, "ctl_objective: ctl_assemble has not been called");
    CTL_CHECK(h->cfg.world == 1, CTL_ERR_ARG, "ctl_objective: single-rank only (use ctl_objective_host)");
    CTL_CUDA(cudaSetDevice(h->cfg.device));
    const int n = h->n, n_t = h->cfg.n_t;
    if (!h->d_M_full) {
        std::vector<double> buf(h->loc_entry.size());
        for (size_t p = 0; p < buf.size(); ++p) buf[p] = h->h_M[h->loc_entry[p]];
        CTL_TRY(ctl_upload(h, &h->d_M_full, buf.data(), buf.size()));
    }
    const int blocks = ceil_div(n, QT);
    double *partial = nullptr, *sums = nullptr;
    CTL_CUDA(cudaMalloc((void **)&partial, sizeof(double) * 2 * (size_t)blocks * n_t));
    CTL_CUDA(cudaMalloc((void **)&sums, sizeof(double) * 2 * n_t));
    quadform_partial_kernel<<<dim3(blocks, n_t), QT, 0, h->stream>>>(h->d_indptr, h->d_indices, h->d_M_full, v, zeta, v_hat,
                                                                    partial, n);
    quadform_finish_kernel<<<n_t, 32, 0, h->stream>>>(partial, sums, blocks);
    h->launches += 2;
    std::vector<double> hs(2 * (size_t)n_t);
    cudaError_t e = cudaMemcpyAsync(hs.data(), sums, sizeof(double) * hs.size(), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(partial);
    cudaFree(sums);
    if (e != cudaSuccess) {
        ctl_set_error(h, std::string("ctl_objective: ") + cudaGetErrorString(e));
        return CTL_ERR_CUDA;
    }
    const double tau = h->cfg.tau, beta = h->cfg.beta;
    double J = 0.0;
    for (int i = 0; i < n_t; ++i) {
        const double w = (h->cfg.CN && (i == 0 || i == n_t - 1)) ? 0.5 * tau : tau;
        J += 0.5 * w * hs[2 * i];
        J += 0.5 / beta * w * hs[2 * i + 1];
    }
    *out = J;
    return CTL_OK;
}
