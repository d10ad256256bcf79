// Host-only check of csrc/amg_setup.cpp (no device calls): the pivoted dense inverse of the coarsest level on
// random non-symmetric matrices, and one complete hierarchy setup.  Built and run by tests/test_host_native.py with
// -I control_b200/csrc; prints one line per check, exit code 1 on failure.
#include "amg_setup.cpp"
#include <random>
#include <cstdio>
int main() {
    std::mt19937 g(1); std::uniform_real_distribution<double> u(-1, 1);
    for (int n : {7, 64, 301}) {
        HostCSR A; A.n_rows = A.n_cols = n; A.indptr.assign(n + 1, 0);
        std::vector<std::vector<double>> D(n, std::vector<double>(n, 0.0));
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) if (i == j || u(g) > 0.7) D[i][j] = u(g) + (i == j ? 0.2 : 0.0);
        for (int i = 0; i < n; ++i) { for (int j = 0; j < n; ++j) if (D[i][j] != 0.0) { A.indices.push_back(j); A.values.push_back(D[i][j]); } A.indptr[i + 1] = (int)A.indices.size(); }
        std::vector<double> inv; dense_inverse(A, {}, inv);
        double err = 0;
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { double s = 0; for (int k = 0; k < n; ++k) s += inv[(size_t)i * n + k] * D[k][j]; err = std::max(err, std::fabs(s - (i == j))); }
        printf("n=%d |inv*A - I|_max = %.3e\n", n, err);
        if (!(err < 1e-11)) return 1;
    }
    // full setup on a 1-D Laplacian-like matrix: levels and Galerkin sizes
    int n = 5000; HostCSR A; A.n_rows = A.n_cols = n; A.indptr.assign(n + 1, 0);
    for (int i = 0; i < n; ++i) { if (i) { A.indices.push_back(i - 1); A.values.push_back(-1); } A.indices.push_back(i); A.values.push_back(2.001); if (i + 1 < n) { A.indices.push_back(i + 1); A.values.push_back(-1); } A.indptr[i + 1] = (int)A.indices.size(); }
    AmgParams p; p.coarse_max = 600; p.device_inverse = 0; std::vector<AmgLevelHost> L; amg_setup_host(A, p, L, 2);
    for (auto &l : L) printf("level n=%d nnz=%zu rho=%.6f P=%dx%d Ainv=%zu\n", l.A.n_rows, l.A.values.size(), l.rho, l.P.n_rows, l.P.n_cols, l.Ainv.size());
    // coarse inverse check on the last level
    auto &last = L.back(); int nc = last.A.n_rows; double err = 0;
    std::vector<double> Dn((size_t)nc * nc, 0.0);
    for (int i = 0; i < nc; ++i) for (int k = last.A.indptr[i]; k < last.A.indptr[i + 1]; ++k) Dn[(size_t)i * nc + last.A.indices[k]] += last.A.values[k];
    for (int i = 0; i < nc; ++i) for (int j = 0; j < nc; ++j) { double s = 0; for (int k = 0; k < nc; ++k) s += last.Ainv[(size_t)i * nc + k] * Dn[(size_t)k * nc + j]; err = std::max(err, std::fabs(s - (i == j))); }
    printf("coarse inverse residual %.3e (n=%d)\n", err, nc);
    if (!(err < 1e-11) || L.size() != 3 || L[0].P.n_cols != L[1].A.n_rows || last.Ainv.size() != (size_t)nc * nc) return 1;
    return 0;
}
