// Host-only check of csrc/kkt_plan.h: builds the block records of a 2-D 7-point mesh pattern (symmetric and
// non-symmetric value sets, row count not a multiple of the block height), then READS them back the way
// kkt_apply_tma_pipe_kernel<REC> / kkt_apply_tma_ws_kernel do (header ptr[TR+1], (m,k) pairs at HDR, kt behind them,
// byte offsets behind those, entry cnt = zero sentinel) and checks that the products M x, K x, K^T x computed from the
// records and a tile filled by slot equal the plain CSR products.  Built and run by tests/test_host_native.py.
#include <cmath>
#include <cstdio>
#include <random>

#include "kkt_plan.h"

int main()
{
    const int nx = 13, ny = 9, n = nx * ny, ld = 64;
    const size_t row_b = (size_t)ld * 8;
    std::vector<int> ip(n + 1, 0), ix;
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) {
            const int r = j * nx + i;
            const int di[7] = {0, -1, 1, 0, 0, 1, -1}, dj[7] = {0, 0, 0, -1, 1, -1, 1};
            std::vector<int> cols;
            for (int q = 0; q < 7; ++q) {
                const int a = i + di[q], b = j + dj[q];
                if (a >= 0 && a < nx && b >= 0 && b < ny) cols.push_back(b * nx + a);
            }
            std::sort(cols.begin(), cols.end());
            ix.insert(ix.end(), cols.begin(), cols.end());
            ip[r + 1] = (int)ix.size();
        }
    const int nnz = (int)ix.size();
    std::mt19937 g(3);
    std::uniform_real_distribution<double> u(-1, 1);
    std::vector<double> m(nnz), k(nnz), kt(nnz), x(n);
    for (int e = 0; e < nnz; ++e) { m[e] = u(g); k[e] = u(g); kt[e] = u(g); }
    for (double &v : x) v = u(g);
    int fails = 0;
    for (int TR : {16, 32})
        for (int sym = 0; sym < 2; ++sym) {
            // slot plan as api.cu builds it: position of the column among the block's sorted unique columns
            const int nblk = (n + TR - 1) / TR;
            std::vector<uint8_t> slot(nnz);
            std::vector<std::vector<int>> uniq(nblk);
            for (int b = 0; b < nblk; ++b) {
                const int r0 = b * TR, r1 = std::min(n, r0 + TR);
                std::vector<int> t(ix.begin() + ip[r0], ix.begin() + ip[r1]);
                std::sort(t.begin(), t.end());
                t.erase(std::unique(t.begin(), t.end()), t.end());
                for (int e = ip[r0]; e < ip[r1]; ++e) slot[e] = (uint8_t)(std::lower_bound(t.begin(), t.end(), ix[e]) - t.begin());
                uniq[b] = t;
            }
            std::vector<uint8_t> rec;
            std::vector<int> roff;
            const int rec_max = kkt_build_block_records(TR, n, ip.data(), slot.data(), m.data(), k.data(), sym ? nullptr : kt.data(),
                                                        row_b, rec, roff);
            const unsigned HDR = ((TR + 1) * 4 + 15) & ~15u;
            double err = 0.0;
            for (int b = 0; b < nblk; ++b) {
                if (roff[b + 1] * 16 - roff[b] * 16 > rec_max || (roff[b] * 16) % 16) ++fails;
                const uint8_t *r = rec.data() + (size_t)roff[b] * 16;
                const int *ptr = reinterpret_cast<const int *>(r);
                const int cnt = ptr[TR];
                const double *mk = reinterpret_cast<const double *>(r + HDR);
                const double *ktp = reinterpret_cast<const double *>(r + HDR + (size_t)(cnt + 1) * 16);
                const unsigned *off = reinterpret_cast<const unsigned *>(
                    r + HDR + (size_t)(cnt + 1) * 16 + (sym ? 0 : (((size_t)(cnt + 1) * 8 + 15) & ~(size_t)15)));
                if (mk[2 * cnt] != 0.0 || mk[2 * cnt + 1] != 0.0 || off[cnt] != 0u) ++fails;        // sentinel
                std::vector<double> tile(uniq[b].size());
                for (size_t s = 0; s < uniq[b].size(); ++s) tile[s] = x[uniq[b][s]];             // what the bulk copies deliver
                const int r0 = b * TR, nrows = std::min(TR, n - r0);
                for (int lr = 0; lr < nrows; ++lr) {
                    double ym = 0, yk = 0, ykt = 0, zm = 0, zk = 0, zkt = 0;
                    for (int e = ptr[lr]; e < ptr[lr + 1]; ++e) {
                        const double xv = tile[off[e] / row_b];
                        ym += mk[2 * e] * xv;
                        yk += mk[2 * e + 1] * xv;
                        ykt += (sym ? mk[2 * e + 1] : ktp[e]) * xv;
                    }
                    for (int e = ip[r0 + lr]; e < ip[r0 + lr + 1]; ++e) {
                        zm += m[e] * x[ix[e]];
                        zk += k[e] * x[ix[e]];
                        zkt += (sym ? k[e] : kt[e]) * x[ix[e]];
                    }
                    err = std::max(err, std::max(std::fabs(ym - zm), std::max(std::fabs(yk - zk), std::fabs(ykt - zkt))));
                }
                for (int i = nrows; i <= TR; ++i)
                    if (ptr[i] != cnt) ++fails;                                                   // rows past the end
            }
            printf("TR=%d sym=%d blocks=%d rec_max=%d max diff %.3e fails %d\n", TR, sym, nblk, rec_max, err, fails);
            if (err != 0.0) ++fails;
        }
    return fails ? 1 : 0;
}
