// Host-only check of csrc/sell_format.h: every exact format (F64, D16, PK, DICT16, DICT8) of the SELL-32 and
// CSR-vector layouts is built with the library's own builders and READ BACK with the kernels' own decode
// functions (sf_load / sf_decode are __host__ __device__); the products must equal the plain CSR products bit for
// bit (same accumulation order).  Also checks which format the automatic choice picks for a uniform-mesh stencil,
// a prolongation-like rectangular matrix, random values, and columns far from the diagonal (escape entries).
// Built and run by tests/test_host_native.py.
#include <cmath>
#include <cstdio>
#include <random>

#include "sell_format.h"

struct Csr {
    int n_rows, n_cols;
    std::vector<int> ip, ix;
    std::vector<double> v;
};

static std::vector<double> csr_mv(const Csr &A, const std::vector<double> &x)
{
    std::vector<double> y(A.n_rows, 0.0);
    for (int r = 0; r < A.n_rows; ++r) {
        double acc = 0.0;
        for (int k = A.ip[r]; k < A.ip[r + 1]; ++k) acc = fma(A.v[k], x[A.ix[k]], acc);
        y[r] = acc;
    }
    return y;
}

template <int FMT>
static double sell_row(const MatView &A, int row, const std::vector<double> &x)
{
    const int s = row >> 5, lane = row & 31;
    const int2 s0 = A.sp[s];
    const int end = A.sp[s + 1].x;
    double acc = 0.0;
    for (int p = s0.x + lane; p < end; p += 32) {
        const SfRaw raw = sf_load<FMT, false>(A, p);
        int c;
        double v;
        sf_decode<FMT>(A, raw, p, row, s0.y, c, v);
        acc = fma(v, x[c], acc);
    }
    return acc;
}

// FMT_STENCIL: the kernel's two paths (sell.cu::sell_row_dot_stencil)
static double sell_row_stencil(const MatView &A, int row, const std::vector<double> &x, long long *uniform_rows)
{
    const int s = row >> 5, lane = row & 31;
    const int4 s0 = A.sp4[s];
    double acc = 0.0;
    if (s0.z >= 0) {
        ++*uniform_rows;
        for (int k = 0; k < s0.y; ++k) {
            const DictEnt e = A.stab[s0.z + k];
            acc = fma(e.v, x[row + e.delta], acc);
        }
        return acc;
    }
    const uint16_t *code = reinterpret_cast<const uint16_t *>(A.code);
    for (int p = s0.x + lane; p < s0.x + 32 * s0.y; p += 32) {
        const DictEnt e = A.dict[code[p]];
        acc = fma(e.v, x[row + e.delta], acc);
    }
    return acc;
}

template <int FMT>
static double csrv_row(const MatView &A, int row, const std::vector<double> &x)
{
    double acc = 0.0;
    const int base = FMT == FMT_F64 ? 0 : A.rbase[row];
    for (int p = A.ptr[row]; p < A.ptr[row + 1]; ++p) {
        const SfRaw raw = sf_load<FMT, false>(A, p);
        int c;
        double v;
        sf_decode<FMT>(A, raw, p, row, base, c, v);
        acc = fma(v, x[c], acc);
    }
    return acc;
}

static int fails = 0;

static const char *fmt_name(int f)
{
    static const char *n[] = {"F64", "D16", "PK", "DICT16", "DICT8", "STENCIL"};
    return n[f];
}

static void check_sell(const char *what, const Csr &A, int expect_fmt)
{
    std::mt19937 g(11);
    std::uniform_real_distribution<double> u(-1, 1);
    std::vector<double> x(A.n_cols);
    for (double &v : x) v = u(g);
    const std::vector<double> ref = csr_mv(A, x);
    SfCsr a;
    a.n_rows = A.n_rows;
    a.n_cols = A.n_cols;
    a.indptr = A.ip.data();
    a.indices = A.ix.data();
    SfSellLayout L;
    sf_sell_layout(a, L);
    for (int cap = FMT_F64; cap <= FMT_STENCIL; ++cap) {
        SfSellValues V;
        sf_sell_values(L, A.v.data(), cap, V);
        MatView M;
        M.fmt = V.fmt;
        M.n_rows = A.n_rows;
        M.n_own = A.n_cols;
        M.sp = L.sp.data();
        M.cols = L.cols.data();
        M.dcol = L.dcol.empty() ? nullptr : L.dcol.data();
        M.vals = V.vals.data();
        M.vcode = V.vcode.data();
        M.vdict = V.vdict.data();
        M.code = V.fmt == FMT_DICT8 ? (const void *)V.code8.data() : (const void *)V.code16.data();
        M.dict = V.dict.data();
        M.sp4 = V.sp4.data();
        M.stab = V.stab.data();
        long long uniform_rows = 0;
        double worst = 0.0;
        for (int r = 0; r < A.n_rows; ++r) {
            double y;
            switch (V.fmt) {
            case FMT_F64: y = sell_row<FMT_F64>(M, r, x); break;
            case FMT_D16: y = sell_row<FMT_D16>(M, r, x); break;
            case FMT_PK: y = sell_row<FMT_PK>(M, r, x); break;
            case FMT_DICT16: y = sell_row<FMT_DICT16>(M, r, x); break;
            case FMT_STENCIL: y = sell_row_stencil(M, r, x, &uniform_rows); break;
            default: y = sell_row<FMT_DICT8>(M, r, x); break;
            }
            worst = std::max(worst, std::fabs(y - ref[r]));
        }
        if (worst != 0.0) {
            printf("FAIL %s cap %s got %s: max diff %.3e\n", what, fmt_name(cap), fmt_name(V.fmt), worst);
            ++fails;
        }
        if (cap == FMT_STENCIL) {
            printf("%-28s SELL: automatic format %-7s %5.2f bytes / nonzero (stored %lld, nnz %lld, rows in uniform slices %.0f %%)\n",
                   what, fmt_name(V.fmt), (double)V.bytes_per_pass / (double)L.nnz, (long long)L.n_stored, (long long)L.nnz,
                   100.0 * uniform_rows / A.n_rows);
            if (expect_fmt >= 0 && V.fmt != expect_fmt) {
                printf("FAIL %s: expected %s\n", what, fmt_name(expect_fmt));
                ++fails;
            }
        }
    }
    // CSR-vector layout
    for (int cap = FMT_F64; cap <= FMT_PK; ++cap) {
        SfCsrvData D;
        sf_csrv_data(a, A.v.data(), cap, D);
        MatView M;
        M.fmt = D.fmt;
        M.n_rows = A.n_rows;
        M.n_own = A.n_cols;
        M.ptr = A.ip.data();
        M.rbase = D.rbase.data();
        M.cols = A.ix.data();
        M.dcol = D.dcol.data();
        M.vals = A.v.data();
        M.vcode = D.vcode.data();
        M.vdict = D.vdict.data();
        double worst = 0.0;
        for (int r = 0; r < A.n_rows; ++r) {
            double y;
            switch (D.fmt) {
            case FMT_F64: y = csrv_row<FMT_F64>(M, r, x); break;
            case FMT_D16: y = csrv_row<FMT_D16>(M, r, x); break;
            default: y = csrv_row<FMT_PK>(M, r, x); break;
            }
            worst = std::max(worst, std::fabs(y - ref[r]));
        }
        if (worst != 0.0) {
            printf("FAIL %s CSR-vector cap %s got %s: max diff %.3e\n", what, fmt_name(cap), fmt_name(D.fmt), worst);
            ++fails;
        }
    }
}

static Csr mesh_stencil(int nx, int ny, bool random_values, int ghost_cols)
{
    // 7-point stencil; ghost_cols > 0: the last row of nodes references columns appended behind the owned ones
    // (the local numbering of a row-partitioned matrix)
    Csr A;
    const int n = nx * ny;
    A.n_rows = n;
    A.n_cols = n + ghost_cols;
    A.ip.assign(n + 1, 0);
    std::mt19937 g(2);
    std::uniform_real_distribution<double> u(-1, 1);
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) {
            const int di[7] = {0, -1, 1, 0, 0, 1, -1}, dj[7] = {0, 0, 0, -1, 1, -1, 1};
            const double val[7] = {4.0 + 1.0 / 3.0, -1.0, -1.0, -1.0, -1.0, 0.0, 0.0};      // structural zeros kept
            std::vector<std::pair<int, double>> row;
            for (int q = 0; q < 7; ++q) {
                const int a = i + di[q], b = j + dj[q];
                if (a < 0 || a >= nx || b < 0) continue;
                int col;
                if (b >= ny) {
                    if (ghost_cols == 0) continue;
                    col = n + a % ghost_cols;
                } else {
                    col = b * nx + a;
                }
                row.emplace_back(col, random_values ? u(g) : val[q]);
            }
            std::sort(row.begin(), row.end());
            for (auto &e : row) {
                A.ix.push_back(e.first);
                A.v.push_back(e.second);
            }
            A.ip[j * nx + i + 1] = (int)A.ix.size();
        }
    return A;
}

int main()
{
    check_sell("uniform 7-point stencil", mesh_stencil(300, 217, false, 0), FMT_STENCIL);
    check_sell("stencil + ghost columns", mesh_stencil(300, 230, false, 300), FMT_STENCIL);
    check_sell("random values (small)", mesh_stencil(97, 53, true, 0), FMT_D16);      // more pairs than a dictionary may hold
    {
        // a few hundred distinct (offset, value) pairs, rows all different: per-entry dictionary codes
        Csr A = mesh_stencil(120, 90, false, 0);
        for (int r = 0; r < A.n_rows; ++r)
            for (int k = A.ip[r]; k < A.ip[r + 1]; ++k) A.v[k] = 0.25 * (double)((r * 31 + 7 * (k - A.ip[r])) % 300) - 20.0;
        check_sell("300 values x 7 offsets", A, FMT_DICT16);
    }
    check_sell("random values", mesh_stencil(397, 253, true, 0), FMT_D16);
    {
        // prolongation-like: n x n/6, 3 entries per row, 40 distinct values
        Csr P;
        P.n_rows = 50000;
        P.n_cols = 50000 / 6 + 400;
        P.ip.assign(P.n_rows + 1, 0);
        for (int r = 0; r < P.n_rows; ++r) {
            for (int q = 0; q < 3; ++q) {
                P.ix.push_back(r / 6 + 130 * q);
                P.v.push_back(0.1 * ((r * 7 + q * 3) % 40) - 1.3);
            }
            P.ip[r + 1] = (int)P.ix.size();
        }
        check_sell("prolongation-like", P, FMT_PK);
    }
    {
        // a few columns far from the diagonal: escape entries in every 16-bit format
        Csr A = mesh_stencil(400, 300, true, 0);
        for (int r = 0; r < A.n_rows; r += 1000) A.ix[A.ip[r + 1] - 1] = A.n_cols - 1 - (r % 7);
        check_sell("escape entries", A, FMT_D16);
    }
    {
        // more than 2 % escapes: 16-bit offsets are not worth it
        Csr A = mesh_stencil(500, 300, true, 0);
        for (int r = 0; r < A.n_rows; r += 3) A.ix[A.ip[r + 1] - 1] = (A.ix[A.ip[r + 1] - 1] + 75000) % A.n_cols;
        check_sell("many far columns", A, FMT_F64);
    }
    printf("sell format check: %d failures\n", fails);
    return fails ? 1 : 0;
}
