// Host-only check of csrc/halo_geom.h: plays ALL ranks of a row partition in one process.
//  * builds a smoothed-aggregation hierarchy of a 2-D mesh operator with the library's own host set-up,
//  * distributes it for world = 2, 3, 4 with amg_distribute_host (every rank computes its own picture; the pictures
//    must agree),
//  * runs one V-cycle the way amg.cu schedules it - local row blocks in local numbering, owner-computes restriction,
//    ghost values delivered only through the send chunks / positions of the exchange geometry, the first replicated
//    level numbered own-rows-first - and compares with the serial cycle on the global matrices.
// Built and run by tests/test_host_native.py.
#include <cmath>
#include <cstdio>
#include <random>

#include "amg_setup.cpp"
#include "cheb_coefficients.h"
#include "halo_geom.h"

static std::vector<double> spmv(const HostCSR &A, const std::vector<double> &x)
{
    std::vector<double> y(A.n_rows, 0.0);
    for (int r = 0; r < A.n_rows; ++r)
        for (int k = A.indptr[r]; k < A.indptr[r + 1]; ++k) y[r] += A.values[k] * x[A.indices[k]];
    return y;
}

// ---- serial reference: oracle/amg.py::vcycle on the global matrices
static std::vector<double> cheb_serial(const AmgLevelHost &L, const AmgParams &p, const std::vector<double> &b,
                                       const std::vector<double> *x0)
{
    double scale;
    std::vector<double> om;
    cheb_coefficients(p.lo * L.rho, p.hi * L.rho, p.nu, &scale, om);
    const int n = L.A.n_rows;
    std::vector<double> prev(n, 0.0), cur(n), next(n);
    if (x0) {
        prev = *x0;
        const std::vector<double> ax = spmv(L.A, prev);
        for (int i = 0; i < n; ++i) cur[i] = prev[i] + scale * L.dinv[i] * (b[i] - ax[i]);
    } else {
        for (int i = 0; i < n; ++i) cur[i] = scale * L.dinv[i] * b[i];
    }
    for (int k = 2; k <= p.nu; ++k) {
        const double w = om[k - 2];
        const std::vector<double> ax = spmv(L.A, cur);
        for (int i = 0; i < n; ++i) next[i] = (1.0 - w) * prev[i] + w * cur[i] + w * scale * L.dinv[i] * (b[i] - ax[i]);
        prev = cur;
        cur = next;
    }
    return cur;
}

static std::vector<double> vcycle_serial(const std::vector<AmgLevelHost> &H, const AmgParams &p, int l,
                                         const std::vector<double> &b, const std::vector<double> *x0)
{
    const AmgLevelHost &L = H[l];
    const int n = L.A.n_rows;
    if (l == (int)H.size() - 1) {
        if (!L.Ainv.empty()) {
            std::vector<double> x(n, 0.0);
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < n; ++j) x[i] += L.Ainv[(size_t)i * n + j] * b[j];
            return x;
        }
        return cheb_serial(L, p, b, x0);
    }
    std::vector<double> x = cheb_serial(L, p, b, x0);
    const std::vector<double> ax = spmv(L.A, x);
    std::vector<double> r(n);
    for (int i = 0; i < n; ++i) r[i] = b[i] - ax[i];
    const std::vector<double> xc = vcycle_serial(H, p, l + 1, spmv(L.R, r), nullptr);
    const std::vector<double> pxc = spmv(L.P, xc);
    for (int i = 0; i < n; ++i) x[i] += pxc[i];
    return cheb_serial(L, p, b, &x);
}

// ---- all ranks in lock step
struct Ranks {
    int world;
    std::vector<DistHierarchy> D;                 // [rank]
    std::shared_ptr<HaloGeom> mesh(int r) { return D[r].levels[0].space; }
};

using Vecs = std::vector<std::vector<double>>;    // [rank][local index]

// deliver the rows every rank pushes into the ghost arrays of the receivers, exactly through chunks / positions
static void exchange(int world, const std::vector<std::shared_ptr<HaloGeom>> &g, const Vecs &own, Vecs &ghost)
{
    for (int r = 0; r < world; ++r) ghost[r].assign((size_t)g[r]->ghosts[r].size(), std::nan(""));
    for (int q = 0; q < world; ++q) {
        const HaloGeom &G = *g[q];
        for (const HaloGeom::Chunk &c : G.chunks[q])
            for (size_t d = 0; d < c.dst.size(); ++d)
                for (int i = 0; i < c.count; ++i) ghost[c.dst[d]][c.pos[d * 32 + i]] = own[q][G.send_rows[q][c.start + i]];
    }
}

// y = A [own | ghost]
static std::vector<double> spmv_local(const HostCSR &A, int n_own, const std::vector<double> &own, const std::vector<double> &ghost)
{
    std::vector<double> y(A.n_rows, 0.0);
    const int no = n_own > 0 ? n_own : A.n_cols;
    for (int r = 0; r < A.n_rows; ++r)
        for (int k = A.indptr[r]; k < A.indptr[r + 1]; ++k) {
            const int c = A.indices[k];
            y[r] += A.values[k] * (c < no ? own[c] : ghost[c - no]);
        }
    return y;
}

static int fails = 0;
#define CHECK(cond, ...)                 \
    do {                                 \
        if (!(cond)) {                   \
            printf("FAIL: " __VA_ARGS__); \
            printf("\n");                \
            ++fails;                     \
        }                                \
    } while (0)

struct Emu {
    int world;
    const std::vector<AmgLevelHost> &H;
    const AmgParams &p;
    std::vector<DistHierarchy> &D;
    HostCSR A0loc[8];      // level-0 local blocks (the library uses the handle's mesh pattern here)

    std::vector<std::shared_ptr<HaloGeom>> spaces(int l)
    {
        std::vector<std::shared_ptr<HaloGeom>> g(world);
        for (int r = 0; r < world; ++r) g[r] = D[r].levels[l].space;
        return g;
    }
    const HostCSR &Aof(int r, int l) { return l == 0 ? A0loc[r] : D[r].levels[l].A; }

    Vecs cheb(int l, const Vecs &b, const Vecs *x0)
    {
        const bool dist = D[0].levels[l].distributed;
        double scale;
        std::vector<double> om;
        cheb_coefficients(p.lo * H[l].rho, p.hi * H[l].rho, p.nu, &scale, om);
        Vecs prev(world), cur(world), next(world), ghost(world);
        auto ax = [&](const Vecs &v) {
            if (dist) exchange(world, spaces(l), v, ghost);
            Vecs y(world);
            for (int r = 0; r < world; ++r) y[r] = spmv_local(Aof(r, l), D[r].levels[l].A_own, v[r], ghost[r]);
            return y;
        };
        for (int r = 0; r < world; ++r) {
            const DistLevel &L = D[r].levels[l];
            prev[r].assign(L.n, 0.0);
            cur[r].resize(L.n);
            next[r].resize(L.n);
        }
        if (x0) {
            prev = *x0;
            const Vecs a = ax(prev);
            for (int r = 0; r < world; ++r)
                for (int i = 0; i < D[r].levels[l].n; ++i)
                    cur[r][i] = prev[r][i] + scale * D[r].levels[l].dinv[i] * (b[r][i] - a[r][i]);
        } else {
            for (int r = 0; r < world; ++r)
                for (int i = 0; i < D[r].levels[l].n; ++i) cur[r][i] = scale * D[r].levels[l].dinv[i] * b[r][i];
        }
        for (int k = 2; k <= p.nu; ++k) {
            const double w = om[k - 2];
            const Vecs a = ax(cur);
            for (int r = 0; r < world; ++r)
                for (int i = 0; i < D[r].levels[l].n; ++i)
                    next[r][i] = (1.0 - w) * prev[r][i] + w * cur[r][i] + w * scale * D[r].levels[l].dinv[i] * (b[r][i] - a[r][i]);
            prev = cur;
            cur = next;
        }
        return cur;
    }

    Vecs vcycle(int l, const Vecs &b, const Vecs *x0)
    {
        const int nl = (int)H.size();
        const int L_rep = D[0].L_rep;
        if (l == nl - 1) {
            if (!H[l].Ainv.empty()) {
                Vecs x(world);
                for (int r = 0; r < world; ++r) {
                    const DistLevel &L = D[r].levels[l];
                    x[r].assign(L.n, 0.0);
                    for (int i = 0; i < L.n; ++i)
                        for (int j = 0; j < L.n; ++j) x[r][i] += L.Ainv[(size_t)i * L.n + j] * b[r][j];
                }
                return x;
            }
            return cheb(l, b, x0);
        }
        const bool dist = D[0].levels[l].distributed;
        Vecs x = cheb(l, b, x0);
        // restricted residual; rows: own slice of level l + 1 while that level takes part in an exchange
        Vecs cb(world), ghost(world), gb(world);
        {
            Vecs res(world);
            if (dist) exchange(world, spaces(l), x, ghost);
            for (int r = 0; r < world; ++r) {
                const DistLevel &L = D[r].levels[l];
                const std::vector<double> a = spmv_local(Aof(r, l), L.A_own, x[r], ghost[r]);
                res[r].resize(L.n);
                for (int i = 0; i < L.n; ++i) res[r][i] = b[r][i] - a[i];
            }
            if (dist) {
                std::vector<std::shared_ptr<HaloGeom>> g(world);
                for (int r = 0; r < world; ++r) g[r] = l == 0 ? D[r].space_r0 : D[r].levels[l].space;
                exchange(world, g, res, ghost);
            }
            for (int r = 0; r < world; ++r) cb[r] = spmv_local(D[r].levels[l].R, D[r].levels[l].R_own, res[r], ghost[r]);
        }
        if (l + 1 == L_rep) {
            // replicating exchange: slot = [own rows | the others in global order]
            Vecs others(world);
            exchange(world, spaces(l + 1), cb, others);
            for (int r = 0; r < world; ++r) cb[r].insert(cb[r].end(), others[r].begin(), others[r].end());
        }
        Vecs xc = vcycle(l + 1, cb, nullptr);
        // prolongation
        Vecs gxc(world);
        if (l + 1 < L_rep) exchange(world, spaces(l + 1), xc, gxc);
        for (int r = 0; r < world; ++r) {
            const DistLevel &L = D[r].levels[l];
            const std::vector<double> pxc = spmv_local(L.P, L.P_own, xc[r], gxc[r]);
            for (int i = 0; i < L.n; ++i) x[r][i] += pxc[i];
        }
        return cheb(l, b, &x);
    }
};

int main()
{
    // 2-D 7-point mesh operator K + 0.05 M-like (SPD), Dirichlet rows as identity
    const int nx = 61, ny = 47, n = nx * ny;
    HostCSR A;
    A.n_rows = A.n_cols = n;
    A.indptr.assign(n + 1, 0);
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) {
            const int r = j * nx + i;
            const bool bd = i == 0 || j == 0 || i == nx - 1 || j == ny - 1;
            const int di[7] = {0, -1, 1, 0, 0, 1, -1}, dj[7] = {0, 0, 0, -1, 1, -1, 1};
            std::vector<std::pair<int, double>> row;
            for (int q = 0; q < 7; ++q) {
                const int a = i + di[q], b = j + dj[q];
                if (a < 0 || a >= nx || b < 0 || b >= ny) continue;
                const bool bd2 = a == 0 || b == 0 || a == nx - 1 || b == ny - 1;
                double v = q == 0 ? 4.2 : (q < 5 ? -1.0 : 0.01);
                if (bd || bd2) v = (q == 0 && bd) ? 1.0 : 0.0;
                row.emplace_back(b * nx + a, v);
            }
            std::sort(row.begin(), row.end());
            for (auto &e : row) {
                A.indices.push_back(e.first);
                A.values.push_back(e.second);
            }
            A.indptr[r + 1] = (int)A.indices.size();
        }
    std::mt19937 gen(5);
    std::uniform_real_distribution<double> u(-1, 1);
    std::vector<double> b(n);
    for (double &v : b) v = u(gen);

    for (int fuse = 0; fuse < 1; ++fuse) {
        AmgParams p;
        p.coarse_max = 40;
        p.device_inverse = 0;
        std::vector<AmgLevelHost> H;
        amg_setup_host(A, p, H, 1);
        const int nl = (int)H.size();
        const std::vector<double> ref = vcycle_serial(H, p, 0, b, nullptr);
        const std::vector<double> ref2 = vcycle_serial(H, p, 0, b, &ref);
        double refmax = 0.0;
        for (double v : ref2) refmax = std::max(refmax, std::fabs(v));
        for (int world : {2, 3, 4})
            for (int rep_min : {1, 600, 100000000}) {
                std::vector<DistHierarchy> D(world);
                const std::vector<int> part0 = halo_even_split(n, world);
                for (int me = 0; me < world; ++me) {
                    auto mesh = std::make_shared<HaloGeom>();
                    halo_geometry(world, me, part0, {halo_consumer(A, part0)}, false, *mesh);
                    amg_distribute_host(world, me, H, rep_min, mesh, D[me]);
                }
                // every rank must have computed the same picture
                for (int me = 1; me < world; ++me)
                    for (int l = 0; l < nl; ++l) {
                        const auto &a = D[0].levels[l].space, &c = D[me].levels[l].space;
                        CHECK((a == nullptr) == (c == nullptr), "space presence differs");
                        if (a && c) {
                            CHECK(a->ghosts == c->ghosts && a->send_rows == c->send_rows && a->n_flags == c->n_flags, "pictures differ (world %d level %d)", world, l);
                        }
                    }
                if (rep_min == 1) {
                    printf("   world %d ghost entries per rank:", world);
                    for (int me = 0; me < world; ++me) {
                        printf(" [rank %d: r0 %d", me, D[me].space_r0 ? D[me].space_r0->n_ghost() : 0);
                        for (int l = 0; l < nl; ++l)
                            if (D[me].levels[l].space && !D[me].levels[l].space->replicate) printf(", level %d %d of %d", l, D[me].levels[l].n_ghost, D[me].levels[l].n);
                        printf("]");
                    }
                    printf("\n");
                }
                {
                    // the coarsest-level inverse left to the device (AmgParams::device_inverse): the level must still
                    // be replicated, carry the request instead of values, and the inverse of its LOCAL block -- what
                    // dense_inverse.cu computes on every rank -- must be the host inverse in local numbering
                    AmgParams pd = p;
                    pd.device_inverse = 1;
                    std::vector<AmgLevelHost> Hd;
                    amg_setup_host(A, pd, Hd, 1);
                    CHECK(Hd.back().Ainv.empty() && Hd.back().coarse_inverse == AMG_COARSE_INVERSE + 1 && Hd.back().has_inverse(),
                          "deferred inverse not requested");
                    for (int me = 0; me < world; ++me) {
                        auto mesh = std::make_shared<HaloGeom>();
                        halo_geometry(world, me, part0, {halo_consumer(A, part0)}, false, *mesh);
                        DistHierarchy Dd;
                        amg_distribute_host(world, me, Hd, rep_min, mesh, Dd);
                        const DistLevel &Ll = Dd.levels[nl - 1], &Lr = D[me].levels[nl - 1];
                        CHECK(Dd.L_rep == D[me].L_rep && !Ll.distributed && Ll.Ainv.empty() && Ll.coarse_inverse == AMG_COARSE_INVERSE + 1,
                              "deferred inverse lost on the way (world %d rank %d)", world, me);
                        CHECK(Ll.A.n_rows == Ll.n && Ll.A.n_cols == Ll.n, "level with the dense inverse is not square / replicated");
                        std::vector<double> inv;
                        dense_inverse(Ll.A, {}, inv);
                        double d = 0.0, m = 0.0;
                        for (size_t q = 0; q < inv.size(); ++q) {
                            d = std::max(d, std::fabs(inv[q] - Lr.Ainv[q]));
                            m = std::max(m, std::fabs(Lr.Ainv[q]));
                        }
                        CHECK(inv.size() == Lr.Ainv.size() && d <= 1e-11 * m, "inverse of the local block differs from the permuted inverse (%.2e)", d / m);
                    }
                }
                Emu E{world, H, p, D};
                for (int me = 0; me < world; ++me) {
                    const std::shared_ptr<HaloGeom> mesh = D[me].levels[0].space;
                    E.A0loc[me] = halo_extract(A, part0[me + 1] - part0[me], [&](int i) { return part0[me] + i; },
                                               [&](int g) { return mesh->local_col(g); }, mesh->n_own() + mesh->n_ghost());
                    // skip range: rows inside it must not touch ghosts
                    int lo, hi;
                    halo_skip_range(E.A0loc[me], mesh->n_own(), &lo, &hi);
                    for (int r = lo; r < hi; ++r)
                        for (int k = E.A0loc[me].indptr[r]; k < E.A0loc[me].indptr[r + 1]; ++k)
                            CHECK(E.A0loc[me].indices[k] < mesh->n_own(), "skip range contains a row that gathers a ghost");
                }
                Vecs bl(world);
                for (int me = 0; me < world; ++me) bl[me].assign(b.begin() + part0[me], b.begin() + part0[me + 1]);
                const Vecs x1 = E.vcycle(0, bl, nullptr);
                const Vecs x2 = E.vcycle(0, bl, &x1);
                double err = 0.0;
                for (int me = 0; me < world; ++me)
                    for (int i = 0; i < (int)x2[me].size(); ++i) err = std::max(err, std::fabs(x2[me][i] - ref2[part0[me] + i]));
                printf("world %d rep_min %6d: levels %d, first replicated %d, max diff vs serial cycle %.3e\n", world,
                       rep_min, nl, D[0].L_rep, err / refmax);
                CHECK(err <= 1e-12 * refmax, "distributed cycle differs from the serial one");
            }
    }
    printf("halo geometry check: %d failures\n", fails);
    return fails ? 1 : 0;
}
