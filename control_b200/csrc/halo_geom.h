// Pure host geometry of the device-initiated halo exchange and of the distributed AMG hierarchy (no CUDA calls):
// ghost lists, send chunks, flag numbering, local numbering and local row blocks.  Header-only so that
// tests/native/halo_geom_check.cu can play all ranks of a partition in one process.
#pragma once
#include <algorithm>
#include <cstdint>
#include <functional>
#include <memory>
#include <vector>

#include "amg_setup.h"

// rows of a global CSR pattern, distributed over the ranks by row_part[world + 1]; its columns live in the space
struct SpaceConsumer {
    int n_rows = 0;
    const int *indptr = nullptr, *indices = nullptr;
    const int *row_part = nullptr;
};

// Who gathers what in one vector space (a hierarchy level, or the mesh).  Every rank computes the WHOLE picture
// (all ghost lists, all send lists): the hierarchy set-up is replicated on the host, and the flag layout on a rank
// depends on the chunking of every sender.
struct HaloGeom {
    int world = 1, me = 0;
    bool replicate = false;                       // every rank receives every row (first replicated AMG level)
    std::vector<int> part;                        // world + 1 row offsets
    std::vector<std::vector<int>> ghosts;         // [rank] sorted global ids gathered but not owned
    struct Chunk {
        int start = 0, count = 0;                 // segment of the sender's row list
        std::vector<int> dst;                     // receiving ranks
        std::vector<int> flag;                    // per dst: flag index on the receiver
        std::vector<int> pos;                     // per dst x 32: position of each row in the receiver's ghost list
    };
    std::vector<std::vector<int>> send_rows;      // [rank] local rows, chunk after chunk
    std::vector<std::vector<Chunk>> chunks;       // [rank]
    std::vector<int> n_flags;                     // [rank] flags the rank waits on

    int n_own(int r) const { return part[r + 1] - part[r]; }
    int n_own() const { return n_own(me); }
    int n_ghost() const { return (int)ghosts[me].size(); }
    long long stride(int r) const { return replicate ? (long long)part[world] : (long long)ghosts[r].size(); }
    // this rank's numbering: owned first, ghosts in global order behind
    int local_col(int g) const
    {
        const int b = part[me], e = part[me + 1];
        if (g >= b && g < e) return g - b;
        const std::vector<int> &gh = ghosts[me];
        return (e - b) + (int)(std::lower_bound(gh.begin(), gh.end(), g) - gh.begin());
    }
};

inline std::vector<int> halo_even_split(int n, int world)
{
    std::vector<int> part(world + 1, 0);
    const int base = n / world, rem = n % world;      // PETSc's ownership split
    for (int r = 0; r < world; ++r) part[r + 1] = part[r] + base + (r < rem ? 1 : 0);
    return part;
}

inline SpaceConsumer halo_consumer(const HostCSR &G, const std::vector<int> &row_part)
{
    SpaceConsumer c;
    c.n_rows = G.n_rows;
    c.indptr = G.indptr.data();
    c.indices = G.indices.data();
    c.row_part = row_part.data();
    return c;
}

// [skip_lo, skip_hi): the largest run of rows of a local matrix that gather no ghost column
inline void halo_skip_range(const HostCSR &local, int n_own, int *skip_lo, int *skip_hi)
{
    int best_lo = 0, best_hi = 0, prev = -1;
    for (int r = 0; r <= local.n_rows; ++r) {
        bool touches = (r == local.n_rows);
        if (!touches)
            for (int k = local.indptr[r]; k < local.indptr[r + 1]; ++k)
                if (local.indices[k] >= n_own) {
                    touches = true;
                    break;
                }
        if (!touches) continue;
        if (r - (prev + 1) > best_hi - best_lo) {
            best_lo = prev + 1;
            best_hi = r;
        }
        prev = r;
    }
    *skip_lo = best_lo;
    *skip_hi = best_hi;
}

inline void halo_geometry(int world, int me, const std::vector<int> &part, const std::vector<SpaceConsumer> &consumers,
                          bool replicate, HaloGeom &sp)
{
    sp.world = world;
    sp.me = me;
    sp.replicate = replicate;
    sp.part = part;
    sp.ghosts.assign(world, std::vector<int>());
    if (replicate) {
        for (int r = 0; r < world; ++r) {
            std::vector<int> &g = sp.ghosts[r];
            g.reserve(part[world] - (part[r + 1] - part[r]));
            for (int i = 0; i < part[world]; ++i)
                if (i < part[r] || i >= part[r + 1]) g.push_back(i);
        }
    } else {
        for (const SpaceConsumer &c : consumers)
            for (int r = 0; r < world; ++r) {
                std::vector<int> &g = sp.ghosts[r];
                const int cb = part[r], ce = part[r + 1];
                for (int row = c.row_part[r]; row < c.row_part[r + 1]; ++row)
                    for (int k = c.indptr[row]; k < c.indptr[row + 1]; ++k) {
                        const int col = c.indices[k];
                        if (col < cb || col >= ce) g.push_back(col);
                    }
            }
        for (int r = 0; r < world; ++r) {
            std::vector<int> &g = sp.ghosts[r];
            std::sort(g.begin(), g.end());
            g.erase(std::unique(g.begin(), g.end()), g.end());
        }
    }
    // send side of every rank: rows grouped by the set of ranks that gather them, chunks of <= 32 rows
    sp.send_rows.assign(world, std::vector<int>());
    sp.chunks.assign(world, std::vector<HaloGeom::Chunk>());
    for (int q = 0; q < world; ++q) {
        struct Item { int row, dst, pos; };
        std::vector<Item> items;
        for (int r = 0; r < world; ++r) {
            if (r == q) continue;
            const std::vector<int> &g = sp.ghosts[r];
            const int lo = (int)(std::lower_bound(g.begin(), g.end(), part[q]) - g.begin());
            const int hi = (int)(std::lower_bound(g.begin(), g.end(), part[q + 1]) - g.begin());
            for (int i = lo; i < hi; ++i) items.push_back({g[i] - part[q], r, i});
        }
        std::sort(items.begin(), items.end(), [](const Item &a, const Item &b) { return a.row != b.row ? a.row < b.row : a.dst < b.dst; });
        std::vector<int> &rows = sp.send_rows[q];
        std::vector<HaloGeom::Chunk> &chunks = sp.chunks[q];
        size_t i = 0;
        while (i < items.size()) {
            size_t j = i;
            std::vector<int> dst, pos;
            while (j < items.size() && items[j].row == items[i].row) {
                dst.push_back(items[j].dst);
                pos.push_back(items[j].pos);
                ++j;
            }
            if (chunks.empty() || chunks.back().count == 32 || chunks.back().dst != dst) {
                HaloGeom::Chunk c;
                c.start = (int)rows.size();
                c.dst = dst;
                c.pos.assign(dst.size() * 32, 0);
                chunks.push_back(c);
            }
            HaloGeom::Chunk &c = chunks.back();
            for (size_t d = 0; d < dst.size(); ++d) c.pos[d * 32 + c.count] = pos[d];
            c.count++;
            rows.push_back(items[i].row);
            i = j;
        }
    }
    sp.n_flags.assign(world, 0);
    for (int q = 0; q < world; ++q)
        for (HaloGeom::Chunk &c : sp.chunks[q]) {
            c.flag.resize(c.dst.size());
            for (size_t d = 0; d < c.dst.size(); ++d) c.flag[d] = sp.n_flags[c.dst[d]]++;
        }
}

// rows (global ids, in local order) of G with the columns renumbered by colmap
template <typename RowFn, typename ColFn>
HostCSR halo_extract(const HostCSR &G, int n_rows, RowFn global_row, ColFn colmap, int n_cols)
{
    HostCSR L;
    L.n_rows = n_rows;
    L.n_cols = n_cols;
    L.indptr.assign(n_rows + 1, 0);
    for (int i = 0; i < n_rows; ++i) {
        const int g = global_row(i);
        L.indptr[i + 1] = L.indptr[i] + (G.indptr[g + 1] - G.indptr[g]);
    }
    L.indices.resize(L.indptr[n_rows]);
    L.values.resize(L.indptr[n_rows]);
    for (int i = 0; i < n_rows; ++i) {
        const int g = global_row(i);
        int q = L.indptr[i];
        for (int k = G.indptr[g]; k < G.indptr[g + 1]; ++k, ++q) {
            L.indices[q] = colmap(G.indices[k]);
            L.values[q] = G.values[k];
        }
    }
    return L;
}

// ---------------------------------------------------------------------------------------------------------
// The hierarchy of one rank: which levels are distributed, the exchange geometry of each vector space, and the
// local row blocks of every matrix in local numbering.
//   levels 0 .. L_rep - 1   distributed by rows (level 0: the handle's partition; others: ownership follows the
//                           aggregates, coarse unknowns renumbered rank after rank)
//   level  L_rep            the first replicated level: its right-hand side is computed by owners and replicated;
//                           numbered own rows first, the others in global order behind (= the slot layout)
//   levels > L_rep          replicated, natural numbering
// ---------------------------------------------------------------------------------------------------------
struct DistLevel {
    bool distributed = false;
    int n = 0, n_ghost = 0;
    HostCSR A, P, R;                 // local blocks (P, R empty where the hierarchy has none)
    int A_own = 0, P_own = 0, R_own = 0;      // owned columns of each block's column space (0: all columns are local)
    std::vector<double> dinv;        // own rows, then the ghosts
    std::vector<double> Ainv;        // dense inverse in local numbering (last level), or, when it is left to the device:
    int coarse_inverse = 0;          // what to compute from A (AmgLevelHost::coarse_inverse)
    std::vector<double> coarse_shift;   // kernel vector of the pseudo-inverse in local numbering
    std::shared_ptr<HaloGeom> space; // vectors of this level (levels < L_rep: halo; level L_rep: replicating)
};

struct DistHierarchy {
    int L_rep = 0;
    std::vector<std::vector<int>> part;
    std::vector<DistLevel> levels;
    std::shared_ptr<HaloGeom> space_r0;      // level-0 residual as the restriction gathers it
};

// rep_nnz: a coarse level with fewer matrix entries per rank than this (and everything below it) is replicated
inline void amg_distribute_host(int world, int me, const std::vector<AmgLevelHost> &host_in, int64_t rep_nnz,
                                const std::shared_ptr<HaloGeom> &mesh_space, DistHierarchy &D)
{
    const int nl = (int)host_in.size();
    int L_rep = nl;
    for (int l = 1; l < nl; ++l)
        if (host_in[l].A.nnz() < rep_nnz * world || host_in[l].has_inverse()) {
            L_rep = l;
            break;
        }
    D.L_rep = L_rep;
    // Ownership of the coarse levels FOLLOWS the fine level: an aggregate belongs to the rank that owns most of its
    // members, and the coarse unknowns are renumbered rank after rank (the set-up numbers the aggregates of its
    // third pass behind all others, so contiguous blocks of the original numbering would scatter a rank's coarse
    // rows over the whole mesh: long ghost lists, every CTA waiting).  perm[l][old] = new.
    D.part.assign(nl, std::vector<int>());
    D.part[0] = halo_even_split(host_in[0].A.n_rows, world);
    std::vector<std::vector<int>> perm(nl);
    for (int l = 1; l < nl; ++l) {
        const int n_l = host_in[l].A.n_rows;
        if (l > L_rep) {
            D.part[l] = halo_even_split(n_l, world);      // replicated, natural numbering: the partition is not used
            continue;
        }
        const std::vector<int> &agg = host_in[l - 1].agg;
        const std::vector<int> &pf = D.part[l - 1];
        std::vector<int> votes((size_t)n_l * world, 0);
        for (int i = 0; i < (int)agg.size(); ++i) {
            if (agg[i] < 0) continue;
            const int inew = perm[l - 1].empty() ? i : perm[l - 1][i];
            const int o = (int)(std::upper_bound(pf.begin(), pf.end(), inew) - pf.begin()) - 1;
            votes[(size_t)agg[i] * world + o]++;
        }
        std::vector<int> owner(n_l, 0);
        D.part[l].assign(world + 1, 0);
        for (int J = 0; J < n_l; ++J) {
            int best = 0;
            for (int r = 1; r < world; ++r)
                if (votes[(size_t)J * world + r] > votes[(size_t)J * world + best]) best = r;
            owner[J] = best;
            D.part[l][best + 1]++;
        }
        for (int r = 0; r < world; ++r) D.part[l][r + 1] += D.part[l][r];
        std::vector<int> next(D.part[l].begin(), D.part[l].end() - 1);
        perm[l].resize(n_l);
        for (int J = 0; J < n_l; ++J) perm[l][J] = next[owner[J]]++;
    }
    // the hierarchy in the new numbering (levels without a permutation are taken as they are)
    auto permuted = [](const HostCSR &G, const std::vector<int> &rp, const std::vector<int> &cp) {
        HostCSR Q;
        Q.n_rows = G.n_rows;
        Q.n_cols = G.n_cols;
        Q.indptr.assign(G.n_rows + 1, 0);
        for (int r = 0; r < G.n_rows; ++r) Q.indptr[(rp.empty() ? r : rp[r]) + 1] = G.indptr[r + 1] - G.indptr[r];
        for (int r = 0; r < G.n_rows; ++r) Q.indptr[r + 1] += Q.indptr[r];
        Q.indices.resize(G.indices.size());
        Q.values.resize(G.values.size());
        for (int r = 0; r < G.n_rows; ++r) {
            int q = Q.indptr[rp.empty() ? r : rp[r]];
            for (int k = G.indptr[r]; k < G.indptr[r + 1]; ++k, ++q) {
                Q.indices[q] = cp.empty() ? G.indices[k] : cp[G.indices[k]];
                Q.values[q] = G.values[k];
            }
        }
        return Q;
    };
    std::vector<AmgLevelHost> hostp(nl);
    for (int l = 0; l < nl; ++l) {
        const AmgLevelHost &Li = host_in[l];
        AmgLevelHost &Lo = hostp[l];
        const std::vector<int> none;
        const std::vector<int> &pl = perm[l], &pn = (l + 1 < nl) ? perm[l + 1] : none;
        Lo.rho = Li.rho;
        if (pl.empty() && pn.empty() && l > 0) {
            Lo.A = Li.A;
            Lo.P = Li.P;
            Lo.R = Li.R;
            Lo.dinv = Li.dinv;
            Lo.Ainv = Li.Ainv;
            Lo.coarse_inverse = Li.coarse_inverse;
            Lo.coarse_shift = Li.coarse_shift;
            continue;
        }
        if (l == 0) {      // level 0 keeps its numbering and is read through the handle's mesh pattern: sizes only
            Lo.A.n_rows = Li.A.n_rows;
            Lo.A.n_cols = Li.A.n_cols;
        } else {
            Lo.A = pl.empty() ? Li.A : permuted(Li.A, pl, pl);
        }
        if (l + 1 < nl) {
            Lo.P = permuted(Li.P, pl, pn);
            Lo.R = permuted(Li.R, pn, pl);
        }
        Lo.dinv.resize(Li.dinv.size());
        for (size_t i = 0; i < Li.dinv.size(); ++i) Lo.dinv[pl.empty() ? i : pl[i]] = Li.dinv[i];
        Lo.coarse_inverse = Li.coarse_inverse;
        Lo.coarse_shift.resize(Li.coarse_shift.size());
        for (size_t i = 0; i < Li.coarse_shift.size(); ++i) Lo.coarse_shift[pl.empty() ? i : pl[i]] = Li.coarse_shift[i];
        if (!Li.Ainv.empty()) {
            const int n = Li.A.n_rows;
            Lo.Ainv.resize(Li.Ainv.size());
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < n; ++j)
                    Lo.Ainv[(size_t)(pl.empty() ? i : pl[i]) * n + (pl.empty() ? j : pl[j])] = Li.Ainv[(size_t)i * n + j];
        }
    }
    const std::vector<AmgLevelHost> &host = hostp;
    const std::vector<std::vector<int>> &part = D.part;
    D.levels.assign(nl, DistLevel());
    std::vector<std::shared_ptr<HaloGeom>> space(nl);
    space[0] = mesh_space;
    if (nl > 1) {
        D.space_r0 = std::make_shared<HaloGeom>();
        halo_geometry(world, me, part[0], {halo_consumer(host[0].R, part[1])}, false, *D.space_r0);
    }
    for (int l = 1; l < std::min(L_rep, nl); ++l) {
        std::vector<SpaceConsumer> cons;
        cons.push_back(halo_consumer(host[l].A, part[l]));
        cons.push_back(halo_consumer(host[l - 1].P, part[l - 1]));
        if (l + 1 < nl) {
            cons.push_back(halo_consumer(host[l].R, part[l + 1]));
        }
        space[l] = std::make_shared<HaloGeom>();
        halo_geometry(world, me, part[l], cons, false, *space[l]);
    }
    if (L_rep < nl) {
        space[L_rep] = std::make_shared<HaloGeom>();
        halo_geometry(world, me, part[L_rep], {}, true, *space[L_rep]);
    }
    // numbering of a level's vectors on this rank
    auto colmap = [&](int l) {
        return [&space, l](int g) { return space[l] ? space[l]->local_col(g) : g; };
    };
    auto n_local_cols = [&](int l) {
        if (l < L_rep) return space[l]->n_own() + space[l]->n_ghost();
        return host[l].A.n_rows;
    };
    std::vector<int> inv_rep;      // replicated level L_rep: local index -> global row
    if (L_rep < nl) {
        inv_rep.resize(host[L_rep].A.n_rows);
        for (int g = 0; g < (int)inv_rep.size(); ++g) inv_rep[space[L_rep]->local_col(g)] = g;
    }
    auto n_rows_of = [&](int l) { return l < L_rep ? part[l][me + 1] - part[l][me] : host[l].A.n_rows; };
    auto row_of = [&](int l) {
        return [&part, &inv_rep, L_rep, me, l](int i) { return l < L_rep ? part[l][me] + i : (l == L_rep ? inv_rep[i] : i); };
    };
    // rows of a restriction-type matrix (level l + 1 rows): the own slice while level l + 1 takes part in an
    // exchange (distributed, or the first replicated level: owner computes, then replicates)
    auto n_rrows = [&](int l) { return (l + 1 <= L_rep) ? part[l + 1][me + 1] - part[l + 1][me] : host[l + 1].A.n_rows; };
    auto rrow_of = [&](int l) { return [&part, L_rep, me, l](int i) { return (l + 1 <= L_rep) ? part[l + 1][me] + i : i; }; };

    for (int l = 0; l < nl; ++l) {
        const AmgLevelHost &Lh = host[l];
        DistLevel &Ld = D.levels[l];
        const bool dist = l < L_rep;
        Ld.distributed = dist;
        Ld.n = n_rows_of(l);
        Ld.n_ghost = dist ? space[l]->n_ghost() : 0;
        Ld.space = space[l];
        if (l > 0) {      // level 0 uses the handle's local mesh pattern
            Ld.A = halo_extract(Lh.A, Ld.n, row_of(l), colmap(l), n_local_cols(l));
        }
        Ld.A_own = dist ? Ld.n : 0;
        Ld.dinv.resize((size_t)Ld.n + Ld.n_ghost);
        {
            const auto rf = row_of(l);
            for (int i = 0; i < Ld.n; ++i) Ld.dinv[i] = Lh.dinv[rf(i)];
            if (dist)
                for (int g = 0; g < Ld.n_ghost; ++g) Ld.dinv[Ld.n + g] = Lh.dinv[space[l]->ghosts[me][g]];
        }
        if (l + 1 < nl) {
            Ld.P = halo_extract(Lh.P, Ld.n, row_of(l), colmap(l + 1), n_local_cols(l + 1));
            Ld.P_own = l + 1 < L_rep ? space[l + 1]->n_own() : 0;
            const int nr = n_rrows(l);
            const int ncols_l = (l == 0) ? part[0][me + 1] - part[0][me] + D.space_r0->n_ghost() : n_local_cols(l);
            if (l == 0) {
                const std::shared_ptr<HaloGeom> s0 = D.space_r0;
                Ld.R = halo_extract(Lh.R, nr, rrow_of(l), [s0](int g) { return s0->local_col(g); }, ncols_l);
            } else {
                Ld.R = halo_extract(Lh.R, nr, rrow_of(l), colmap(l), ncols_l);
            }
            Ld.R_own = dist ? Ld.n : 0;
        } else if (!Lh.Ainv.empty()) {
            const int n = Ld.n;
            Ld.Ainv.resize((size_t)n * n);
            const auto rf = row_of(l);
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < n; ++j) Ld.Ainv[(size_t)i * n + j] = Lh.Ainv[(size_t)rf(i) * n + rf(j)];
        } else if (Lh.coarse_inverse) {
            // left to the device: the inverse of the LOCAL block Ld.A is the inverse in local numbering
            Ld.coarse_inverse = Lh.coarse_inverse;
            const auto rf = row_of(l);
            Ld.coarse_shift.resize(Lh.coarse_shift.size());
            for (size_t i = 0; i < Lh.coarse_shift.size(); ++i) Ld.coarse_shift[i] = Lh.coarse_shift[rf((int)i)];
        }
    }
}
