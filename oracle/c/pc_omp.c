/* C / OpenMP restatement of the reference's hot path for a time-independent K (test infrastructure / CPU baseline
 * only; never linked into the product).
 *
 * What it restates, step for step, from the numpy oracle (which cites the reference line by line):
 *   oracle_kkt_apply      oracle/kkt.py::kkt_apply_fused        <- preconditioner/preconditioner.py:375-543
 *   oracle_pc_apply       oracle/pc.py::construct_pc (CN, BE), construct_pc_diagonal
 *                                                               <- control/control.py:1943-2440
 *   oracle_amg_solve      oracle/amg.py::solve / vcycle         (stand-in for BoomerAMG, control.py:2056-2067)
 *   oracle_cheb*          oracle/cheb.py::chebyshev             <- control/control.py:1967-1982
 * The numpy oracle stays the checker (tests/test_oracle_fast.py compares the two); this file exists so that the CPU
 * arm of bench.py runs the reference's algorithm at the speed the host allows: every loop over the spatial unknowns is
 * an OpenMP loop, sparse product and vector update fused per row, all N time blocks of a row together where the
 * algorithm is batched.  Vectors are block major, (N, n) row-major, as in the reference's mixed PETSc vectors.
 * Compile with -ffp-contract=off: the arithmetic follows the oracle's operation order. */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int32_t n_rows, n_cols;
    const int32_t *ip, *ix;
    const double *v;
} OCsr;

typedef struct {
    int32_t n;
    OCsr A, P, R;
    const double *dinv;
    double rho;
    const double *ainv;          /* dense inverse (last level) or NULL */
    double *x, *b, *r, *t0, *t1; /* work vectors, n each (x, b unused on level 0) */
} OLevel;

typedef struct {
    int32_t nl, nu, nu_fine, cycles;
    double lo, hi;
    OLevel *lv;
} OHier;

typedef struct {
    int32_t n, N, CN, mode;      /* mode 0: block lower-triangular (reference), 1: block-diagonal SPD variant (CN only) */
    double tau, beta, eps;
    OCsr M, K;                   /* operator blocks, no boundary conditions applied */
    OCsr Mbc;                    /* assemble(M, bcs): solver_0 */
    const double *mdinv;         /* 1 / diag(Mbc) */
    double emin, emax;           /* Chebyshev bounds; emax <= 0: one Jacobi sweep */
    int32_t cheb_steps;
    const uint8_t *bc;           /* n: 1 = constrained dof */
    int32_t n_hier;
    OHier *hier;
    const int32_t *fwd_h, *bwd_h;/* N each: hierarchy of the diagonal block of every forward / backward step */
    OCsr off;                    /* CN: h K + (c - 1) M, sub-diagonal block of L_hat (and of its transpose: K symmetric) */
} OPc;

int oracle_omp_threads(void) { return omp_get_max_threads(); }
void oracle_omp_set_threads(int n) { omp_set_num_threads(n > 0 ? n : 1); }

static inline double row_dot(const OCsr *A, int r, const double *x)
{
    double acc = 0.0;
    for (int k = A->ip[r]; k < A->ip[r + 1]; ++k) acc += A->v[k] * x[A->ix[k]];
    return acc;
}

/* ---------------------------------------------------------------- Chebyshev on D^-1 A (one column) */
static void cheb_coeff(double emin, double emax, int steps, double *scale, double *om)
{
    *scale = 2.0 / (emax + emin);
    const double alpha = 1.0 - *scale * emin, mu = 1.0 / alpha, omegaprod = 2.0 / alpha;
    double c_prev = 1.0, c_cur = mu;
    for (int k = 2; k <= steps; ++k) {
        const double c_next = 2.0 * mu * c_cur - c_prev;
        om[k - 2] = omegaprod * c_cur / c_next;
        c_prev = c_cur;
        c_cur = c_next;
    }
}

/* x <- `steps` iterations for A x = b; zero != 0: zero initial guess, else x holds the guess.  w0, w1: work (n). */
static void cheb(const OCsr *A, const double *dinv, const double *b, double *x, int zero, double emin, double emax,
                 int steps, double *w0, double *w1)
{
    const int n = A->n_rows;
    double scale, om[64];
    cheb_coeff(emin, emax, steps, &scale, om);
    double *prev = w0, *cur = w1;
    if (zero) {
#pragma omp parallel for schedule(static)
        for (int i = 0; i < n; ++i) {
            prev[i] = 0.0;
            cur[i] = scale * (dinv[i] * b[i]);
        }
    } else {
#pragma omp parallel for schedule(static)
        for (int i = 0; i < n; ++i) {
            prev[i] = x[i];
            cur[i] = x[i] + scale * (dinv[i] * (b[i] - row_dot(A, i, x)));
        }
    }
    double *next = x;      /* x is free once p_1 is formed; the three buffers rotate */
    for (int k = 2; k <= steps; ++k) {
        const double w = om[k - 2], ws = w * scale;
#pragma omp parallel for schedule(static)
        for (int i = 0; i < n; ++i) {
            const double r = b[i] - row_dot(A, i, cur);
            next[i] = ((1.0 - w) * prev[i] + w * cur[i]) + ws * (dinv[i] * r);
        }
        double *t = prev;
        prev = cur;
        cur = next;
        next = t;
    }
    if (cur != x) {
#pragma omp parallel for schedule(static)
        for (int i = 0; i < n; ++i) x[i] = cur[i];
    }
}

void oracle_cheb(const OCsr *A, const double *dinv, const double *b, double *x, int zero, double emin, double emax, int steps,
                 double *w0, double *w1)
{
    cheb(A, dinv, b, x, zero, emin, emax, steps, w0, w1);
}

/* ---------------------------------------------------------------- AMG V-cycle (oracle/amg.py::vcycle) */
static void vcycle(const OHier *H, int l, const double *b, double *x, int zero)
{
    OLevel *L = &H->lv[l];
    const int n = L->n;
    const int nu = (l == 0 && H->nu_fine > 0) ? H->nu_fine : H->nu;
    if (l == H->nl - 1) {
        if (L->ainv) {
#pragma omp parallel for schedule(static)
            for (int i = 0; i < n; ++i) {
                double acc = 0.0;
                const double *a = L->ainv + (size_t)i * n;
                for (int j = 0; j < n; ++j) acc += a[j] * b[j];
                x[i] = acc;
            }
            return;
        }
        cheb(&L->A, L->dinv, b, x, zero, H->lo * L->rho, H->hi * L->rho, nu, L->t0, L->t1);
        return;
    }
    OLevel *C = &H->lv[l + 1];
    cheb(&L->A, L->dinv, b, x, zero, H->lo * L->rho, H->hi * L->rho, nu, L->t0, L->t1);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) L->r[i] = b[i] - row_dot(&L->A, i, x);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < C->n; ++i) C->b[i] = row_dot(&L->R, i, L->r);
    vcycle(H, l + 1, C->b, C->x, 1);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) x[i] = x[i] + row_dot(&L->P, i, C->x);
    cheb(&L->A, L->dinv, b, x, 0, H->lo * L->rho, H->hi * L->rho, nu, L->t0, L->t1);
}

void oracle_amg_solve(const OHier *H, const double *b, double *x)
{
    for (int c = 0; c < H->cycles; ++c) vcycle(H, 0, b, x, c == 0);
}

/* ---------------------------------------------------------------- KKT operator (time-independent K) */
void oracle_kkt_apply(const OPc *P, const double *x0, const double *x1, double *y0, double *y1)
{
    const int n = P->n, N = P->N;
    const double tau = P->tau, beta = P->beta, h = 0.5 * tau;
#pragma omp parallel
    {
        double *MV = (double *)malloc(sizeof(double) * 4 * (size_t)N), *KV = MV + N, *MZ = KV + N, *KZ = MZ + N;
        double *r0 = (double *)malloc(sizeof(double) * 2 * (size_t)N), *r1 = r0 + N;
#pragma omp for schedule(static)
        for (int i = 0; i < n; ++i) {
            /* the four batched products of this row, all N time blocks (constrained columns masked) */
            for (int j = 0; j < N; ++j) MV[j] = KV[j] = MZ[j] = KZ[j] = 0.0;
            for (int k = P->M.ip[i]; k < P->M.ip[i + 1]; ++k) {
                const int c = P->M.ix[k];
                if (P->bc[c]) continue;
                const double m = P->M.v[k], kk = P->K.v[k];
                for (int j = 0; j < N; ++j) {
                    const double a = x0[(size_t)j * n + c], z = x1[(size_t)j * n + c];
                    MV[j] += m * a;
                    KV[j] += kk * a;
                    MZ[j] += m * z;
                    KZ[j] += kk * z;      /* K symmetric: K^T = K on the shared pattern */
                }
            }
            if (P->CN) {
                for (int j = 0; j < N; ++j) {
                    r0[j] = h * MV[j] + h * KZ[j] + MZ[j];
                    r1[j] = h * KV[j] + MV[j] - (h / beta) * MZ[j];
                }
                for (int j = 1; j < N; ++j) {
                    r0[j] += h * MV[j - 1];
                    r1[j] += h * KV[j - 1] - MV[j - 1];
                }
                for (int j = 0; j + 1 < N; ++j) {
                    r0[j] += h * KZ[j + 1] - MZ[j + 1];
                    r1[j] += -(h / beta) * MZ[j + 1];
                }
                for (int j = 0; j < N; ++j) {
                    y0[(size_t)j * n + i] = r0[j] + (j + 1 < N ? r0[j + 1] : 0.0);      /* T_1 */
                    y1[(size_t)j * n + i] = r1[j] + (j > 0 ? r1[j - 1] : 0.0);          /* T_2 */
                }
            } else {
                for (int j = 0; j < N; ++j) {
                    double a = tau * KZ[j] + MZ[j];
                    if (j + 1 < N) a += tau * MV[j] - MZ[j + 1];
                    double c = tau * KV[j] + MV[j];
                    if (j > 0) c += -MV[j - 1] - (tau / beta) * MZ[j];
                    y0[(size_t)j * n + i] = a;
                    y1[(size_t)j * n + i] = c;
                }
            }
            if (P->bc[i])
                for (int j = 0; j < N; ++j) {
                    y0[(size_t)j * n + i] = x0[(size_t)j * n + i];
                    y1[(size_t)j * n + i] = x1[(size_t)j * n + i];
                }
        }
        free(MV);
        free(r0);
    }
}

/* ---------------------------------------------------------------- preconditioner */
/* solver_0 on all N blocks at once: Chebyshev-20 / Jacobi on assemble(M, bcs), zero guess.  B, U: (N, n); W: 2 N n */
static void solver0(const OPc *P, const double *B, double *U, double *W)
{
    const int n = P->n, N = P->N;
    const size_t sz = (size_t)N * n;
    if (P->emax <= 0.0) {
#pragma omp parallel for schedule(static)
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < N; ++j) U[(size_t)j * n + i] = B[(size_t)j * n + i] * P->mdinv[i];
        return;
    }
    double scale, om[64];
    cheb_coeff(P->emin, P->emax, P->cheb_steps, &scale, om);
    double *prev = W, *cur = W + sz, *next = U;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < N; ++j) {
            prev[(size_t)j * n + i] = 0.0;
            cur[(size_t)j * n + i] = scale * (P->mdinv[i] * B[(size_t)j * n + i]);
        }
    for (int k = 2; k <= P->cheb_steps; ++k) {
        const double w = om[k - 2], ws = w * scale;
#pragma omp parallel for schedule(static)
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < N; ++j) {
                const double *cj = cur + (size_t)j * n;
                const double r = B[(size_t)j * n + i] - row_dot(&P->Mbc, i, cj);
                next[(size_t)j * n + i] = ((1.0 - w) * prev[(size_t)j * n + i] + w * cj[i]) + ws * (P->mdinv[i] * r);
            }
        double *t = prev;
        prev = cur;
        cur = next;
        next = t;
    }
    if (cur != U) memcpy(U, cur, sz * sizeof(double));
}

/* y_i (-)= alpha A x over the free rows, constrained rows of y set to zero (bc_apply) */
static void spmv_update(const OPc *P, const OCsr *A, const double *x, double *y, double alpha, int accumulate)
{
    const int n = P->n;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
        if (P->bc[i]) {
            y[i] = 0.0;
            continue;
        }
        const double ax = alpha * row_dot(A, i, x);
        y[i] = accumulate ? y[i] + ax : ax;
    }
}

/* u_0, u_1 <- pc(b_0, b_1); work: 4 N n doubles */
void oracle_pc_apply(const OPc *P, const double *b0, const double *b1, double *u0, double *u1, double *work)
{
    const int n = P->n, N = P->N;
    const size_t sz = (size_t)N * n;
    const double tau = P->tau, eps = P->eps, h = 0.5 * tau;
    double *T = work, *W = work + sz, *B = work + 3 * sz;
    if (P->CN) {
        /* (1,1) block: u_0 = (2 / tau) T_2^-1 M~^-1 T_1^-1 b_0 */
#pragma omp parallel for schedule(static)
        for (int i = 0; i < n; ++i) {
            double acc = 0.0;
            for (int j = N - 1; j >= 0; --j) {
                acc = b0[(size_t)j * n + i] - acc;
                T[(size_t)j * n + i] = acc;
            }
        }
        solver0(P, T, u0, W);
#pragma omp parallel for schedule(static)
        for (int i = 0; i < n; ++i) {
            double acc = 0.0;
            for (int j = 0; j < N; ++j) {
                acc = u0[(size_t)j * n + i] * (2.0 / tau) - acc;
                u0[(size_t)j * n + i] = acc;
            }
        }
        if (P->mode == 0) {
            /* b = T_2^-1 (T_2 (L u_0) - b_1), bc-masked: control.py:2017-2053 */
#pragma omp parallel
            {
                double *row = (double *)malloc(sizeof(double) * 2 * (size_t)N), *kv = row + N;
#pragma omp for schedule(static)
                for (int i = 0; i < n; ++i) {
                    if (P->bc[i]) {
                        for (int j = 0; j < N; ++j) B[(size_t)j * n + i] = 0.0;
                        continue;
                    }
                    for (int j = 0; j < N; ++j) {
                        const double *uj = u0 + (size_t)j * n;
                        row[j] = row_dot(&P->M, i, uj);
                        kv[j] = row_dot(&P->K, i, uj);
                    }
                    double prev_t2 = 0.0, acc = 0.0;
                    for (int j = 0; j < N; ++j) {
                        double lj = h * kv[j] + row[j];
                        if (j > 0) lj += h * kv[j - 1] - row[j - 1];
                        const double t2 = lj + prev_t2;      /* T_2 */
                        prev_t2 = lj;
                        const double bj = t2 - b1[(size_t)j * n + i];
                        acc = bj - acc;                       /* T_2^-1 */
                        B[(size_t)j * n + i] = acc;
                    }
                }
                free(row);
            }
        } else {
#pragma omp parallel for schedule(static)
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < N; ++j) B[(size_t)j * n + i] = P->bc[i] ? 0.0 : b1[(size_t)j * n + i];
        }
        /* forward sweep */
        for (int j = 0; j < N; ++j) {
            double *bj = B + (size_t)j * n;
            if (j > 0) spmv_update(P, &P->off, u1 + (size_t)(j - 1) * n, bj, -1.0, 1);
            oracle_amg_solve(&P->hier[P->fwd_h[j]], bj, u1 + (size_t)j * n);
        }
        /* b = h M (T_2 u_1) (triangular) or h M u_1 (diagonal), bc-masked */
        for (int j = N - 1; j >= 0; --j) {
            double *bj = B + (size_t)j * n;
            const double *uj = u1 + (size_t)j * n, *up = (P->mode == 0 && j > 0) ? u1 + (size_t)(j - 1) * n : NULL;
#pragma omp parallel for schedule(static)
            for (int i = 0; i < n; ++i) {
                if (P->bc[i]) {
                    bj[i] = 0.0;
                    continue;
                }
                double acc = 0.0;
                for (int k = P->M.ip[i]; k < P->M.ip[i + 1]; ++k) {
                    const int cc = P->M.ix[k];
                    acc += P->M.v[k] * (up ? uj[cc] + up[cc] : uj[cc]);
                }
                bj[i] = h * acc;
            }
        }
        /* backward sweep */
        for (int j = N - 1; j >= 0; --j) {
            double *bj = B + (size_t)j * n;
            if (j + 1 < N) spmv_update(P, &P->off, u1 + (size_t)(j + 1) * n, bj, -1.0, 1);
            oracle_amg_solve(&P->hier[P->bwd_h[j]], bj, u1 + (size_t)j * n);
        }
    } else {
        /* backward Euler: control.py:2193-2438 */
        solver0(P, b0, u0, W);
#pragma omp parallel for schedule(static)
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < N; ++j) u0[(size_t)j * n + i] *= (j == N - 1) ? (1.0 / tau) * (1.0 / eps) : 1.0 / tau;
#pragma omp parallel
        {
            double *row = (double *)malloc(sizeof(double) * 2 * (size_t)N), *kv = row + N;
#pragma omp for schedule(static)
            for (int i = 0; i < n; ++i) {
                if (P->bc[i]) {
                    for (int j = 0; j < N; ++j) B[(size_t)j * n + i] = 0.0;
                    continue;
                }
                for (int j = 0; j < N; ++j) {
                    const double *uj = u0 + (size_t)j * n;
                    row[j] = row_dot(&P->M, i, uj);
                    kv[j] = row_dot(&P->K, i, uj);
                }
                for (int j = 0; j < N; ++j) {
                    double lj = tau * kv[j] + row[j];
                    if (j > 0) lj -= row[j - 1];
                    B[(size_t)j * n + i] = lj - b1[(size_t)j * n + i];
                }
            }
            free(row);
        }
        for (int j = 0; j < N; ++j) {
            double *bj = B + (size_t)j * n;
            if (j > 0) spmv_update(P, &P->M, u1 + (size_t)(j - 1) * n, bj, 1.0, 1);      /* b_j -= (-M) u_{j-1} */
            oracle_amg_solve(&P->hier[P->fwd_h[j]], bj, u1 + (size_t)j * n);
        }
        for (int j = N - 1; j >= 0; --j) {
            double *bj = B + (size_t)j * n;
            spmv_update(P, &P->M, u1 + (size_t)j * n, bj, (j == N - 1) ? eps * tau : tau, 0);
        }
        for (int j = N - 1; j >= 0; --j) {
            double *bj = B + (size_t)j * n;
            if (j + 1 < N) spmv_update(P, &P->M, u1 + (size_t)(j + 1) * n, bj, 1.0, 1);
            oracle_amg_solve(&P->hier[P->bwd_h[j]], bj, u1 + (size_t)j * n);
        }
    }
}

/* ---------------------------------------------------------------- Krylov vector work */
double oracle_dot(int64_t n, const double *x, const double *y)
{
    double acc = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : acc)
    for (int64_t i = 0; i < n; ++i) acc += x[i] * y[i];
    return acc;
}

void oracle_axpby(int64_t n, double a, const double *x, double b, double *y)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) y[i] = a * x[i] + b * y[i];
}
