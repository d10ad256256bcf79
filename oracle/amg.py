"""Smoothed-aggregation AMG (oracle; test infrastructure only).

The reference's inner solves of the Schur-complement sweeps are hypre BoomerAMG, two
V-cycles from a zero guess, set up from scratch at every time step of every
preconditioner application (control/control.py:2056-2067, 2098-2109, 2139-2150,
2172-2183 CN; 2242-2252 ... 2421-2431 BE).  hypre is a third-party dependency that is
not in /root/reference and cannot be installed here, and a classical C/F AMG cannot be
reproduced by an aggregation AMG, so this module does NOT restate BoomerAMG: it defines
the deterministic smoothed-aggregation hierarchy and V-cycle that ``control_b200``
implements in C++ (setup, host) and CUDA (cycle), and is the oracle for THAT algorithm.
Parity of the AMG with hypre is unpinned (DESIGN.md section "AMG").

Algorithm (identical, step for step, in control_b200/csrc/amg_setup.cpp):
  strength   |a_ij|^2 >= theta_l^2 |a_ii a_jj|, j != i, theta_l = theta * theta_decay^level
  aggregate  three greedy passes in natural order (root + strong neighbours; join the
             aggregate of the strongest already-aggregated strong neighbour; leftovers);
             rows without strong neighbours (e.g. Dirichlet identity rows) stay out
  prolong    P = (I - omega D^-1 A) T,  T = the near-kernel candidate (constants on the finest
             level) restricted to each aggregate and normalised; the coarse candidate is the
             vector of those norms, so T cand_c = cand on every level
             omega = 4 / (3 rho),  rho = bound on lambda_max(D^-1 A): the smaller of the
             Gershgorin bound max_i sum_j |a_ij| / |a_ii| and 1.2 x a 30-step power-iteration
             estimate from a fixed start vector (Gershgorin alone overestimates the Galerkin
             operators by 30-50 %, which detunes both omega and the smoother interval: measured
             V-cycle energy contraction at 512^2 0.34 -> 0.17)
  coarse     A_c = P^T A P; stop at n <= coarse_max (dense inverse) or max_levels
  smoother   Chebyshev (oracle/cheb.py) of degree nu (nu_fine on level 0) on D^-1 A over [lo*rho, hi*rho]
  cycle      V(nu, nu); ``solve`` = ``cycles`` V-cycles from a zero guess, optionally
             Chebyshev-accelerated over [acc_lo, acc_hi] (see ``solve``)
"""
import numpy as np
import scipy.sparse as sp

from .cheb import chebyshev
from .fastmv import mv

try:                                       # the loops below are plain Python; numba only
    import numba                           # makes large oracle runs (CPU baseline) bearable
    _jit = numba.njit(cache=True)
except Exception:                          # pragma: no cover
    def _jit(f):
        return f


DEFAULTS = dict(theta=0.08, theta_decay=0.5, max_levels=10, coarse_max=2500, nu=3, nu_fine=0, lo=0.25, hi=1.0, cycles=3,
                acc_lo=0.0, acc_hi=1.0, coarse="inverse")


@_jit
def _aggregate(n, indptr, indices, data, theta):
    diag = np.zeros(n)
    for i in range(n):
        for k in range(indptr[i], indptr[i + 1]):
            if indices[k] == i:
                diag[i] = data[k]
    strong = np.zeros(indices.shape[0], dtype=np.bool_)
    th2 = theta * theta
    for i in range(n):
        for k in range(indptr[i], indptr[i + 1]):
            j = indices[k]
            if j != i:
                a = data[k]
                if a != 0.0 and a * a >= th2 * abs(diag[i] * diag[j]):
                    strong[k] = True
    agg = np.full(n, -1, dtype=np.int64)
    n_agg = 0
    # pass 1: a node whose strong neighbourhood is entirely free roots an aggregate
    for i in range(n):
        if agg[i] != -1:
            continue
        has_nbr = False
        free = True
        for k in range(indptr[i], indptr[i + 1]):
            if strong[k]:
                has_nbr = True
                if agg[indices[k]] != -1:
                    free = False
                    break
        if has_nbr and free:
            agg[i] = n_agg
            for k in range(indptr[i], indptr[i + 1]):
                if strong[k]:
                    agg[indices[k]] = n_agg
            n_agg += 1
    # pass 2: join the aggregate of the strongest strong neighbour aggregated in pass 1
    agg1 = agg.copy()
    for i in range(n):
        if agg1[i] != -1:
            continue
        best = -1
        best_val = -1.0
        for k in range(indptr[i], indptr[i + 1]):
            if strong[k] and agg1[indices[k]] != -1:
                a = abs(data[k])
                if a > best_val:
                    best_val = a
                    best = agg1[indices[k]]
        if best != -1:
            agg[i] = best
    # pass 3: leftovers with strong neighbours form new aggregates among themselves
    for i in range(n):
        if agg[i] != -1:
            continue
        has_nbr = False
        for k in range(indptr[i], indptr[i + 1]):
            if strong[k]:
                has_nbr = True
                break
        if not has_nbr:
            continue
        agg[i] = n_agg
        for k in range(indptr[i], indptr[i + 1]):
            if strong[k] and agg[indices[k]] == -1:
                agg[indices[k]] = n_agg
        n_agg += 1
    return agg, n_agg


def aggregate(A, theta):
    A = sp.csr_matrix(A)
    A.sort_indices()
    agg, n_agg = _aggregate(A.shape[0], A.indptr.astype(np.int64), A.indices.astype(np.int64),
                            A.data.astype(np.float64), float(theta))
    return np.asarray(agg), int(n_agg)


def gershgorin_rho(A):
    A = sp.csr_matrix(A)
    d = np.abs(A.diagonal())
    rowsum = np.asarray(abs(A).sum(axis=1)).ravel()
    return float(np.max(rowsum / d))


POWER_ITS = 30
POWER_SAFETY = 1.2


def power_rho(A, dinv, stop_at=0.0):
    """POWER_ITS steps of the power method on D^-1 A from a fixed start vector; returns the last
    Rayleigh-type quotient ||D^-1 A x|| / ||x|| (a lower estimate of lambda_max).  With
    ``stop_at`` > 0 it ends as soon as POWER_SAFETY x the estimate reaches ``stop_at``: the caller
    takes min(stop_at, POWER_SAFETY x estimate), which is ``stop_at`` from then on."""
    n = A.shape[0]
    i = np.arange(n, dtype=np.float64)
    x = np.sin(0.37 * i + 0.1) + 0.5 * np.cos(1.3 * i)
    lam = 0.0
    for _ in range(POWER_ITS):
        y = dinv * (A @ x)
        ny = float(np.sqrt(np.dot(y, y)))
        if ny == 0.0:
            return 0.0
        lam = ny / float(np.sqrt(np.dot(x, x)))
        if stop_at > 0.0 and POWER_SAFETY * lam >= stop_at:
            return lam
        x = y / ny
    return lam


def spectral_bound(A, dinv):
    g = gershgorin_rho(A)
    return min(g, POWER_SAFETY * power_rho(A, dinv, g))


class Level:
    __slots__ = ("A", "dinv", "rho", "P", "R", "agg", "Ainv")


class Hierarchy:
    def __init__(self, levels, params):
        self.levels = levels
        self.params = params

    def sizes(self):
        return [(L.A.shape[0], L.A.nnz) for L in self.levels]


def setup(A, **kw):
    params = dict(DEFAULTS)
    params.update(kw)
    levels = []
    A = sp.csr_matrix(A).astype(np.float64)
    A.sort_indices()
    cand = np.ones(A.shape[0])
    while True:
        L = Level()
        L.A = A
        d = A.diagonal()
        L.dinv = 1.0 / d
        L.rho = spectral_bound(A, L.dinv)
        L.P = L.R = L.agg = L.Ainv = None
        levels.append(L)
        n = A.shape[0]
        if n <= params["coarse_max"] or len(levels) >= params["max_levels"]:
            break
        agg, n_agg = aggregate(A, params["theta"] * params["theta_decay"] ** (len(levels) - 1))
        if n_agg == 0 or n_agg >= 0.9 * n:
            break
        L.agg = agg
        # tentative prolongator from the near-kernel candidate (constants on the finest level):
        # column j = the candidate restricted to aggregate j, normalised; the coarse candidate is
        # the vector of those norms, so that T cand_coarse = cand exactly on every level
        rows = np.flatnonzero(agg >= 0)
        norms = np.sqrt(np.bincount(agg[rows], weights=cand[rows] ** 2, minlength=n_agg))
        T = sp.csr_matrix((cand[rows] / norms[agg[rows]], (rows, agg[rows])), shape=(n, n_agg))
        cand = norms
        omega = 4.0 / (3.0 * L.rho)
        Az = A.copy()
        Az.eliminate_zeros()
        P = (T - sp.diags(omega * L.dinv) @ (Az @ T)).tocsr()
        P.sort_indices()
        L.P = P
        L.R = P.T.tocsr()
        L.R.sort_indices()
        A = (L.R @ (A @ P)).tocsr()
        A.sort_indices()
    last = levels[-1]
    # coarse = "inverse": dense inverse; "pinv_constant": pseudo-inverse of a symmetric operator
    # whose kernel is the constants (the Neumann pressure Laplacian of the Stokes preconditioner:
    # the smoothed prolongators reproduce constants, so every Galerkin operator keeps that kernel),
    # (A + e e^T)^-1 - e e^T with e = the normalised coarse image of the constants; "smooth": the coarsest level is only smoothed
    if last.A.shape[0] <= 4096 and params["coarse"] == "inverse":
        last.Ainv = np.linalg.inv(last.A.toarray())
    elif last.A.shape[0] <= 4096 and params["coarse"] == "pinv_constant":
        e = cand / np.linalg.norm(cand)           # kernel of the coarsest Galerkin operator
        E = np.outer(e, e)
        last.Ainv = np.linalg.inv(last.A.toarray() + E) - E
    return Hierarchy(levels, params)


def vcycle(H, lvl, b, x=None):
    """One V(nu, nu) cycle on level ``lvl``; ``x`` None means zero initial guess."""
    L = H.levels[lvl]
    p = H.params
    nu = p["nu_fine"] if (lvl == 0 and p["nu_fine"] > 0) else p["nu"]
    if lvl == len(H.levels) - 1:
        if L.Ainv is not None:
            return L.Ainv @ b
        return chebyshev(L.A, L.dinv, b, p["lo"] * L.rho, p["hi"] * L.rho, nu, x)
    x = chebyshev(L.A, L.dinv, b, p["lo"] * L.rho, p["hi"] * L.rho, nu, x)
    r = b - mv(L.A, x)
    xc = vcycle(H, lvl + 1, mv(L.R, r), None)
    x = x + mv(L.P, xc)
    x = chebyshev(L.A, L.dinv, b, p["lo"] * L.rho, p["hi"] * L.rho, nu, x)
    return x


def solve(H, b, cycles=None):
    """``cycles`` V-cycles from a zero guess (the stand-in for
    ``pc_hypre_boomeramg_max_iter``: 2, control/control.py:2065).  The default is THREE
    V(3,3) cycles, not two: this cycle contracts the energy norm by about 0.2 per cycle at
    1024^2 (a BoomerAMG cycle: about 0.1), and the forward/backward time substitution of the
    triangular preconditioner amplifies inner-solve errors.  Measured at config C2 (FGMRES,
    rtol 1e-6): 24 / 11 / 8 outer iterations with 2 / 3 / 4 cycles, 6-7 with exact inner
    solves; three cycles minimise the solve time (DESIGN.md section "AMG").

    With ``acc_lo > 0`` the cycles are Chebyshev-accelerated: the V-cycle B (zero guess) is
    the preconditioner of a ``cycles``-step Chebyshev semi-iteration on B A with spectrum
    bounds [acc_lo, acc_hi] (a symmetric V-cycle has spec(B A) in (0, 1]).  Still a fixed
    linear, symmetric operator, so it is legal inside GMRES and MINRES."""
    p = H.params
    cycles = p["cycles"] if cycles is None else cycles
    if p["acc_lo"] <= 0.0:
        x = None
        for _ in range(cycles):
            x = vcycle(H, 0, b, x)
        return x
    from .cheb import chebyshev_coefficients
    A = H.levels[0].A
    scale, omegas = chebyshev_coefficients(p["acc_lo"], p["acc_hi"], cycles)
    p_prev = np.zeros_like(b)
    p_cur = scale * vcycle(H, 0, b, None)
    for omega in omegas:
        r = b - mv(A, p_cur)
        p_next = (1.0 - omega) * p_prev + omega * p_cur + (omega * scale) * vcycle(H, 0, r, None)
        p_prev, p_cur = p_cur, p_next
    return p_cur
