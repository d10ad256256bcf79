"""In-built block preconditioner of ``Control.Instationary`` (oracle; test infra only).

``construct_pc`` restates control/control.py:1943-2440 step for step on (N, n) arrays.
Inner solvers are injected:
  * ``solver_0`` ~ M^-1: Chebyshev-20/Jacobi with user bounds (1967-1982), one Jacobi
    sweep without bounds (1984-1991), or the AMG stand-in when ``Multigrid`` (1954-1965)
  * ``inner`` ~ BoomerAMG x2 on (block_ii + shift M) assembled with bcs (2056-2067 ...):
    ``"amg"`` (oracle/amg.py, what the CUDA library implements) or ``"exact"`` (sparse LU:
    the reference-independent yardstick for iteration counts).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import amg as _amg
from .cheb import chebyshev
from synthetic.fem import assemble_bc
from .kkt import apply_T_1_inv, apply_T_2, apply_T_2_inv, n_blocks


class InnerSolverCache:
    """The reference re-assembles and re-sets-up every shifted matrix on every use; the
    result only depends on the matrix, so the oracle caches by (level matrix id, shift)."""

    def __init__(self, kind="amg", amg_params=None):
        self.kind = kind
        self.amg_params = amg_params or {}
        self._cache = {}
        self.hierarchies = {}

    def get(self, key, build_matrix):
        if key not in self._cache:
            A = build_matrix()
            if self.kind == "exact":
                lu = spla.splu(sp.csc_matrix(A))
                self._cache[key] = lu.solve
            elif self.kind == "amg":
                H = _amg.setup(A, **self.amg_params)
                self.hierarchies[key] = H
                self._cache[key] = (lambda b, H=H: _amg.solve(H, b))
            else:
                raise ValueError(self.kind)
        return self._cache[key]


def make_solver_0(M, bdofs, lambda_v_bounds=None, Multigrid=False, amg_params=None, steps=20):
    """control/control.py:1953-1991.  Returns f(B) acting on every row of B (N, n)."""
    M_bc = assemble_bc(M, bdofs)
    dinv = 1.0 / M_bc.diagonal()
    if Multigrid:
        H = _amg.setup(M_bc, **(amg_params or {}))
        return lambda B: np.stack([_amg.solve(H, b) for b in B])
    if lambda_v_bounds is not None:
        e_min, e_max = lambda_v_bounds
        return lambda B: chebyshev(M_bc, dinv, B.T, e_min, e_max, steps).T
    return lambda B: B * dinv[None, :]


def construct_pc(M, K_levels, tau, beta, n_t, CN, bdofs, *, lambda_v_bounds=None,
                 Multigrid=False, inner="amg", amg_params=None, epsilon=1e-3,
                 cache=None):
    """Returns ``pc_linear(b_0, b_1) -> (u_0, u_1)`` (arrays (N, n)); the reference's
    in-place ``pc_fn(u_0, u_1, b_0, b_1)`` with the outputs returned instead."""
    if sp.issparse(K_levels):
        K_levels = [K_levels] * n_t
    N = n_blocks(n_t, CN)
    solver_0 = make_solver_0(M, bdofs, lambda_v_bounds, Multigrid, amg_params)
    cache = cache if cache is not None else InnerSolverCache(inner, amg_params)

    def bc_apply(b):
        b[..., bdofs] = 0.0

    def inner_solver(level, transposed, shift):
        # matrix = assemble(block_ii + shift * M, bcs) with block_ii = w K_level(^T) + M
        w = 0.5 * tau if CN else tau
        key = (id(K_levels[level]), bool(transposed), float(shift))

        def build():
            Kl = K_levels[level].T if transposed else K_levels[level]
            return assemble_bc((w * Kl + (1.0 + shift) * M).tocsr(), bdofs)
        return cache.get(key, build)

    if CN:
        h = 0.5 * tau
        c = 0.5 * tau / beta ** 0.5                                     # my_const, 2051

        def pc_linear(b_0, b_1):
            # (1,1) block: 1997-2014
            u_0 = solver_0(apply_T_1_inv(b_0)) * (2.0 / tau)
            u_0 = apply_T_2_inv(u_0)
            # b = T_2 (L u_0) - b_1 : 2017-2048
            b = np.zeros_like(b_1)
            b[0] = h * (K_levels[1] @ u_0[0]) + M @ u_0[0]
            for i in range(1, N):
                b[i] = (h * (K_levels[i + 1] @ u_0[i]) + M @ u_0[i]) \
                    + (h * (K_levels[i] @ u_0[i - 1]) - M @ u_0[i - 1])
            bc_apply(b)
            b = apply_T_2(b)
            b -= b_1
            bc_apply(b)
            # forward sweep: 2053-2116
            b = apply_T_2_inv(b)
            u_1 = np.zeros_like(b_1)
            u_1[0] = inner_solver(1, False, c)(b[0])
            for i in range(1, N):
                b[i] -= h * (K_levels[i] @ u_1[i - 1]) - M @ u_1[i - 1]
                b[i] -= c * (M @ u_1[i - 1])
                bc_apply(b[i])
                u_1[i] = inner_solver(i + 1, False, c)(b[i])
            # 2118-2133
            u_1 = apply_T_2(u_1)
            b = h * (M @ u_1.T).T
            bc_apply(b)
            # backward sweep: 2135-2189
            u_1[N - 1] = inner_solver(N - 1, True, c)(b[N - 1])
            for i in range(N - 2, -1, -1):
                b[i] -= (h * (K_levels[i + 1].T @ u_1[i + 1]) - M @ u_1[i + 1]) \
                    + c * (M @ u_1[i + 1])
                bc_apply(b[i])
                u_1[i] = inner_solver(i, True, c)(b[i])
            return u_0, u_1
    else:
        s = tau / beta ** 0.5
        se = epsilon ** 0.5 * s

        def pc_linear(b_0, b_1):
            # (1,1) block: 2193-2206
            u_0 = solver_0(b_0) * (1.0 / tau)
            u_0[N - 1] *= 1.0 / epsilon
            # b = L u_0 - b_1 : 2209-2237
            b = np.zeros_like(b_1)
            b[0] = tau * (K_levels[0] @ u_0[0]) + M @ u_0[0]
            for i in range(1, N):
                b[i] = (tau * (K_levels[i] @ u_0[i]) + M @ u_0[i]) - M @ u_0[i - 1]
            b -= b_1
            bc_apply(b)
            # forward sweep: 2241-2328
            u_1 = np.zeros_like(b_1)
            u_1[0] = inner_solver(0, False, 0.0)(b[0])
            for i in range(1, N):
                b[i] -= -(M @ u_1[i - 1])
                bc_apply(b[i])
                shift = s if i < N - 1 else se
                u_1[i] = inner_solver(i, False, shift)(b[i])
            # 2330-2350
            b = tau * (M @ u_1.T).T
            b[N - 1] *= epsilon
            bc_apply(b)
            # backward sweep: 2352-2438
            u_1[N - 1] = inner_solver(N - 1, True, se)(b[N - 1])
            for i in range(N - 2, -1, -1):
                b[i] -= -(M @ u_1[i + 1])
                bc_apply(b[i])
                shift = s if i > 0 else 0.0
                u_1[i] = inner_solver(i, True, shift)(b[i])
            return u_0, u_1

    pc_linear.cache = cache
    return pc_linear


def construct_pc_diagonal(M, K, tau, beta, n_t, bdofs, *, lambda_v_bounds=None,
                          inner="amg", amg_params=None, cache=None):
    """Block-diagonal SPD variant diag(A_hat, S_hat) of the CN preconditioner for MINRES
    (not in the reference, whose in-built PC is block lower-triangular: SURVEY.md 7 H2).
    Same building blocks as ``construct_pc`` with the A_10 u_0 coupling dropped, the sign
    of the Schur block flipped to positive, and the mathematically redundant T_2 / T_2^-1
    pair around the forward sweep removed so the operator is symmetric to rounding:
        u_0 = (2/tau) T_2^-1 M~^-1 T_1^-1 b_0
        u_1 = L_hat^-T (tau/2 M) L_hat^-1 b_1
    Requires a time-independent symmetric K."""
    N = n_t - 1
    h = 0.5 * tau
    c = 0.5 * tau / beta ** 0.5
    solver_0 = make_solver_0(M, bdofs, lambda_v_bounds, False, amg_params)
    cache = cache if cache is not None else InnerSolverCache(inner, amg_params)
    A_d = assemble_bc((h * K + (1.0 + c) * M).tocsr(), bdofs)
    solve_d = cache.get(("diag", float(c)), lambda: A_d)
    off = (h * K + (c - 1.0) * M).tocsr()

    def pc_diag(b_0, b_1):
        u_0 = solver_0(apply_T_1_inv(b_0)) * (2.0 / tau)
        u_0 = apply_T_2_inv(u_0)
        b = b_1.copy()
        b[:, bdofs] = 0.0
        u_1 = np.zeros_like(b_1)
        u_1[0] = solve_d(b[0])
        for i in range(1, N):
            b[i] -= off @ u_1[i - 1]
            b[i, bdofs] = 0.0
            u_1[i] = solve_d(b[i])
        b = h * (M @ u_1.T).T
        b[:, bdofs] = 0.0
        u_1[N - 1] = solve_d(b[N - 1])
        for i in range(N - 2, -1, -1):
            b[i] -= off.T @ u_1[i + 1]
            b[i, bdofs] = 0.0
            u_1[i] = solve_d(b[i])
        return u_0, u_1

    pc_diag.cache = cache
    return pc_diag
