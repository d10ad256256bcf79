"""Instationary Stokes control: ``Control.Instationary.incompressible_linear_solve`` restated
on assembled matrices (oracle; test infrastructure only).

  * block structure and outer operator   control/control.py:3750-3957, 4273-4289 and
    ``MultiBlockSystemMatrix.mult`` with sub-block T transforms,
    preconditioner/preconditioner.py:375-543 (471-525 for the ``sub_n_blocks`` branch)
  * ``ConstantNullspace``                preconditioner/preconditioner.py:133-155
  * in-built pressure-Schur preconditioner  control/control.py:4299-4513 (CN), 4515-4687 (BE)
  * solve driver                         control/control.py:4273-4297, 4688-4725

Unknowns: ``x_0`` = (2N, n_v) blocks [v | zeta], ``x_1`` = (2N, n_p) blocks [mu | p] (the first
N pressure blocks multiply B^T in the adjoint-equation rows, control/control.py:3765-3766).
Pinned: ``stokes_apply_literal`` and ``ConstantNullspace`` reproduce the reference's Stokes
known-answer test (test/test_control.py:232-358, the stationary system = the same block structure
with N = 1 and tau = 1) to 1e-13 (tests/test_oracle.py); the fused form equals the literal one to
rounding.  Unpinned: everything the real PETSc/hypre stack decides (iteration counts, histories).
The inner solves on ``K_p`` are six cycles of the stand-in AMG (ONE hypre BoomerAMG cycle in the
reference, control/control.py:4300-4309): parity with hypre is unpinned, as for the heat path.
"""
import numpy as np

from . import amg as _amg
from . import kkt, krylov
from .cheb import chebyshev
from .control import system_solve
from .pc import construct_pc


class ConstantNullspace:
    """preconditioner/preconditioner.py:133-155 (alpha = 1)."""

    def __init__(self, alpha=1.0):
        self.alpha = alpha

    @staticmethod
    def _mean(x):
        return x.sum(axis=-1, keepdims=True) / float(x.shape[-1])

    def project(self, x):
        x -= self._mean(x)

    def pre_mult_corrected_lhs(self, x):
        xc = x.copy()
        self.project(xc)
        return xc

    def post_mult_correct_lhs(self, x, y):
        self.project(y)
        y += self.alpha * self._mean(x)

    def pc_pre_mult_corrected(self, b):
        bc = b.copy()
        self.project(bc)
        return bc

    def pc_post_mult_correct(self, u, b):
        self.project(u)
        u += self._mean(b)


def _T(y, N, first, second):
    y[:N] = first(y[:N])
    y[N:] = second(y[N:])


def stokes_apply_literal(heat_blocks, B, tau, N, CN, ns_v, ns_p, x0, x1):
    """``mult`` of the outer system, block by block."""
    b00, b01, b10, b11 = heat_blocks
    xc0 = ns_v.pre_mult_corrected_lhs(x0)
    xc1 = ns_p.pre_mult_corrected_lhs(x1)
    y0 = np.zeros_like(x0)
    y1 = np.zeros_like(x1)
    for (i, j), A in b00.items():              # block_00 = the heat-type KKT with index offsets
        if A is not None:
            y0[i] += A @ xc0[j]
    for (i, j), A in b01.items():
        if A is not None:
            y0[i] += A @ xc0[N + j]
    for (i, j), A in b10.items():
        if A is not None:
            y0[N + i] += A @ xc0[j]
    for (i, j), A in b11.items():
        if A is not None:
            y0[N + i] += A @ xc0[N + j]
    for i in range(2 * N):                     # block_01 = diag(tau B^T), block_10 = diag(tau B)
        y0[i] += tau * (B.T @ xc1[i])
        y1[i] += tau * (B @ xc0[i])
    if CN:                                     # preconditioner.py:471-525
        _T(y0, N, kkt.apply_T_1, kkt.apply_T_2)
        _T(y1, N, kkt.apply_T_2, kkt.apply_T_1)
    ns_v.post_mult_correct_lhs(x0, y0)
    ns_p.post_mult_correct_lhs(x1, y1)
    return y0, y1


def stokes_apply_fused(M_v, K_v, B, tau, beta, n_t, CN, bdofs_v, x0, x1):
    """The same operator in the form the CUDA library implements: the fused heat-type KKT apply
    on the velocity blocks plus the (transformed) divergence couplings."""
    N = kkt.n_blocks(n_t, CN)
    ns_p = ConstantNullspace()
    xc0 = x0.copy()
    xc0[:, bdofs_v] = 0.0
    xc1 = ns_p.pre_mult_corrected_lhs(x1)
    y0a, y0b = kkt.kkt_apply_fused(M_v, K_v, tau, beta, n_t, CN, bdofs_v, x0[:N], x0[N:])
    g0 = tau * (B.T @ xc1.T).T                 # tau B^T x_1, block by block
    y1 = tau * (B @ xc0.T).T
    if CN:
        _T(g0, N, kkt.apply_T_1, kkt.apply_T_2)
        _T(y1, N, kkt.apply_T_2, kkt.apply_T_1)
    g0[:, bdofs_v] = 0.0
    y0 = np.concatenate([y0a, y0b]) + g0
    ns_p.post_mult_correct_lhs(x1, y1)
    return y0, y1


def make_solver_p(M_p, K_p, lambda_p_bounds, amg_params=None):
    """``solver_K_p`` (AMG cycles on the Neumann Laplacian; ONE BoomerAMG cycle in the reference,
    control/control.py:4300-4309; six of the stand-in AMG here) and
    ``solver_M_p`` (Chebyshev-20/Jacobi or one Jacobi sweep, 4311-4333), acting on (k, n_p)."""
    params = dict(cycles=6, coarse="pinv_constant")       # see ctl_stokes_pc_default_options
    params.update(amg_params or {})
    H = _amg.setup(K_p, **params)
    dinv = 1.0 / M_p.diagonal()

    def K_solve(Bk):
        return np.stack([_amg.solve(H, b) for b in Bk])

    if lambda_p_bounds is not None:
        e_min, e_max = lambda_p_bounds

        def M_solve(Bk):
            return chebyshev(M_p, dinv, Bk.T, e_min, e_max, 20).T
    else:
        def M_solve(Bk):
            return Bk * dinv[None, :]
    return K_solve, M_solve, H


def construct_stokes_pc(M_v, K_v, B, M_p, K_p, tau, beta, n_t, CN, bdofs_v, *, lambda_v_bounds=None,
                        lambda_p_bounds=None, inner="amg", amg_params=None, amg_params_p=None, epsilon=1e-3,
                        D_p=None, Multigrid=False):
    """``pc_fn(b_0, b_1) -> (u_0, u_1)`` of control/control.py:4337-4513 (CN) / 4515-4687 (BE).
    ``K_p``: the Laplacian ``solver_K_p`` inverts (3746, 4300-4309); ``D_p``: the forward form on the pressure
    space per time level (``D_p_i``, 3787-3789, 3926-3928; one matrix or n_t matrices) of the pressure-space
    KKT multiply -- default ``K_p``, which is what the Stokes forward operator gives."""
    N = kkt.n_blocks(n_t, CN)
    ns_v = kkt.DirichletBCNullspace(bdofs_v)
    vparams = dict(cycles=6)                      # ctl_stokes_pc_default_options: six cycles on the P2 operator
    vparams.update(amg_params or {})
    heat_pc = construct_pc(M_v, K_v, tau, beta, n_t, CN, bdofs_v, lambda_v_bounds=lambda_v_bounds,
                           Multigrid=Multigrid, inner=inner, amg_params=vparams, epsilon=epsilon)
    K_solve, M_solve, _ = make_solver_p(M_p, K_p, lambda_p_bounds, amg_params_p)
    pblocks = kkt.build_blocks(M_p, K_p if D_p is None else D_p, tau, beta, n_t, CN)       # block_*_int_p, 3805-3957
    inner_parameters = {"preconditioner": True, "linear_solver": "gmres", "maximum_iterations": 5,
                        "relative_tolerance": 0.0, "absolute_tolerance": 0.0}      # 4355-4361

    def apply_heat(x0, x1):
        return kkt.kkt_apply_fused(M_v, K_v, tau, beta, n_t, CN, bdofs_v, x0, x1)

    def pc_fn(b_0, b_1):
        n_v, n_p = M_v.shape[0], M_p.shape[0]
        v, zeta, _ = system_solve(apply_heat, ns_v, np.zeros((N, n_v)), np.zeros((N, n_v)), b_0[:N], b_0[N:],
                                  solver_parameters=inner_parameters, pc_fn=heat_pc)
        u_0 = np.concatenate([v, zeta])
        # u_1 = -b_1 + block_10 u_0, then the pressure Schur-complement approximation
        h0 = tau * (B @ u_0[:N].T).T
        h1 = tau * (B @ u_0[N:].T).T
        if CN:
            h0 = kkt.apply_T_2(h0)
            h1 = kkt.apply_T_1(h1)
        h0 -= b_1[:N]
        h1 -= b_1[N:]
        h0 *= 1.0 / tau ** 2
        h1 *= 1.0 / tau ** 2
        if CN:
            h0 = kkt.apply_T_2_inv(h0)
            h1 = kkt.apply_T_1_inv(h1)
        u_p = K_solve(h0)
        u_m = K_solve(h1)
        p00, p01, p10, p11 = pblocks
        c0 = np.zeros((N, n_p))
        c1 = np.zeros((N, n_p))
        for (i, j), A in p00.items():
            if A is not None:
                c0[i] += A @ u_p[j]
        for (i, j), A in p01.items():
            if A is not None:
                c0[i] += A @ u_m[j]
        for (i, j), A in p10.items():
            if A is not None:
                c1[i] += A @ u_p[j]
        for (i, j), A in p11.items():
            if A is not None:
                c1[i] += A @ u_m[j]
        u_1 = np.concatenate([M_solve(c0), M_solve(c1)])
        return u_0, u_1

    return pc_fn


class _PairNullspace:
    """Velocity blocks: Dirichlet; pressure blocks: constant (full_nullspace_0 / _1, 3628-3652)."""

    def __init__(self, bdofs_v):
        self.v = kkt.DirichletBCNullspace(bdofs_v)
        self.p = ConstantNullspace()


def stokes_solve(M_v, K_v, B, M_p, K_p, *, beta, n_t, CN, time_interval=(0.0, 1.0), bdofs_v, b_0, b_1,
                 solver_parameters=None, lambda_v_bounds=None, lambda_p_bounds=None, inner="amg",
                 amg_params=None, amg_params_p=None, pc_fn=None, D_p=None, Multigrid=False):
    """``MultiBlockSystem.solve`` of the outer Stokes system from a zero initial guess
    (control/control.py:4273-4297, 4688-4693).  ``b_0`` (2N, n_v), ``b_1`` (2N, n_p) are the
    final right-hand sides (T transforms already applied).  Returns (u_0, u_1, KSPResult)."""
    t_0, T_f = time_interval
    tau = (T_f - t_0) / (n_t - 1.0)
    N = kkt.n_blocks(n_t, CN)
    n_v, n_p = M_v.shape[0], M_p.shape[0]
    ns = _PairNullspace(bdofs_v)
    if pc_fn is None:
        pc_fn = construct_stokes_pc(M_v, K_v, B, M_p, K_p, tau, beta, n_t, CN, bdofs_v,
                                    lambda_v_bounds=lambda_v_bounds, lambda_p_bounds=lambda_p_bounds,
                                    inner=inner, amg_params=amg_params, amg_params_p=amg_params_p, D_p=D_p,
                                    Multigrid=Multigrid)
    if solver_parameters is None:                                # control/control.py:4291-4297
        solver_parameters = {"linear_solver": "fgmres", "maximum_iterations": 100,
                             "relative_tolerance": 1.0e-6, "absolute_tolerance": 0.0}
    sp_ = solver_parameters
    L0, L1 = 2 * N * n_v, 2 * N * n_p

    def unpack(x):
        return x[:L0].reshape(2 * N, n_v).copy(), x[L0:].reshape(2 * N, n_p).copy()

    def pack(a0, a1):
        return np.concatenate([a0.ravel(), a1.ravel()])

    def A(x):
        x0, x1 = unpack(x)
        return pack(*stokes_apply_fused(M_v, K_v, B, tau, beta, n_t, CN, bdofs_v, x0, x1))

    def P(x):                                                     # Preconditioner.apply, 562-656
        c0, c1 = unpack(x)
        d0 = ns.v.pc_pre_mult_corrected(c0)
        d1 = ns.p.pc_pre_mult_corrected(c1)
        w0, w1 = pc_fn(d0, d1)
        w0, w1 = w0.copy(), w1.copy()
        ns.v.pc_post_mult_correct(w0, c0)
        ns.p.pc_post_mult_correct(w1, c1)
        return pack(w0, w1)

    c0, c1 = b_0.copy(), b_1.copy()
    ns.v.project(c0)
    ns.p.project(c1)
    kw = dict(rtol=sp_["relative_tolerance"], atol=sp_["absolute_tolerance"], max_it=sp_.get("maximum_iterations", 1000))
    ksp_type = sp_.get("linear_solver", "fgmres")
    x, res = krylov.gmres(A, pack(c0, c1), np.zeros(L0 + L1), pc=P, flexible=(ksp_type == "fgmres"),
                          restart=sp_.get("gmres_restart", 30), **kw)
    u0, u1 = unpack(x)
    ns.v.project(u0)
    ns.p.project(u1)
    if not sp_.get("preconditioner", False) and res.reason <= 0:
        raise RuntimeError("Solver failed to converge")
    return u0, u1, res


def incompressible_linear_solve(M_v, K_v, B, M_p, K_p, *, beta, n_t, CN, time_interval=(0.0, 1.0), bdofs_v, v_d, f,
                                v_0=None, div_v=None, div_zeta=None, solver_parameters=None, lambda_v_bounds=None,
                                lambda_p_bounds=None, inner="amg", amg_params=None, amg_params_p=None,
                                check_v_d=True, check_f=True, bc_values=None, D_p=None, Multigrid=False):
    """``Control.Instationary.incompressible_linear_solve`` for homogeneous Dirichlet velocity data
    (control/control.py:3592-4725): right-hand sides (3961-4243; the velocity rows are those of
    the heat problem, the pressure rows are zero unless div_v / div_zeta are given), outer solve,
    unpacking (4705-4725).  ``v_d``, ``f``: (n_t, n_v) cofunction values, or ready (N, n_v) blocks
    when check_v_d / check_f are False (the ``v_d=`` / ``f=`` keywords of the reference).
    ``bc_values`` (n_t, len(bdofs_v)): inhomogeneous time-dependent velocity data, lifted into the
    velocity rows as in the heat problem and into the divergence rows, ``b_1_0[i] -= tau B v_inhom``
    (control/control.py:4107-4119 BE, 4207-4219 CN), and re-imposed on the solution.  Returns
    (v, zeta, p, mu, KSPResult) with v, zeta of n_t levels and p, mu of N levels."""
    t_0, T_f = time_interval
    tau = (T_f - t_0) / (n_t - 1.0)
    N = kkt.n_blocks(n_t, CN)
    n_v, n_p = M_v.shape[0], M_p.shape[0]
    if v_0 is None:
        v_0 = np.zeros(n_v)
    b_0_0, b_0_1 = kkt.build_rhs(M_v, K_v, tau, n_t, CN, bdofs_v, v_d, f, v_0, check_v_d=check_v_d, check_f=check_f,
                                 bc_values=bc_values)
    b_1_0 = np.zeros((N, n_p)) if div_v is None else np.array(div_v, dtype=float)
    if div_v is None and bc_values is not None:
        g = np.zeros((n_t, n_v))
        g[:, bdofs_v] = bc_values
        b_1_0 -= tau * (B @ (g[1:] if CN else g).T).T
    b_1_1 = np.zeros((N, n_p)) if div_zeta is None else np.array(div_zeta, dtype=float)
    if CN:                                                         # 4233-4234
        b_1_0 = kkt.apply_T_2(b_1_0)
        b_1_1 = kkt.apply_T_1(b_1_1)
    u_0, u_1, res = stokes_solve(M_v, K_v, B, M_p, K_p, beta=beta, n_t=n_t, CN=CN, time_interval=time_interval,
                                 bdofs_v=bdofs_v, b_0=np.concatenate([b_0_0, b_0_1]),
                                 b_1=np.concatenate([b_1_0, b_1_1]), solver_parameters=solver_parameters,
                                 lambda_v_bounds=lambda_v_bounds, lambda_p_bounds=lambda_p_bounds, inner=inner,
                                 amg_params=amg_params, amg_params_p=amg_params_p, D_p=D_p, Multigrid=Multigrid)
    if CN:
        v = np.zeros((n_t, n_v))
        zeta = np.zeros((n_t, n_v))
        if check_v_d and check_f:
            v[0] = v_0
        v[1:] = u_0[:N]
        zeta[:-1] = u_0[N:]
    else:
        v, zeta = u_0[:N].copy(), u_0[N:].copy()
    zeta[:, bdofs_v] = 0.0
    if bc_values is not None:
        v[:, bdofs_v] = bc_values
    return v, zeta, u_1[N:].copy(), u_1[:N].copy(), res


def incompressible_non_linear_solve(M_v, D_v, B, M_p, K_p, D_p, *, beta, n_t, CN, time_interval=(0.0, 1.0), bdofs_v,
                                    v_d, f, v_0=None, bc_values=None, solver_parameters=None, lambda_v_bounds=None,
                                    lambda_p_bounds=None, inner="amg", amg_params=None, amg_params_p=None,
                                    max_non_linear_iter=10, relative_non_linear_tol=1e-5,
                                    absolute_non_linear_tol=1e-8):
    """``Control.Instationary.incompressible_non_linear_solve`` (control/control.py:4886-5219): Picard loop of
    Navier-Stokes control.  ``D_v(v_i, t)`` / ``D_p(v_i, t)``: the matrix of ``construct_D_v`` on the velocity /
    pressure space at the velocity state ``v_i`` (1887-1896 with ``v_trial, v_test`` / ``p_trial, p_test``,
    3780-3789).  ``v_d``, ``f``: (n_t, n_v) assembled desired state / force.  Returns a dict with the final
    iterates, the residual-norm history and the inner iteration counts."""
    from .control import non_linear_res_eval
    t_0, T_f = time_interval
    tau = (T_f - t_0) / (n_t - 1.0)
    times = t_0 + tau * np.arange(n_t)
    N = kkt.n_blocks(n_t, CN)
    n_v, n_p = M_v.shape[0], M_p.shape[0]
    v_0 = np.zeros(n_v) if v_0 is None else v_0
    v_old = np.zeros((n_t, n_v))
    zeta_old = np.zeros((n_t, n_v))
    p_old = np.zeros((N, n_p))
    mu_old = np.zeros((N, n_p))
    if CN:
        v_old[0] = v_0                                           # 4968-4969

    def res_eval():                                              # 4979-5072
        r00, r01 = non_linear_res_eval(M_v, D_v, times, tau, beta, n_t, CN, bdofs_v, v_old, zeta_old, v_0, v_d, f)
        r00 -= tau * (B.T @ mu_old.T).T
        r01 -= tau * (B.T @ p_old.T).T
        r00[:, bdofs_v] = 0.0
        r01[:, bdofs_v] = 0.0
        r10 = -(B @ (v_old[1:] if CN else v_old).T).T            # CN: v_old.sub(i + 1), zeta_old.sub(i)
        r11 = -(B @ (zeta_old[:-1] if CN else zeta_old).T).T
        return r00, r01, r10, r11

    def norm(parts):
        return float(np.sqrt(sum((a ** 2).sum() for a in parts)))

    r = res_eval()
    norm_0 = norm(r)
    norm_k, k = norm_0, 0
    history, inner_its = [norm_0], []
    while norm_k > relative_non_linear_tol * norm_0 and norm_k > absolute_non_linear_tol:
        K_levels = [D_v(v_old[i], times[i]) for i in range(n_t)]
        P_levels = [D_p(v_old[i], times[i]) for i in range(n_t)]
        dv, dzeta, dp, dmu, res = incompressible_linear_solve(
            M_v, K_levels, B, M_p, K_p, beta=beta, n_t=n_t, CN=CN, time_interval=time_interval, bdofs_v=bdofs_v,
            v_d=r[0], f=r[1], div_v=tau * r[2], div_zeta=tau * r[3], check_v_d=False, check_f=False,
            solver_parameters=solver_parameters, lambda_v_bounds=lambda_v_bounds, lambda_p_bounds=lambda_p_bounds,
            inner=inner, amg_params=amg_params, amg_params_p=amg_params_p, D_p=P_levels)
        inner_its.append(res.its)
        v_old = v_old + dv                                       # 5127-5160
        if bc_values is not None:
            v_old[:, bdofs_v] = bc_values
        p_old = p_old + dp
        zeta_old = zeta_old + dzeta
        zeta_old[:, bdofs_v] = 0.0
        mu_old = mu_old + dmu
        r = res_eval()
        norm_k = norm(r)
        k += 1
        history.append(norm_k)
        if k + 1 > max_non_linear_iter:
            break
    return dict(v=v_old, zeta=zeta_old, p=p_old, mu=mu_old, history=history, iterations=k, inner_its=inner_its)
