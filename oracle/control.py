"""``MultiBlockSystem.solve`` and ``Control.Instationary.linear_solve`` (oracle; test
infrastructure only), restated on spatial matrices instead of UFL forms.

  * ``system_solve``  preconditioner/preconditioner.py:337-786
  * ``linear_solve``  control/control.py:2820-3375 (numerics only; no output / plots)
  * ``objective``     the discrete objective defined in SURVEY.md section 8c
"""
import numpy as np
import scipy.sparse as sp

from . import kkt, krylov
from .pc import construct_pc, construct_pc_diagonal

DEFAULT_SOLVER_PARAMETERS = {"linear_solver": "gmres",        # control/control.py:3260-3266
                             "gmres_restart": 10,
                             "maximum_iterations": 50,
                             "relative_tolerance": 1.0e-6,
                             "absolute_tolerance": 0.0}


def system_solve(apply_A, nullspace, u_0, u_1, b_0, b_1, *, solver_parameters, pc_fn=None):
    """``MultiBlockSystem.solve``.  ``apply_A(x0, x1) -> (y0, y1)`` is the operator
    (already including the nullspace handling of ``mult``); ``pc_fn(b_0, b_1) ->
    (u_0, u_1)``.  Returns (u_0, u_1, KSPResult)."""
    N, n = b_0.shape
    if pc_fn is None:
        def pc_fn(b_0, b_1):
            return b_0.copy(), b_1.copy()

    def pack(a0, a1):
        return np.concatenate([a0.ravel(), a1.ravel()])

    def unpack(x):
        return x[:N * n].reshape(N, n).copy(), x[N * n:].reshape(N, n).copy()

    def A(x):
        y0, y1 = apply_A(*unpack(x))
        return pack(y0, y1)

    def P(x):
        # Preconditioner.apply: preconditioner/preconditioner.py:562-656
        b0, b1 = unpack(x)
        b0c = nullspace.pc_pre_mult_corrected(b0)
        b1c = nullspace.pc_pre_mult_corrected(b1)
        v0, v1 = pc_fn(b0c, b1c)
        v0 = v0.copy()
        v1 = v1.copy()
        nullspace.pc_post_mult_correct(v0, b0)
        nullspace.pc_post_mult_correct(v1, b1)
        return pack(v0, v1)

    u0 = u_0.copy()
    u1 = u_1.copy()
    nullspace.project(u0)              # correct_soln, 658-678
    nullspace.project(u1)
    c0 = b_0.copy()
    c1 = b_1.copy()
    nullspace.project(c0)              # correct_rhs, 680-704
    nullspace.project(c1)

    sp_ = solver_parameters
    ksp_type = sp_.get("linear_solver", "fgmres")
    kw = dict(rtol=sp_["relative_tolerance"], atol=sp_["absolute_tolerance"],
              max_it=sp_.get("maximum_iterations", 1000))
    if sp_.get("divergence limit") is not None:
        kw["divtol"] = sp_["divergence limit"]
    monitor = sp_.get("monitor")
    if ksp_type in ("gmres", "fgmres"):
        x, res = krylov.gmres(A, pack(c0, c1), pack(u0, u1), pc=P,
                              flexible=(ksp_type == "fgmres"),
                              restart=sp_.get("gmres_restart", 30), monitor=monitor, **kw)
    elif ksp_type == "minres":
        x, res = krylov.minres(A, pack(c0, c1), pack(u0, u1), pc=P, monitor=monitor, **kw)
    else:
        raise ValueError(f"unsupported linear_solver {ksp_type!r}")
    u0, u1 = unpack(x)
    nullspace.project(u0)              # 761-766
    nullspace.project(u1)
    if not sp_.get("preconditioner", False) and res.reason <= 0:
        raise RuntimeError("Solver failed to converge")         # 768-770
    return u0, u1, res


def linear_solve(M, K_levels, *, beta, n_t, CN, time_interval=(0.0, 1.0), bdofs,
                 v_d=None, f=None, v_0=None, check_v_d=True, check_f=True,
                 P=None, solver_parameters=None, Multigrid=False, lambda_v_bounds=None,
                 inner="amg", amg_params=None, pc_mode="triangular", literal=False, bc_values=None):
    """``Instationary.linear_solve`` on assembled matrices.  Returns a dict with the
    unpacked n_t-level ``v``/``zeta``, the raw block solution and the KSP result.
    ``bc_values`` (n_t, len(bdofs)): inhomogeneous Dirichlet data of the state, lifted into the
    right-hand sides and re-imposed on the solution (``set_v``, control/control.py:1836-1845)."""
    t_0, T_f = time_interval
    tau = (T_f - t_0) / (n_t - 1.0)
    n = M.shape[0]
    if v_0 is None:
        v_0 = np.zeros(n)
    if sp.issparse(K_levels):
        K_list = [K_levels] * n_t
    else:
        K_list = list(K_levels)
    nullspace = kkt.DirichletBCNullspace(bdofs)
    b_0, b_1 = kkt.build_rhs(M, K_list, tau, n_t, CN, bdofs, v_d, f, v_0,
                             check_v_d=check_v_d, check_f=check_f, bc_values=bc_values)
    if literal:
        blocks = kkt.build_blocks(M, K_list, tau, beta, n_t, CN)

        def apply_A(x0, x1):
            return kkt.kkt_apply_literal(blocks, nullspace, CN, x0, x1)
    else:
        def apply_A(x0, x1):
            return kkt.kkt_apply_fused(M, K_list, tau, beta, n_t, CN, bdofs, x0, x1)
    if P is None:
        if pc_mode == "triangular":
            pc_fn = construct_pc(M, K_list, tau, beta, n_t, CN, bdofs,
                                 lambda_v_bounds=lambda_v_bounds, Multigrid=Multigrid,
                                 inner=inner, amg_params=amg_params)
        elif pc_mode == "diagonal":
            assert CN
            pc_fn = construct_pc_diagonal(M, K_list[0], tau, beta, n_t, bdofs,
                                          lambda_v_bounds=lambda_v_bounds, inner=inner,
                                          amg_params=amg_params)
        else:
            raise ValueError(pc_mode)
    else:
        pc_fn = P
    if solver_parameters is None:
        solver_parameters = dict(DEFAULT_SOLVER_PARAMETERS)
    N = kkt.n_blocks(n_t, CN)
    v, zeta, res = system_solve(apply_A, nullspace, np.zeros((N, n)), np.zeros((N, n)),
                                b_0, b_1, solver_parameters=solver_parameters, pc_fn=pc_fn)
    y0, y1 = apply_A(v, zeta)
    c0, c1 = b_0.copy(), b_1.copy()
    nullspace.project(c0)
    nullspace.project(c1)
    kkt_res = np.sqrt(np.linalg.norm(c0 - y0) ** 2 + np.linalg.norm(c1 - y1) ** 2)
    v_full, zeta_full = kkt.unpack_solution(v, zeta, n_t, CN,
                                            v_0 if (check_f and check_v_d) else None)
    if bc_values is not None:
        v_full[:, bdofs] = bc_values
    return dict(v=v_full, zeta=zeta_full, v_blocks=v, zeta_blocks=zeta, ksp=res,
                kkt_residual=kkt_res, b_0=b_0, b_1=b_1, tau=tau, pc_fn=pc_fn)


def objective(M, v, zeta, v_hat, tau, beta, CN):
    """J_h = 1/2 sum_i w_i (v_i - vhat_i)^T M (v_i - vhat_i) + 1/(2 beta) sum_i w_i
    zeta_i^T M zeta_i over the n_t levels, trapezoid weights for CN, tau for BE (SURVEY.md
    section 8c; the reference never evaluates J, control/control.py:1876-1885)."""
    n_t = v.shape[0]
    w = np.full(n_t, tau)
    if CN:
        w[0] = w[-1] = 0.5 * tau
    d = v - v_hat
    J = 0.0
    for i in range(n_t):
        J += 0.5 * w[i] * float(d[i] @ (M @ d[i]))
        J += 0.5 / beta * w[i] * float(zeta[i] @ (M @ zeta[i]))
    return J


def non_linear_res_eval(M, D_v, times, tau, beta, n_t, CN, bdofs, v_old, zeta_old, v_0, v_d, f):
    """``Instationary.non_linear_res_eval`` (control/control.py:2442-2818): the right-hand side
    minus the KKT operator applied to the iterate, block row by block row (untransformed rows:
    ``linear_solve`` applies T_1 / T_2 itself), with ``D_v`` evaluated at the iterate.
    ``D_v(v_i, t)`` returns the matrix of ``construct_D_v`` (control/control.py:1887-1896)."""
    n = M.shape[0]
    D = [D_v(v_old[i], times[i]) for i in range(n_t)]
    if CN:
        h = 0.5 * tau
        rhs_0 = np.zeros((n_t - 1, n))
        rhs_1 = np.zeros((n_t - 1, n))
        for i in range(n_t - 1):                                # 2621-2814
            rhs_0[i] = h * (v_d[i] + v_d[i + 1]) - h * (M @ v_old[i]) - h * (M @ v_old[i + 1]) \
                - (h * (D[i].T @ zeta_old[i]) + M @ zeta_old[i]) \
                - (h * (D[i + 1].T @ zeta_old[i + 1]) - M @ zeta_old[i + 1])
            rhs_1[i] = h * (f[i] + f[i + 1]) - (h * (D[i] @ v_old[i]) - M @ v_old[i]) \
                - (h * (D[i + 1] @ v_old[i + 1]) + M @ v_old[i + 1]) \
                + (h / beta) * (M @ zeta_old[i]) + (h / beta) * (M @ zeta_old[i + 1])
    else:
        rhs_0 = np.zeros((n_t, n))
        rhs_1 = np.zeros((n_t, n))
        D_v_0 = D_v(v_0, times[0])
        for i in range(n_t):                                    # 2457-2620
            if i < n_t - 1:
                rhs_0[i] = tau * v_d[i] - tau * (M @ v_old[i]) \
                    - (tau * (D[i].T @ zeta_old[i]) + M @ zeta_old[i]) + M @ zeta_old[i + 1]
            else:
                rhs_0[i] = -(tau * (D[i].T @ zeta_old[i]) + M @ zeta_old[i])
            if i == 0:
                rhs_1[0] = (tau * (D_v_0 @ v_0) + M @ v_0) - (tau * (D[0] @ v_old[0]) + M @ v_old[0])
            else:
                rhs_1[i] = tau * f[i] - (tau * (D[i] @ v_old[i]) + M @ v_old[i]) + M @ v_old[i - 1] \
                    + (tau / beta) * (M @ zeta_old[i])
    rhs_0[:, bdofs] = 0.0
    rhs_1[:, bdofs] = 0.0
    return rhs_0, rhs_1


def non_linear_solve(M, D_v, *, beta, n_t, CN, time_interval=(0.0, 1.0), bdofs, v_d, f, v_0=None,
                     solver_parameters=None, lambda_v_bounds=None, inner="amg", amg_params=None,
                     max_non_linear_iter=10, relative_non_linear_tol=1e-5, absolute_non_linear_tol=1e-8,
                     bc_values=None):
    """``Instationary.non_linear_solve`` (control/control.py:3377-3590).  Returns a dict with
    the final iterate, the residual-norm history and the inner iteration counts.  ``bc_values``
    (n_t, len(bdofs)): inhomogeneous Dirichlet data, re-imposed on the iterate after every update
    (3480-3483)."""
    t_0, T_f = time_interval
    tau = (T_f - t_0) / (n_t - 1.0)
    times = t_0 + tau * np.arange(n_t)
    n = M.shape[0]
    v_0 = np.zeros(n) if v_0 is None else v_0
    v_old = np.zeros((n_t, n))
    zeta_old = np.zeros((n_t, n))
    if CN:
        v_old[0] = v_0
    rhs_0, rhs_1 = non_linear_res_eval(M, D_v, times, tau, beta, n_t, CN, bdofs, v_old, zeta_old, v_0, v_d, f)
    norm_0 = float(np.sqrt((rhs_0 ** 2).sum() + (rhs_1 ** 2).sum()))
    norm_k = norm_0
    history = [norm_0]
    inner_its = []
    k = 0
    while norm_k > relative_non_linear_tol * norm_0 and norm_k > absolute_non_linear_tol:
        K_levels = [D_v(v_old[i], times[i]) for i in range(n_t)]
        out = linear_solve(M, K_levels, beta=beta, n_t=n_t, CN=CN, time_interval=time_interval, bdofs=bdofs,
                           v_d=rhs_0, f=rhs_1, check_v_d=False, check_f=False,
                           solver_parameters=solver_parameters, lambda_v_bounds=lambda_v_bounds,
                           inner=inner, amg_params=amg_params)
        inner_its.append(out["ksp"].its)
        v_old = v_old + out["v"]
        if bc_values is not None:
            v_old[:, bdofs] = bc_values
        zeta_old = zeta_old + out["zeta"]
        zeta_old[:, bdofs] = 0.0
        rhs_0, rhs_1 = non_linear_res_eval(M, D_v, times, tau, beta, n_t, CN, bdofs, v_old, zeta_old, v_0, v_d, f)
        norm_k = float(np.sqrt((rhs_0 ** 2).sum() + (rhs_1 ** 2).sum()))
        k += 1
        history.append(norm_k)
        if k + 1 > max_non_linear_iter:
            break
    return dict(v=v_old, zeta=zeta_old, history=history, iterations=k, inner_its=inner_its)
