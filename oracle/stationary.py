"""``Control.Stationary`` heat-type drivers (oracle; test infrastructure only), restated on assembled
matrices: SURVEY.md section 8f rank 4 ("Stationary problems as the N = 1 case").

  * ``construct_pc``       control/control.py:351-450
  * ``linear_solve``       control/control.py:489-628
  * ``non_linear_res_eval``control/control.py:452-487
  * ``non_linear_solve``   control/control.py:630-800

The block system is written down LITERALLY, [[M, D_v^T], [D_v, -M/beta]] with one block per
dict (control/control.py:549-560), and solved through the same restatement of
``MultiBlockSystem.solve`` as the instationary path.  It deliberately does NOT go through the
trapezoidal block tables: the product maps the stationary system onto its instationary handle
(n_t = 2, tau = 2, K' = K - M), and the tests compare that mapping with this direct statement.

Pinned by the reference's ``test_stationary_linear_control`` (test/test_control.py:26-119, analytic
solution to 1e-13) run through this driver with the test's own solver parameters.  Preconditioner
outputs and iteration counts of the real stack: parity unpinned (BoomerAMG -> ``oracle/amg.py``).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import amg as _amg
from . import kkt
from .control import system_solve
from .pc import make_solver_0
from synthetic.fem import assemble_bc

DEFAULT_SOLVER_PARAMETERS = {"linear_solver": "gmres",        # control/control.py:562-568
                             "gmres_restart": 10,
                             "maximum_iterations": 50,
                             "relative_tolerance": 1.0e-6,
                             "absolute_tolerance": 0.0}


def _inner(A, kind, amg_params):
    if kind == "exact":
        return spla.splu(sp.csc_matrix(A)).solve
    H = _amg.setup(A, **(amg_params or {}))
    return lambda b: _amg.solve(H, b)


def construct_pc(M, D_v, beta, bdofs, *, lambda_v_bounds=None, Multigrid=False, inner="amg", amg_params=None):
    """``Stationary.construct_pc`` -> ``pc_linear(b_0, b_1) -> (u_0, u_1)`` on arrays (1, n)."""
    solver_0 = make_solver_0(M, bdofs, lambda_v_bounds, Multigrid, amg_params)          # 357-393
    c = 1.0 / beta ** 0.5
    solver_1 = _inner(assemble_bc((D_v + c * M).tocsr(), bdofs), inner, amg_params)      # 395-405
    solver_2 = _inner(assemble_bc((D_v.T + c * M).tocsr(), bdofs), inner, amg_params)    # 407-417

    def pc_linear(b_0, b_1):
        u_0 = solver_0(b_0)                                     # 420-423
        b = (D_v @ u_0.T).T - b_1                               # 425-429
        b[..., bdofs] = 0.0                                     # 433-434
        u_1 = np.stack([solver_1(r) for r in b])                # 435-438
        b = (M @ u_1.T).T                                       # 442
        b[..., bdofs] = 0.0
        u_1 = np.stack([solver_2(r) for r in b])                # 445-448
        return u_0, u_1
    return pc_linear


def apply_A(M, D_v, beta, nullspace, x0, x1):
    """``MultiBlockSystemMatrix.mult`` for the one-block dicts of 549-560 (no T transform: CN unset)."""
    blocks = ({(0, 0): M}, {(0, 0): D_v.T.tocsr()}, {(0, 0): D_v}, {(0, 0): (-(1.0 / beta) * M).tocsr()})
    return kkt.kkt_apply_literal(blocks, nullspace, False, x0, x1)


def linear_solve(M, D_v, *, beta, bdofs, v_d, f, check_v_d=True, check_f=True, bc_values=None, P=None,
                 solver_parameters=None, Multigrid=False, lambda_v_bounds=None, inner="amg", amg_params=None):
    """``Stationary.linear_solve``.  ``v_d``, ``f``: cofunction values (n,) -- with check_*=True the
    assembled desired state / force (``M @ nodal``), lifted here when ``bc_values`` (values of the state at
    ``bdofs``) is given (326-349); with check_*=False ready right-hand sides, used as they are."""
    n = M.shape[0]
    bdofs = np.asarray(bdofs, dtype=np.int64)
    nullspace = kkt.DirichletBCNullspace(bdofs)
    v_inhom = None
    if bc_values is not None:                                   # 520-526
        v_inhom = np.zeros(n)
        v_inhom[bdofs] = bc_values
    b_0 = np.array(v_d, dtype=float)
    b_1 = np.array(f, dtype=float)
    if v_inhom is not None and check_f:                         # construct_f, 326-336
        b_1 -= D_v @ v_inhom
        b_1[bdofs] = 0.0
    if v_inhom is not None and check_v_d:                       # construct_v_d, 338-349
        b_0 -= M @ v_inhom
        b_0[bdofs] = 0.0
    pc_fn = P if P is not None else construct_pc(M, D_v, beta, bdofs, lambda_v_bounds=lambda_v_bounds,
                                                 Multigrid=Multigrid, inner=inner, amg_params=amg_params)
    if solver_parameters is None:
        solver_parameters = dict(DEFAULT_SOLVER_PARAMETERS)
    v, zeta, res = system_solve(lambda x0, x1: apply_A(M, D_v, beta, nullspace, x0, x1), nullspace,
                                np.zeros((1, n)), np.zeros((1, n)), b_0[None], b_1[None],
                                solver_parameters=solver_parameters, pc_fn=pc_fn)
    v, zeta = v[0], zeta[0]
    if v_inhom is not None:                                     # 586-589
        v = v + v_inhom
    if bc_values is not None:                                   # set_v / set_zeta re-apply the bcs, 264-283
        v[bdofs] = bc_values
    else:
        v[bdofs] = 0.0
    zeta[bdofs] = 0.0
    return dict(v=v, zeta=zeta, ksp=res, b_0=b_0, b_1=b_1, pc_fn=pc_fn)


def non_linear_res_eval(M, D_v, beta, bdofs, v_d, f, v_old, zeta_old):
    """452-487: rhs_0 = v_d - M v - D_v^T zeta, rhs_1 = f - D_v v + (1/beta) M zeta, bcs applied."""
    rhs_0 = v_d - M @ v_old - D_v.T @ zeta_old
    rhs_1 = f - D_v @ v_old + (1.0 / beta) * (M @ zeta_old)
    rhs_0[bdofs] = 0.0
    rhs_1[bdofs] = 0.0
    return rhs_0, rhs_1


def non_linear_solve(M, D_v_of, *, beta, bdofs, v_d, f, v_init=None, zeta_init=None, bc_values=None,
                     solver_parameters=None, lambda_v_bounds=None, inner="amg", amg_params=None,
                     max_non_linear_iter=10, relative_non_linear_tol=1e-5, absolute_non_linear_tol=1e-8):
    """``Stationary.non_linear_solve`` (630-800).  ``D_v_of(v)`` -> the matrix of ``construct_D_v`` at the
    state ``v`` (314-324).  The increment solves use homogeneous data; inhomogeneous values are re-imposed
    on the iterate (690-693)."""
    n = M.shape[0]
    bdofs = np.asarray(bdofs, dtype=np.int64)
    v_old = np.zeros(n) if v_init is None else np.array(v_init, dtype=float)
    zeta_old = np.zeros(n) if zeta_init is None else np.array(zeta_init, dtype=float)
    D_v = D_v_of(v_old)
    rhs_0, rhs_1 = non_linear_res_eval(M, D_v, beta, bdofs, v_d, f, v_old, zeta_old)
    norm_0 = float(np.sqrt(rhs_0 @ rhs_0 + rhs_1 @ rhs_1))
    norm_k, k = norm_0, 0
    history, inner_its = [norm_0], []
    while norm_k > relative_non_linear_tol * norm_0 and norm_k > absolute_non_linear_tol:
        # linear_solve is called with ready right-hand sides (v_d=rhs_0, f=rhs_1) but still adds v_inhom
        # to its solution when the data are inhomogeneous (586-589) -- restated as is
        out = linear_solve(M, D_v, beta=beta, bdofs=bdofs, v_d=rhs_0, f=rhs_1, check_v_d=False, check_f=False,
                           bc_values=bc_values, solver_parameters=solver_parameters, lambda_v_bounds=lambda_v_bounds,
                           inner=inner, amg_params=amg_params)
        inner_its.append(out["ksp"].its)
        v_old = v_old + out["v"]
        if bc_values is not None:
            v_old[bdofs] = bc_values
        zeta_old = zeta_old + out["zeta"]
        zeta_old[bdofs] = 0.0
        D_v = D_v_of(v_old)
        rhs_0, rhs_1 = non_linear_res_eval(M, D_v, beta, bdofs, v_d, f, v_old, zeta_old)
        norm_k = float(np.sqrt(rhs_0 @ rhs_0 + rhs_1 @ rhs_1))
        k += 1
        history.append(norm_k)
        if k + 1 > max_non_linear_iter:
            break
    return dict(v=v_old, zeta=zeta_old, history=history, iterations=k, inner_its=inner_its)


# --------------------------------------------------------------------------------------------------------------
# Stationary Stokes control: ``Control.Stationary.incompressible_linear_solve`` (control/control.py:802-1201)
# --------------------------------------------------------------------------------------------------------------
def stokes_apply(M_v, D_v, B, beta, ns_v, ns_p, x0, x1):
    """``mult`` of the outer system of 885-925: block_00 = [[M_v, D_v^T], [D_v, -M_v/beta]], block_01 =
    diag(B^T), block_10 = diag(B), no time scaling -- the literal operator of oracle/stokes.py with N = 1,
    tau = 1 (the form pinned by the reference's stationary Stokes known-answer test)."""
    from .stokes import stokes_apply_literal
    heat_blocks = ({(0, 0): M_v}, {(0, 0): D_v.T.tocsr()}, {(0, 0): D_v}, {(0, 0): (-(1.0 / beta) * M_v).tocsr()})
    return stokes_apply_literal(heat_blocks, B, 1.0, 1, False, ns_v, ns_p, x0, x1)


def construct_stokes_pc(M_v, D_v, B, M_p, K_p, D_p, beta, bdofs_v, *, lambda_v_bounds=None, lambda_p_bounds=None,
                        inner="amg", amg_params=None, amg_params_p=None):
    """``pc_fn`` of 986-1110: five GMRES iterations on the velocity KKT block with ``Stationary.construct_pc``,
    ``solver_K_p`` on ``B u_0 - b_1``, the pressure-space KKT multiply [[M_p, D_p^T], [D_p, -M_p/beta]]
    (block_*_p, 968-984; ``D_p`` = the forward form on the pressure space), ``solver_M_p``."""
    from .stokes import make_solver_p
    ns_v = kkt.DirichletBCNullspace(bdofs_v)
    vparams = dict(cycles=6)                   # as oracle/stokes.py: six cycles on the vector operator
    vparams.update(amg_params or {})
    inner_pc = construct_pc(M_v, D_v, beta, bdofs_v, lambda_v_bounds=lambda_v_bounds, inner=inner, amg_params=vparams)
    K_solve, M_solve, _ = make_solver_p(M_p, K_p, lambda_p_bounds, amg_params_p)
    inner_parameters = {"preconditioner": True, "linear_solver": "gmres", "maximum_iterations": 5,
                        "relative_tolerance": 0.0, "absolute_tolerance": 0.0}          # 1000-1005
    n_v = M_v.shape[0]

    def pc_fn(b_0, b_1):
        v, zeta, _ = system_solve(lambda x0, x1: apply_A(M_v, D_v, beta, ns_v, x0, x1), ns_v, np.zeros((1, n_v)),
                                  np.zeros((1, n_v)), b_0[:1], b_0[1:], solver_parameters=inner_parameters,
                                  pc_fn=inner_pc)
        u_0 = np.concatenate([v, zeta])
        h = (B @ u_0.T).T - b_1                                 # 1023-1034
        w = K_solve(h)                                          # 1042-1052
        c0 = M_p @ w[0] + D_p.T @ w[1]                          # 1064-1069
        c1 = D_p @ w[0] - (1.0 / beta) * (M_p @ w[1])
        return u_0, M_solve(np.stack([c0, c1]))                 # 1074-1084
    return pc_fn


def incompressible_linear_solve(M_v, D_v, B, M_p, K_p, *, beta, bdofs_v, v_d, f, div_v=None, div_zeta=None, D_p=None,
                                check_v_d=True, check_f=True, bc_values=None, solver_parameters=None,
                                lambda_v_bounds=None, lambda_p_bounds=None, inner="amg", amg_params=None,
                                amg_params_p=None):
    """``Stationary.incompressible_linear_solve``.  Returns (v, zeta, p, mu, KSPResult)."""
    from . import krylov
    from .stokes import ConstantNullspace
    n_v, n_p = M_v.shape[0], M_p.shape[0]
    bdofs_v = np.asarray(bdofs_v, dtype=np.int64)
    ns_v, ns_p = kkt.DirichletBCNullspace(bdofs_v), ConstantNullspace()
    D_p = K_p if D_p is None else D_p
    v_inhom = None
    if bc_values is not None:
        v_inhom = np.zeros(n_v)
        v_inhom[bdofs_v] = bc_values
    b00, b01 = np.array(v_d, dtype=float), np.array(f, dtype=float)
    if v_inhom is not None and check_f:                         # construct_f, 326-336
        b01 -= D_v @ v_inhom
        b01[bdofs_v] = 0.0
    if v_inhom is not None and check_v_d:                       # construct_v_d, 338-349
        b00 -= M_v @ v_inhom
        b00[bdofs_v] = 0.0
    if div_v is None:                                           # 866-873
        b10 = np.zeros(n_p) if v_inhom is None else -(B @ v_inhom)
    else:
        b10 = np.array(div_v, dtype=float)
    b11 = np.zeros(n_p) if div_zeta is None else np.array(div_zeta, dtype=float)
    pc_fn = construct_stokes_pc(M_v, D_v, B, M_p, K_p, D_p, beta, bdofs_v, lambda_v_bounds=lambda_v_bounds,
                                lambda_p_bounds=lambda_p_bounds, inner=inner, amg_params=amg_params,
                                amg_params_p=amg_params_p)
    if solver_parameters is None:                               # 1088-1094
        solver_parameters = {"linear_solver": "fgmres", "maximum_iterations": 50, "relative_tolerance": 1.0e-6,
                             "absolute_tolerance": 0.0}
    sp_ = solver_parameters
    L0 = 2 * n_v

    def unpack(x):
        return x[:L0].reshape(2, n_v).copy(), x[L0:].reshape(2, n_p).copy()

    def pack(a0, a1):
        return np.concatenate([a0.ravel(), a1.ravel()])

    def A(x):
        return pack(*stokes_apply(M_v, D_v, B, beta, ns_v, ns_p, *unpack(x)))

    def P(x):                                                   # Preconditioner.apply, preconditioner.py:562-656
        c0, c1 = unpack(x)
        w0, w1 = pc_fn(ns_v.pc_pre_mult_corrected(c0), ns_p.pc_pre_mult_corrected(c1))
        w0, w1 = w0.copy(), w1.copy()
        ns_v.pc_post_mult_correct(w0, c0)
        ns_p.pc_post_mult_correct(w1, c1)
        return pack(w0, w1)

    c0, c1 = np.stack([b00, b01]), np.stack([b10, b11])
    ns_v.project(c0)
    ns_p.project(c1)
    x, res = krylov.gmres(A, pack(c0, c1), np.zeros(L0 + 2 * n_p), pc=P,
                          flexible=(sp_.get("linear_solver", "fgmres") == "fgmres"), restart=sp_.get("gmres_restart", 30),
                          rtol=sp_["relative_tolerance"], atol=sp_["absolute_tolerance"],
                          max_it=sp_.get("maximum_iterations", 1000))
    u0, u1 = unpack(x)
    ns_v.project(u0)
    ns_p.project(u1)
    if not sp_.get("preconditioner", False) and res.reason <= 0:
        raise RuntimeError("Solver failed to converge")
    v, zeta = u0[0], u0[1]
    if v_inhom is not None:                                     # 1107-1110
        v = v + v_inhom
    v[bdofs_v] = 0.0 if bc_values is None else bc_values        # set_v / set_zeta, 264-283
    zeta[bdofs_v] = 0.0
    return v, zeta, u1[1].copy(), u1[0].copy(), res


def incompressible_non_linear_solve(M_v, D_v_of, B, M_p, K_p, D_p_of, *, beta, bdofs_v, v_d, f, bc_values=None,
                                    solver_parameters=None, lambda_v_bounds=None, lambda_p_bounds=None, inner="amg",
                                    amg_params=None, amg_params_p=None, max_non_linear_iter=10,
                                    relative_non_linear_tol=1e-5, absolute_non_linear_tol=1e-8):
    """``Stationary.incompressible_non_linear_solve`` (control/control.py:1203-1486): Picard loop of stationary
    Navier-Stokes control.  ``D_v_of(v)`` / ``D_p_of(v)``: ``construct_D_v`` on the velocity / pressure space at
    the velocity state ``v``."""
    n_v, n_p = M_v.shape[0], M_p.shape[0]
    bdofs_v = np.asarray(bdofs_v, dtype=np.int64)
    v_old, zeta_old = np.zeros(n_v), np.zeros(n_v)
    p_old, mu_old = np.zeros(n_p), np.zeros(n_p)

    def res_eval(D_v):                                          # 1272-1319
        r00, r01 = non_linear_res_eval(M_v, D_v, beta, bdofs_v, v_d, f, v_old, zeta_old)
        r00 = r00 - B.T @ mu_old
        r01 = r01 - B.T @ p_old
        r00[bdofs_v] = 0.0
        r01[bdofs_v] = 0.0
        return r00, r01, -(B @ v_old), -(B @ zeta_old)

    def norm(parts):
        return float(np.sqrt(sum(a @ a for a in parts)))

    D_v = D_v_of(v_old)
    r = res_eval(D_v)
    norm_0 = norm(r)
    norm_k, k = norm_0, 0
    history, inner_its = [norm_0], []
    while norm_k > relative_non_linear_tol * norm_0 and norm_k > absolute_non_linear_tol:
        dv, dzeta, dp, dmu, res = incompressible_linear_solve(
            M_v, D_v, B, M_p, K_p, beta=beta, bdofs_v=bdofs_v, v_d=r[0], f=r[1], div_v=r[2], div_zeta=r[3],
            D_p=D_p_of(v_old), check_v_d=False, check_f=False, bc_values=bc_values, solver_parameters=solver_parameters,
            lambda_v_bounds=lambda_v_bounds, lambda_p_bounds=lambda_p_bounds, inner=inner, amg_params=amg_params,
            amg_params_p=amg_params_p)
        inner_its.append(res.its)
        v_old = v_old + dv
        if bc_values is not None:
            v_old[bdofs_v] = bc_values
        zeta_old = zeta_old + dzeta
        zeta_old[bdofs_v] = 0.0
        p_old = p_old + dp
        mu_old = mu_old + dmu
        D_v = D_v_of(v_old)
        r = res_eval(D_v)
        norm_k = norm(r)
        k += 1
        history.append(norm_k)
        if k + 1 > max_non_linear_iter:
            break
    return dict(v=v_old, zeta=zeta_old, p=p_old, mu=mu_old, history=history, iterations=k, inner_its=inner_its)
