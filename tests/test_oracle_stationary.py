"""CPU tests of the ``Control.Stationary`` restatement (oracle/stationary.py) and of the mapping the product
uses to run stationary problems on its instationary device handle (control_b200/stationary.py)."""
import numpy as np
import pytest

import kat
from oracle import kkt, stationary
from oracle.pc import construct_pc
from synthetic import fem


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("inner", ["exact", "amg"])
def test_reference_stationary_known_answer_through_the_driver(inner):
    """test/test_control.py:26-119 with the test's own call: Q2 on 8x8 quads, no boundary conditions, ready
    right-hand sides, Chebyshev bounds (0.25, 1.5625), FGMRES to 1e-14; the reference asserts 1e-13."""
    M, L, coords, _ = fem.assemble_q2_2d(8, 8)
    K = (L + M).tocsr()
    beta = 1e-3
    X0, X1 = coords[:, 0], coords[:, 1]
    v_ref = X0 * np.exp(X1)
    zeta_ref = np.sin(np.pi * X0) * np.sin(2.0 * np.pi * X1)
    b_0 = M @ v_ref + K @ zeta_ref
    b_1 = K @ v_ref - (1.0 / beta) * (M @ zeta_ref)
    sp_ = {"linear_solver": "fgmres", "fgmres_restart": 10, "maximum_iterations": 500, "relative_tolerance": 1e-14,
           "absolute_tolerance": 1e-14}
    r = stationary.linear_solve(M, K, beta=beta, bdofs=[], v_d=b_0, f=b_1, check_v_d=False, check_f=False,
                                lambda_v_bounds=(0.25, 1.5625), solver_parameters=sp_, inner=inner)
    assert r["ksp"].reason > 0
    assert kat.l2_error(M, r["v"][None], v_ref[None]) < 1e-13
    assert kat.l2_error(M, r["zeta"][None], zeta_ref[None]) < 1e-13


@pytest.mark.parametrize("symmetric", [True, False])
def test_stationary_system_is_the_one_block_trapezoidal_system(symmetric):
    """The mapping of control_b200/stationary.py: [[M, D_v^T], [D_v, -M/beta]] and ``Stationary.construct_pc``
    (control/control.py:351-450) equal the trapezoidal operator and preconditioner (control/control.py:1995-2189,
    2929-2958) with n_t = 2, tau = 2 and the forward matrix D_v - M.  Compared with the Dirichlet nullspace,
    for symmetric and non-symmetric D_v, operator and preconditioner, exact and AMG inner solves."""
    M, L, coords, bd = fem.assemble_p1_2d(9, 7, 1.0, 1.0)
    D_v = (L + 2.0 * M).tocsr()
    if not symmetric:
        D_v = D_v.copy()
        D_v.data = D_v.data * (1.0 + 0.2 * np.random.default_rng(3).standard_normal(D_v.nnz))
    beta = 1e-2
    n = M.shape[0]
    Kp = (D_v - M).tocsr()
    rng = np.random.default_rng(0)
    x0, x1 = rng.standard_normal((1, n)), rng.standard_normal((1, n))
    ns = kkt.DirichletBCNullspace(bd)
    y0, y1 = stationary.apply_A(M, D_v, beta, ns, x0, x1)
    z0, z1 = kkt.kkt_apply_fused(M, Kp, 2.0, beta, 2, True, bd, x0, x1)
    assert _rel(z0, y0) < 1e-14 and _rel(z1, y1) < 1e-14
    for inner, tol in (("exact", 1e-12), ("amg", 1e-12)):
        b0, b1 = rng.standard_normal((1, n)), rng.standard_normal((1, n))
        b0[:, bd] = 0.0
        b1[:, bd] = 0.0
        u0, u1 = stationary.construct_pc(M, D_v, beta, bd, lambda_v_bounds=(0.5, 2.0), inner=inner)(b0, b1)
        w0, w1 = construct_pc(M, Kp, 2.0, beta, 2, True, bd, lambda_v_bounds=(0.5, 2.0), inner=inner)(b0, b1)
        assert _rel(w0, u0) < tol and _rel(w1, u1) < tol


def test_reference_mms_stationary_poisson_control_convergence():
    """test/test_control.py:122-229 (degree 1): Poisson control on the unit square, zero Dirichlet data,
    beta = 1e-3, manufactured v = sin sin exp(x+y), zeta = sin(2 pi x) sin(2 pi y), the test's FGMRES
    parameters and DEFAULT preconditioner (Jacobi on M: no Chebyshev bounds given).  The reference prints the
    observed orders; here second order is asserted."""
    beta = 1e-3
    errs = []
    for N in (8, 16, 32):
        M, L, coords, bd = fem.assemble_p1_2d(N, N, 1.0, 1.0)
        x, y = coords[:, 0], coords[:, 1]
        v = np.sin(np.pi * x) * np.sin(np.pi * y) * np.exp(x + y)
        zeta = np.sin(2.0 * np.pi * x) * np.sin(2.0 * np.pi * y)
        # -div grad of the manufactured fields (the interpolated expressions of 140-163)
        lap_zeta = -8.0 * np.pi ** 2 * zeta
        sx, cx, sy, cy = np.sin(np.pi * x), np.cos(np.pi * x), np.sin(np.pi * y), np.cos(np.pi * y)
        e = np.exp(x + y)
        lap_v = e * (2.0 * (1.0 - np.pi ** 2) * sx * sy + 2.0 * np.pi * (cx * sy + sx * cy))
        v_hat = -lap_zeta + v
        f_nodal = -lap_v - zeta / beta
        sp_ = {"linear_solver": "fgmres", "fgmres_restart": 10, "maximum_iterations": 500, "relative_tolerance": 1e-6,
               "absolute_tolerance": 1e-6}
        r = stationary.linear_solve(M, L, beta=beta, bdofs=bd, v_d=M @ v_hat, f=M @ f_nodal, solver_parameters=sp_)
        assert r["ksp"].reason > 0
        errs.append((kat.l2_error(M, r["v"][None], v[None]), kat.l2_error(M, r["zeta"][None], zeta[None])))
    e = np.array(errs)
    orders = np.log(e[:-1] / e[1:]) / np.log(2.0)
    assert (orders[0] > 1.6).all() and (orders[1] > 1.85).all(), orders      # measured 1.68 / 1.78 and 1.92 / 1.95


def test_stationary_inhomogeneous_lifting_against_direct_solve():
    """Lifting of inhomogeneous Dirichlet data (control/control.py:326-349, 520-526, 586-589): the driver's
    solution equals a direct solve of the un-eliminated system with the boundary rows replaced."""
    M, L, coords, bd = fem.assemble_p1_2d(7, 6, 1.0, 1.0)
    D_v = (L + M).tocsr()
    beta = 1e-2
    n = M.shape[0]
    rng = np.random.default_rng(5)
    g = rng.standard_normal(bd.size)
    v_d, f = M @ rng.standard_normal(n), M @ rng.standard_normal(n)
    sp_ = {"linear_solver": "fgmres", "maximum_iterations": 200, "relative_tolerance": 1e-13, "absolute_tolerance": 0.0,
           "gmres_restart": 200}
    r = stationary.linear_solve(M, D_v, beta=beta, bdofs=bd, v_d=v_d, f=f, bc_values=g, solver_parameters=sp_,
                                lambda_v_bounds=(0.5, 2.0), inner="exact")
    A = np.block([[M.toarray(), D_v.T.toarray()], [D_v.toarray(), -M.toarray() / beta]])
    rhs = np.concatenate([v_d, f])
    for d in bd:                                # v = g and zeta = 0 on the boundary, rows replaced
        for row, val in ((d, g[list(bd).index(d)]), (n + d, 0.0)):
            A[row, :] = 0.0
            A[row, row] = 1.0
            rhs[row] = val
    # the adjoint row of a constrained state dof and the state row of a constrained adjoint dof are dropped, the
    # remaining rows keep their couplings to the known boundary values -- what the lifting encodes
    x = np.linalg.solve(A, rhs)
    assert _rel(r["v"], x[:n]) < 1e-10 and _rel(r["zeta"], x[n:]) < 1e-10


@pytest.mark.parametrize("gauss_newton", [False, True])
def test_stationary_non_linear_loop_converges_and_residual_is_consistent(gauss_newton):
    """``Stationary.non_linear_solve`` (control/control.py:630-800) on the non-linear diffusion operator
    (1 + v^2) grad.grad: the residual history decreases to the reference's default tolerances, and at the
    fixed point the KKT equations with D_v evaluated at the solution hold."""
    nx = 8
    M, L, coords, bd = fem.assemble_p1_2d(nx, nx, 1.0, 1.0)
    D = fem.nonlinear_diffusion_p1_2d(nx, nx, 1.0, 1.0)
    x, y = coords[:, 0], coords[:, 1]
    v_hat = np.sin(np.pi * x) * np.sin(np.pi * y) * np.exp(x + y)
    beta = 1e-2
    sp_ = {"linear_solver": "fgmres", "maximum_iterations": 200, "relative_tolerance": 1e-10, "absolute_tolerance": 0.0}
    out = stationary.non_linear_solve(M, lambda v: D(v, gauss_newton), beta=beta, bdofs=bd, v_d=M @ v_hat,
                                      f=np.zeros(M.shape[0]), solver_parameters=sp_, lambda_v_bounds=(0.5, 2.0),
                                      max_non_linear_iter=30)
    h = out["history"]
    assert h[-1] <= 1e-5 * h[0] and out["iterations"] < 30
    r0, r1 = stationary.non_linear_res_eval(M, D(out["v"], gauss_newton), beta, bd, M @ v_hat, np.zeros(M.shape[0]),
                                            out["v"], out["zeta"])
    assert np.sqrt(r0 @ r0 + r1 @ r1) == pytest.approx(h[-1], rel=1e-12)
    assert np.abs(out["v"]).max() > 0.05          # the state is far enough from zero for the non-linearity to act


def test_stationary_kkt_solution_minimises_the_reduced_functional():
    """Formulation check in the spirit of test/test_control.py:554-707 (which compares with an L-BFGS minimisation
    of the reduced functional through tlm_adjoint): the solution (v, zeta) of the block system, with the control
    u = zeta / beta (test/test_control.py:625), is the minimiser of
        J(u) = 1/2 (v(u) - v_hat)^T M (v(u) - v_hat) + beta/2 u^T M u,   K v(u) = f + M u  (interior rows, v = 0 on
    the boundary) -- its gradient vanishes and J grows in every direction.  Same operator as the reference's test:
    grad.grad + 2 mass, beta = 1, zero Dirichlet data, zero force."""
    import scipy.sparse.linalg as spla
    M, L, coords, bd = fem.assemble_p1_2d(8, 8, 1.0, 1.0)
    K = (L + 2.0 * M).tocsr()
    n = M.shape[0]
    inter = np.setdiff1d(np.arange(n), bd)
    x, y = coords[:, 0], coords[:, 1]
    v_hat = np.sin(np.pi * x) * np.sin(np.pi * y) * np.exp(x + y)
    beta = 1.0
    sp_ = {"linear_solver": "fgmres", "maximum_iterations": 500, "relative_tolerance": 1e-14, "absolute_tolerance": 1e-14}
    r = stationary.linear_solve(M, K, beta=beta, bdofs=bd, v_d=M @ v_hat, f=np.zeros(n), solver_parameters=sp_,
                                inner="exact")
    lu = spla.splu(K[inter][:, inter].tocsc())
    M_II = M[inter][:, inter]

    def state(u_I):
        v = np.zeros(n)
        v[inter] = lu.solve(M_II @ u_I)
        return v

    def J(u_I):
        d = state(u_I) - v_hat
        return 0.5 * d @ (M @ d) + 0.5 * beta * u_I @ (M_II @ u_I)
    u_I = r["zeta"][inter] / beta
    assert np.abs(state(u_I) - r["v"]).max() < 1e-11 * np.abs(r["v"]).max()      # the state equation row
    # gradient: M_II K_II^-T (M (v - v_hat))_I + beta M_II u
    d = state(u_I) - v_hat
    g = M_II @ lu.solve((M @ d)[inter], trans="T") + beta * (M_II @ u_I)
    assert np.abs(g).max() < 1e-11 * np.abs(beta * (M_II @ u_I)).max()
    rng = np.random.default_rng(0)
    J0 = J(u_I)
    for _ in range(5):
        assert J(u_I + 1e-3 * rng.standard_normal(u_I.size)) > J0


def test_reference_mms_stationary_stokes_control_convergence():
    """test/test_control.py:361-551 (degree 2; Q2 - Q1 here): manufactured stationary Stokes control problem with
    inhomogeneous Dirichlet velocity data, beta = 1e-3, the test's solver parameters and Chebyshev bounds.  The
    reference prints the observed orders; asserted here: at least third order for the velocity and its adjoint,
    second order for the pressures.  Exercises the lifting of the velocity data into all four right-hand sides
    (control/control.py:326-349, 866-873)."""
    errs = []
    for N in (2, 4, 8):
        m = kat.stokes_mms_fields(N)
        sq, M, bd = m["sq"], m["M"], m["sq"]["bdofs_v"]
        sp_ = {"linear_solver": "fgmres", "fgmres_restart": 10, "maximum_iterations": 200, "relative_tolerance": 1e-10,
               "absolute_tolerance": 1e-10}
        v, zeta, p, mu, res = stationary.incompressible_linear_solve(
            M, sq["L_v"], sq["B"], sq["M_p"], sq["L_p"], beta=m["beta"], bdofs_v=bd, v_d=M @ m["v_hat"],
            f=M @ m["f_nodal"], bc_values=m["v"][bd], solver_parameters=sp_, lambda_v_bounds=(0.3924, 2.0598),
            lambda_p_bounds=(0.5, 2.0))
        assert res.reason > 0
        Mp = sq["M_p"]
        one = np.ones(Mp.shape[0])

        def centred(q):
            return q - (one @ (Mp @ q)) / (one @ (Mp @ one))
        errs.append((kat.l2_error(M, v[None], m["v"][None]), kat.l2_error(M, zeta[None], m["zeta"][None]),
                     kat.l2_error(Mp, centred(p)[None], centred(m["p"])[None]),
                     kat.l2_error(Mp, centred(mu)[None], centred(m["mu"])[None])))
    e = np.array(errs)
    orders = np.log(e[:-1] / e[1:]) / np.log(2.0)
    # measured (nodal errors in the mass norm, so the polynomial fields of degree 4 converge faster than the a-priori
    # rates): v 3.6 / 4.0, zeta 5.1 / 4.7, p 4.1 / 3.8, mu 4.2 / 4.1; errors at N = 8: 1.5e-4, 6.9e-7, 1.8e-3, 5.9e-5
    assert (orders[-1, :2] > 2.7).all() and (orders[-1, 2:] > 1.7).all(), orders
    assert e[-1, 0] < 3e-4 and e[-1, 2] < 4e-3


def test_reference_mms_stationary_navier_stokes_control():
    """test/test_control.py:1095-1240 (degree 2; Q2 - Q1 here): manufactured stationary Navier-Stokes control
    problem -- nu = 1/100, v = (x y^3, (x^4 - y^4) / 4) on (-1, 1)^2 as desired state, exact state and Dirichlet
    data, zeta = 0, force -nu/2 div(grad v + grad v^T) + (grad v) v, beta = 1e-3 -- through the Picard loop
    ``Stationary.incompressible_non_linear_solve`` with the test's tolerances (1e-9, at most 10 iterations).  The
    reference prints the observed orders; asserted here: the loop converges, the velocity error decreases at
    least with third order and the adjoint stays at the level of the velocity error times beta."""
    import scipy.sparse as sp
    nu, beta = 1.0 / 100.0, 1e-3
    errs = []
    for N in (2, 4, 8):
        sq = fem.assemble_q2q1_stokes_2d(N, N, 2.0, 2.0)
        M, bd = sq["M_v"], sq["bdofs_v"]
        x, y = sq["coords_v"][:, 0] - 1.0, sq["coords_v"][:, 1] - 1.0
        v1, v2 = x * y ** 3, 0.25 * (x ** 4 - y ** 4)
        v = np.zeros(M.shape[0])
        v[0::2], v[1::2] = v1, v2
        f_nodal = np.zeros_like(v)
        f_nodal[0::2] = -0.5 * nu * (6.0 * x * y) + (y ** 3 * v1 + 3.0 * x * y ** 2 * v2)
        f_nodal[1::2] = -0.5 * nu * (3.0 * x ** 2 - 3.0 * y ** 2) + (x ** 3 * v1 - y ** 3 * v2)
        conv_v, conv_p = fem.convection_q2_2d(N, N, 2.0, 2.0), fem.convection_q1_q2wind_2d(N, N, 2.0, 2.0)
        I2 = sp.identity(2, format="csr")
        L_v, L_p = sq["L_v"], sq["L_p"]

        def D_v(w):
            C = sp.kron(conv_v(w[0::2], w[1::2]), I2, format="csr")
            C.sort_indices()
            return sp.csr_matrix((nu * L_v.data + C.data, M.indices, M.indptr), shape=M.shape)

        def D_p(w):
            C = conv_p(w[0::2], w[1::2])
            return sp.csr_matrix((nu * L_p.data + C.data, L_p.indices, L_p.indptr), shape=L_p.shape)
        sp_ = {"linear_solver": "fgmres", "fgmres_restart": 10, "maximum_iterations": 500, "relative_tolerance": 1e-10,
               "absolute_tolerance": 1e-10}
        out = stationary.incompressible_non_linear_solve(
            M, D_v, sq["B"], sq["M_p"], L_p, D_p, beta=beta, bdofs_v=bd, v_d=M @ v, f=M @ f_nodal, bc_values=v[bd],
            solver_parameters=sp_, lambda_v_bounds=(0.3924, 2.0598), lambda_p_bounds=(0.5, 2.0), max_non_linear_iter=10,
            relative_non_linear_tol=1e-9, absolute_non_linear_tol=1e-9)
        h = out["history"]
        assert h[-1] < 1e-5 * h[0], h
        errs.append((kat.l2_error(M, out["v"][None], v[None]), kat.l2_error(M, out["zeta"][None], 0.0 * v[None]),
                     out["iterations"]))
    e = np.array(errs)
    orders = np.log(e[:-1, 0] / e[1:, 0]) / np.log(2.0)
    assert (orders > 3.0).all(), orders                 # measured 4.09, 4.00 (nodal norm); errors 2.4e-2, 1.4e-3, 8.7e-5
    assert (e[:, 1] < 2e-5).all() and e[2, 1] < e[0, 1]     # the adjoint tends to zero (1.3e-5, 8.4e-6, 2.7e-6)
    assert (e[:, 2] <= 6).all()                          # 4, 5, 6 Picard iterations to 1e-9
