"""GPU tests added at the end of round 1 (convection-diffusion, ``Control.Stationary``, Navier-Stokes Picard
loops).  The file name sorts last on purpose: these cases were written when almost no GPU time was left, so
``pytest -x`` reaches every long-verified test before them.  All of them passed on a B200
(gpurun_out/late_tests.log of that run is quoted in DESIGN.md)."""
import numpy as np
import pytest

import kat
from oracle import control as ocontrol

pytestmark = pytest.mark.gpu


def test_reference_mms_convection_diffusion_problem_on_gpu():
    """The manufactured convection-diffusion control problem of the reference's convergence studies
    (test/test_control.py:2675-2857) on a 16 x 16 mesh with n_t = 100: 100 distinct non-symmetric ``K_i``
    (value panels, ``K_iᵀ`` in the adjoint rows, one AMG hierarchy per level and orientation), inhomogeneous
    Dirichlet data.  Same discretisation error as the oracle, same iteration count."""
    from control_b200 import Control
    q = kat.mms_convection_diffusion_problem(16, 100, True)

    def level(t):
        return int(round(t / q["tau"]))
    c = Control.Instationary(q["M"], lambda v_i, t, gauss_newton: q["K_levels"][level(t)],
                             desired_state=lambda t: (q["v_d"][level(t)], q["v_hat"][level(t)]),
                             force_f=lambda t: q["f"][level(t)], beta=q["beta"], n_t=q["n_t"], CN=True,
                             time_interval=q["time_interval"], bc_dofs=q["bdofs"],
                             bc_values=lambda t: q["bc_values"][level(t)], initial_condition=q["v_0"])
    sp_ = {"linear_solver": "fgmres", "gmres_restart": 100, "maximum_iterations": 200, "relative_tolerance": 1e-10,
           "absolute_tolerance": 1e-10}
    info = c.linear_solve(solver_parameters=sp_, lambda_v_bounds=(0.5, 2.0), print_error=False)
    assert info.reason > 0
    ref = ocontrol.linear_solve(q["M"], q["K_levels"], beta=q["beta"], n_t=q["n_t"], CN=True,
                                time_interval=q["time_interval"], bdofs=q["bdofs"], v_d=q["v_d"], f=q["f"],
                                v_0=q["v_0"], bc_values=q["bc_values"], solver_parameters=sp_,
                                lambda_v_bounds=(0.5, 2.0))
    assert abs(info.its - ref["ksp"].its) <= 1
    assert np.abs(c._v - ref["v"]).max() < 1e-7 * np.abs(ref["v"]).max()
    assert np.abs(c._zeta - ref["zeta"]).max() < 1e-7 * np.abs(ref["zeta"]).max()
    ev = np.sqrt(q["tau"]) * kat.l2_error(q["M"], c._v, q["v_exact"])
    ez = np.sqrt(q["tau"]) * kat.l2_error(q["M"], c._zeta, q["zeta_exact"])
    assert abs(ev - 0.022182639553203404) < 1e-6 and abs(ez - 0.05137767135076495) < 1e-6     # the oracle's errors
    c.close()


def test_reference_stationary_known_answer_on_gpu():
    """test/test_control.py:26-119 (``test_stationary_linear_control``) through ``Control.Stationary`` on the
    device: Q2 on 8x8 quads, no boundary conditions, ready right-hand sides, the test's solver parameters and
    Chebyshev bounds; the reference asserts 1e-13 (5e-13 here, as for the instationary KATs on the GPU)."""
    from control_b200 import Control
    from oracle import stationary
    from synthetic import fem
    M, L, coords, _ = fem.assemble_q2_2d(8, 8)
    K = (L + M).tocsr()
    beta = 1e-3
    X0, X1 = coords[:, 0], coords[:, 1]
    v_ref = X0 * np.exp(X1)
    zeta_ref = np.sin(np.pi * X0) * np.sin(2.0 * np.pi * X1)
    b_0 = M @ v_ref + K @ zeta_ref
    b_1 = K @ v_ref - (1.0 / beta) * (M @ zeta_ref)
    sp_ = {"linear_solver": "fgmres", "fgmres_restart": 10, "maximum_iterations": 500, "relative_tolerance": 1e-14,
           "absolute_tolerance": 1e-14, "monitor_convergence": False}
    c = Control.Stationary(M, K, beta=beta)
    info = c.linear_solve(lambda_v_bounds=(0.25, 1.5625), solver_parameters=sp_, v_d=b_0, f=b_1, print_error=False)
    assert info.reason > 0
    assert kat.l2_error(M, c._v[None], v_ref[None]) < 5e-13
    assert kat.l2_error(M, c._zeta[None], zeta_ref[None]) < 5e-13
    ref = stationary.linear_solve(M, K, beta=beta, bdofs=[], v_d=b_0, f=b_1, check_v_d=False, check_f=False,
                                  lambda_v_bounds=(0.25, 1.5625), solver_parameters=sp_)
    assert abs(info.its - ref["ksp"].its) <= 2                  # rtol 1e-14: on the rounding floor
    c.close()


def test_stationary_inhomogeneous_and_non_linear_on_gpu():
    """``Control.Stationary`` with inhomogeneous Dirichlet data (linear solve) and the Picard / Gauss-Newton loop
    (control/control.py:630-800) against the oracle's direct restatement of the stationary drivers."""
    from control_b200 import Control
    from oracle import stationary
    from synthetic import fem
    nx = 8
    M, L, coords, bd = fem.assemble_p1_2d(nx, nx, 1.0, 1.0)
    n = M.shape[0]
    x, y = coords[:, 0], coords[:, 1]
    v_hat = np.sin(np.pi * x) * np.sin(np.pi * y) * np.exp(x + y)
    beta = 1e-2
    sp_ = {"linear_solver": "fgmres", "maximum_iterations": 200, "relative_tolerance": 1e-10, "absolute_tolerance": 0.0}
    # linear, inhomogeneous data
    D_v = (L + 2.0 * M).tocsr()
    g = np.cos(3.0 * x[bd]) + y[bd]
    f = M @ (x * y)
    c = Control.Stationary(M, D_v, desired_state=lambda: (M @ v_hat, v_hat), force_f=lambda: f, beta=beta,
                           bc_dofs=bd, bc_values=g)
    info = c.linear_solve(solver_parameters=sp_, lambda_v_bounds=(0.5, 2.0), print_error=False)
    ref = stationary.linear_solve(M, D_v, beta=beta, bdofs=bd, v_d=M @ v_hat, f=f, bc_values=g, solver_parameters=sp_,
                                  lambda_v_bounds=(0.5, 2.0))
    assert info.reason > 0 and abs(info.its - ref["ksp"].its) <= 1
    assert np.abs(c._v - ref["v"]).max() < 1e-7 * np.abs(ref["v"]).max()
    assert np.abs(c._zeta - ref["zeta"]).max() < 1e-7 * np.abs(ref["zeta"]).max()
    assert np.array_equal(c._v[bd], g)
    c.close()
    # non-linear diffusion, Picard and Gauss-Newton
    D = fem.nonlinear_diffusion_p1_2d(nx, nx, 1.0, 1.0)
    for gauss_newton in (False, True):
        c = Control.Stationary(M, D, desired_state=lambda: (M @ v_hat, v_hat), beta=beta, Gauss_Newton=gauss_newton,
                               bc_dofs=bd)
        k = c.non_linear_solve(solver_parameters=sp_, lambda_v_bounds=(0.5, 2.0), max_non_linear_iter=30,
                               print_error_non_linear=False)
        out = stationary.non_linear_solve(M, lambda v: D(v, gauss_newton), beta=beta, bdofs=bd, v_d=M @ v_hat,
                                          f=np.zeros(n), solver_parameters=sp_, lambda_v_bounds=(0.5, 2.0),
                                          max_non_linear_iter=30)
        assert k == out["iterations"]
        assert np.allclose(c.non_linear_history, out["history"], rtol=1e-4)
        assert np.abs(c._v - out["v"]).max() < 1e-5 * np.abs(out["v"]).max()
        c.close()


def test_navier_stokes_picard_loop_on_gpu():
    """``Control.Instationary.incompressible_non_linear_solve`` (control/control.py:4886-5219) on the problem of
    the reference's Navier-Stokes tests (test/test_control.py:4271-4368), 4 x 4 Q2-Q1 cells, CN: per-level
    non-symmetric ``D_v_i`` on the velocity handle AND ``D_p_i`` on the pressure handle, the Laplacian of
    ``solver_K_p`` set separately (``ctl_stokes_set_laplacian_p``).  Residual history and iterates against the
    oracle's restatement of the loop."""
    from control_b200 import Control
    from oracle import stokes
    q = kat.reference_navier_stokes_problem(True, nx=4, n_t=6)
    sq = q["sq"]

    def level(t):
        return int(round(t / q["tau"]))
    c = Control.Instationary(q["M"], lambda v_i, t, gauss_newton: q["D_v"](v_i, t),
                             desired_state=lambda t: (q["v_d"][level(t)], q["v_hat"][level(t)]), beta=q["beta"],
                             n_t=q["n_t"], CN=True, time_interval=q["time_interval"], bc_dofs=q["bdofs"],
                             bc_values=lambda t: q["bc_values"][level(t)])
    space_p = dict(B=q["B"], M_p=sq["M_p"], K_p=sq["L_p"], forward_matrix_p=lambda v_i, t, gauss_newton: q["D_p"](v_i, t))
    sp_ = {"linear_solver": "fgmres", "maximum_iterations": 200, "relative_tolerance": 1e-8, "absolute_tolerance": 0.0}
    k = c.incompressible_non_linear_solve("constant", space_p=space_p, solver_parameters=sp_,
                                          lambda_v_bounds=q["lambda_v_bounds"], lambda_p_bounds=q["lambda_p_bounds"],
                                          print_error_non_linear=False)
    out = stokes.incompressible_non_linear_solve(
        q["M"], q["D_v"], q["B"], sq["M_p"], sq["L_p"], q["D_p"], beta=q["beta"], n_t=q["n_t"], CN=True,
        time_interval=q["time_interval"], bdofs_v=q["bdofs"], v_d=q["v_d"], f=q["f"], bc_values=q["bc_values"],
        solver_parameters=sp_, lambda_v_bounds=q["lambda_v_bounds"], lambda_p_bounds=q["lambda_p_bounds"])
    assert k == out["iterations"]
    assert np.allclose(c.non_linear_history, out["history"], rtol=1e-3)
    assert np.abs(c._v - out["v"]).max() < 1e-5 * np.abs(out["v"]).max()
    assert max(abs(a - b) for a, b in zip(c.inner_iterations, out["inner_its"])) <= 2
    c.close()


def test_reference_stationary_stokes_known_answer_on_gpu():
    """test/test_control.py:232-358 (``test_stationary_incompressible_linear_control``) through
    ``Control.Stationary.incompressible_linear_solve`` on the device (the outer system mapped onto the
    instationary Stokes handle with one time block): vector Q2 - Q1 on 4 x 4 quads, forward form grad.grad +
    mass on both spaces, analytic fields, the test's tolerances for the solve (1e-14)."""
    from control_b200 import Control
    from synthetic import fem
    sq = fem.assemble_q2q1_stokes_2d(4, 4)
    M, L, B, Mp, Lp, bd = sq["M_v"], sq["L_v"], sq["B"], sq["M_p"], sq["L_p"], sq["bdofs_v"]
    K = (L + M).tocsr()
    beta = 1e-3
    x, y = sq["coords_v"][:, 0], sq["coords_v"][:, 1]
    px, py = sq["coords_p"][:, 0], sq["coords_p"][:, 1]

    def vec(cx, cy):
        a = np.zeros(M.shape[0])
        a[0::2], a[1::2] = cx, cy
        return a
    v_ref = vec(x * np.exp(y) * np.sin(np.pi * x) * np.sin(2.0 * np.pi * y), np.sin(3.0 * np.pi * x) * np.sin(4.0 * np.pi * y))
    zeta_ref = vec(np.sin(np.pi * x) * np.sin(2.0 * np.pi * y), np.sin(3.0 * np.pi * x) * np.sin(4.0 * np.pi * y))
    p_ref = np.sin(np.pi * px) * np.sin(2.0 * np.pi * py)
    mu_ref = px * np.exp(py)
    sp_ = {"linear_solver": "fgmres", "fgmres_restart": 10, "maximum_iterations": 500, "relative_tolerance": 1e-14,
           "absolute_tolerance": 1e-14}
    c = Control.Stationary(M, K, beta=beta, bc_dofs=bd)
    space_p = dict(B=B, M_p=Mp, K_p=Lp, forward_matrix_p=(Lp + Mp).tocsr())
    info = c.incompressible_linear_solve("constant", space_p=space_p, solver_parameters=sp_, print_error=False,
                                         lambda_v_bounds=(0.3924, 2.0598), lambda_p_bounds=(0.5, 2.0),
                                         v_d=M @ v_ref + K @ zeta_ref + B.T @ mu_ref,
                                         f=K @ v_ref - (1.0 / beta) * (M @ zeta_ref) + B.T @ p_ref,
                                         div_v=B @ v_ref, div_zeta=B @ zeta_ref)
    assert info.reason > 0

    def shift(q):
        return q - np.ones(Mp.shape[0]) @ (Mp @ q)
    assert kat.l2_error(M, c._v[None], v_ref[None]) < 5e-13
    assert kat.l2_error(M, c._zeta[None], zeta_ref[None]) < 5e-13
    assert kat.l2_error(Mp, shift(c._p)[None], shift(p_ref)[None]) < 3e-12
    assert kat.l2_error(Mp, shift(c._mu)[None], shift(mu_ref)[None]) < 3e-12
    c.close()


def test_stokes_control_with_multigrid_mass_solver():
    """``incompressible_linear_solve(Multigrid=True)``: the (1,1) block of the inner heat-type preconditioner is
    solved with AMG cycles on ``M_v`` instead of Chebyshev (control/control.py:1954-1965 inside 4346-4353)."""
    from control_b200 import Control
    from oracle import stokes
    q = kat.stokes_problem(6, 6, True)
    th = q["th"]
    times = q["tau"] * np.arange(q["n_t"])
    lookup = {round(float(t), 12): i for i, t in enumerate(times)}
    c = Control.Instationary(q["M"], q["K"], desired_state=lambda t: (q["v_d"][lookup[round(float(t), 12)]],
                                                                    q["v_hat"][lookup[round(float(t), 12)]]),
                             force_f=lambda t: q["f"][lookup[round(float(t), 12)]], beta=q["beta"], CN=True, n_t=q["n_t"],
                             time_interval=q["time_interval"], bc_dofs=q["bdofs"])
    sp_ = {"linear_solver": "fgmres", "maximum_iterations": 200, "relative_tolerance": 1e-8, "absolute_tolerance": 0.0,
           "gmres_restart": 100}
    amg, amg_p = dict(coarse_max=40), dict(coarse_max=20)
    info = c.incompressible_linear_solve("constant", space_p=dict(B=th["B"], M_p=th["M_p"], K_p=th["K_p"]),
                                         solver_parameters=sp_, Multigrid=True, lambda_p_bounds=q["lambda_p_bounds"],
                                         amg=amg, amg_p=amg_p)
    v, zeta, p, mu, res = stokes.incompressible_linear_solve(
        th["M_v"], th["K_v"], th["B"], th["M_p"], th["K_p"], beta=q["beta"], n_t=q["n_t"], CN=True,
        time_interval=q["time_interval"], bdofs_v=q["bdofs"], v_d=q["v_d"], f=q["f"], solver_parameters=sp_,
        Multigrid=True, lambda_p_bounds=q["lambda_p_bounds"], amg_params=amg, amg_params_p=amg_p)
    assert info.reason == res.reason > 0
    assert abs(info.its - res.its) <= max(1, int(0.03 * res.its))
    assert np.abs(c._v - v).max() < 1e-5 * np.abs(v).max() and np.abs(c._zeta - zeta).max() < 1e-5 * np.abs(zeta).max()
    c.close()
