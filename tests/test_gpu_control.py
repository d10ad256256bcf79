"""GPU parity of the host-side ``Control.Instationary`` mirror (linear_solve and the
Picard / Gauss-Newton loop, control/control.py:2820-3375 and 3377-3590) against the oracle's
restatement of the same drivers."""
import numpy as np
import pytest

import kat
from oracle import control as ocontrol
from synthetic import fem

pytestmark = pytest.mark.gpu


def _callables(q):
    times = q["time_interval"][0] + q["tau"] * np.arange(q["n_t"])

    def desired_state(t):
        i = int(round((t - times[0]) / q["tau"]))
        return q["v_d"][i], q["v_hat"][i]

    def force_f(t):
        i = int(round((t - times[0]) / q["tau"]))
        return q["f"][i]
    return desired_state, force_f


@pytest.mark.parametrize("CN", [True, False])
def test_linear_solve_readme_problem(CN):
    """BASELINE config C1 through the reference's call shape (README.md:24-60)."""
    from control_b200 import Control
    q = kat.heat_problem(10, 10, CN)
    desired_state, force_f = _callables(q)
    c = Control.Instationary(q["M"], q["K"], desired_state=desired_state, force_f=force_f, beta=q["beta"],
                             n_t=q["n_t"], CN=CN, time_interval=q["time_interval"], bc_dofs=q["bdofs"])
    # BE: the epsilon-regularised system stalls near 1e-8 relative residual (rounding floor of
    # classical Gram-Schmidt on an ill-conditioned operator); counts are comparable above it
    sp_ = {"linear_solver": "fgmres", "maximum_iterations": 200, "relative_tolerance": 1e-8 if CN else 1e-6,
           "absolute_tolerance": 0.0, "gmres_restart": 100}
    ksp = c.linear_solve(lambda_v_bounds=q["lambda_v_bounds"], solver_parameters=sp_, print_error=False)
    ref = ocontrol.linear_solve(q["M"], q["K"], beta=q["beta"], n_t=q["n_t"], CN=CN,
                                time_interval=q["time_interval"], bdofs=q["bdofs"], v_d=q["v_d"], f=q["f"],
                                lambda_v_bounds=q["lambda_v_bounds"], solver_parameters=sp_)
    assert ksp.getConvergedReason() > 0 and abs(ksp.getIterationNumber() - ref["ksp"].its) <= 1
    tol = 1e-6 if CN else 1e-4
    assert np.abs(c._v - ref["v"]).max() < tol * np.abs(ref["v"]).max()
    assert np.abs(c._zeta - ref["zeta"]).max() < tol * np.abs(ref["zeta"]).max()
    # default solver_parameters of the reference: GMRES(10), 50 its, rtol 1e-6 (control.py:3260-3266)
    c2 = Control.Instationary(q["M"], q["K"], desired_state=desired_state, force_function=force_f,
                              beta=q["beta"], n_t=q["n_t"], CN=CN, time_interval=q["time_interval"],
                              bc_dofs=q["bdofs"])
    ksp2 = c2.linear_solve(lambda_v_bounds=q["lambda_v_bounds"], print_error=False)
    ref2 = ocontrol.linear_solve(q["M"], q["K"], beta=q["beta"], n_t=q["n_t"], CN=CN,
                                 time_interval=q["time_interval"], bdofs=q["bdofs"], v_d=q["v_d"], f=q["f"],
                                 lambda_v_bounds=q["lambda_v_bounds"])
    assert abs(ksp2.its - ref2["ksp"].its) <= 1
    with pytest.raises(TypeError):
        Control.Instationary(q["M"], q["K"], force_f=force_f, force_function=force_f)
    c.close()
    c2.close()


@pytest.mark.parametrize("CN,gauss_newton,max_it", [(True, False, 6), (False, False, 4), (True, True, 3)])
def test_non_linear_solve_matches_oracle(CN, gauss_newton, max_it):
    """Config C5 in small: non-linear diffusion (1 + v^2), per-level K_i(v_i) (non-symmetric
    with Gauss_Newton=True); outer residual history and inner iteration counts vs the oracle."""
    from control_b200 import Control
    nx, n_t = 12, 6
    q = kat.heat_problem(nx, n_t, CN, beta=1e-2)
    Dv = fem.nonlinear_diffusion_p1_2d(nx, nx, 2.0, 2.0)
    desired_state, force_f = _callables(q)
    sp_ = {"linear_solver": "fgmres", "maximum_iterations": 200, "relative_tolerance": 1e-7 if CN else 1e-6,
           "absolute_tolerance": 0.0, "gmres_restart": 100}
    c = Control.Instationary(q["M"], lambda v, t, gn: Dv(v, gn), desired_state=desired_state, force_f=force_f,
                             beta=q["beta"], Gauss_Newton=gauss_newton, n_t=n_t, CN=CN,
                             time_interval=q["time_interval"], bc_dofs=q["bdofs"])
    k = c.non_linear_solve(lambda_v_bounds=q["lambda_v_bounds"], solver_parameters=sp_,
                           max_non_linear_iter=max_it, relative_non_linear_tol=1e-9,
                           print_error_non_linear=False)
    ref = ocontrol.non_linear_solve(q["M"], lambda v, t: Dv(v, gauss_newton), beta=q["beta"], n_t=n_t, CN=CN,
                                    time_interval=q["time_interval"], bdofs=q["bdofs"], v_d=q["v_d"], f=q["f"],
                                    solver_parameters=sp_, lambda_v_bounds=q["lambda_v_bounds"],
                                    max_non_linear_iter=max_it, relative_non_linear_tol=1e-9)
    assert k == ref["iterations"]
    assert np.allclose(c.non_linear_history, ref["history"], rtol=1e-4)
    assert np.abs(c._v - ref["v"]).max() < 1e-5 * np.abs(ref["v"]).max()
    assert np.abs(c._zeta - ref["zeta"]).max() < 1e-5 * np.abs(ref["zeta"]).max()
    if not gauss_newton:
        assert c.non_linear_history[-1] < 0.05 * c.non_linear_history[1]      # Picard contracts
    c.close()


@pytest.mark.parametrize("CN", [True, False])
@pytest.mark.parametrize("with_v0", [False, True])
def test_device_rhs_from_nodal_data_matches_oracle(CN, with_v0):
    """ctl_build_rhs: the right-hand sides of linear_solve
    (control/control.py:2980-3243) assembled on the device from nodal v_hat / f, against the oracle's
    restatement fed with the cofunctions M v_hat / M f; then the solve from that device vector."""
    import torch
    from control_b200 import MultiBlockSystem
    from oracle import kkt
    q = kat.heat_problem(14, 7, CN, beta=1e-2)
    M, K, bd, n_t = q["M"], q["K"], q["bdofs"], q["n_t"]
    rng = np.random.default_rng(11)
    f_nodal = rng.standard_normal((n_t, M.shape[0]))               # general data, non-zero on the boundary
    v_hat = q["v_hat"] + 0.1 * rng.standard_normal(q["v_hat"].shape)
    v_0 = None
    if with_v0:
        v_0 = rng.standard_normal(M.shape[0])
        v_0[bd] = 0.0
    s = MultiBlockSystem(M, K, n_t=n_t, beta=q["beta"], CN=CN, time_interval=q["time_interval"], bc_dofs=bd)
    b = s.build_rhs_device(torch.from_numpy(v_hat).to(s.device), torch.from_numpy(f_nodal).to(s.device), v_0)
    b0, b1 = s.to_host_blocks(b)
    r0, r1 = kkt.build_rhs(M, K, q["tau"], n_t, CN, bd, (M @ v_hat.T).T, (M @ f_nodal.T).T,
                           np.zeros(M.shape[0]) if v_0 is None else v_0)
    scale = max(np.abs(r0).max(), np.abs(r1).max())
    assert np.abs(b0 - r0).max() <= 1e-13 * scale and np.abs(b1 - r1).max() <= 1e-13 * scale
    # and straight into a device solve
    s.setup_preconditioner(lambda_v_bounds=q["lambda_v_bounds"], coarse_max=60)
    u = s.new_vector()
    sp_ = {"linear_solver": "fgmres", "maximum_iterations": 200, "relative_tolerance": 1e-8, "absolute_tolerance": 0.0,
           "gmres_restart": 100}
    info = s.solve_device(b, u, solver_parameters=sp_)
    assert info.reason > 0
    assert s.residual_norm(b, u) <= 1e-6 * float(b.norm())
    s.close()


@pytest.mark.parametrize("CN", [True, False])
def test_linear_solve_with_inhomogeneous_dirichlet_data(CN):
    """Time-dependent inhomogeneous Dirichlet data through Control.Instationary.linear_solve on the GPU
    against the oracle (whose lifting is checked against a direct un-eliminated solve in
    tests/test_oracle.py)."""
    from control_b200 import Control
    q = kat.heat_problem(10, 6, CN, beta=1e-2)
    M, K, bd, n_t = q["M"], q["K"], q["bdofs"], q["n_t"]
    x, y = q["coords"][bd, 0], q["coords"][bd, 1]
    times = q["tau"] * np.arange(n_t)

    def bc_values(t):
        return (1.0 + t) * np.sin(x + 2.0 * y) + t
    g = np.stack([bc_values(t) for t in times])
    v_0 = np.zeros(M.shape[0])
    v_0[bd] = g[0]
    desired_state, force_f = _callables(q)
    c = Control.Instationary(M, K, desired_state=desired_state, force_f=force_f, beta=q["beta"], n_t=n_t, CN=CN,
                             time_interval=q["time_interval"], bc_dofs=bd, bc_values=bc_values, initial_condition=v_0)
    # BE: below ~1e-8 the iteration sits on the rounding floor of classical Gram-Schmidt, so the
    # counts are only comparable at a looser tolerance
    sp_ = {"linear_solver": "fgmres", "maximum_iterations": 300, "relative_tolerance": 1e-11 if CN else 1e-7,
           "absolute_tolerance": 0.0, "gmres_restart": 100}
    info = c.linear_solve(solver_parameters=sp_, lambda_v_bounds=q["lambda_v_bounds"], print_error=False, coarse_max=60)
    ref = ocontrol.linear_solve(M, K, beta=q["beta"], n_t=n_t, CN=CN, time_interval=q["time_interval"], bdofs=bd,
                                v_d=q["v_d"], f=q["f"], v_0=v_0, bc_values=g, solver_parameters=sp_,
                                lambda_v_bounds=q["lambda_v_bounds"], amg_params=dict(coarse_max=60))
    assert info.reason == ref["ksp"].reason > 0 and abs(info.its - ref["ksp"].its) <= 1
    tol = 1e-7 if CN else 1e-4
    assert np.abs(c._v - ref["v"]).max() < tol * np.abs(ref["v"]).max()
    assert np.abs(c._zeta - ref["zeta"]).max() < tol * np.abs(ref["zeta"]).max()
    assert np.array_equal(c._v[1:, bd], g[1:])
    c.close()


def test_reference_mms_heat_problem_on_gpu():
    """The manufactured heat-control problem of the reference's convergence studies
    (test/test_control.py:1983-2138) on a 16 x 16 mesh with n_t = 100 -- 99 time blocks, i.e. the wide
    (ld = 128) kernels -- and inhomogeneous Dirichlet data: same discretisation error as the oracle."""
    from control_b200 import Control
    q = kat.mms_heat_problem(16, 100, True)
    times = q["tau"] * np.arange(q["n_t"])

    def level(t):
        return int(round(t / q["tau"]))
    c = Control.Instationary(q["M"], q["K"], desired_state=lambda t: (q["v_d"][level(t)], q["v_hat"][level(t)]),
                             force_f=lambda t: q["f"][level(t)], beta=q["beta"], n_t=q["n_t"], CN=True,
                             time_interval=q["time_interval"], bc_dofs=q["bdofs"],
                             bc_values=lambda t: q["bc_values"][level(t)], initial_condition=q["v_0"])
    sp_ = {"linear_solver": "fgmres", "gmres_restart": 100, "maximum_iterations": 200, "relative_tolerance": 1e-10,
           "absolute_tolerance": 1e-10}
    info = c.linear_solve(solver_parameters=sp_, lambda_v_bounds=(0.5, 2.0), print_error=False)
    assert info.reason > 0
    ev = np.sqrt(q["tau"]) * kat.l2_error(q["M"], c._v, q["v_exact"])
    ez = np.sqrt(q["tau"]) * kat.l2_error(q["M"], c._zeta, q["zeta_exact"])
    assert abs(ev - 0.02218281663521956) < 1e-6 and abs(ez - 0.05136980559114903) < 1e-6      # the oracle's errors
    assert len(times) == 100
    c.close()
