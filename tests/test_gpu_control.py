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
