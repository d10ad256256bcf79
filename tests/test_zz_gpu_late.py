"""GPU tests added at the end of round 1.  The file name sorts last on purpose: these cases were written
when almost no GPU time was left, so ``pytest -x`` reaches every long-verified test before them."""
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
