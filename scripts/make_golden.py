"""Generates the fixtures under tests/golden/ (committed together with this script).

The reference cannot run in this image (Firedrake / PETSc / hypre are absent), so these are NOT
outputs of the reference: they are the converged solutions of the reference's README problem
(BASELINE config C1) and of the Stokes control problem in small, computed by the CPU oracle with EXACT
inner solves to rtol 1e-13 -- i.e. the discrete KKT solutions themselves, independent of the
preconditioner, the AMG stand-in and the Krylov method.  Tests check the oracle (regression) and
the CUDA path (through the C ABI) against them.

    python scripts/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import control as ocontrol          # noqa: E402
from oracle import stokes as ostokes            # noqa: E402
from synthetic import problems                  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
SP = {"linear_solver": "fgmres", "gmres_restart": 200, "maximum_iterations": 400, "relative_tolerance": 1e-13,
      "absolute_tolerance": 0.0}


def heat(CN):
    q = problems.heat_problem(10, 10, CN)
    r = ocontrol.linear_solve(q["M"], q["K"], beta=q["beta"], n_t=q["n_t"], CN=CN, time_interval=q["time_interval"],
                              bdofs=q["bdofs"], v_d=q["v_d"], f=q["f"], lambda_v_bounds=q["lambda_v_bounds"],
                              solver_parameters=SP, inner="exact")
    J = ocontrol.objective(q["M"], r["v"], r["zeta"], q["v_hat"], q["tau"], q["beta"], CN)
    assert r["ksp"].reason > 0
    np.savez_compressed(os.path.join(OUT, f"c1_heat_{'cn' if CN else 'be'}.npz"), v=r["v"], zeta=r["zeta"], J=J,
                        nx=10, n_t=10, beta=q["beta"])
    print("heat", CN, r["ksp"].its, J)


def stokes(CN):
    q = problems.stokes_problem(4, 5, CN)
    th = q["th"]
    v, zeta, p, mu, res = ostokes.incompressible_linear_solve(
        th["M_v"], th["K_v"], th["B"], th["M_p"], th["K_p"], beta=q["beta"], n_t=q["n_t"], CN=CN,
        time_interval=q["time_interval"], bdofs_v=q["bdofs"], v_d=q["v_d"], f=q["f"], solver_parameters=SP,
        lambda_v_bounds=q["lambda_v_bounds"], lambda_p_bounds=q["lambda_p_bounds"], inner="exact",
        amg_params_p=dict(coarse_max=10 ** 6))
    assert res.reason > 0
    np.savez_compressed(os.path.join(OUT, f"stokes_4x4_{'cn' if CN else 'be'}.npz"), v=v, zeta=zeta, p=p, mu=mu,
                        nx=4, n_t=5, beta=q["beta"])
    print("stokes", CN, res.its)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for CN in (True, False):
        heat(CN)
        stokes(CN)
