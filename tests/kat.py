"""Known-answer problems re-created from the reference's own tests (shared by the CPU
oracle tests and the GPU parity tests).

``instationary_kat`` follows test/test_control.py:1243-1444 (BE) and 1447-1655 (CN):
Q2 on an 8x8 quadrilateral unit square, n_t = 5, tau = 0.25, beta = 1e-3, K = Laplacian,
homogeneous Dirichlet data, analytic v_ref / zeta_ref per time level, right-hand sides
built row by row from the block stencils (untransformed rows: ``linear_solve`` applies
T_1 / T_2 itself, control/control.py:3242-3243), FGMRES to 1e-14, Chebyshev bounds
(0.25, 1.5625).
"""
import numpy as np

from synthetic import fem
from synthetic.problems import heat_problem, heat_problem_3d, stokes_problem  # noqa: F401


def instationary_kat(CN, mesh_size=3):
    nx = 2 ** mesh_size
    M, K, coords, bdofs = fem.assemble_q2_2d(nx, nx)
    X0, X1 = coords[:, 0], coords[:, 1]
    beta = 1e-3
    n_t = 5
    tau = 0.25
    pi, sin, exp = np.pi, np.sin, np.exp
    n = M.shape[0]
    v_ref = np.zeros((n_t, n))
    zeta_ref = np.zeros((n_t, n))
    v_ref[1] = tau * sin(3.0 * pi * X0) * sin(4.0 * pi * X1)
    v_ref[2] = tau ** 2.0 * X0 * exp(X1) * sin(pi * X0) * sin(2.0 * pi * X1)
    v_ref[3] = tau ** 3.0 * sin(3.0 * pi * X0) * sin(4.0 * pi * X1)
    v_ref[4] = tau ** 4.0 * X0 * exp(X1) * sin(pi * X0) * sin(2.0 * pi * X1)
    zeta_ref[0] = sin(pi * X0) * sin(2.0 * pi * X1)
    zeta_ref[1] = tau * sin(3.0 * pi * X0) * sin(4.0 * pi * X1)
    zeta_ref[2] = tau ** 2.0 * sin(pi * X0) * sin(2.0 * pi * X1)
    zeta_ref[3] = tau ** 3.0 * sin(3.0 * pi * X0) * sin(4.0 * pi * X1)
    if not CN:
        N = n_t
        b_0 = np.zeros((N, n))
        b_1 = np.zeros((N, n))
        for i in range(4):          # test/test_control.py:1331-1362
            b_0[i] = tau * (M @ v_ref[i]) + tau * (K @ zeta_ref[i]) + M @ zeta_ref[i] \
                - M @ zeta_ref[i + 1]
        b_0[4] = tau * (K @ zeta_ref[4]) + M @ zeta_ref[4]                    # 1363-1366
        b_1[0] = tau * (K @ v_ref[0]) + M @ v_ref[0]                           # 1373-1376
        for i in range(1, 5):       # 1377-1408
            b_1[i] = tau * (K @ v_ref[i]) + M @ v_ref[i] - M @ v_ref[i - 1] \
                - (tau / beta) * (M @ zeta_ref[i])
    else:
        N = n_t - 1
        h = 0.5 * tau
        b_0 = np.zeros((N, n))
        b_1 = np.zeros((N, n))
        # test/test_control.py:1538-1576
        b_0[0] = h * (M @ v_ref[1]) + h * (K @ zeta_ref[0]) + M @ zeta_ref[0] \
            + h * (K @ zeta_ref[1]) - M @ zeta_ref[1]
        for i in (1, 2):
            b_0[i] = h * (M @ v_ref[i + 1]) + h * (M @ v_ref[i]) + h * (K @ zeta_ref[i]) \
                + M @ zeta_ref[i] + h * (K @ zeta_ref[i + 1]) - M @ zeta_ref[i + 1]
        b_0[3] = h * (M @ v_ref[4]) + h * (M @ v_ref[3]) + h * (K @ zeta_ref[3]) + M @ zeta_ref[3]
        # 1582-1620 (line 1613 assigns zeta_ref.sub(3), which equals v_ref.sub(3))
        b_1[0] = h * (K @ v_ref[1]) + M @ v_ref[1] - (h / beta) * (M @ zeta_ref[0]) \
            - (h / beta) * (M @ zeta_ref[1])
        for i in (1, 2):
            b_1[i] = h * (K @ v_ref[i + 1]) + M @ v_ref[i + 1] + h * (K @ v_ref[i]) \
                - M @ v_ref[i] - (h / beta) * (M @ zeta_ref[i]) - (h / beta) * (M @ zeta_ref[i + 1])
        b_1[3] = h * (K @ v_ref[4]) + M @ v_ref[4] + h * (K @ zeta_ref[3]) - M @ zeta_ref[3] \
            - (h / beta) * (M @ zeta_ref[3])
    solver_parameters = {"linear_solver": "fgmres",            # 1418-1423 / 1629-1634
                         "fgmres_restart": 10,
                         "maximum_iterations": 500,
                         "relative_tolerance": 1.0e-14,
                         "absolute_tolerance": 1.0e-14,
                         "monitor_convergence": False}
    return dict(M=M, K=K, coords=coords, bdofs=bdofs, beta=beta, n_t=n_t, tau=tau, CN=CN,
                v_ref=v_ref, zeta_ref=zeta_ref, b_0=b_0, b_1=b_1,
                lambda_v_bounds=(0.25, 1.5625), solver_parameters=solver_parameters)


def l2_error(M, a, b):
    """sqrt(|assemble(inner(a - b, a - b) * dx)|) summed over the time levels, as at
    test/test_control.py:1438-1444 (mixed-space inner product = sum over blocks)."""
    d = a - b
    return float(np.sqrt(abs(sum(di @ (M @ di) for di in d))))
