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
import scipy.sparse as sp

from synthetic import fem
from synthetic.problems import heat_problem, heat_problem_3d, stokes_problem  # noqa: F401


def _kat_fields(X0, X1, tau, n_t=5):
    """v_ref / zeta_ref of test/test_control.py:1263-1330 (BE) = 1467-1534 (CN) at the nodes."""
    pi, sin, exp = np.pi, np.sin, np.exp
    n = X0.size
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
    return v_ref, zeta_ref


def _kat_rows(M, K, v_ref, zeta_ref, CN, tau, beta):
    """Untransformed right-hand-side rows built from the block stencils, exactly as the reference's
    tests write them (test/test_control.py:1331-1408 BE, 1538-1620 CN)."""
    n_t, n = v_ref.shape
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
        # 1582-1620
        b_1[0] = h * (K @ v_ref[1]) + M @ v_ref[1] - (h / beta) * (M @ zeta_ref[0]) \
            - (h / beta) * (M @ zeta_ref[1])
        for i in (1, 2):
            b_1[i] = h * (K @ v_ref[i + 1]) + M @ v_ref[i + 1] + h * (K @ v_ref[i]) \
                - M @ v_ref[i] - (h / beta) * (M @ zeta_ref[i]) - (h / beta) * (M @ zeta_ref[i + 1])
        # (the reference's line 1613 writes zeta_ref.sub(3) in the two middle terms; its fields have
        # v_ref.sub(3) == zeta_ref.sub(3), and the block stencil is the one with v_ref)
        b_1[3] = h * (K @ v_ref[4]) + M @ v_ref[4] + h * (K @ v_ref[3]) - M @ v_ref[3] \
            - (h / beta) * (M @ zeta_ref[3])
    return b_0, b_1


def instationary_kat(CN, mesh_size=3):
    nx = 2 ** mesh_size
    M, K, coords, bdofs = fem.assemble_q2_2d(nx, nx)
    beta = 1e-3
    n_t = 5
    tau = 0.25
    v_ref, zeta_ref = _kat_fields(coords[:, 0], coords[:, 1], tau, n_t)
    b_0, b_1 = _kat_rows(M, K, v_ref, zeta_ref, CN, tau, beta)
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


def instationary_stokes_kat(CN, nx=4, beta=1e-2):
    """A known-answer problem for the instationary Stokes system in the style of the reference's own
    KATs (it ships none with assertions for this system: test/test_control.py:3045-3302 only run):
    vector Q2 - Q1 on nx x nx quads of the unit square (the spaces of test/test_control.py:232-240),
    n_t = 5, tau = 0.25, the analytic v_ref / zeta_ref of the heat KATs in both components, analytic
    p_ref / mu_ref per time block, and right-hand sides built ROW BY ROW from the block stencils
    (velocity rows: ``_kat_rows`` + tau B^T mu_i / tau B^T p_i; pressure rows tau B v_i / tau B zeta_i,
    control/control.py:3750-3769), handed over as ready blocks (v_d=, f=, div_v=, div_zeta=), which
    ``incompressible_linear_solve`` T-transforms itself (control/control.py:4230-4234)."""
    sq = fem.assemble_q2q1_stokes_2d(nx, nx)
    M, K, B, bd = sq["M_v"], sq["L_v"], sq["B"], sq["bdofs_v"]
    n_t, tau = 5, 0.25
    N = n_t - 1 if CN else n_t
    x, y = sq["coords_v"][:, 0], sq["coords_v"][:, 1]
    px, py = sq["coords_p"][:, 0], sq["coords_p"][:, 1]
    va, za = _kat_fields(x, y, tau, n_t)
    v_ref = np.zeros((n_t, M.shape[0]))
    zeta_ref = np.zeros((n_t, M.shape[0]))
    v_ref[:, 0::2], v_ref[:, 1::2] = va, 0.5 * za[::-1]          # second component: another mix of the same fields,
    v_ref[0] = 0.0                                                # zero initial condition as in the heat KATs
    zeta_ref[:, 0::2], zeta_ref[:, 1::2] = za, 0.5 * va[::-1]
    if CN:
        zeta_ref[n_t - 1] = 0.0                                   # CN unknowns: v_1..v_N, zeta_0..zeta_{N-1}
    b_0, b_1 = _kat_rows(M, K, v_ref, zeta_ref, CN, tau, beta)
    p_ref = np.stack([tau ** i * np.sin(np.pi * px) * np.sin(2.0 * np.pi * py) for i in range(N)])
    mu_ref = np.stack([tau ** i * px * np.exp(py) for i in range(N)])
    v_unknown = v_ref[1:] if CN else v_ref                        # unknown block i of the state
    z_unknown = zeta_ref[:-1] if CN else zeta_ref
    b_0 = b_0 + tau * (B.T @ mu_ref.T).T                          # block_01[(i, i)] = tau B^T on the first N rows
    b_1 = b_1 + tau * (B.T @ p_ref.T).T
    div_v = tau * (B @ v_unknown.T).T                             # block_10[(i, i)] = tau B
    div_zeta = tau * (B @ z_unknown.T).T
    solver_parameters = {"linear_solver": "fgmres", "gmres_restart": 100, "maximum_iterations": 500,
                         "relative_tolerance": 1.0e-13, "absolute_tolerance": 1.0e-14}
    return dict(sq=sq, M=M, K=K, B=B, bdofs=bd, beta=beta, n_t=n_t, tau=tau, CN=CN, N=N, v_ref=v_ref, zeta_ref=zeta_ref,
                p_ref=p_ref, mu_ref=mu_ref, v_unknown=v_unknown, z_unknown=z_unknown, v_d=b_0, f=b_1, div_v=div_v,
                div_zeta=div_zeta, lambda_v_bounds=(0.25, 1.5625), lambda_p_bounds=(0.25, 2.25),
                solver_parameters=solver_parameters)


def reference_stokes_exact_problem(CN):
    """The problem of the reference's two instationary Stokes tests with an exact solution
    (test/test_control.py:3045-3172 BE: 8x8 quads, n_t = 20; 3175-3302 CN: 16x16 quads, n_t = 10): vector
    Q2 - Q1 on (0, 2)^2, beta = 1, K = vector Laplacian, the analytic desired state and force of the test
    (interpolated, then tested with the mass matrix), time-dependent inhomogeneous Dirichlet data = the
    exact velocity, initial condition = the exact velocity at t = 0, Chebyshev bounds (0.25, 1.5625) /
    (0.25, 2.25).  The reference only runs these (no assertion); the exact velocity is returned so that
    the discretisation error can be looked at."""
    nx = 16 if CN else 8
    n_t = 10 if CN else 20
    sq = fem.assemble_q2q1_stokes_2d(nx, nx, 2.0, 2.0)
    M = sq["M_v"]
    x, y = sq["coords_v"][:, 0] - 1.0, sq["coords_v"][:, 1] - 1.0
    T_f, beta = 1.0, 1.0
    tau = T_f / (n_t - 1.0)
    times = tau * np.arange(n_t)

    def vec(cx, cy):
        a = np.zeros(M.shape[0])
        a[0::2], a[1::2] = cx, cy
        return a
    v_d_help = vec(4.0 * beta * y * (2.0 * (3.0 * x * x - 1.0) * (y * y - 1.0) + 3.0 * (x * x - 1.0) ** 2),
                   -4.0 * beta * x * (3.0 * (y * y - 1.0) ** 2 + 2.0 * (x * x - 1.0) * (3.0 * y * y - 1.0)))
    f_help = vec(2.0 * y * (x ** 2 - 1.0) ** 2 * (y ** 2 - 1.0), -2.0 * x * (x ** 2 - 1.0) * (y ** 2 - 1.0) ** 2)
    v_hat, f_nodal, true_v = [], [], []
    for t in times:
        e = np.exp(T_f - t)
        v_hat.append(vec(e * (x * y ** 3 + 2.0 * beta * y * (((x * x - 1.0) ** 2) * (y * y - 7.0)
                                                          - 4.0 * (3.0 * x * x - 1.0) * (y * y - 1.0) + 2.0)),
                         e * (0.25 * (x ** 4 - y ** 4) - 2.0 * beta * x * (((y * y - 1.0) ** 2) * (x * x - 7.0)
                                                                       - 4.0 * (x * x - 1.0) * (3.0 * y * y - 1.0) - 2.0)))
                     + v_d_help)
        f_nodal.append(vec(e * (-x * y ** 3 - 2.0 * y * (x * x - 1.0) ** 2 * (y * y - 1.0)),
                           e * (0.25 * (y ** 4 - x ** 4) + 2.0 * x * (x * x - 1.0) * (y * y - 1.0) ** 2)) + f_help)
        true_v.append(vec(e * x * y ** 3, 0.25 * e * (x ** 4 - y ** 4)))
    v_hat, f_nodal, true_v = np.stack(v_hat), np.stack(f_nodal), np.stack(true_v)
    bd = sq["bdofs_v"]
    return dict(sq=sq, M=M, K=sq["L_v"], B=sq["B"], bdofs=bd, beta=beta, n_t=n_t, tau=tau, CN=CN,
                time_interval=(0.0, T_f), v_hat=v_hat, v_d=(M @ v_hat.T).T, f=(M @ f_nodal.T).T, true_v=true_v,
                bc_values=true_v[:, bd], v_0=true_v[0], lambda_v_bounds=(0.25, 1.5625), lambda_p_bounds=(0.25, 2.25))


def mms_heat_problem(N, n_t=100, CN=True):
    """The manufactured solution of the reference's heat-control convergence studies
    (test/test_control.py:1983-2138 CN / 1658-1826 BE, degree 1): P1 on an N x N mesh of (0, 2)^2,
    beta = 1, t_f = 2, v = 1 + (c_1 + c_2(t)) cos cos, zeta = (e^{t_f} - e^t) cos cos, f = 0, Dirichlet
    data v = 1 on the boundary (inhomogeneous), initial condition = v(0).  The reference prints observed
    orders without asserting them."""
    M, K, coords, bd = fem.assemble_p1_2d(N, N, 2.0, 2.0)
    x, y = coords[:, 0] - 1.0, coords[:, 1] - 1.0
    beta, t_f = 1.0, 2.0
    tau = t_f / (n_t - 1.0)
    times = tau * np.arange(n_t)
    cc = np.cos(0.5 * np.pi * x) * np.cos(0.5 * np.pi * y)
    pi2 = np.pi * np.pi

    def v_exact(t):
        return 1.0 + ((2.0 / (pi2 * beta)) * np.exp(t_f) - (2.0 / ((2.0 + pi2) * beta)) * np.exp(t)) * cc

    def zeta_exact(t):
        return (np.exp(t_f) - np.exp(t)) * cc

    def v_hat(t):                                   # desired state, test/test_control.py:2016-2022
        c = (2.0 / (pi2 * beta) + 0.5 * pi2) * np.exp(t_f) + (1.0 - 2.0 / ((2.0 + pi2) * beta) - 0.5 * pi2) * np.exp(t)
        return 1.0 + c * cc
    vh = np.stack([v_hat(t) for t in times])
    return dict(M=M, K=K, coords=coords, bdofs=bd, beta=beta, n_t=n_t, tau=tau, CN=CN, time_interval=(0.0, t_f),
                v_hat=vh, v_d=(M @ vh.T).T, f=np.zeros((n_t, M.shape[0])), v_0=v_exact(0.0),
                bc_values=np.ones((n_t, bd.size)), v_exact=np.stack([v_exact(t) for t in times]),
                zeta_exact=np.stack([zeta_exact(t) for t in times]))


def mms_convection_diffusion_problem(N, n_t=100, CN=True):
    """The manufactured solution of the reference's convection-diffusion control studies
    (test/test_control.py:2675-2857 CN / 2297-2492 BE, degree 1): the fields of ``mms_heat_problem`` with the
    time-dependent, divergence-free wind cos(pi t / 2) (2 y (1 - x^2), -2 x (1 - y^2)) in the forward
    operator, i.e. one NON-SYMMETRIC matrix K_i = L + C(t_i) per time level; desired state and force carry
    the convection of zeta and v (2707-2717, 2763-2774)."""
    q = mms_heat_problem(N, n_t, CN)
    M, L, coords = q["M"], q["K"], q["coords"]
    x, y = coords[:, 0] - 1.0, coords[:, 1] - 1.0
    tau, t_f, beta = q["tau"], 2.0, q["beta"]
    times = tau * np.arange(n_t)
    pi2 = np.pi * np.pi

    def wind_at(t):
        a = np.cos(0.5 * np.pi * t)
        return lambda X, Y: (a * 2.0 * (Y - 1.0) * (1.0 - (X - 1.0) ** 2), -a * 2.0 * (X - 1.0) * (1.0 - (Y - 1.0) ** 2))
    K_levels = []
    for t in times:
        C = fem.assemble_convection_p1_2d(N, N, 2.0, 2.0, wind_at(t))
        assert np.array_equal(C.indices, M.indices) and np.array_equal(L.indices, M.indices)
        K_levels.append(sp.csr_matrix((L.data + C.data, M.indices, M.indptr), shape=M.shape))
    sx, cx = np.sin(0.5 * np.pi * x), np.cos(0.5 * np.pi * x)
    sy, cy = np.sin(0.5 * np.pi * y), np.cos(0.5 * np.pi * y)
    gx, gy = -0.5 * np.pi * sx * cy, -0.5 * np.pi * cx * sy          # gradient of cos cos
    v_hat, f_nodal = [], []
    for t in times:
        a = np.cos(0.5 * np.pi * t)
        wx, wy = a * 2.0 * y * (1.0 - x * x), -a * 2.0 * x * (1.0 - y * y)
        conv_cc = gx * wx + gy * wy                                  # wind . grad(cos cos)
        cz = np.exp(t_f) - np.exp(t)
        cv = (2.0 / (pi2 * beta)) * np.exp(t_f) - (2.0 / ((2.0 + pi2) * beta)) * np.exp(t)
        v_hat.append(q["v_hat"][len(v_hat)] - cz * conv_cc)          # 2707-2724
        f_nodal.append(cv * conv_cc)                                 # 2769-2774
    v_hat, f_nodal = np.stack(v_hat), np.stack(f_nodal)
    q = dict(q)
    q.update(K_levels=K_levels, v_hat=v_hat, v_d=(M @ v_hat.T).T, f=(M @ f_nodal.T).T)
    return q


def reference_navier_stokes_problem(CN, nx=8, n_t=10):
    """The problem of the reference's instationary Navier-Stokes control tests (test/test_control.py:4171-4268
    BE, 4271-4368 CN; they run ``incompressible_non_linear_solve`` and assert nothing): (0, 2)^2, n_t = 10 on
    (0, 2), beta = 1e-3, nu = 1/100, forward form ``nu grad.grad + dot(grad(trial), u) . test`` (Picard), lid
    velocity (min(t, 1), 0) on the top edge and no-slip elsewhere, the two counter-rotating vortices as desired
    state, zero force and initial condition.  Vector Q2 - Q1 on quadrilaterals here (the reference: P2 - P1
    on triangles).  ``D_v(v_i, t)`` / ``D_p(v_i, t)`` are the matrices of ``construct_D_v`` on the velocity /
    pressure space at the velocity iterate."""
    sq = fem.assemble_q2q1_stokes_2d(nx, nx, 2.0, 2.0)
    M = sq["M_v"]
    n_v = M.shape[0]
    xs, ys = sq["coords_v"][:, 0], sq["coords_v"][:, 1]
    x, y = xs - 1.0, ys - 1.0
    T_f, beta, nu = 2.0, 1.0e-3, 1.0 / 100.0
    tau = T_f / (n_t - 1.0)
    times = tau * np.arange(n_t)
    a, b = (100.0 / 49.0) ** 2, (100.0 / 99.0) ** 2
    c_1 = 1.0 - np.sqrt(a * (x - 0.5) ** 2 + b * y ** 2)
    c_2 = 1.0 - np.sqrt(a * (x + 0.5) ** 2 + b * y ** 2)
    shape = np.zeros(n_v)                                          # 4304-4328
    shape[0::2] = np.where(c_1 >= 0.0, c_1 * b * y, np.where(c_2 >= 0.0, -c_2 * b * y, 0.0))
    shape[1::2] = np.where(c_1 >= 0.0, -c_1 * a * (x - 0.5), np.where(c_2 >= 0.0, c_2 * a * (x + 0.5), 0.0))
    v_hat = np.cos(0.5 * np.pi * times)[:, None] * shape[None, :]
    bd = sq["bdofs_v"]
    top = np.abs(ys[bd // 2] - 2.0) < 1e-12                       # boundary id 4 of RectangleMesh: y = ly
    g = np.zeros((n_t, bd.size))
    g[:, top & (bd % 2 == 0)] = np.minimum(times, 1.0)[:, None]    # 4282-4289
    conv_v = fem.convection_q2_2d(nx, nx, 2.0, 2.0)
    conv_p = fem.convection_q1_q2wind_2d(nx, nx, 2.0, 2.0)
    I2 = sp.identity(2, format="csr")
    L_v, L_p = sq["L_v"], sq["L_p"]

    def D_v(v_i, t):
        C = sp.kron(conv_v(v_i[0::2], v_i[1::2]), I2, format="csr")
        C.sort_indices()
        assert np.array_equal(C.indices, M.indices)
        return sp.csr_matrix((nu * L_v.data + C.data, M.indices, M.indptr), shape=M.shape)

    def D_p(v_i, t):
        C = conv_p(v_i[0::2], v_i[1::2])
        return sp.csr_matrix((nu * L_p.data + C.data, L_p.indices, L_p.indptr), shape=L_p.shape)
    return dict(sq=sq, M=M, B=sq["B"], D_v=D_v, D_p=D_p, bdofs=bd, beta=beta, n_t=n_t, tau=tau, CN=CN,
                time_interval=(0.0, T_f), v_hat=v_hat, v_d=(M @ v_hat.T).T, f=np.zeros((n_t, n_v)), bc_values=g,
                lambda_v_bounds=(0.3924, 2.0598), lambda_p_bounds=(0.5, 2.0))


def mms_heat_problem_linear_in_time(N, n_t=10):
    """The manufactured solution of ``test_MMS_instationary_heat_control_BE_convergence_FE``
    (test/test_control.py:1658-1826, degree 1): v = 1 + (t_f - t) cos cos, zeta = (t_f - t) cos cos on (0, 2)^2,
    beta = 1, t_f = 2, n_t = 10, backward Euler -- LINEAR in time, so the time discretisation is exact and the
    error is the spatial one.  Desired state ``zeta_space - lapl(zeta) + v`` (1695-1708), force
    ``-v_space - lapl(v) - zeta / beta`` (1726-1739), Dirichlet data v = 1, initial condition v(0)."""
    M, K, coords, bd = fem.assemble_p1_2d(N, N, 2.0, 2.0)
    x, y = coords[:, 0] - 1.0, coords[:, 1] - 1.0
    beta, t_f = 1.0, 2.0
    tau = t_f / (n_t - 1.0)
    times = tau * np.arange(n_t)
    cc = np.cos(0.5 * np.pi * x) * np.cos(0.5 * np.pi * y)
    lam = 0.5 * np.pi * np.pi                                     # -lapl(cos cos) = lam cos cos
    v_exact = np.stack([1.0 + (t_f - t) * cc for t in times])
    zeta_exact = np.stack([(t_f - t) * cc for t in times])
    v_hat = np.stack([cc + lam * (t_f - t) * cc + 1.0 + (t_f - t) * cc for t in times])
    f_nodal = np.stack([-cc + lam * (t_f - t) * cc - (t_f - t) * cc / beta for t in times])
    return dict(M=M, K=K, coords=coords, bdofs=bd, beta=beta, n_t=n_t, tau=tau, CN=False, time_interval=(0.0, t_f),
                v_hat=v_hat, v_d=(M @ v_hat.T).T, f=(M @ f_nodal.T).T, v_0=v_exact[0],
                bc_values=np.ones((n_t, bd.size)), v_exact=v_exact, zeta_exact=zeta_exact)


def stokes_mms_fields(N, beta=1e-3):
    """Fields of test/test_control.py:361-551 on vector Q2 - Q1, (0, 2)^2 (coordinates shifted to (-1, 1)^2)."""
    sq = fem.assemble_q2q1_stokes_2d(N, N, 2.0, 2.0)
    M = sq["M_v"]
    x, y = sq["coords_v"][:, 0] - 1.0, sq["coords_v"][:, 1] - 1.0
    px, py = sq["coords_p"][:, 0] - 1.0, sq["coords_p"][:, 1] - 1.0

    def vec(cx, cy):
        a = np.zeros(M.shape[0])
        a[0::2], a[1::2] = cx, cy
        return a
    v = vec(x * y ** 3, 0.25 * (x ** 4 - y ** 4))                                  # div v = 0, -lapl v + grad p = 0
    zeta = vec(2.0 * beta * y * (x ** 2 - 1.0) ** 2 * (y ** 2 - 1.0), -2.0 * beta * x * (x ** 2 - 1.0) * (y ** 2 - 1.0) ** 2)
    lap_zeta = vec(2.0 * beta * (y * (y ** 2 - 1.0) * (12.0 * x ** 2 - 4.0) + 6.0 * y * (x ** 2 - 1.0) ** 2),
                   -2.0 * beta * (x * (x ** 2 - 1.0) * (12.0 * y ** 2 - 4.0) + 6.0 * x * (y ** 2 - 1.0) ** 2))
    grad_mu = vec(4.0 * beta * y, 4.0 * beta * x)
    v_hat = -lap_zeta + grad_mu + v                                                # 403-405
    f_nodal = -zeta / beta                                                          # -lapl v + grad p - zeta / beta, 424-425
    return dict(sq=sq, M=M, v=v, zeta=zeta, lap_zeta=lap_zeta, grad_mu=grad_mu, p=3.0 * px ** 2 * py - py ** 3, mu=4.0 * beta * px * py, v_hat=v_hat,
                f_nodal=f_nodal, beta=beta)


def mms_convection_diffusion_problem_be(N, n_t=10):
    """``test_MMS_instationary_convection_diffusion_control_BE_convergence_FE`` (test/test_control.py:2297-2492,
    degree 1): the linear-in-time fields of ``mms_heat_problem_linear_in_time`` with the time-dependent wind of the
    convection-diffusion studies in the forward operator (one non-symmetric ``K_i`` per level); the desired state
    loses, the force gains the convection of zeta / v (2342-2358, 2376-2395)."""
    q = dict(mms_heat_problem_linear_in_time(N, n_t))
    M, L, coords = q["M"], q["K"], q["coords"]
    x, y = coords[:, 0] - 1.0, coords[:, 1] - 1.0
    tau, t_f = q["tau"], 2.0
    times = tau * np.arange(n_t)
    sx, cx = np.sin(0.5 * np.pi * x), np.cos(0.5 * np.pi * x)
    sy, cy = np.sin(0.5 * np.pi * y), np.cos(0.5 * np.pi * y)
    gx, gy = -0.5 * np.pi * sx * cy, -0.5 * np.pi * cx * sy          # gradient of cos cos
    K_levels, v_hat, f_nodal = [], q["v_hat"].copy(), np.zeros_like(q["v_hat"])
    lam = 0.5 * np.pi * np.pi
    cc = cx * cy
    for i, t in enumerate(times):
        a = np.cos(0.5 * np.pi * t)
        C = fem.assemble_convection_p1_2d(
            N, N, 2.0, 2.0, lambda X, Y, a=a: (a * 2.0 * (Y - 1.0) * (1.0 - (X - 1.0) ** 2),
                                               -a * 2.0 * (X - 1.0) * (1.0 - (Y - 1.0) ** 2)))
        assert np.array_equal(C.indices, M.indices)
        K_levels.append(sp.csr_matrix((L.data + C.data, M.indices, M.indptr), shape=M.shape))
        conv_cc = gx * (a * 2.0 * y * (1.0 - x * x)) + gy * (-a * 2.0 * x * (1.0 - y * y))
        v_hat[i] -= (t_f - t) * conv_cc
        f_nodal[i] = -cc + lam * (t_f - t) * cc + (t_f - t) * conv_cc - (t_f - t) * cc / q["beta"]
    q.update(K_levels=K_levels, v_hat=v_hat, v_d=(M @ v_hat.T).T, f=(M @ f_nodal.T).T)
    return q
