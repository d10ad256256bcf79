"""The BASELINE.json configurations as concrete synthetic problems (matrices + tested data)."""
import numpy as np

from . import fem


def heat_problem(nx, n_t, CN=True, beta=1e-4, length=2.0, T=2.0):
    """README heat-control problem (README.md:24-60, read as in SURVEY.md section 2.4):
    P1 on (0, length)^2, zero Dirichlet data, vhat = t cos(pi (x-1)/2) cos(pi (y-1)/2),
    f = cos(pi (x-1)/2) cos(pi (y-1)/2).  BASELINE configs C1 (nx=10, n_t=10) and
    C2 (nx=1024, n_t=64)."""
    M, K, coords, bdofs = fem.assemble_p1_2d(nx, nx, length, length)
    x, y = coords[:, 0], coords[:, 1]
    shape = np.cos(np.pi * (x - 1.0) / 2.0) * np.cos(np.pi * (y - 1.0) / 2.0)
    tau = T / (n_t - 1.0)
    t = tau * np.arange(n_t)
    v_hat = t[:, None] * shape[None, :]
    f_nodal = np.tile(shape, (n_t, 1))
    v_d = (M @ v_hat.T).T          # assemble(inner(v_d, test) * dx) with v_d interpolated
    f = (M @ f_nodal.T).T
    return dict(M=M, K=K, coords=coords, bdofs=bdofs, beta=beta, n_t=n_t, CN=CN,
                time_interval=(0.0, T), tau=tau, v_hat=v_hat, v_d=v_d, f=f,
                lambda_v_bounds=(0.5, 2.0))


def heat_problem_3d(nx, n_t, CN=False, beta=1e-4, length=1.0, T=1.0):
    """BASELINE config C3 family: 3-D heat control, P1 tetrahedra on the unit cube, zero Dirichlet
    data (nx=128, n_t=32, backward Euler in BASELINE.json).  Chebyshev bounds (0.5, 2.5): Wathen's
    bound for D^-1 M of P1 tetrahedra (not from the reference, which has no 3-D test)."""
    M, K, coords, bdofs = fem.assemble_p1_3d(nx, nx, nx, length, length, length)
    x, y, z = coords[:, 0], coords[:, 1], coords[:, 2]
    shape = np.sin(np.pi * x / length) * np.sin(np.pi * y / length) * np.sin(np.pi * z / length)
    tau = T / (n_t - 1.0)
    t = tau * np.arange(n_t)
    v_hat = t[:, None] * shape[None, :]
    f_nodal = np.tile(shape, (n_t, 1))
    return dict(M=M, K=K, coords=coords, bdofs=bdofs, beta=beta, n_t=n_t, CN=CN,
                time_interval=(0.0, T), tau=tau, v_hat=v_hat, v_d=(M @ v_hat.T).T, f=(M @ f_nodal.T).T,
                lambda_v_bounds=(0.5, 2.5))


def stokes_problem(nx, n_t, CN=True, beta=1.0, length=2.0, T=1.0):
    """Instationary Stokes control with homogeneous Dirichlet velocity data on (0, length)^2,
    Taylor-Hood P2-P1 (BASELINE config C4 family).  Fields follow the structure of the
    reference's exact-solution test (test/test_control.py:3045-3172): the boundary-vanishing,
    divergence-free parts of its desired state and force, with its exp(T_f - t) time factor
    (the reference's inhomogeneous boundary data are outside this round's scope)."""
    th = fem.assemble_taylor_hood_2d(nx, nx, length, length)
    xy = th["coords_v"]
    x, y = xy[:, 0] - 1.0, xy[:, 1] - 1.0
    n_v = th["M_v"].shape[0]
    tau = T / (n_t - 1.0)
    t = tau * np.arange(n_t)

    def field(cx, cy):
        a = np.zeros(n_v)
        a[0::2], a[1::2] = cx, cy
        return a
    # curl of psi = (x^2 - 1)^2 (y^2 - 1)^2 / 2: zero on the boundary of (-1, 1)^2, divergence free
    shape = field(2.0 * y * (x ** 2 - 1.0) ** 2 * (y ** 2 - 1.0), -2.0 * x * (x ** 2 - 1.0) * (y ** 2 - 1.0) ** 2)
    v_hat = np.exp(T - t)[:, None] * shape[None, :]
    f_nodal = (1.0 + t)[:, None] * shape[None, :]
    M = th["M_v"]
    return dict(th=th, M=M, K=th["K_v"], bdofs=th["bdofs_v"], beta=beta, n_t=n_t, CN=CN, time_interval=(0.0, T),
                tau=tau, v_hat=v_hat, v_d=(M @ v_hat.T).T, f=(M @ f_nodal.T).T,
                lambda_v_bounds=(0.3924, 2.0598), lambda_p_bounds=(0.5, 2.0))
