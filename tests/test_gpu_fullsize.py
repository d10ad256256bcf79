"""BASELINE config C2 at FULL size (2-D heat control, P1 on 1024x1024, n_t = 64, CN) through the
C ABI: direct comparison of the fused KKT apply with the oracle (one application is seconds on the
host), and the size-independent properties of the path -- linearity and symmetry of the operator,
symmetry and positivity of the block-diagonal preconditioner, a converged solve whose true
residual meets the tolerance, and J_h consistent with the objective evaluated by the oracle."""
import numpy as np
import pytest
import torch

import kat
from oracle import control as ocontrol
from oracle import kkt

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2():
    from control_b200 import MultiBlockSystem
    q = kat.heat_problem(1024, 64, True)
    s = MultiBlockSystem(q["M"], q["K"], n_t=q["n_t"], beta=q["beta"], CN=True, time_interval=q["time_interval"],
                         bc_dofs=q["bdofs"])
    yield q, s
    s.close()


def _dot(a, b):
    return float(torch.dot(a, b))


def test_c2_apply_matches_oracle_and_is_linear_and_symmetric(c2):
    q, s = c2
    g = torch.Generator(device=s.device).manual_seed(0)
    x = torch.randn(s.vec_len(), dtype=torch.float64, device=s.device, generator=g)
    y = torch.randn(s.vec_len(), dtype=torch.float64, device=s.device, generator=g)
    Ax, Ay = s.apply(x), s.apply(y)
    # direct parity with the oracle's fused restatement on the same input (1.06 GB vectors)
    x0, x1 = s.to_host_blocks(x)
    r0, r1 = kkt.kkt_apply_fused(q["M"], q["K"], q["tau"], q["beta"], q["n_t"], True, q["bdofs"], x0, x1)
    g0, g1 = s.to_host_blocks(Ax)
    scale = max(np.abs(r0).max(), np.abs(r1).max())
    assert np.abs(g0 - r0).max() <= 1e-13 * scale and np.abs(g1 - r1).max() <= 1e-13 * scale
    # linearity
    z = s.apply(2.0 * x - 3.0 * y)
    assert float((z - (2.0 * Ax - 3.0 * Ay)).abs().max()) <= 1e-12 * float(z.abs().max())
    # symmetry on the constrained subspace (the transformed KKT matrix is symmetric)
    mask = torch.ones(s.n, dtype=torch.float64, device=s.device)
    mask[torch.from_numpy(q["bdofs"].astype(np.int64)).to(s.device)] = 0.0
    xm = (x.view(2 * s.N, s.n) * mask).reshape(-1)
    ym = (y.view(2 * s.N, s.n) * mask).reshape(-1)
    lhs, rhs = _dot(s.apply(xm), ym), _dot(xm, s.apply(ym))
    assert abs(lhs - rhs) <= 1e-10 * abs(lhs)


def test_c2_diagonal_preconditioner_is_symmetric_positive(c2):
    q, s = c2
    s.setup_preconditioner(lambda_v_bounds=q["lambda_v_bounds"], mode="diagonal")
    g = torch.Generator(device=s.device).manual_seed(1)
    x = torch.randn(s.vec_len(), dtype=torch.float64, device=s.device, generator=g)
    y = torch.randn(s.vec_len(), dtype=torch.float64, device=s.device, generator=g)
    Px, Py = s.pc_apply(x), s.pc_apply(y)
    lhs, rhs = _dot(Px, y), _dot(x, Py)
    assert abs(lhs - rhs) <= 1e-9 * abs(lhs)          # fixed polynomial smoothers and cycles: a symmetric operator
    assert _dot(Px, x) > 0.0 and _dot(Py, y) > 0.0


def test_c2_solve_meets_tolerance_and_objective(c2):
    from control_b200.control import build_rhs
    q, s = c2
    s.setup_preconditioner(lambda_v_bounds=q["lambda_v_bounds"], mode="triangular")
    b0, b1 = build_rhs(q["M"], q["K"], q["tau"], q["n_t"], True, q["bdofs"], q["v_d"], q["f"], np.zeros(s.n))
    b = s.to_device(b0, b1)
    u = s.new_vector()
    sp_ = {"linear_solver": "fgmres", "gmres_restart": 30, "maximum_iterations": 60, "relative_tolerance": 1e-8,
           "absolute_tolerance": 0.0}
    info = s.solve_device(b, u, solver_parameters=sp_, pc="builtin")
    assert info.reason > 0 and info.its <= 25
    bnorm = float(b.norm())
    assert s.residual_norm(b, u) <= 2e-8 * bnorm        # true residual against the recurrence (CGS drift)
    v_blocks, z_blocks = s.to_host_blocks(u)
    v = np.concatenate([np.zeros((1, s.n)), v_blocks])
    zeta = np.concatenate([z_blocks, np.zeros((1, s.n))])
    J_gpu = s.objective(v, zeta, q["v_hat"])
    J_ref = ocontrol.objective(q["M"], v, zeta, q["v_hat"], q["tau"], q["beta"], True)
    assert abs(J_gpu - J_ref) <= 1e-12 * abs(J_ref)
    # the same on the device (ctl_objective): 64 levels x 1.05 M dofs without leaving HBM
    J_dev = s.objective_device(torch.from_numpy(v).to(s.device), torch.from_numpy(zeta).to(s.device),
                               torch.from_numpy(np.ascontiguousarray(q["v_hat"])).to(s.device))
    assert abs(J_dev - J_ref) <= 1e-12 * abs(J_ref)
    # the optimal state tracks the desired state: J is far below J(v = 0, zeta = 0)
    J0 = ocontrol.objective(q["M"], 0 * v, 0 * zeta, q["v_hat"], q["tau"], q["beta"], True)
    assert J_gpu < 0.5 * J0
