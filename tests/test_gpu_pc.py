"""GPU parity of the in-built block preconditioner, its AMG inner solver and the Krylov
solve against the oracle (oracle/pc.py, oracle/amg.py, oracle/krylov.py), through the
C ABI.  fp64; tolerances are written next to each assertion."""
import numpy as np
import pytest
import torch

import kat
from oracle import amg as oamg
from oracle import control as ocontrol
from oracle import kkt
from synthetic import fem
from oracle import pc as opc

pytestmark = pytest.mark.gpu


def _system(q, CN, **kw):
    from control_b200 import MultiBlockSystem
    return MultiBlockSystem(q["M"], q.get("K_levels", q["K"]), n_t=q["n_t"], beta=q["beta"], CN=CN,
                            time_interval=q["time_interval"], bc_dofs=q["bdofs"], **kw)


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


# ------------------------------------------------------------------ AMG
def test_amg_hierarchy_matches_oracle_setup():
    q = kat.heat_problem(40, 8, True)
    s = _system(q, True)
    s.setup_preconditioner(lambda_v_bounds=q["lambda_v_bounds"], coarse_max=50)
    tau, beta = s.tau, q["beta"]
    c = 0.5 * tau / beta ** 0.5
    A = fem.assemble_bc((0.5 * tau * q["K"] + (1 + c) * q["M"]).tocsr(), q["bdofs"])
    H = oamg.setup(A, coarse_max=50)
    G = s.amg_hierarchy(0)
    assert len(G) >= 3
    assert [e["n"] for e in G] == [L.A.shape[0] for L in H.levels]
    for e, L in zip(G, H.levels):
        assert abs(e["A"] - L.A).max() <= 1e-13 * abs(L.A).max()
        if L.P is not None:
            assert np.array_equal(e["agg"], L.agg)          # aggregates: bit-exact indexing
            assert abs(e["P"] - L.P).max() <= 1e-13 * abs(L.P).max()
    # two V-cycles on a random right-hand side
    rng = np.random.default_rng(0)
    b = rng.standard_normal(A.shape[0])
    b[q["bdofs"]] = 0.0
    x_ref = oamg.solve(H, b)
    x = s.amg_solve(torch.from_numpy(b).to(s.device)).cpu().numpy()
    assert _rel(x, x_ref) < 1e-11
    s.close()


# ------------------------------------------------------------------ pc_fn
def _oracle_pc(q, CN, **kw):
    K = q.get("K_levels", q["K"])
    return opc.construct_pc(q["M"], K, q["tau"], q["beta"], q["n_t"], CN, q["bdofs"], **kw)


@pytest.mark.parametrize("CN", [True, False])
@pytest.mark.parametrize("s0", ["chebyshev", "jacobi", "multigrid"])
def test_pc_fn_matches_oracle(CN, s0):
    q = kat.heat_problem(24, 7, CN, beta=1e-3)
    s = _system(q, CN)
    bounds = q["lambda_v_bounds"] if s0 == "chebyshev" else None
    s.setup_preconditioner(lambda_v_bounds=bounds, Multigrid=(s0 == "multigrid"))
    pc = _oracle_pc(q, CN, lambda_v_bounds=bounds, Multigrid=(s0 == "multigrid"))
    rng = np.random.default_rng(1)
    b0 = rng.standard_normal((s.N, s.n))
    b1 = rng.standard_normal((s.N, s.n))
    b0[:, q["bdofs"]] = 0.0
    b1[:, q["bdofs"]] = 0.0
    r0, r1 = pc(b0, b1)
    g0, g1 = s.to_host_blocks(s.pc_apply(s.to_device(b0, b1), raw=True))
    assert _rel(g0, r0) < 1e-11 and _rel(g1, r1) < 1e-11
    # Preconditioner.apply: constrained entries of u take b's values
    b0w = b0 + 0.0
    b1w = b1 + 0.0
    b0w[:, q["bdofs"]] = 3.0
    b1w[:, q["bdofs"]] = -2.0
    w0, w1 = s.to_host_blocks(s.pc_apply(s.to_device(b0w, b1w)))
    assert np.array_equal(w0[:, q["bdofs"]], b0w[:, q["bdofs"]])
    assert np.array_equal(w1[:, q["bdofs"]], b1w[:, q["bdofs"]])
    mask = np.ones(s.n, bool)
    mask[q["bdofs"]] = False
    assert _rel(w0[:, mask], r0[:, mask]) < 1e-11 and _rel(w1[:, mask], r1[:, mask]) < 1e-11
    # the same through the P-compatible host callable
    u0, u1 = np.zeros_like(b0), np.zeros_like(b1)
    s.pc_fn()(u0, u1, b0, b1)
    assert np.array_equal(u0, g0) and np.array_equal(u1, g1)
    s.close()


@pytest.mark.parametrize("CN", [True, False])
def test_pc_fn_time_dependent_nonsymmetric_K(CN):
    q = kat.heat_problem(16, 6, CN, beta=1e-2)
    rng = np.random.default_rng(7)
    Ks = []
    for _ in range(q["n_t"]):
        Ki = q["K"].copy()
        Ki.data = Ki.data * (1.0 + 0.1 * rng.standard_normal(Ki.nnz))
        Ks.append(Ki)
    q["K_levels"] = Ks
    s = _system(q, CN)
    s.setup_preconditioner(lambda_v_bounds=q["lambda_v_bounds"])
    pc = _oracle_pc(q, CN, lambda_v_bounds=q["lambda_v_bounds"])
    b0 = rng.standard_normal((s.N, s.n))
    b1 = rng.standard_normal((s.N, s.n))
    b0[:, q["bdofs"]] = 0.0
    b1[:, q["bdofs"]] = 0.0
    r0, r1 = pc(b0, b1)
    g0, g1 = s.to_host_blocks(s.pc_apply(s.to_device(b0, b1), raw=True))
    assert _rel(g0, r0) < 1e-11 and _rel(g1, r1) < 1e-10
    s.close()


def test_pc_diagonal_mode_matches_oracle_and_is_symmetric():
    q = kat.heat_problem(20, 9, True)
    s = _system(q, True)
    s.setup_preconditioner(lambda_v_bounds=q["lambda_v_bounds"], mode="diagonal")
    pc = opc.construct_pc_diagonal(q["M"], q["K"], q["tau"], q["beta"], q["n_t"], q["bdofs"],
                                   lambda_v_bounds=q["lambda_v_bounds"])
    rng = np.random.default_rng(2)
    vecs = []
    for _ in range(2):
        b0 = rng.standard_normal((s.N, s.n))
        b1 = rng.standard_normal((s.N, s.n))
        b0[:, q["bdofs"]] = 0.0
        b1[:, q["bdofs"]] = 0.0
        vecs.append((b0, b1))
    outs = []
    for b0, b1 in vecs:
        r0, r1 = pc(b0, b1)
        g0, g1 = s.to_host_blocks(s.pc_apply(s.to_device(b0, b1), raw=True))
        assert _rel(g0, r0) < 1e-11 and _rel(g1, r1) < 1e-11
        outs.append((g0, g1))
    lhs = (vecs[0][0] * outs[1][0]).sum() + (vecs[0][1] * outs[1][1]).sum()
    rhs = (vecs[1][0] * outs[0][0]).sum() + (vecs[1][1] * outs[0][1]).sum()
    assert abs(lhs - rhs) < 1e-9 * abs(lhs)
    s.close()


# ------------------------------------------------------------------ solve
@pytest.mark.parametrize("CN", [False, True])
def test_reference_known_answer_on_gpu(CN):
    """test/test_control.py:1243-1444 (BE) / 1447-1655 (CN) through the CUDA path."""
    p = kat.instationary_kat(CN)
    from control_b200 import MultiBlockSystem
    s = MultiBlockSystem(p["M"], p["K"], n_t=p["n_t"], beta=p["beta"], CN=CN, bc_dofs=p["bdofs"])
    s.setup_preconditioner(lambda_v_bounds=p["lambda_v_bounds"])
    b0, b1 = p["b_0"], p["b_1"]
    if CN:
        b0, b1 = kkt.apply_T_1(b0), kkt.apply_T_2(b1)          # control/control.py:3242-3243
    u0 = np.zeros((s.N, s.n))
    u1 = np.zeros((s.N, s.n))
    info = s.solve(u0, u1, b0, b1, solver_parameters=p["solver_parameters"], pc_fn="builtin")
    assert info.reason > 0
    v, zeta = kkt.unpack_solution(u0, u1, p["n_t"], CN)
    assert kat.l2_error(p["M"], v, p["v_ref"]) < 5e-13
    assert kat.l2_error(p["M"], zeta, p["zeta_ref"]) < 5e-13
    ref = ocontrol.linear_solve(p["M"], p["K"], beta=p["beta"], n_t=p["n_t"], CN=CN, bdofs=p["bdofs"],
                                v_d=p["b_0"], f=p["b_1"], check_v_d=False, check_f=False,
                                solver_parameters=p["solver_parameters"],
                                lambda_v_bounds=p["lambda_v_bounds"])
    # at rtol 1e-14 the last iterations sit on the rounding floor (the monitored residual stagnates for twenty and
    # more iterations between 1e-11 and 1e-14 of its start, and where it crosses a threshold there differs between
    # two summation orders): +-2 on the total here; +-1 is asserted wherever the count is a property of the method --
    # test_heat_control_solve_matches_oracle below (rtol 1e-9), tests/test_gpu_fullsize.py (256^2 x 64, C2, C3)
    assert abs(info.its - ref["ksp"].its) <= 2
    s.close()


@pytest.mark.parametrize("ksp,mode", [("gmres", "triangular"), ("fgmres", "triangular"), ("minres", "diagonal")])
@pytest.mark.parametrize("nx,n_t", [(10, 10), (32, 17)])
def test_heat_control_solve_matches_oracle(ksp, mode, nx, n_t):
    """BASELINE config C1 (and a larger sibling): iteration counts within +-1, residual
    history, solution, KKT residual and objective against the oracle running the same
    preconditioner."""
    q = kat.heat_problem(nx, n_t, True)
    sp_ = {"linear_solver": ksp, "gmres_restart": 10, "maximum_iterations": 100,
           "relative_tolerance": 1e-9, "absolute_tolerance": 0.0}
    ref = ocontrol.linear_solve(q["M"], q["K"], beta=q["beta"], n_t=n_t, CN=True,
                                time_interval=q["time_interval"], bdofs=q["bdofs"], v_d=q["v_d"],
                                f=q["f"], lambda_v_bounds=q["lambda_v_bounds"], solver_parameters=sp_,
                                pc_mode=mode)
    s = _system(q, True)
    s.setup_preconditioner(lambda_v_bounds=q["lambda_v_bounds"], mode=mode)
    u0 = np.zeros((s.N, s.n))
    u1 = np.zeros((s.N, s.n))
    info = s.solve(u0, u1, ref["b_0"], ref["b_1"], solver_parameters=sp_, pc_fn="builtin")
    assert info.reason == ref["ksp"].reason
    assert abs(info.its - ref["ksp"].its) <= 1
    k = min(len(info.history), len(ref["ksp"].history)) - 1
    assert np.allclose(info.history[:k], ref["ksp"].history[:k], rtol=1e-6)
    assert _rel(u0, ref["v_blocks"]) < 1e-7 and _rel(u1, ref["zeta_blocks"]) < 1e-7
    res = s.residual_norm(s.to_device(ref["b_0"], ref["b_1"]), s.to_device(u0, u1))
    assert res <= 10 * max(ref["kkt_residual"], 1e-12 * np.linalg.norm(ref["b_1"]))
    v, zeta = kkt.unpack_solution(u0, u1, n_t, True, np.zeros(s.n))
    J = s.objective(v, zeta, q["v_hat"])
    J_ref = ocontrol.objective(q["M"], ref["v"], ref["zeta"], q["v_hat"], q["tau"], q["beta"], True)
    assert abs(J - J_ref) <= 1e-8 * abs(J_ref)
    s.close()


def test_backward_euler_solve_matches_oracle():
    q = kat.heat_problem(16, 8, False, beta=1e-2)
    # rtol 1e-7: below ~1e-8 the BE histories sit on the rounding floor of classical
    # Gram-Schmidt (the system is ill-conditioned through the epsilon-regularised last block)
    # and iteration counts stop being comparable between two summation orders
    sp_ = {"linear_solver": "fgmres", "maximum_iterations": 200, "relative_tolerance": 1e-7,
           "absolute_tolerance": 0.0, "gmres_restart": 100}
    ref = ocontrol.linear_solve(q["M"], q["K"], beta=q["beta"], n_t=q["n_t"], CN=False,
                                time_interval=q["time_interval"], bdofs=q["bdofs"], v_d=q["v_d"],
                                f=q["f"], lambda_v_bounds=q["lambda_v_bounds"], solver_parameters=sp_)
    s = _system(q, False)
    s.setup_preconditioner(lambda_v_bounds=q["lambda_v_bounds"])
    u0 = np.zeros((s.N, s.n))
    u1 = np.zeros((s.N, s.n))
    info = s.solve(u0, u1, ref["b_0"], ref["b_1"], solver_parameters=sp_, pc_fn="builtin")
    assert info.reason > 0 and abs(info.its - ref["ksp"].its) <= 1
    assert _rel(u0, ref["v_blocks"]) < 1e-5 and _rel(u1, ref["zeta_blocks"]) < 1e-5
    s.close()


def test_user_preconditioner_hook_and_identity():
    """`P=` hook (control/control.py:3245-3258): any host callable pc_fn(u_0, u_1, b_0, b_1)."""
    q = kat.heat_problem(8, 5, True, beta=1e-2)
    sp_ = {"linear_solver": "fgmres", "maximum_iterations": 60, "relative_tolerance": 1e-8,
           "absolute_tolerance": 0.0}
    pc = _oracle_pc(q, True, lambda_v_bounds=q["lambda_v_bounds"], inner="exact")
    ref = ocontrol.linear_solve(q["M"], q["K"], beta=q["beta"], n_t=q["n_t"], CN=True,
                                time_interval=q["time_interval"], bdofs=q["bdofs"], v_d=q["v_d"],
                                f=q["f"], solver_parameters=sp_, P=pc)
    calls = []

    def P(u_0, u_1, b_0, b_1):
        calls.append(1)
        r0, r1 = pc(b_0, b_1)
        u_0[:] = r0
        u_1[:] = r1

    s = _system(q, True)
    u0 = np.zeros((s.N, s.n))
    u1 = np.zeros((s.N, s.n))
    info = s.solve(u0, u1, ref["b_0"], ref["b_1"], solver_parameters=sp_, pc_fn=P)
    assert len(calls) == info.n_pc and info.its == ref["ksp"].its
    assert _rel(u0, ref["v_blocks"]) < 1e-7
    # identity preconditioner + too few iterations: the reference's error policy
    sp2 = dict(sp_, maximum_iterations=3)
    with pytest.raises(RuntimeError, match="Solver failed to converge"):
        s.solve(np.zeros((s.N, s.n)), np.zeros((s.N, s.n)), ref["b_0"], ref["b_1"], solver_parameters=sp2)
    info = s.solve(np.zeros((s.N, s.n)), np.zeros((s.N, s.n)), ref["b_0"], ref["b_1"],
                   solver_parameters=dict(sp2, preconditioner=True))
    assert info.reason == -3 and info.its == 3          # DIVERGED_ITS, swallowed (preconditioner.py:756)

    def bad(u_0, u_1, b_0, b_1):
        raise ValueError("boom")
    with pytest.raises(RuntimeError):
        s.solve(np.zeros((s.N, s.n)), np.zeros((s.N, s.n)), ref["b_0"], ref["b_1"], solver_parameters=sp_, pc_fn=bad)
    s.close()


def test_shell_contexts_follow_the_petsc_python_protocol():
    """mult(A, x, y) / apply(pc, x, y) on host vectors (preconditioner.py:376, 563)."""
    q = kat.heat_problem(6, 4, True, beta=1e-2)
    s = _system(q, True)
    s.setup_preconditioner(lambda_v_bounds=q["lambda_v_bounds"])
    rng = np.random.default_rng(0)
    x = rng.standard_normal(2 * s.N * s.n)
    y = np.zeros_like(x)
    s.matshell().mult(None, x, y)
    y0, y1 = kkt.kkt_apply_fused(q["M"], q["K"], s.tau, q["beta"], q["n_t"], True, q["bdofs"],
                                 x[:s.N * s.n].reshape(s.N, s.n), x[s.N * s.n:].reshape(s.N, s.n))
    assert _rel(y, np.concatenate([y0.ravel(), y1.ravel()])) < 1e-13
    z = np.zeros_like(x)
    s.pcshell().apply(None, x, z)
    z2 = s.pc_apply(s.to_device(x)).cpu().numpy()
    assert np.array_equal(z, z2)
    s.close()


def test_accelerated_amg_solve_matches_oracle():
    """Chebyshev-accelerated V-cycles (oracle/amg.py::solve with acc_lo > 0)."""
    q = kat.heat_problem(40, 6, True)
    s = _system(q, True)
    params = dict(cycles=4, nu=3, acc_lo=0.5, acc_hi=1.0, coarse_max=50)
    s.setup_preconditioner(lambda_v_bounds=q["lambda_v_bounds"], **params)
    c = 0.5 * s.tau / q["beta"] ** 0.5
    A = fem.assemble_bc((0.5 * s.tau * q["K"] + (1 + c) * q["M"]).tocsr(), q["bdofs"])
    H = oamg.setup(A, **params)
    rng = np.random.default_rng(3)
    b = rng.standard_normal(A.shape[0])
    b[q["bdofs"]] = 0.0
    x_ref = oamg.solve(H, b)
    x = s.amg_solve(torch.from_numpy(b).to(s.device)).cpu().numpy()
    assert _rel(x, x_ref) < 1e-11
    xs = np.linalg.solve(A.toarray(), b)
    assert np.linalg.norm(x - xs) / np.linalg.norm(xs) < 5e-3
    # and inside the preconditioner
    pc = _oracle_pc(q, True, lambda_v_bounds=q["lambda_v_bounds"], amg_params=params)
    b0 = rng.standard_normal((s.N, s.n))
    b1 = rng.standard_normal((s.N, s.n))
    b0[:, q["bdofs"]] = 0.0
    b1[:, q["bdofs"]] = 0.0
    r0, r1 = pc(b0, b1)
    g0, g1 = s.to_host_blocks(s.pc_apply(s.to_device(b0, b1), raw=True))
    assert _rel(g0, r0) < 1e-11 and _rel(g1, r1) < 1e-11
    s.close()


@pytest.mark.parametrize("params", [dict(nu=3, nu_fine=2, cycles=2, coarse_max=50),
                                    dict(nu=2, nu_fine=4, cycles=1, coarse_max=50),
                                    dict(nu=4, nu_fine=1, cycles=3, coarse_max=50)])
def test_amg_with_separate_fine_level_degree(params):
    """nu_fine: smoother degree on level 0 differs from the coarse levels (buffer rotation of the
    smoother depends on the degree)."""
    q = kat.heat_problem(40, 6, True)
    s = _system(q, True)
    s.setup_preconditioner(lambda_v_bounds=q["lambda_v_bounds"], **params)
    c = 0.5 * s.tau / q["beta"] ** 0.5
    A = fem.assemble_bc((0.5 * s.tau * q["K"] + (1 + c) * q["M"]).tocsr(), q["bdofs"])
    H = oamg.setup(A, **params)
    rng = np.random.default_rng(5)
    b = rng.standard_normal(A.shape[0])
    b[q["bdofs"]] = 0.0
    x = s.amg_solve(torch.from_numpy(b).to(s.device)).cpu().numpy()
    assert _rel(x, oamg.solve(H, b)) < 1e-11
    s.close()


@pytest.mark.parametrize("CN,n_t", [(True, 81), (False, 70), (True, 140)])
def test_many_time_blocks_pc_and_solve(CN, n_t):
    """N > 64: batched preconditioner kernels with 4 / 8 columns per lane, full solve."""
    q = kat.heat_problem(9, n_t, CN, beta=1e-2)
    s = _system(q, CN)
    s.setup_preconditioner(lambda_v_bounds=q["lambda_v_bounds"])
    pc = _oracle_pc(q, CN, lambda_v_bounds=q["lambda_v_bounds"])
    rng = np.random.default_rng(4)
    b0 = rng.standard_normal((s.N, s.n))
    b1 = rng.standard_normal((s.N, s.n))
    b0[:, q["bdofs"]] = 0.0
    b1[:, q["bdofs"]] = 0.0
    r0, r1 = pc(b0, b1)
    g0, g1 = s.to_host_blocks(s.pc_apply(s.to_device(b0, b1), raw=True))
    assert _rel(g0, r0) < 1e-10 and _rel(g1, r1) < 1e-10
    sp_ = {"linear_solver": "fgmres", "maximum_iterations": 200, "relative_tolerance": 1e-6,
           "absolute_tolerance": 0.0, "gmres_restart": 100}
    ref = ocontrol.linear_solve(q["M"], q["K"], beta=q["beta"], n_t=n_t, CN=CN, time_interval=q["time_interval"],
                                bdofs=q["bdofs"], v_d=q["v_d"], f=q["f"], lambda_v_bounds=q["lambda_v_bounds"],
                                solver_parameters=sp_)
    u0 = np.zeros((s.N, s.n))
    u1 = np.zeros((s.N, s.n))
    info = s.solve(u0, u1, ref["b_0"], ref["b_1"], solver_parameters=sp_, pc_fn="builtin")
    assert info.reason > 0 and abs(info.its - ref["ksp"].its) <= 1
    assert _rel(u0, ref["v_blocks"]) < 1e-4
    s.close()
