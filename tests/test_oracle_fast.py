"""The C / OpenMP restatement of the hot path (oracle/c/pc_omp.c, the CPU arm of bench.py) against the numpy oracle:
operator, AMG solve and the three preconditioner variants.  No GPU."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import kat  # noqa: E402
from oracle import amg as oamg  # noqa: E402
from oracle import kkt  # noqa: E402
from oracle import pc as opc  # noqa: E402

fastpc = pytest.importorskip("oracle.fastpc")


def rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


@pytest.mark.parametrize("CN,mode,cycles", [(True, "triangular", {}), (True, "diagonal", {}), (False, "triangular", {}),
                                            (True, "diagonal", dict(cycles=2, nu=4))])      # last: bench.py's C2 arm
def test_c_port_matches_numpy_oracle(CN, mode, cycles):
    try:
        fastpc.lib()
    except ImportError:
        pytest.skip("oracle/_build/liboracle.so not built")
    fastpc.set_threads(4)
    q = kat.heat_problem(30, 7, CN, beta=1e-3)
    amg = dict(coarse_max=40, **cycles)
    M, K, bd = q["M"], q["K"], q["bdofs"]
    f = fastpc.FastPc(M, K, q["tau"], q["beta"], q["n_t"], CN, bd, lambda_v_bounds=q["lambda_v_bounds"], mode=mode,
                      amg_params=amg)
    rng = np.random.default_rng(1)
    x0 = rng.standard_normal((f.N, f.n))
    x1 = rng.standard_normal((f.N, f.n))
    y0, y1 = kkt.kkt_apply_fused(M, K, q["tau"], q["beta"], q["n_t"], CN, bd, x0, x1)
    g0, g1 = f.kkt_apply(x0, x1)
    assert max(rel(g0, y0), rel(g1, y1)) < 1e-13
    # one AMG solve
    H = f.hierarchies[0]
    b = rng.standard_normal(f.n)
    assert rel(f.amg_solve(0, b), oamg.solve(H, b)) < 1e-12
    # preconditioner (right-hand sides projected, as Preconditioner.apply hands them over)
    b0, b1 = x0.copy(), x1.copy()
    b0[:, bd] = 0.0
    b1[:, bd] = 0.0
    if mode == "diagonal":
        pc = opc.construct_pc_diagonal(M, K, q["tau"], q["beta"], q["n_t"], bd, lambda_v_bounds=q["lambda_v_bounds"],
                                       amg_params=amg)
    else:
        pc = opc.construct_pc(M, K, q["tau"], q["beta"], q["n_t"], CN, bd, lambda_v_bounds=q["lambda_v_bounds"],
                              amg_params=amg)
    r0, r1 = pc(b0.copy(), b1.copy())
    u0, u1 = f.pc_apply(b0, b1)
    assert rel(u0, r0) < 1e-11 and rel(u1, r1) < 1e-10
    assert fastpc.lib().oracle_omp_threads() == 4
