"""GPU parity: the fused KKT-apply kernel against the oracle's literal block-by-block
``MultiBlockSystemMatrix.mult`` (preconditioner/preconditioner.py:375-543), through the
C ABI.  fp64, tolerance 1e-13 relative to the largest output entry (only the summation
order differs)."""
import numpy as np
import pytest
import torch

import kat
from oracle import kkt
from synthetic import fem

pytestmark = pytest.mark.gpu


def _system(M, K, **kw):
    from control_b200 import MultiBlockSystem
    return MultiBlockSystem(M, K, **kw)


def _check_apply(M, K_levels, n_t, CN, bd, tau_interval=(0.0, 1.0), beta=1e-3, seed=0):
    s = _system(M, K_levels, n_t=n_t, beta=beta, CN=CN, time_interval=tau_interval, bc_dofs=bd)
    N, n = s.N, M.shape[0]
    rng = np.random.default_rng(seed)
    x0 = rng.standard_normal((N, n))
    x1 = rng.standard_normal((N, n))
    Kl = K_levels if isinstance(K_levels, list) else [K_levels] * n_t
    blocks = kkt.build_blocks(M, Kl, s.tau, beta, n_t, CN)
    ns = kkt.DirichletBCNullspace(bd)
    y0, y1 = kkt.kkt_apply_literal(blocks, ns, CN, x0, x1)
    yd = s.apply(s.to_device(x0, x1))
    g0, g1 = s.to_host_blocks(yd)
    scale = max(np.abs(y0).max(), np.abs(y1).max())
    assert np.abs(g0 - y0).max() <= 1e-13 * scale
    assert np.abs(g1 - y1).max() <= 1e-13 * scale
    # constrained rows return x exactly
    assert np.array_equal(g0[:, bd], x0[:, bd]) and np.array_equal(g1[:, bd], x1[:, bd])
    s.close()


@pytest.mark.parametrize("CN", [True, False])
@pytest.mark.parametrize("n_t", [2, 3, 6, 10, 17, 33, 64])
def test_apply_matches_literal_operator(CN, n_t):
    M, K, _, bd = fem.assemble_p1_2d(7, 5, 2.0, 1.0)
    _check_apply(M, K, n_t, CN, bd, tau_interval=(0.0, 2.0), seed=n_t)


@pytest.mark.parametrize("CN", [True, False])
def test_apply_max_blocks(CN):
    M, K, _, bd = fem.assemble_p1_2d(4, 4)
    _check_apply(M, K, 65 if CN else 64, CN, bd)


@pytest.mark.parametrize("CN", [True, False])
def test_apply_time_dependent_nonsymmetric_K(CN):
    M, K, _, bd = fem.assemble_p1_2d(6, 6)
    rng = np.random.default_rng(5)
    n_t = 9
    Ks = []
    for _ in range(n_t):
        Ki = K.copy()
        Ki.data = Ki.data * (1.0 + 0.3 * rng.standard_normal(Ki.nnz)) + 0.01 * rng.standard_normal(Ki.nnz)
        Ks.append(Ki)
    _check_apply(M, Ks, n_t, CN, bd)
    # one non-symmetric K for all levels
    _check_apply(M, Ks[0], n_t, CN, bd)


@pytest.mark.parametrize("CN", [True, False])
def test_apply_q2_kat_matrices_and_3d(CN):
    p = kat.instationary_kat(CN)
    _check_apply(p["M"], p["K"], p["n_t"], CN, p["bdofs"], beta=p["beta"])
    M, K, _, bd = fem.assemble_p1_3d(5, 4, 3)
    _check_apply(M, K, 8, CN, bd)


def test_apply_without_dirichlet_dofs_and_single_block():
    M, K, _, _ = fem.assemble_p1_2d(3, 3)
    _check_apply(M, K, 2, True, np.zeros(0, dtype=np.int32))
    _check_apply(M, K, 2, False, np.zeros(0, dtype=np.int32))


def test_layout_round_trip_and_padding():
    from control_b200 import _lib as L
    M, K, _, bd = fem.assemble_p1_2d(9, 7)
    s = _system(M, K, n_t=11, beta=1e-2, CN=True, bc_dofs=bd)
    x = torch.randn(s.vec_len(), dtype=torch.float64, device=s.device)
    tf = s.convert(x, L.CTL_LAYOUT_BLOCK_MAJOR, L.CTL_LAYOUT_TIME_FASTEST)
    assert tf.numel() == 2 * s.n * s.ld
    panel = tf.view(2, s.n, s.ld)
    assert torch.count_nonzero(panel[:, :, s.N:]) == 0
    assert torch.equal(panel[0, :, :s.N].t().contiguous().view(-1), x[:s.N * s.n])
    back = s.convert(tf, L.CTL_LAYOUT_TIME_FASTEST, L.CTL_LAYOUT_BLOCK_MAJOR)
    assert torch.equal(back, x)
    s.close()


def test_apply_linearity_at_scale():
    """Size-independent property on a larger mesh: A(a x + b y) = a A x + b A y, and the
    operator is symmetric for symmetric K: <A x, y> = <x, A y>."""
    M, K, _, bd = fem.assemble_p1_2d(96, 96, 2.0, 2.0)
    s = _system(M, K, n_t=64, beta=1e-4, CN=True, time_interval=(0.0, 2.0), bc_dofs=bd)
    g = torch.Generator(device=s.device).manual_seed(0)
    x = torch.randn(s.vec_len(), dtype=torch.float64, device=s.device, generator=g)
    y = torch.randn(s.vec_len(), dtype=torch.float64, device=s.device, generator=g)
    Ax, Ay = s.apply(x), s.apply(y)
    Axy = s.apply(0.5 * x - 2.0 * y)
    ref = 0.5 * Ax - 2.0 * Ay
    assert (Axy - ref).abs().max() <= 1e-12 * ref.abs().max()
    assert abs(torch.dot(Ax, y) - torch.dot(x, Ay)) <= 1e-11 * abs(torch.dot(Ax, y))
    s.close()


def test_errors_are_reported_not_swallowed():
    from control_b200 import CtlError
    M, K, _, bd = fem.assemble_p1_2d(3, 3)
    with pytest.raises(CtlError):
        _system(M, K, n_t=300, beta=1e-2, CN=True, bc_dofs=bd)      # more than 256 blocks
    K2 = K.copy()
    K2.eliminate_zeros()
    if K2.nnz != M.nnz:
        with pytest.raises(ValueError):
            _system(M, K2, n_t=4, beta=1e-2, CN=True, bc_dofs=bd)   # different pattern


@pytest.mark.parametrize("CN", [True, False])
@pytest.mark.parametrize("n_t", [66, 100, 129, 200, 256])
def test_apply_more_than_64_time_blocks(CN, n_t):
    """Wide kernels: every lane owns 4 (ld = 128) or 8 (ld = 256) consecutive time columns."""
    M, K, _, bd = fem.assemble_p1_2d(6, 5, 2.0, 1.0)
    _check_apply(M, K, n_t, CN, bd, tau_interval=(0.0, 2.0), seed=n_t)
    if n_t == 100:
        rng = np.random.default_rng(2)
        Ks = []
        for _ in range(n_t):
            Ki = K.copy()
            Ki.data = Ki.data * (1.0 + 0.2 * rng.standard_normal(Ki.nnz))
            Ks.append(Ki)
        _check_apply(M, Ks, n_t, CN, bd)
