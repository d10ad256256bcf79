"""All-at-once KKT operator of ``Control.Instationary`` (oracle; test infrastructure only).

Vectors are numpy arrays of shape ``(N, n)``: block-major / time-slowest, the layout of
the reference's mixed PETSc vectors (preconditioner/preconditioner.py:276-287).
"""
import numpy as np
import scipy.sparse as sp


# --------------------------------------------------------------------------------------
# T_1 = I + S (add next block), T_2 = I + S^T (add previous block) and their inverses.
# control/control.py:26-96 (duplicated at preconditioner/preconditioner.py:33-60)
# --------------------------------------------------------------------------------------
def apply_T_1(x):
    y = x.copy()
    y[:-1] += x[1:]
    return y


def apply_T_2(x):
    y = x.copy()
    y[1:] += x[:-1]
    return y


def apply_T_1_inv(x):
    y = x.copy()
    for i in range(x.shape[0] - 2, -1, -1):
        y[i] -= y[i + 1]
    return y


def apply_T_2_inv(x):
    y = x.copy()
    for i in range(1, x.shape[0]):
        y[i] -= y[i - 1]
    return y


# --------------------------------------------------------------------------------------
# Block tables: control/control.py:2889-2978
# --------------------------------------------------------------------------------------
def n_blocks(n_t, CN):
    return n_t - 1 if CN else n_t


def build_blocks(M, K_levels, tau, beta, n_t, CN):
    """Dicts block_00/01/10/11 keyed (i, j) -> scipy CSR or None, exactly as
    ``Instationary.linear_solve`` builds them.  ``K_levels[i]`` is the matrix of
    ``forward_form`` at time level i (``D_v_i``); its adjoint is the transpose."""
    if sp.issparse(K_levels):
        K_levels = [K_levels] * n_t
    assert len(K_levels) == n_t
    N = n_blocks(n_t, CN)
    b00 = {(i, j): None for i in range(N) for j in range(N)}
    b01 = dict(b00)
    b10 = dict(b00)
    b11 = dict(b00)
    if CN:
        # control/control.py:2929-2958
        for i in range(n_t - 1):
            D_v_i, D_v_ip = K_levels[i], K_levels[i + 1]
            if i - 1 >= 0:
                b00[(i, i - 1)] = 0.5 * tau * M
                b10[(i, i - 1)] = 0.5 * tau * D_v_i - M
            b00[(i, i)] = 0.5 * tau * M
            b01[(i, i)] = 0.5 * tau * D_v_i.T + M
            b10[(i, i)] = 0.5 * tau * D_v_ip + M
            b11[(i, i)] = -0.5 * tau / beta * M
            if i + 1 < N:
                b01[(i, i + 1)] = 0.5 * tau * D_v_ip.T - M
                b11[(i, i + 1)] = -0.5 * tau / beta * M
    else:
        # control/control.py:2894-2928 and 2960-2978
        for i in range(n_t - 1):
            D_v_i = K_levels[i]
            if i - 1 >= 0:
                b10[(i, i - 1)] = -1.0 * M
            b00[(i, i)] = tau * M
            b01[(i, i)] = tau * D_v_i.T + M
            b10[(i, i)] = tau * D_v_i + M
            b01[(i, i + 1)] = -1.0 * M
            b11[(i + 1, i + 1)] = -(tau / beta) * M
        D_v_i = K_levels[n_t - 1]
        b01[(n_t - 1, n_t - 1)] = tau * D_v_i.T + M
        b10[(n_t - 1, n_t - 2)] = -1.0 * M
        b10[(n_t - 1, n_t - 1)] = tau * D_v_i + M
    return b00, b01, b10, b11


def count_blocks(blocks):
    return sum(1 for d in blocks for v in d.values() if v is not None)


# --------------------------------------------------------------------------------------
# DirichletBCNullspace: preconditioner/preconditioner.py:158-197 (+ base 92-116)
# --------------------------------------------------------------------------------------
class DirichletBCNullspace:
    def __init__(self, bdofs, alpha=1.0):
        self.bdofs = np.asarray(bdofs, dtype=np.int64)
        self.alpha = alpha

    def project(self, x):              # apply_nullspace_transformation_lhs_{right,left}
        x[..., self.bdofs] = 0.0

    def pre_mult_corrected_lhs(self, x):
        xc = x.copy()
        self.project(xc)
        return xc

    def post_mult_correct_lhs(self, x, y):
        self.project(y)
        y[..., self.bdofs] += self.alpha * x[..., self.bdofs]

    def pc_pre_mult_corrected(self, b):
        bc = b.copy()
        self.project(bc)
        return bc

    def pc_post_mult_correct(self, u, b):
        self.project(u)
        u[..., self.bdofs] += b[..., self.bdofs]


# --------------------------------------------------------------------------------------
# MultiBlockSystemMatrix.mult, literal: preconditioner/preconditioner.py:375-543
# --------------------------------------------------------------------------------------
def kkt_apply_literal(blocks, nullspace, CN, x0, x1):
    """y = A x, block by block.  x0, x1: (N, n).  Returns (y0, y1)."""
    b00, b01, b10, b11 = blocks
    N = x0.shape[0]
    xc0 = nullspace.pre_mult_corrected_lhs(x0)
    xc1 = nullspace.pre_mult_corrected_lhs(x1)
    y0 = np.zeros_like(x0)
    y1 = np.zeros_like(x1)
    for (i, j), B in b00.items():
        if B is not None:
            y0[i] += B @ xc0[j]
    for (i, j), B in b01.items():
        if B is not None:
            y0[i] += B @ xc1[j]
    for (i, j), B in b10.items():
        if B is not None:
            y1[i] += B @ xc0[j]
    for (i, j), B in b11.items():
        if B is not None:
            y1[i] += B @ xc1[j]
    if CN:
        y0 = apply_T_1(y0)
        y1 = apply_T_2(y1)
    nullspace.post_mult_correct_lhs(x0, y0)
    nullspace.post_mult_correct_lhs(x1, y1)
    return y0, y1


# --------------------------------------------------------------------------------------
# The same operator in the fused form the CUDA kernel implements.
# --------------------------------------------------------------------------------------
def kkt_apply_fused(M, K_levels, tau, beta, n_t, CN, bdofs, x0, x1, eps_unused=None):
    """Fused restatement: four batched products MV, KV, MZ, KZ, a time-neighbour
    combination per row, then the Dirichlet fix-up.  ``K_levels`` is one CSR matrix
    (time independent) or a list of n_t matrices."""
    N = x0.shape[0]
    if sp.issparse(K_levels):
        K_levels = [K_levels] * n_t
    xc0 = x0.copy()
    xc1 = x1.copy()
    xc0[:, bdofs] = 0.0
    xc1[:, bdofs] = 0.0
    MV = (M @ xc0.T).T
    MZ = (M @ xc1.T).T
    KV = np.empty_like(x0)
    KZ = np.empty_like(x1)
    if CN:
        # column j of x0 is v_{j+1}: multiplied by K_{j+1}; column j of x1 is zeta_j: K_j^T
        for j in range(N):
            KV[j] = K_levels[j + 1] @ xc0[j]
            KZ[j] = K_levels[j].T @ xc1[j]
        h = 0.5 * tau
        r0 = h * MV + h * KZ + MZ
        r0[1:] += h * MV[:-1]
        r0[:-1] += h * KZ[1:] - MZ[1:]
        r1 = h * KV + MV - (h / beta) * MZ
        # block_10[(i, i-1)] applies D_v_i to block i-1 (= v_i): the same K_{j+1} x0[j] as the diagonal term
        r1[1:] += h * KV[:-1] - MV[:-1]
        r1[:-1] += -(h / beta) * MZ[1:]
        y0 = apply_T_1(r0)
        y1 = apply_T_2(r1)
    else:
        for j in range(N):
            KV[j] = K_levels[j] @ xc0[j]
            KZ[j] = K_levels[j].T @ xc1[j]
        y0 = tau * KZ + MZ
        y0[:-1] += tau * MV[:-1] - MZ[1:]
        y1 = tau * KV + MV
        y1[1:] += -MV[:-1] - (tau / beta) * MZ[1:]
    y0[:, bdofs] = x0[:, bdofs]
    y1[:, bdofs] = x1[:, bdofs]
    return y0, y1


# --------------------------------------------------------------------------------------
# Right-hand side: control/control.py:2980-3243
# --------------------------------------------------------------------------------------
def build_rhs(M, K_levels, tau, n_t, CN, bdofs, v_d, f, v_0, check_v_d=True, check_f=True, bc_values=None):
    """v_d, f: (n_t, n) cofunction values (already tested against the basis, i.e.
    ``M @ nodal``), or, when check_* is False, the ready right-hand-side blocks (N, n)
    that are passed through untouched (control/control.py:3008, 3031, 3169, 3216).
    ``bc_values`` (n_t, len(bdofs)): inhomogeneous, time-dependent Dirichlet data of the state; the
    reference lifts them into the right-hand sides (``v_inhom``, control/control.py:2993-3124 BE,
    3137-3212 CN) and solves with homogenised conditions."""
    if sp.issparse(K_levels):
        K_levels = [K_levels] * n_t
    N = n_blocks(n_t, CN)
    n = M.shape[0]
    b_0 = np.zeros((N, n))
    b_1 = np.zeros((N, n))
    g = None
    if bc_values is not None:
        g = np.zeros((n_t, n))
        g[:, bdofs] = bc_values
    if not CN:
        if check_v_d:
            b_0[:n_t - 1] = tau * v_d[:n_t - 1]
            if g is not None:                           # 2993-3001, 3043-3053
                b_0[:n_t - 1] -= tau * (M @ g[:n_t - 1].T).T
            b_0[:, bdofs] = 0.0
        else:
            b_0[:] = v_d
        if check_f:
            b_1[0] = (tau * K_levels[0] + M) @ v_0
            b_1[1:] = tau * f[1:]
            if g is not None:                           # 3014-3023, 3062-3088, 3100-3124
                for i in range(n_t):
                    b_1[i] -= (tau * K_levels[i] + M) @ g[i]
                    if i > 0:
                        b_1[i] += M @ g[i - 1]
            b_1[:, bdofs] = 0.0
        else:
            b_1[:] = f
    else:
        if check_v_d:
            b_0[:] = 0.5 * tau * (v_d[:-1] + v_d[1:])
            if g is not None:                           # 3137-3163
                for i in range(N):
                    b_0[i] -= 0.5 * tau * (M @ g[i + 1])
                    if i > 0:
                        b_0[i] -= 0.5 * tau * (M @ g[i])
            b_0[:, bdofs] = 0.0
        else:
            b_0[:] = v_d
        if check_f:
            b_1[:] = 0.5 * tau * (f[:-1] + f[1:])
            if g is not None:                           # 3173-3212
                for i in range(N):
                    b_1[i] -= (0.5 * tau * K_levels[i + 1] + M) @ g[i + 1]
                    if i > 0:
                        b_1[i] -= (0.5 * tau * K_levels[i] - M) @ g[i]
            b_1[:, bdofs] = 0.0
        else:
            b_1[:] = f
        if check_v_d:
            b_0[0] -= 0.5 * tau * (M @ v_0)
            b_0[0, bdofs] = 0.0
        if check_f:
            b_1[0] -= (0.5 * tau * K_levels[0] - M) @ v_0
            b_1[0, bdofs] = 0.0
        b_0 = apply_T_1(b_0)
        b_1 = apply_T_2(b_1)
    return b_0, b_1


def unpack_solution(v, zeta, n_t, CN, v_0=None):
    """control/control.py:3299-3315: CN shifts v by one level."""
    if not CN:
        return v.copy(), zeta.copy()
    n = v.shape[1]
    v_new = np.zeros((n_t, n))
    zeta_new = np.zeros((n_t, n))
    if v_0 is not None:
        v_new[0] = v_0
    v_new[1:] = v
    zeta_new[:-1] = zeta
    return v_new, zeta_new
