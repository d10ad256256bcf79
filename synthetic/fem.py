"""Structured-mesh finite-element assemblers (input generator: stands in for Firedrake assembly).

Stand-ins for ``assemble(inner(trial, test) * dx)`` (mass ``M``, control/control.py:1562)
and ``assemble(forward_form)`` for the Laplacian ``inner(grad(trial), grad(test)) * dx``
(stiffness ``K``, e.g. test/test_control.py:1251-1253) on the meshes BASELINE.md names.
Dof numbering is lexicographic (x fastest).  ``M`` and ``K`` always share one sparsity
pattern (structural zeros of ``K`` are kept) and have sorted column indices.
"""
import numpy as np
import scipy.sparse as sp


def _csr_same_pattern(n, rows, cols, vals_list):
    """Assemble several COO value sets over the same (rows, cols) into CSR matrices that
    share indptr/indices exactly (explicit zeros kept)."""
    order = np.lexsort((cols, rows))
    rows = rows[order]
    cols = cols[order]
    key_change = np.empty(rows.size, dtype=bool)
    key_change[0] = True
    key_change[1:] = (rows[1:] != rows[:-1]) | (cols[1:] != cols[:-1])
    starts = np.flatnonzero(key_change)
    u_rows = rows[starts]
    u_cols = cols[starts].astype(np.int32)
    indptr = np.zeros(n + 1, dtype=np.int32)
    np.add.at(indptr, u_rows + 1, 1)
    indptr = np.cumsum(indptr).astype(np.int32)
    out = []
    for vals in vals_list:
        data = np.add.reduceat(vals[order], starts)
        out.append(sp.csr_matrix((data, u_cols.copy(), indptr.copy()), shape=(n, n)))
    return out


def p1_triangle_mesh(nx, ny, lx=1.0, ly=1.0):
    """Vertices and triangles of an nx-by-ny rectangle mesh, each cell split by the
    diagonal joining its lower-right and upper-left corners."""
    xs = np.linspace(0.0, lx, nx + 1)
    ys = np.linspace(0.0, ly, ny + 1)
    X, Y = np.meshgrid(xs, ys, indexing="xy")
    coords = np.stack([X.ravel(), Y.ravel()], axis=1)
    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    v00 = (j * (nx + 1) + i).ravel()
    v10 = v00 + 1
    v01 = v00 + (nx + 1)
    v11 = v01 + 1
    tris = np.concatenate([np.stack([v00, v10, v01], axis=1),
                           np.stack([v10, v11, v01], axis=1)], axis=0)
    return coords, tris


def assemble_p1_2d(nx, ny, lx=1.0, ly=1.0):
    """P1 mass and Laplacian stiffness on triangles.  Returns (M, K, coords, bdofs)."""
    coords, tris = p1_triangle_mesh(nx, ny, lx, ly)
    n = coords.shape[0]
    p = coords[tris]                      # (nt, 3, 2)
    e1 = p[:, 1] - p[:, 0]
    e2 = p[:, 2] - p[:, 0]
    det = e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]
    area = 0.5 * np.abs(det)
    # gradients of the barycentric basis functions
    g = np.empty((tris.shape[0], 3, 2))
    g[:, 1, 0] = e2[:, 1] / det
    g[:, 1, 1] = -e2[:, 0] / det
    g[:, 2, 0] = -e1[:, 1] / det
    g[:, 2, 1] = e1[:, 0] / det
    g[:, 0] = -g[:, 1] - g[:, 2]
    Ke = area[:, None, None] * np.einsum("tad,tbd->tab", g, g)
    Me = area[:, None, None] / 12.0 * (np.ones((3, 3)) + np.eye(3))[None]
    rows = np.repeat(tris, 3, axis=1).ravel()
    cols = np.tile(tris, (1, 3)).ravel()
    M, K = _csr_same_pattern(n, rows, cols, [Me.ravel(), Ke.ravel()])
    x, y = coords[:, 0], coords[:, 1]
    tol = 1e-12 * max(lx, ly)
    bd = np.flatnonzero((x < tol) | (x > lx - tol) | (y < tol) | (y > ly - tol))
    return M, K, coords, bd.astype(np.int32)


def _p2_1d(nel, length):
    """1-D P2 mass/stiffness on a uniform mesh, nodes ordered left to right."""
    h = length / nel
    Me = h / 30.0 * np.array([[4.0, 2.0, -1.0], [2.0, 16.0, 2.0], [-1.0, 2.0, 4.0]])
    Ke = 1.0 / (3.0 * h) * np.array([[7.0, -8.0, 1.0], [-8.0, 16.0, -8.0], [1.0, -8.0, 7.0]])
    n = 2 * nel + 1
    M = sp.lil_matrix((n, n))
    K = sp.lil_matrix((n, n))
    for e in range(nel):
        idx = np.arange(2 * e, 2 * e + 3)
        M[np.ix_(idx, idx)] += Me
        K[np.ix_(idx, idx)] += Ke
    return M.tocsr(), K.tocsr()


def assemble_q2_2d(nx, ny, lx=1.0, ly=1.0):
    """Q2 (tensor-product biquadratic) mass and Laplacian stiffness on a uniform
    quadrilateral mesh: the element of the reference's instationary known-answer tests
    (test/test_control.py:1245-1247).  Returns (M, K, coords, bdofs)."""
    Mx, Kx = _p2_1d(nx, lx)
    My, Ky = _p2_1d(ny, ly)
    # the 1-D factors share one pattern, so every Kronecker product below has the same
    # (row, col) list; sum the value sets over it (scipy's "+" would prune exact zeros)
    MM = sp.kron(My, Mx, format="coo")
    KM = sp.kron(My, Kx, format="coo")
    MK = sp.kron(Ky, Mx, format="coo")
    n = MM.shape[0]
    rows = np.concatenate([MM.row, KM.row, MK.row]).astype(np.int64)
    cols = np.concatenate([MM.col, KM.col, MK.col]).astype(np.int64)
    zM = np.zeros(KM.nnz + MK.nnz)
    M, K = _csr_same_pattern(
        n, rows, cols,
        [np.concatenate([MM.data, zM]),
         np.concatenate([np.zeros(MM.nnz), KM.data, MK.data])])
    xs = np.linspace(0.0, lx, 2 * nx + 1)
    ys = np.linspace(0.0, ly, 2 * ny + 1)
    X, Y = np.meshgrid(xs, ys, indexing="xy")
    coords = np.stack([X.ravel(), Y.ravel()], axis=1)
    x, y = coords[:, 0], coords[:, 1]
    tol = 1e-12 * max(lx, ly)
    bd = np.flatnonzero((x < tol) | (x > lx - tol) | (y < tol) | (y > ly - tol))
    return M, K, coords, bd.astype(np.int32)


def assemble_p1_3d(nx, ny, nz, lx=1.0, ly=1.0, lz=1.0):
    """P1 mass and Laplacian stiffness on tetrahedra (each cube split into six Kuhn
    tetrahedra sharing the main diagonal).  Returns (M, K, coords, bdofs)."""
    xs = np.linspace(0.0, lx, nx + 1)
    ys = np.linspace(0.0, ly, ny + 1)
    zs = np.linspace(0.0, lz, nz + 1)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    coords = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    n = coords.shape[0]
    sx, sy, sz = 1, nx + 1, (nx + 1) * (ny + 1)
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    base = (k * sz + j * sy + i).ravel()
    import itertools
    tets = []
    for perm in itertools.permutations((sx, sy, sz)):
        a = base
        b = a + perm[0]
        c = b + perm[1]
        d = c + perm[2]
        tets.append(np.stack([a, b, c, d], axis=1))
    tets = np.concatenate(tets, axis=0)
    p = coords[tets]                     # (nt, 4, 3)
    J = np.stack([p[:, 1] - p[:, 0], p[:, 2] - p[:, 0], p[:, 3] - p[:, 0]], axis=1)  # rows = edges
    det = np.linalg.det(J)
    vol = np.abs(det) / 6.0
    Jinv = np.linalg.inv(J)              # columns of Jinv = gradients of lambda_1..3
    g = np.empty((tets.shape[0], 4, 3))
    g[:, 1:] = np.transpose(Jinv, (0, 2, 1))
    g[:, 0] = -g[:, 1] - g[:, 2] - g[:, 3]
    Ke = vol[:, None, None] * np.einsum("tad,tbd->tab", g, g)
    Me = vol[:, None, None] / 20.0 * (np.ones((4, 4)) + np.eye(4))[None]
    rows = np.repeat(tets, 4, axis=1).ravel()
    cols = np.tile(tets, (1, 4)).ravel()
    M, K = _csr_same_pattern(n, rows, cols, [Me.ravel(), Ke.ravel()])
    x, y, z = coords[:, 0], coords[:, 1], coords[:, 2]
    tol = 1e-12 * max(lx, ly, lz)
    bd = np.flatnonzero((x < tol) | (x > lx - tol) | (y < tol) | (y > ly - tol)
                        | (z < tol) | (z > lz - tol))
    return M, K, coords, bd.astype(np.int32)


def assemble_bc(A, bdofs):
    """``assemble(a, bcs=...)``: zero constrained rows and columns, unit diagonal
    (Firedrake convention relied on at control/control.py:1971-1972, 2057-2059).
    The sparsity pattern of ``A`` is kept (zeros stay as explicit entries)."""
    A = A.tocsr().copy()
    n = A.shape[0]
    mask = np.zeros(n, dtype=bool)
    mask[bdofs] = True
    row_of = np.repeat(np.arange(n), np.diff(A.indptr))
    kill = mask[row_of] | mask[A.indices]
    A.data[kill] = 0.0
    diag = (row_of == A.indices) & mask[row_of]
    A.data[diag] = 1.0
    return A


def nonlinear_diffusion_p1_2d(nx, ny, lx=1.0, ly=1.0):
    """Assembler for the non-linear diffusion operator of BASELINE config C5,
    forward_form(trial, test, v) = (1 + v^2) grad(trial) . grad(test) dx, on the P1 mesh of
    ``assemble_p1_2d`` (one-point quadrature at the centroid).  Returns ``D_v(v, gauss_newton)``:
    the matrix of the form at state v (Picard, symmetric) or of its derivative
    ufl.derivative(form(v, test, v), v, trial) (Gauss-Newton, non-symmetric), the two choices of
    ``construct_D_v`` (control/control.py:1887-1896).  Both on the pattern of M."""
    coords, tris = p1_triangle_mesh(nx, ny, lx, ly)
    n = coords.shape[0]
    p = coords[tris]
    e1 = p[:, 1] - p[:, 0]
    e2 = p[:, 2] - p[:, 0]
    det = e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]
    area = 0.5 * np.abs(det)
    g = np.empty((tris.shape[0], 3, 2))
    g[:, 1, 0] = e2[:, 1] / det
    g[:, 1, 1] = -e2[:, 0] / det
    g[:, 2, 0] = -e1[:, 1] / det
    g[:, 2, 1] = e1[:, 0] / det
    g[:, 0] = -g[:, 1] - g[:, 2]
    GG = np.einsum("tad,tbd->tab", g, g)
    rows = np.repeat(tris, 3, axis=1).ravel()
    cols = np.tile(tris, (1, 3)).ravel()

    def D_v(v, gauss_newton=False):
        vt = v[tris]                                   # (nt, 3)
        vc = vt.mean(axis=1)
        Ke = (area * (1.0 + vc ** 2))[:, None, None] * GG
        if gauss_newton:
            gradv = np.einsum("ta,tad->td", vt, g)     # grad v_h per triangle
            gv = np.einsum("td,tid->ti", gradv, g)     # grad v . grad phi_i
            Ke = Ke + (area * 2.0 * vc / 3.0)[:, None, None] * gv[:, :, None] * np.ones((1, 1, 3))
        return _csr_same_pattern(n, rows, cols, [Ke.ravel()])[0]

    return D_v


def assemble_taylor_hood_2d(nx, ny, lx=1.0, ly=1.0):
    """Taylor-Hood P2 (vector) - P1 on the triangle mesh of ``p1_triangle_mesh``: the spaces of
    the reference's Stokes control tests (``space_v`` = VectorFunctionSpace P2, ``space_p`` = P1,
    test/test_control.py:3045-3172; BASELINE config C4).  P2 nodes are the points of the
    (2nx+1) x (2ny+1) lattice, velocity dofs are interleaved by component (dof = 2 node + comp,
    Firedrake's block size 2).  Returns a dict with M_v, K_v (vector mass / vector Laplacian),
    B (n_p x n_v, ``-inner(div(trial_v), test_p) * dx``, control/control.py:3708), M_p, K_p, the
    velocity boundary dofs and the node coordinates."""
    mx, my = 2 * nx + 1, 2 * ny + 1
    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    i, j = i.ravel(), j.ravel()

    def node(a, b):                      # lattice point (2i + a, 2j + b)
        return (2 * j + b) * mx + (2 * i + a)

    def vert(a, b):                      # P1 vertex (i + a, j + b)
        return (j + b) * (nx + 1) + (i + a)
    # lower-left triangles (v00, v10, v01) and upper-right ones (v10, v11, v01); local order:
    # vertices 0,1,2 then midpoints of edges (1,2), (0,2), (0,1)
    t2 = np.concatenate([
        np.stack([node(0, 0), node(2, 0), node(0, 2), node(1, 1), node(0, 1), node(1, 0)], axis=1),
        np.stack([node(2, 0), node(2, 2), node(0, 2), node(1, 2), node(1, 1), node(2, 1)], axis=1)], axis=0)
    t1 = np.concatenate([np.stack([vert(0, 0), vert(1, 0), vert(0, 1)], axis=1),
                         np.stack([vert(1, 0), vert(1, 1), vert(0, 1)], axis=1)], axis=0)
    xs = np.linspace(0.0, lx, mx)
    ys = np.linspace(0.0, ly, my)
    X, Y = np.meshgrid(xs, ys, indexing="xy")
    coords2 = np.stack([X.ravel(), Y.ravel()], axis=1)
    p = coords2[t2[:, :3]]
    e1 = p[:, 1] - p[:, 0]
    e2 = p[:, 2] - p[:, 0]
    det = e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]
    area = 0.5 * np.abs(det)
    g = np.empty((t2.shape[0], 3, 2))    # gradients of the barycentric coordinates
    g[:, 1, 0] = e2[:, 1] / det
    g[:, 1, 1] = -e2[:, 0] / det
    g[:, 2, 0] = -e1[:, 1] / det
    g[:, 2, 1] = e1[:, 0] / det
    g[:, 0] = -g[:, 1] - g[:, 2]
    # degree-4 quadrature on the reference triangle (6 points), barycentric coordinates
    a1, a2 = 0.445948490915965, 0.091576213509771
    w1, w2 = 0.223381589678011, 0.109951743655322
    lam = np.array([[1 - 2 * a1, a1, a1], [a1, 1 - 2 * a1, a1], [a1, a1, 1 - 2 * a1],
                    [1 - 2 * a2, a2, a2], [a2, 1 - 2 * a2, a2], [a2, a2, 1 - 2 * a2]])
    wq = np.array([w1, w1, w1, w2, w2, w2])
    nq = lam.shape[0]
    phi = np.empty((nq, 6))              # P2 basis at the quadrature points
    dphi = np.empty((nq, 6, 3))          # d phi / d lambda_k
    for q in range(nq):
        l0, l1, l2 = lam[q]
        phi[q] = [l0 * (2 * l0 - 1), l1 * (2 * l1 - 1), l2 * (2 * l2 - 1), 4 * l1 * l2, 4 * l0 * l2, 4 * l0 * l1]
        dphi[q] = [[4 * l0 - 1, 0, 0], [0, 4 * l1 - 1, 0], [0, 0, 4 * l2 - 1],
                   [0, 4 * l2, 4 * l1], [4 * l2, 0, 4 * l0], [4 * l1, 4 * l0, 0]]
    grad = np.einsum("qak,tkd->tqad", dphi, g)                   # (nt, nq, 6, 2) physical gradients
    Ms = np.einsum("q,qa,qb->ab", wq, phi, phi)[None] * area[:, None, None]
    Ks = np.einsum("q,tqad,tqbd->tab", wq, grad, grad) * area[:, None, None]
    # B[p, (a, d)] = - int psi_p d phi_a / d x_d
    Bs = -np.einsum("q,qp,tqad->tpad", wq, lam, grad) * area[:, None, None, None]
    n2 = mx * my
    n1 = (nx + 1) * (ny + 1)
    rows = np.repeat(t2, 6, axis=1).ravel()
    cols = np.tile(t2, (1, 6)).ravel()
    Msc, Ksc = _csr_same_pattern(n2, rows, cols, [np.broadcast_to(Ms, Ks.shape).ravel(), Ks.ravel()])
    eye2 = sp.identity(2, format="csr")
    M_v = sp.kron(Msc, eye2, format="csr")
    K_v = sp.kron(Ksc, eye2, format="csr")
    # keep the shared pattern exactly (kron of explicit zeros keeps them)
    for A in (M_v, K_v):
        A.sort_indices()
        A.indices = A.indices.astype(np.int32)
        A.indptr = A.indptr.astype(np.int32)
    M_v = sp.csr_matrix((M_v.data, M_v.indices, M_v.indptr), shape=M_v.shape)
    K_v = sp.csr_matrix((K_v.data, K_v.indices, K_v.indptr), shape=K_v.shape)
    brow = np.repeat(t1[:, :, None, None], 6, axis=2).repeat(2, axis=3)
    bcol = 2 * t2[:, None, :, None] + np.arange(2)[None, None, None, :]
    bcol = np.broadcast_to(bcol, brow.shape)
    B = sp.coo_matrix((Bs.ravel(), (brow.ravel(), bcol.ravel())), shape=(n1, 2 * n2)).tocsr()
    B.sort_indices()
    B = sp.csr_matrix((B.data, B.indices.astype(np.int32), B.indptr.astype(np.int32)), shape=B.shape)
    M_p, K_p, coords1, _ = assemble_p1_2d(nx, ny, lx, ly)
    x, y = coords2[:, 0], coords2[:, 1]
    tol = 1e-12 * max(lx, ly)
    bn = np.flatnonzero((x < tol) | (x > lx - tol) | (y < tol) | (y > ly - tol))
    bd_v = np.sort(np.concatenate([2 * bn, 2 * bn + 1])).astype(np.int32)
    return dict(M_v=M_v, K_v=K_v, B=B, M_p=M_p, K_p=K_p, bdofs_v=bd_v, coords_v=coords2, coords_p=coords1,
                M_scalar=Msc, K_scalar=Ksc)


def _p1_1d(nel, length):
    """1-D P1 mass/stiffness on a uniform mesh."""
    h = length / nel
    Me = h / 6.0 * np.array([[2.0, 1.0], [1.0, 2.0]])
    Ke = 1.0 / h * np.array([[1.0, -1.0], [-1.0, 1.0]])
    n = nel + 1
    M = sp.lil_matrix((n, n))
    K = sp.lil_matrix((n, n))
    for e in range(nel):
        idx = np.arange(e, e + 2)
        M[np.ix_(idx, idx)] += Me
        K[np.ix_(idx, idx)] += Ke
    return M.tocsr(), K.tocsr()


def _p1_p2_1d(nel, length):
    """1-D mixed matrices between P1 (rows) and P2 (columns) on the same uniform mesh:
    G[c, a] = int psi_c phi_a dx, D[c, a] = int psi_c phi_a' dx (exact: 3-point Gauss)."""
    h = length / nel
    gp = 0.5 + 0.5 * np.array([-np.sqrt(0.6), 0.0, np.sqrt(0.6)])      # points on (0, 1)
    gw = 0.5 * np.array([5.0 / 9.0, 8.0 / 9.0, 5.0 / 9.0])
    psi = np.stack([1.0 - gp, gp])                                      # P1 on the reference interval
    phi = np.stack([(1.0 - gp) * (1.0 - 2.0 * gp), 4.0 * gp * (1.0 - gp), gp * (2.0 * gp - 1.0)])
    dphi = np.stack([4.0 * gp - 3.0, 4.0 - 8.0 * gp, 4.0 * gp - 1.0])  # d/ds
    Ge = h * np.einsum("q,cq,aq->ca", gw, psi, phi)
    De = np.einsum("q,cq,aq->ca", gw, psi, dphi)                        # h * (1/h)
    G = sp.lil_matrix((nel + 1, 2 * nel + 1))
    D = sp.lil_matrix((nel + 1, 2 * nel + 1))
    for e in range(nel):
        G[np.ix_(np.arange(e, e + 2), np.arange(2 * e, 2 * e + 3))] += Ge
        D[np.ix_(np.arange(e, e + 2), np.arange(2 * e, 2 * e + 3))] += De
    return G.tocsr(), D.tocsr()


def assemble_q2q1_stokes_2d(nx, ny, lx=1.0, ly=1.0):
    """Vector Q2 - Q1 on a uniform quadrilateral mesh: the spaces of the reference's Stokes
    known-answer test (test/test_control.py:232-240).  Velocity dofs interleaved by component
    (dof = 2 node + comp).  Returns a dict with M_v, L_v (vector mass / vector Laplacian), B
    (n_p x n_v, ``-inner(div(trial_v), test_p) * dx``), M_p, L_p, the velocity boundary dofs and the
    node coordinates of both spaces."""
    M_s, L_s, coords_v, bd_nodes = assemble_q2_2d(nx, ny, lx, ly)
    I2 = sp.identity(2, format="csr")
    M_v = sp.kron(M_s, I2, format="csr")
    L_v = sp.kron(L_s, I2, format="csr")
    for A in (M_v, L_v):
        A.sort_indices()
    Mx1, Kx1 = _p1_1d(nx, lx)
    My1, Ky1 = _p1_1d(ny, ly)
    MM = sp.kron(My1, Mx1, format="coo")           # one shared pattern, explicit zeros kept (as assemble_q2_2d)
    KM = sp.kron(My1, Kx1, format="coo")
    MK = sp.kron(Ky1, Mx1, format="coo")
    rows = np.concatenate([MM.row, KM.row, MK.row]).astype(np.int64)
    cols = np.concatenate([MM.col, KM.col, MK.col]).astype(np.int64)
    M_p, L_p = _csr_same_pattern(MM.shape[0], rows, cols,
                                 [np.concatenate([MM.data, np.zeros(KM.nnz + MK.nnz)]),
                                  np.concatenate([np.zeros(MM.nnz), KM.data, MK.data])])
    Gx, Dx = _p1_p2_1d(nx, lx)
    Gy, Dy = _p1_p2_1d(ny, ly)
    Bx = -sp.kron(Gy, Dx, format="csr")            # -int d_x(phi) psi
    By = -sp.kron(Dy, Gx, format="csr")            # -int d_y(phi) psi
    n_p, n_s = Bx.shape
    B = sp.lil_matrix((n_p, 2 * n_s))
    B[:, 0::2] = Bx
    B[:, 1::2] = By
    B = B.tocsr()
    B.sort_indices()
    xs = np.linspace(0.0, lx, nx + 1)
    ys = np.linspace(0.0, ly, ny + 1)
    X, Y = np.meshgrid(xs, ys, indexing="xy")
    coords_p = np.stack([X.ravel(), Y.ravel()], axis=1)
    bdofs_v = np.sort(np.concatenate([2 * bd_nodes, 2 * bd_nodes + 1])).astype(np.int32)
    return dict(M_v=M_v, L_v=L_v, B=B, M_p=M_p, L_p=L_p, bdofs_v=bdofs_v, coords_v=coords_v, coords_p=coords_p)


def assemble_convection_p1_2d(nx, ny, lx, ly, wind):
    """P1 convection matrix ``inner(dot(grad(trial), wind), test) * dx`` on the mesh of
    ``assemble_p1_2d`` (same sparsity pattern, explicit zeros kept), e.g. the forward operator of the
    reference's convection-diffusion studies (test/test_control.py:2693-2705).  ``wind(x, y)`` returns the
    two components at the given points; edge-midpoint quadrature (exact for quadratic integrands)."""
    coords, tris = p1_triangle_mesh(nx, ny, lx, ly)
    n = coords.shape[0]
    p = coords[tris]
    e1 = p[:, 1] - p[:, 0]
    e2 = p[:, 2] - p[:, 0]
    det = e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]
    area = 0.5 * np.abs(det)
    g = np.empty((tris.shape[0], 3, 2))
    g[:, 1, 0] = e2[:, 1] / det
    g[:, 1, 1] = -e2[:, 0] / det
    g[:, 2, 0] = -e1[:, 1] / det
    g[:, 2, 1] = e1[:, 0] / det
    g[:, 0] = -g[:, 1] - g[:, 2]
    bary = np.array([[0.5, 0.5, 0.0], [0.0, 0.5, 0.5], [0.5, 0.0, 0.5]])       # quadrature points = edge midpoints
    Ce = np.zeros((tris.shape[0], 3, 3))
    for q in range(3):
        xq = np.einsum("a,tad->td", bary[q], p)
        wx, wy = wind(xq[:, 0], xq[:, 1])
        wg = wx[:, None] * g[:, :, 0] + wy[:, None] * g[:, :, 1]                # wind . grad(phi_j), (nt, 3)
        Ce += (area / 3.0)[:, None, None] * bary[q][None, :, None] * wg[:, None, :]   # phi_i(x_q) * (w . grad phi_j)
    rows = np.repeat(tris, 3, axis=1).ravel()
    cols = np.tile(tris, (1, 3)).ravel()
    (C,) = _csr_same_pattern(n, rows, cols, [Ce.ravel()])
    return C


def convection_q2_2d(nx, ny, lx=1.0, ly=1.0):
    """Assembler of the Q2 convection matrix ``inner(dot(grad(trial), w), test) * dx`` for a NODAL (Q2) wind
    ``w`` on the mesh and numbering of ``assemble_q2_2d`` -- the Picard linearisation of the Navier-Stokes
    forward operator in the reference's tests (test/test_control.py:4297-4302).  Returns ``C(wx, wy)`` -> scalar
    CSR on the pattern of the Q2 mass matrix (explicit zeros kept); the vector-valued operator is
    ``kron(C, I_2)`` on the pattern of ``assemble_q2q1_stokes_2d``'s ``M_v``.  4 x 4 Gauss points per cell
    (exact: the integrand has degree 6 per direction)."""
    hx, hy = lx / nx, ly / ny
    gp, gw = np.polynomial.legendre.leggauss(4)
    s = 0.5 * (gp + 1.0)                                   # points on (0, 1)
    w1 = 0.5 * gw
    # 1-D quadratic shape functions at nodes 0, 1/2, 1 and their derivatives, evaluated at s
    N1 = np.stack([(1.0 - s) * (1.0 - 2.0 * s), 4.0 * s * (1.0 - s), s * (2.0 * s - 1.0)])        # (3, q)
    D1 = np.stack([4.0 * s - 3.0, 4.0 - 8.0 * s, 4.0 * s - 1.0])
    # 2-D shape functions, local node a = 3 * ay + ax, at quadrature point (qy, qx)
    phi = np.einsum("bq,ap->bapq", N1, N1).reshape(9, 4, 4)            # N_ay(s_qy) N_ax(s_qx)
    dphix = np.einsum("bq,ap->bapq", N1, D1).reshape(9, 4, 4) / hx
    dphiy = np.einsum("bq,ap->bapq", D1, N1).reshape(9, 4, 4) / hy
    wq = np.einsum("q,p->qp", w1, w1) * hx * hy
    mx = 2 * nx + 1
    ex, ey = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    base = (2 * ey.ravel()) * mx + 2 * ex.ravel()                      # lower-left node of every cell
    loc = (np.arange(3)[:, None] * mx + np.arange(3)[None, :]).ravel()  # a = 3 ay + ax
    cells = base[:, None] + loc[None, :]                               # (ncell, 9) global nodes
    n = mx * (2 * ny + 1)
    rows = np.repeat(cells, 9, axis=1).ravel()
    cols = np.tile(cells, (1, 9)).ravel()
    T_x = np.einsum("iqp,jqp,kqp,qp->kij", phi, dphix, phi, wq)        # sum_q phi_i dphix_j phi_k w
    T_y = np.einsum("iqp,jqp,kqp,qp->kij", phi, dphiy, phi, wq)

    def C(wx, wy):
        Ce = np.einsum("ck,kij->cij", wx[cells], T_x) + np.einsum("ck,kij->cij", wy[cells], T_y)
        return _csr_same_pattern(n, rows, cols, [Ce.ravel()])[0]
    return C


def convection_q1_q2wind_2d(nx, ny, lx=1.0, ly=1.0):
    """The same convection form on the Q1 (pressure) space with the Q2 nodal wind of the velocity space:
    ``construct_D_v(p_trial, p_test, v_n_help, t)`` of the reference's pressure-space blocks
    (control/control.py:3787-3789) for a Navier-Stokes forward operator.  Returns ``C(wx, wy)`` -> CSR on the
    pattern of ``assemble_q2q1_stokes_2d``'s ``M_p``; ``wx, wy`` are the wind components at the Q2 nodes.
    3 x 3 Gauss points per cell (exact: degree 4 per direction)."""
    hx, hy = lx / nx, ly / ny
    gp, gw = np.polynomial.legendre.leggauss(3)
    s = 0.5 * (gp + 1.0)
    w1 = 0.5 * gw
    N2 = np.stack([(1.0 - s) * (1.0 - 2.0 * s), 4.0 * s * (1.0 - s), s * (2.0 * s - 1.0)])        # wind, quadratic
    N1 = np.stack([1.0 - s, s])                                                                   # pressure, linear
    D1 = np.stack([-np.ones_like(s), np.ones_like(s)])
    wind = np.einsum("bq,ap->bapq", N2, N2).reshape(9, 3, 3)
    psi = np.einsum("bq,ap->bapq", N1, N1).reshape(4, 3, 3)
    dpsix = np.einsum("bq,ap->bapq", N1, D1).reshape(4, 3, 3) / hx
    dpsiy = np.einsum("bq,ap->bapq", D1, N1).reshape(4, 3, 3) / hy
    wq = np.einsum("q,p->qp", w1, w1) * hx * hy
    mx2, mx1 = 2 * nx + 1, nx + 1
    ex, ey = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    ex, ey = ex.ravel(), ey.ravel()
    cells2 = ((2 * ey) * mx2 + 2 * ex)[:, None] + (np.arange(3)[:, None] * mx2 + np.arange(3)[None, :]).ravel()[None, :]
    cells1 = (ey * mx1 + ex)[:, None] + (np.arange(2)[:, None] * mx1 + np.arange(2)[None, :]).ravel()[None, :]
    n = mx1 * (ny + 1)
    rows = np.repeat(cells1, 4, axis=1).ravel()
    cols = np.tile(cells1, (1, 4)).ravel()
    T_x = np.einsum("iqp,jqp,kqp,qp->kij", psi, dpsix, wind, wq)
    T_y = np.einsum("iqp,jqp,kqp,qp->kij", psi, dpsiy, wind, wq)

    def C(wx, wy):
        Ce = np.einsum("ck,kij->cij", wx[cells2], T_x) + np.einsum("ck,kij->cij", wy[cells2], T_y)
        return _csr_same_pattern(n, rows, cols, [Ce.ravel()])[0]
    return C
