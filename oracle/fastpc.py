"""ctypes driver of oracle/c/pc_omp.c: the reference's algorithm (operator, in-built block preconditioner with the
AMG stand-in, Krylov vector work) in C / OpenMP for a time-independent, symmetric K.  Test infrastructure / CPU arm of
bench.py only.  The hierarchies are the numpy oracle's own (oracle/amg.py::setup); tests/test_oracle_fast.py checks
operator and preconditioner against oracle/kkt.py and oracle/pc.py."""
import ctypes as C
import os

import numpy as np
import scipy.sparse as sp

from . import amg as _amg
from synthetic.fem import assemble_bc

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "liboracle.so")
_P = C.c_void_p


class OCsr(C.Structure):
    _fields_ = [("n_rows", C.c_int32), ("n_cols", C.c_int32), ("ip", _P), ("ix", _P), ("v", _P)]


class OLevel(C.Structure):
    _fields_ = [("n", C.c_int32), ("A", OCsr), ("P", OCsr), ("R", OCsr), ("dinv", _P), ("rho", C.c_double),
                ("ainv", _P), ("x", _P), ("b", _P), ("r", _P), ("t0", _P), ("t1", _P)]


class OHier(C.Structure):
    _fields_ = [("nl", C.c_int32), ("nu", C.c_int32), ("nu_fine", C.c_int32), ("cycles", C.c_int32),
                ("lo", C.c_double), ("hi", C.c_double), ("lv", C.POINTER(OLevel))]


class OPc(C.Structure):
    _fields_ = [("n", C.c_int32), ("N", C.c_int32), ("CN", C.c_int32), ("mode", C.c_int32),
                ("tau", C.c_double), ("beta", C.c_double), ("eps", C.c_double),
                ("M", OCsr), ("K", OCsr), ("Mbc", OCsr), ("mdinv", _P), ("emin", C.c_double), ("emax", C.c_double),
                ("cheb_steps", C.c_int32), ("bc", _P), ("n_hier", C.c_int32), ("hier", C.POINTER(OHier)),
                ("fwd_h", _P), ("bwd_h", _P), ("off", OCsr)]


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(_PATH):
            raise ImportError(f"{_PATH} is missing: make -C oracle/c")
        os.environ.setdefault("OMP_WAIT_POLICY", "passive")
        L = C.CDLL(_PATH)
        L.oracle_omp_threads.restype = C.c_int
        L.oracle_dot.restype = C.c_double
        L.oracle_dot.argtypes = [C.c_int64, _P, _P]
        L.oracle_axpby.argtypes = [C.c_int64, C.c_double, _P, C.c_double, _P]
        L.oracle_pc_apply.argtypes = [C.POINTER(OPc), _P, _P, _P, _P, _P]
        L.oracle_kkt_apply.argtypes = [C.POINTER(OPc), _P, _P, _P, _P]
        L.oracle_amg_solve.argtypes = [C.POINTER(OHier), _P, _P]
        _LIB = L
    return _LIB


def set_threads(n=None):
    """All host cores unless told otherwise -- explicitly, because torchrun exports OMP_NUM_THREADS=1."""
    n = n or int(os.environ.get("ORACLE_THREADS", "0")) or (os.cpu_count() or 1)
    lib().oracle_omp_set_threads(int(n))
    return lib().oracle_omp_threads()


class FastPc:
    """Operator + preconditioner of one heat-control problem in C.  mode: "triangular" | "diagonal"."""

    def __init__(self, M, K, tau, beta, n_t, CN, bdofs, *, lambda_v_bounds=None, mode="triangular", epsilon=1e-3,
                 amg_params=None, cheb_steps=20):
        self._keep = []
        self.n = n = M.shape[0]
        self.N = N = n_t - 1 if CN else n_t
        self.CN, self.mode = bool(CN), mode
        assert mode == "triangular" or CN, "the block-diagonal variant exists for CN only"
        M = sp.csr_matrix(M).astype(np.float64)
        K = sp.csr_matrix(K).astype(np.float64)
        M.sort_indices()
        K.sort_indices()
        # one shared pattern for M and K (structural zeros kept), as the device library holds them
        pat = (abs(M) + abs(K)).tocsr()
        pat.sort_indices()
        ones = sp.csr_matrix((np.ones(pat.nnz), pat.indices, pat.indptr), shape=pat.shape)
        M = (M + 0 * ones).tocsr() if M.nnz != pat.nnz else M
        K = (K + 0 * ones).tocsr() if K.nnz != pat.nnz else K
        M.sort_indices()
        K.sort_indices()
        assert np.array_equal(M.indices, K.indices) and np.array_equal(M.indptr, K.indptr)
        h = 0.5 * tau
        bc = np.zeros(n, dtype=np.uint8)
        bc[np.asarray(bdofs, dtype=np.int64)] = 1
        Mbc = assemble_bc(M, bdofs).tocsr()
        Mbc.sort_indices()
        p = OPc()
        p.n, p.N, p.CN, p.mode = n, N, int(CN), 0 if mode == "triangular" else 1
        p.tau, p.beta, p.eps = tau, beta, epsilon
        p.M, p.K, p.Mbc = self._csr(M), self._csr(K), self._csr(Mbc)
        p.mdinv = self._arr(1.0 / Mbc.diagonal())
        if lambda_v_bounds is not None:
            p.emin, p.emax = lambda_v_bounds
        else:
            p.emin, p.emax = 0.0, 0.0
        p.cheb_steps = cheb_steps
        p.bc = self._arr(bc)
        # distinct diagonal blocks of L_hat -> hierarchies (numpy oracle's setup), exactly oracle/pc.py's shifts
        self.hierarchies = []
        hmap = {}

        def hier(shift, w):
            key = float(shift)
            if key not in hmap:
                A = assemble_bc((w * K + (1.0 + shift) * M).tocsr(), bdofs)
                H = _amg.setup(A, **(amg_params or {}))
                self.hierarchies.append(H)
                hmap[key] = len(self.hierarchies) - 1
            return hmap[key]
        if CN:
            c = h / beta ** 0.5
            fwd = [hier(c, h)] * N
            bwd = list(fwd)
            p.off = self._csr((h * K + (c - 1.0) * M).tocsr())
        else:
            s = tau / beta ** 0.5
            se = epsilon ** 0.5 * s
            fwd = [hier(0.0 if i == 0 else (s if i < N - 1 else se), tau) for i in range(N)]
            bwd = [hier(se if i == N - 1 else (s if i > 0 else 0.0), tau) for i in range(N)]
            p.off = self._csr(M)
        hs = (OHier * len(self.hierarchies))()
        for k, H in enumerate(self.hierarchies):
            self._fill_hier(hs[k], H)
        self._keep.append(hs)
        p.n_hier, p.hier = len(self.hierarchies), hs
        p.fwd_h = self._arr(np.asarray(fwd, dtype=np.int32))
        p.bwd_h = self._arr(np.asarray(bwd, dtype=np.int32))
        self._p = p
        self._work = np.zeros(4 * N * n)

    # -- plumbing: arrays stay referenced for the lifetime of the object
    def _arr(self, a):
        a = np.ascontiguousarray(a)
        self._keep.append(a)
        return a.ctypes.data

    def _csr(self, A):
        A = sp.csr_matrix(A)
        o = OCsr()
        o.n_rows, o.n_cols = A.shape
        o.ip = self._arr(A.indptr.astype(np.int32))
        o.ix = self._arr(A.indices.astype(np.int32))
        o.v = self._arr(A.data.astype(np.float64))
        return o

    def _fill_hier(self, h, H):
        p = H.params
        lv = (OLevel * len(H.levels))()
        for l, L in enumerate(H.levels):
            n = L.A.shape[0]
            lv[l].n = n
            lv[l].A = self._csr(L.A)
            if L.P is not None:
                lv[l].P = self._csr(L.P)
                lv[l].R = self._csr(L.R)
            lv[l].dinv = self._arr(L.dinv)
            lv[l].rho = L.rho
            lv[l].ainv = self._arr(L.Ainv) if L.Ainv is not None else None
            for name in ("x", "b", "r", "t0", "t1"):
                setattr(lv[l], name, self._arr(np.zeros(n)))
        self._keep.append(lv)
        h.nl, h.nu, h.nu_fine, h.cycles = len(H.levels), p["nu"], p["nu_fine"], p["cycles"]
        h.lo, h.hi, h.lv = p["lo"], p["hi"], lv
        assert p["acc_lo"] <= 0.0

    # -- the three pieces of one Krylov iteration
    def kkt_apply(self, x0, x1):
        y0, y1 = np.empty_like(x0), np.empty_like(x1)
        lib().oracle_kkt_apply(C.byref(self._p), x0.ctypes.data, x1.ctypes.data, y0.ctypes.data, y1.ctypes.data)
        return y0, y1

    def pc_apply(self, b0, b1):
        u0, u1 = np.zeros_like(b0), np.zeros_like(b1)
        lib().oracle_pc_apply(C.byref(self._p), b0.ctypes.data, b1.ctypes.data, u0.ctypes.data, u1.ctypes.data,
                              self._work.ctypes.data)
        return u0, u1

    def amg_solve(self, k, b):
        x = np.zeros_like(b)
        lib().oracle_amg_solve(C.byref(self._p.hier[k]), b.ctypes.data, x.ctypes.data)
        return x


def vector_work(x, y, n_dots, n_axpys):
    """The Krylov vector work of one iteration on the stacked (2 N n) vector: n_dots inner products, n_axpys updates."""
    L = lib()
    acc = 0.0
    for _ in range(n_dots):
        acc += L.oracle_dot(x.size, x.ctypes.data, y.ctypes.data)
    for _ in range(n_axpys):
        L.oracle_axpby(x.size, 1e-3, x.ctypes.data, 1.0, y.ctypes.data)
    return acc
