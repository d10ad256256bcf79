"""Multi-threaded sparse products for the oracle (test infrastructure only).

``mv(A, x)`` = ``A @ x`` for a scipy CSR matrix and a vector or a dense (n, m) block.  Uses the
OpenMP loops of oracle/c/spmv_omp.c when oracle/_build/liboracle.so has been built
(``__graft_entry__.build()``) and ``ORACLE_THREADS`` is not 1; falls back to scipy otherwise.
Both paths sum each row in CSR order, so results agree bit for bit."""
import ctypes as C
import os

import numpy as np
import scipy.sparse as sp

_LIB = None
_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "liboracle.so")


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = False
        if os.path.exists(_PATH) and os.environ.get("ORACLE_THREADS", "") != "1":
            try:
                os.environ.setdefault("OMP_WAIT_POLICY", "passive")   # spinning workers starve numpy on shared hosts
                lib = C.CDLL(_PATH)
                lib.oracle_csr_matvec.argtypes = [C.c_int32] + [C.c_void_p] * 5
                lib.oracle_csr_matmat.argtypes = [C.c_int32, C.c_int32] + [C.c_void_p] * 5
                _LIB = lib
            except OSError:
                _LIB = False
    return _LIB


def threads():
    return (os.cpu_count() or 1) if _lib() else 1


def mv(A, x):
    lib = _lib()
    if not lib or not sp.isspmatrix_csr(A) or A.indices.dtype != np.int32 or A.shape[0] < 20000:
        return A @ x
    x = np.ascontiguousarray(x, dtype=np.float64)
    if x.ndim == 1:
        y = np.empty(A.shape[0])
        lib.oracle_csr_matvec(A.shape[0], A.indptr.ctypes.data, A.indices.ctypes.data, A.data.ctypes.data,
                              x.ctypes.data, y.ctypes.data)
        return y
    y = np.empty((A.shape[0], x.shape[1]))
    lib.oracle_csr_matmat(A.shape[0], x.shape[1], A.indptr.ctypes.data, A.indices.ctypes.data,
                          A.data.ctypes.data, x.ctypes.data, y.ctypes.data)
    return y
