"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol
that include/ctl_b200.h declares (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ctl_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(ctl_[a-z0-9_]+)\s*\(", src)
    typedefs = set(re.findall(r"\(\*(ctl_[a-z0-9_]+)\)", src))
    return sorted(set(names) - typedefs)


def test_header_declares_the_path():
    names = declared_functions()
    for required in ("ctl_create", "ctl_destroy", "ctl_set_pattern", "ctl_set_values", "ctl_set_bc",
                     "ctl_kkt_apply", "ctl_pc_setup", "ctl_pc_apply", "ctl_solve", "ctl_solve_host",
                     "ctl_last_error", "ctl_comm_init", "ctl_stokes_create", "ctl_stokes_apply",
                     "ctl_stokes_pc_setup", "ctl_stokes_solve"):
        assert required in names


def test_library_exports_every_declared_symbol():
    from control_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build the library first (__graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, f"declared in ctl_b200.h but not exported: {missing}"


def test_python_binding_covers_the_header():
    from control_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_functions()
    lib = _lib.load()
    assert lib.ctl_version().startswith(b"ctl_b200")


def test_struct_layouts_match_the_header():
    """sizeof checks against a tiny C program compiled from the header itself."""
    import subprocess
    import tempfile
    from control_b200 import _lib
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "s.c")
        open(src, "w").write(
            '#include <stdio.h>\n#include "ctl_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu\\n",'
            'sizeof(ctl_config),sizeof(ctl_pc_options),sizeof(ctl_krylov_options),'
            'sizeof(ctl_solve_result),sizeof(ctl_stokes_pc_options));return 0;}\n')
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    assert sizes == [ctypes.sizeof(_lib.ctl_config), ctypes.sizeof(_lib.ctl_pc_options),
                     ctypes.sizeof(_lib.ctl_krylov_options), ctypes.sizeof(_lib.ctl_solve_result),
                     ctypes.sizeof(_lib.ctl_stokes_pc_options)]


def test_create_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import numpy as np
    import scipy.sparse as sp
    from control_b200 import CtlError, MultiBlockSystem
    M = sp.identity(4, format="csr")
    with pytest.raises(CtlError):
        MultiBlockSystem(M, M, n_t=3, beta=1e-2, CN=True, bc_dofs=np.array([0]))


@pytest.mark.parametrize("nx,coarse_max", [(24, 40), (48, 60)])
def test_host_amg_setup_matches_oracle_without_a_gpu(nx, coarse_max):
    """The C++ host setup of the aggregation AMG (csrc/amg_setup.cpp) against oracle/amg.py through the
    host-only probe: level sizes and entry counts identical, aggregates bit-identical, spectral bounds equal
    to rounding -- on the shifted heat operator of the Schur sweeps and on the singular pressure Laplacian."""
    import numpy as np
    from control_b200 import _lib
    from oracle import amg as oamg
    from synthetic import fem, problems
    lib = _lib.load()
    q = problems.heat_problem(nx, 8, True)
    c = 0.5 * q["tau"] / q["beta"] ** 0.5
    A_heat = fem.assemble_bc((0.5 * q["tau"] * q["K"] + (1 + c) * q["M"]).tocsr(), q["bdofs"])
    A_kp = fem.assemble_taylor_hood_2d(nx // 2, nx // 2, 2.0, 2.0)["K_p"].tocsr()
    for A in (A_heat, A_kp):
        A.sort_indices()
        H = oamg.setup(A, coarse_max=coarse_max)
        o = _lib.ctl_pc_options()
        assert lib.ctl_pc_default_options(ctypes.byref(o)) == 0
        o.amg_coarse_max = coarse_max
        n = A.shape[0]
        ip, ix, va = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)
        nl = ctypes.c_int32()
        ln = np.zeros(16, dtype=np.int32)
        lz = np.zeros(16, dtype=np.int64)
        lr = np.zeros(16)
        agg = np.full(n, -7, dtype=np.int32)
        rc = lib.ctl_amg_setup_probe(ip.ctypes.data, ix.ctypes.data, va.ctypes.data, n, ctypes.byref(o), ctypes.byref(nl),
                                     ln.ctypes.data, lz.ctypes.data, lr.ctypes.data, agg.ctypes.data)
        assert rc == 0
        assert nl.value == len(H.levels) >= 2
        assert [int(v) for v in ln[:nl.value]] == [L.A.shape[0] for L in H.levels]
        assert [int(v) for v in lz[:nl.value]] == [int((abs(L.A.data) > 1e-13 * abs(L.A.data).max()).sum()) for L in H.levels]
        assert np.array_equal(agg, H.levels[0].agg)
        assert np.allclose(lr[:nl.value], [L.rho for L in H.levels], rtol=1e-12, atol=0.0)


@pytest.mark.parametrize("CN", [True, False])
def test_host_rhs_with_inhomogeneous_dirichlet_data_matches_oracle(CN):
    """control_b200.control.build_rhs (host vector algebra of the caller, control/control.py:2980-3243,
    including the ``v_inhom`` lifting of time-dependent Dirichlet data and per-level K_i) against the
    oracle's restatement -- no GPU involved."""
    import numpy as np
    from control_b200.control import build_rhs
    from oracle import kkt
    from synthetic import fem
    M, K, coords, bd = fem.assemble_p1_2d(7, 6, 2.0, 1.0)
    n, n_t = M.shape[0], 6
    tau = 0.2
    rng = np.random.default_rng(2)
    Ks = [(K + 0.1 * i * M).tocsr() for i in range(n_t)]
    v_d = rng.standard_normal((n_t, n))
    f = rng.standard_normal((n_t, n))
    g = rng.standard_normal((n_t, bd.size))
    v_0 = rng.standard_normal(n)
    for bc_values in (None, g):
        a0, a1 = build_rhs(M, Ks[0], tau, n_t, CN, bd, v_d, f, v_0, bc_values=bc_values, K_levels=Ks)
        r0, r1 = kkt.build_rhs(M, Ks, tau, n_t, CN, bd, v_d, f, v_0, bc_values=bc_values)
        assert np.abs(a0 - r0).max() <= 1e-13 * np.abs(r0).max()
        assert np.abs(a1 - r1).max() <= 1e-13 * np.abs(r1).max()
