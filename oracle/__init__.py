"""CPU oracle for the `control` all-at-once KKT hot path.  TEST INFRASTRUCTURE ONLY.

This package is a numpy/scipy restatement of the algorithm the reference executes
through Firedrake/PETSc/hypre for ``Control.Instationary.linear_solve``:

* (the assembled input matrices come from the neutral ``synthetic`` package)
* ``kkt``      block tables (control/control.py:2889-2978), the literal
               ``MultiBlockSystemMatrix.mult`` (preconditioner/preconditioner.py:375-543),
               ``T_1/T_2`` and inverses (control/control.py:26-96), nullspace
               projections (preconditioner/preconditioner.py:75-213), RHS
               (control/control.py:2980-3243), solution unpack (3299-3315)
* ``cheb``     PETSc KSPCHEBYSHEV+PCJACOBI as configured at control/control.py:1967-1991
* ``amg``      the aggregation AMG that replaces hypre BoomerAMG (control/control.py:2056-2067)
* ``pc``       ``construct_pc`` / ``pc_linear`` (control/control.py:1943-2440)
* ``krylov``   PETSc KSP GMRES / FGMRES / MINRES conventions used by
               preconditioner/preconditioner.py:732-772
* ``control``  ``Instationary.linear_solve`` driver and the discrete objective

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this package, and only as the checker / CPU baseline.  The product
(``control_b200``) never imports it and has no CPU fallback.

Parity status: the operator, T-transforms, nullspace projections and RHS are pinned by
re-creating the reference's analytic known-answer tests
(test/test_control.py:1243-1444 BE, 1447-1655 CN); the N = 1 block structure by
test/test_control.py:26-119, and the Stokes divergence coupling, block ordering and
ConstantNullspace (``oracle/stokes.py::stokes_apply_literal``) by test/test_control.py:232-358,
all four to the reference's own 1e-13.  Iteration counts, preconditioner
outputs and residual histories of the real Firedrake/PETSc/hypre stack are **parity
unpinned** (the reference ships no golden vectors and cannot run in this image;
BoomerAMG is replaced by the aggregation AMG in ``amg``).
"""
