"""Export a reference run to a fixture that pins parity OFF this box.

NOT runnable in this image (it needs Firedrake, petsc4py and the reference checkout on PYTHONPATH);
written against the reference's API as read from its sources and untested here.  A maintainer runs it
once inside a Firedrake environment:

    python scripts/export_from_firedrake.py --nx 10 --n_t 10 --out tests/golden/reference_c1_cn.npz

It solves the README heat-control problem (README.md:24-60; BASELINE config C1 by default) with the
reference's own ``Control.Instationary.linear_solve`` and stores exactly what crosses the boundary of
this library plus what the reference computed from it:

    M, K        CSR of assemble(inner(trial, test) dx) and of forward_form, Firedrake's dof numbering
                (``Mat.getValuesCSR()``; K on M's pattern is checked)
    bc_dofs     nodes of DirichletBC(space, 0, "on_boundary")                 (control/control.py:2851-2862)
    v_hat, f    nodal desired state / force per time level                   (README.md:33-55)
    v, zeta     the reference solution, n_t levels x n                       (control/control.py:3299-3315)
    its, reason, history   KSP iteration count, converged reason, monitored residual norms
    meta        beta, n_t, CN, time interval, solver_parameters, lambda_v_bounds, versions

``tests/test_reference_fixture.py`` picks up every ``tests/golden/reference_*.npz`` and checks the oracle
and the CUDA path against it (operator-level quantities to 1e-8 relative as north_star asks, iteration
counts reported; they depend on hypre vs the aggregation AMG stand-in).
"""
import argparse
import json

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=10)
    ap.add_argument("--n_t", type=int, default=10)
    ap.add_argument("--beta", type=float, default=1e-4)
    ap.add_argument("--be", action="store_true", help="backward Euler instead of the trapezoidal rule")
    ap.add_argument("--rtol", type=float, default=1e-10)
    ap.add_argument("--out", required=True)
    args = ap.parse_args()

    from firedrake import (DirichletBC, Function, FunctionSpace, SpatialCoordinate, TestFunction, TrialFunction,
                           UnitSquareMesh, assemble, cos, dx, grad, inner, pi)
    from firedrake.petsc import PETSc
    from control import Control                      # the reference checkout

    mesh = UnitSquareMesh(args.nx, args.nx, 2.0, 2.0)
    space = FunctionSpace(mesh, "Lagrange", 1)
    T_f = 2.0
    n_t = args.n_t
    tau = T_f / (n_t - 1.0)

    def forw_diff_operator(trial, test, v, t):
        return inner(grad(trial), grad(test)) * dx

    def shape(space_):
        X = SpatialCoordinate(space_.mesh())
        return cos(0.5 * pi * (X[0] - 1.0)) * cos(0.5 * pi * (X[1] - 1.0))

    def desired_state(test, t):
        v_d = Function(test.function_space(), name="v_d")
        v_d.interpolate(t * shape(test.function_space()))
        return inner(v_d, test) * dx, v_d

    def force_f(test, t):
        f = Function(test.function_space(), name="f")
        f.interpolate(shape(test.function_space()))
        return inner(f, test) * dx

    def bc_t(space_, t):
        return DirichletBC(space_, 0.0, "on_boundary")

    solver_parameters = {"linear_solver": "fgmres", "gmres_restart": 30, "maximum_iterations": 200,
                         "relative_tolerance": args.rtol, "absolute_tolerance": 0.0, "monitor_convergence": False}
    lambda_v_bounds = (0.5, 2.0)
    control = Control.Instationary(space, forw_diff_operator, desired_state=desired_state, force_f=force_f,
                                   beta=args.beta, n_t=n_t, CN=not args.be, time_interval=(0.0, T_f), bcs_v=bc_t)
    # residual history through a PETSc monitor would need access to the KSP; the reference returns nothing from
    # linear_solve, so only what it stores on the object is exported
    control.linear_solve(solver_parameters=solver_parameters, lambda_v_bounds=lambda_v_bounds, print_error=False,
                         create_output=False, plots=False)

    trial, test = TrialFunction(space), TestFunction(space)
    M = assemble(inner(trial, test) * dx).petscmat
    K = assemble(forw_diff_operator(trial, test, None, 0.0) + 0.0 * inner(trial, test) * dx).petscmat   # on M's pattern
    m_ip, m_ix, m_v = M.getValuesCSR()
    k_ip, k_ix, k_v = K.getValuesCSR()
    assert np.array_equal(m_ip, k_ip) and np.array_equal(m_ix, k_ix), "K is not on M's pattern"
    bc_dofs = np.asarray(bc_t(space, 0.0).nodes, dtype=np.int32)
    v_hat = np.zeros((n_t, space.dim()))
    f_nodal = np.zeros((n_t, space.dim()))
    for i in range(n_t):
        t = i * tau
        v_hat[i] = desired_state(test, t)[1].dat.data_ro
        fn = Function(space)
        fn.interpolate(shape(space))
        f_nodal[i] = fn.dat.data_ro
    v = np.stack([control._v.sub(i).dat.data_ro.copy() for i in range(n_t)])
    zeta = np.stack([control._zeta.sub(i).dat.data_ro.copy() for i in range(n_t)])
    meta = dict(nx=args.nx, n_t=n_t, beta=args.beta, CN=not args.be, time_interval=[0.0, T_f],
                solver_parameters=solver_parameters, lambda_v_bounds=list(lambda_v_bounds),
                petsc=".".join(str(x) for x in PETSc.Sys.getVersion()), comm_size=mesh.comm.size)
    np.savez_compressed(args.out, indptr=m_ip, indices=m_ix, M=m_v, K=k_v, bc_dofs=bc_dofs, v_hat=v_hat, f_nodal=f_nodal,
                        v=v, zeta=zeta, coords=mesh.coordinates.dat.data_ro.copy(), meta=json.dumps(meta))
    print("wrote", args.out)


if __name__ == "__main__":
    main()
