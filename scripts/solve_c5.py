"""BASELINE config C5: nonlinear-diffusion heat control (operator depends on v), Gauss_Newton=True,
P1 on 512x512, n_t = 32: a few outer iterations of Control.Instationary.non_linear_solve with the
time split host assembly / preconditioner setup / device solve."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=512)
    ap.add_argument("--n_t", type=int, default=32)
    ap.add_argument("--outer", type=int, default=3)
    ap.add_argument("--picard", action="store_true")
    args = ap.parse_args()
    from synthetic import fem, problems
    from control_b200 import Control
    import control_b200.system as sysm
    q = problems.heat_problem(args.nx, args.n_t, True, beta=1e-2)
    Dv = fem.nonlinear_diffusion_p1_2d(args.nx, args.nx, 2.0, 2.0)
    times = q["tau"] * np.arange(q["n_t"])
    idx = {round(float(t), 12): i for i, t in enumerate(times)}
    stats = {"assemble": 0.0, "setup": 0.0, "solve": 0.0}

    def forward(v, t, gn):
        t0 = time.perf_counter()
        A = Dv(v, gn)
        stats["assemble"] += time.perf_counter() - t0
        return A
    orig_setup = sysm.MultiBlockSystem.setup_preconditioner
    orig_solve = sysm.MultiBlockSystem.solve

    def timed_setup(self, **kw):
        t0 = time.perf_counter()
        r = orig_setup(self, **kw)
        stats["setup"] += time.perf_counter() - t0
        return r

    def timed_solve(self, *a, **kw):
        t0 = time.perf_counter()
        r = orig_solve(self, *a, **kw)
        stats["solve"] += time.perf_counter() - t0
        print(f"   linear solve: {r.its} its, device {r.seconds_total:.2f} s", flush=True)
        return r
    sysm.MultiBlockSystem.setup_preconditioner = timed_setup
    sysm.MultiBlockSystem.solve = timed_solve
    c = Control.Instationary(q["M"], forward, desired_state=lambda t: (q["v_d"][idx[round(float(t), 12)]],
                                                                   q["v_hat"][idx[round(float(t), 12)]]),
                             force_f=lambda t: q["f"][idx[round(float(t), 12)]], beta=q["beta"],
                             Gauss_Newton=not args.picard, n_t=q["n_t"], CN=True, time_interval=q["time_interval"],
                             bc_dofs=q["bdofs"])
    sp_ = {"linear_solver": "fgmres", "maximum_iterations": 100, "relative_tolerance": 1e-6, "absolute_tolerance": 0.0,
           "gmres_restart": 30}
    t0 = time.perf_counter()
    k = c.non_linear_solve(lambda_v_bounds=q["lambda_v_bounds"], solver_parameters=sp_, max_non_linear_iter=args.outer,
                           relative_non_linear_tol=1e-9, print_error_non_linear=False)
    wall = time.perf_counter() - t0
    print({"outer": k, "wall_s": wall, **stats, "history": c.non_linear_history})
    c.close()


if __name__ == "__main__":
    main()
