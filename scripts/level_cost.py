"""In-situ cost of each AMG level: time one inner solve with the hierarchy truncated at 1..L levels."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import kat
from control_b200 import MultiBlockSystem
q = kat.heat_problem(1024, 64, True)
s = MultiBlockSystem(q["M"], q["K"], n_t=64, beta=q["beta"], CN=True, time_interval=q["time_interval"], bc_dofs=q["bdofs"])
for ml in (1, 2, 3, 4, 5, 6):
    s.setup_preconditioner(lambda_v_bounds=q["lambda_v_bounds"], max_levels=ml)
    a = s.micro_benchmarks(reps=20, flush_l2=False)
    b = s.micro_benchmarks(reps=20, flush_l2=True)
    print(ml, "levels", s._lib.ctl_amg_num_levels(s._h, 0), "inner_solve_ms warm %.3f flushed-start %.3f kernels %d | cheb warm %.4f cold %.4f resid warm %.4f" % (
        a["inner_solve_ms"], b["inner_solve_ms"], a["inner_solve_kernels"], a["cheb_ms"], b["cheb_ms"], a["residual_ms"]), flush=True)
