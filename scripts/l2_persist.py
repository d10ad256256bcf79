import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import kat
from control_b200 import MultiBlockSystem
q = kat.heat_problem(1024, 64, True)
s = MultiBlockSystem(q["M"], q["K"], n_t=64, beta=q["beta"], CN=True, time_interval=q["time_interval"], bc_dofs=q["bdofs"])
s.setup_preconditioner(lambda_v_bounds=q["lambda_v_bounds"])
a = s.micro_benchmarks(reps=30, flush_l2=False)
print(os.environ.get("CTL_L2_PERSIST"), "inner_solve_ms %.3f cheb warm %.4f resid %.4f" % (a["inner_solve_ms"], a["cheb_ms"], a["residual_ms"]), flush=True)
