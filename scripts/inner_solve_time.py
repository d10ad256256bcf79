"""Time of one AMG inner solve at config C2 (warm, back to back) -- for A/B experiments via env."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from synthetic import problems
from control_b200 import MultiBlockSystem
q = problems.heat_problem(1024, 64, True)
s = MultiBlockSystem(q["M"], q["K"], n_t=64, beta=q["beta"], CN=True, time_interval=q["time_interval"], bc_dofs=q["bdofs"])
s.setup_preconditioner(lambda_v_bounds=q["lambda_v_bounds"])
a = s.micro_benchmarks(reps=30, flush_l2=False)
print("inner_solve_ms %.4f kernels %d" % (a["inner_solve_ms"], a["inner_solve_kernels"]))
