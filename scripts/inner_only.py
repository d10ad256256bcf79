"""N back-to-back AMG inner solves at config C2 through ctl_amg_solve (for ncu launch lists: no flushes, no other kernels)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from synthetic import problems
from control_b200 import MultiBlockSystem
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
q = problems.heat_problem(nx, 64, True)
s = MultiBlockSystem(q["M"], q["K"], n_t=64, beta=q["beta"], CN=True, time_interval=q["time_interval"], bc_dofs=q["bdofs"])
s.setup_preconditioner(lambda_v_bounds=q["lambda_v_bounds"])
b = torch.sin(0.37 * torch.arange(s.n, dtype=torch.float64, device=s.device)) + 0.1
x = torch.zeros_like(b)
torch.cuda.synchronize()
l0 = s.kernel_launches()
for _ in range(reps):
    s._call(s._lib.ctl_amg_solve, 0, b.data_ptr(), x.data_ptr())
torch.cuda.synchronize()
print("kernels per solve", (s.kernel_launches() - l0) / reps, "x norm", float(x.norm()))
