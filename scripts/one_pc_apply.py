"""One preconditioner application at config C2 (for ncu launch lists)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import kat
from control_b200 import MultiBlockSystem
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
q = kat.heat_problem(nx, 64, True)
s = MultiBlockSystem(q["M"], q["K"], n_t=64, beta=q["beta"], CN=True, time_interval=q["time_interval"], bc_dofs=q["bdofs"])
s.setup_preconditioner(lambda_v_bounds=q["lambda_v_bounds"])
b = torch.randn(s.vec_len(), dtype=torch.float64, device=s.device)
u = s.pc_apply(b)
torch.cuda.synchronize()
print("ok", float(u.abs().max()))
