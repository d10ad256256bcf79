"""A few applications of the Stokes outer operator at config C4 size (profiling target for
panel_spmm_kernel: 4 launches per application, B^T twice then B twice)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from synthetic import problems                      # noqa: E402
from control_b200.stokes import StokesSystem        # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 512
q = problems.stokes_problem(nx, 32, True)
th = q["th"]
s = StokesSystem(th["M_v"], th["K_v"], th["B"], th["M_p"], th["K_p"], n_t=32, beta=1.0, CN=True,
                 time_interval=q["time_interval"], bc_dofs_v=q["bdofs"])
x = torch.from_numpy(np.random.default_rng(0).standard_normal(s.vec_len())).to(s.device)
y = torch.empty_like(x)
for _ in range(3):
    s.apply(x, y)
torch.cuda.synchronize()
print("ok", float(y.norm()))
