"""Hypothesis test: does a patch (2-D tile) row ordering speed up the KKT apply?"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from control_b200 import MultiBlockSystem, _lib as L
from synthetic import fem
nx = 1024
M, K, _, bd = fem.assemble_p1_2d(nx, nx, 2.0, 2.0)
n = M.shape[0]
for tile in (0, 8, 16):
    if tile:
        ii, jj = np.meshgrid(np.arange(nx + 1), np.arange(nx + 1), indexing="xy")
        key = ((jj // tile) * ((nx + tile) // tile) + (ii // tile)).ravel() * (tile * tile) + ((jj % tile) * tile + (ii % tile)).ravel()
        perm = np.argsort(key, kind="stable")
        Mp = M[perm][:, perm].tocsr(); Kp = K[perm][:, perm].tocsr()
        Mp.sort_indices(); Kp.sort_indices()
        inv = np.empty(n, dtype=np.int64); inv[perm] = np.arange(n)
        bdp = inv[bd].astype(np.int32)
    else:
        Mp, Kp, bdp = M, K, bd
    s = MultiBlockSystem(Mp, Kp, n_t=64, beta=1e-4, CN=True, time_interval=(0.0, 2.0), bc_dofs=bdp)
    x = torch.randn(s.vec_len(L.CTL_LAYOUT_TIME_FASTEST), dtype=torch.float64, device=s.device)
    y = torch.empty_like(x)
    s.time_apply(x, y, 3)
    ms = s.time_apply(x, y, 20)
    alg = 32.0 * n * s.N + 20.0 * M.nnz + 4.0 * (n + 1)
    print(json.dumps({"tile": tile, "ms": ms, "GBps": alg / ms / 1e6, "frac": alg / ms / 1e6 / 6548.8}), flush=True)
    s.close()
