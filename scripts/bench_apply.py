"""Micro-benchmark of the fused KKT-apply kernel (config C2 by default)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=1024)
    ap.add_argument("--n_t", type=int, default=64)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--be", action="store_true")
    args = ap.parse_args()
    from control_b200 import MultiBlockSystem, _lib as L
    from synthetic import fem
    t = time.time()
    M, K, _, bd = fem.assemble_p1_2d(args.nx, args.nx, 2.0, 2.0)
    print(f"assembled n={M.shape[0]} nnz={M.nnz} in {time.time() - t:.1f}s", flush=True)
    s = MultiBlockSystem(M, K, n_t=args.n_t, beta=1e-4, CN=not args.be, time_interval=(0.0, 2.0), bc_dofs=bd)
    n, N, nnz = s.n, s.N, M.nnz
    x = torch.randn(s.vec_len(L.CTL_LAYOUT_TIME_FASTEST), dtype=torch.float64, device=s.device)
    x.view(2, n, s.ld)[:, :, N:] = 0
    y = torch.empty_like(x)
    s.time_apply(x, y, 3)
    ms = s.time_apply(x, y, args.reps)
    alg = 32.0 * n * N + 20.0 * nnz + 4.0 * (n + 1)
    print(json.dumps({"kernel": "kkt_apply", "n": n, "N": N, "ld": s.ld, "ms": ms,
                      "alg_bytes": alg, "GBps": alg / ms / 1e6,
                      "frac_of_6548.8": alg / ms / 1e6 / 6548.8}))


if __name__ == "__main__":
    main()
