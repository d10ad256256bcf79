"""Correctness and timing of an opt-in KKT-apply variant against the default kernel in ONE process:
``python scripts/compare_apply_variants.py --env CTL_KKT_TMA=2 [--nx 512] [--be]``.  The environment switch is read
when a handle is created, so the two systems below run different kernels."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--env", action="append", default=[], help="NAME=VALUE set while the variant's handle is created")
    ap.add_argument("--nx", type=int, default=512)
    ap.add_argument("--n_t", type=int, default=64)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--be", action="store_true")
    args = ap.parse_args()
    from control_b200 import MultiBlockSystem, _lib as L
    from synthetic import fem
    M, K, _, bd = fem.assemble_p1_2d(args.nx, args.nx, 2.0, 2.0)
    kw = dict(n_t=args.n_t, beta=1e-4, CN=not args.be, time_interval=(0.0, 2.0), bc_dofs=bd)
    base = MultiBlockSystem(M, K, **kw)
    for kv in args.env:
        k, v = kv.split("=", 1)
        os.environ[k] = v
    var = MultiBlockSystem(M, K, **kw)
    n, N = base.n, base.N
    x = torch.randn(base.vec_len(L.CTL_LAYOUT_TIME_FASTEST), dtype=torch.float64, device=base.device)
    x.view(2, n, base.ld)[:, :, N:] = 0
    y0, y1 = torch.empty_like(x), torch.empty_like(x)
    base.time_apply(x, y0, 1)
    var.time_apply(x, y1, 1)
    torch.cuda.synchronize()
    err = float((y0 - y1).abs().max() / y0.abs().max())
    out = {"variant": args.env, "n": n, "N": N, "rel_diff_vs_default": err}
    print(json.dumps(out), flush=True)
    out["default_ms"] = base.time_apply(x, y0, args.reps)
    out["variant_ms"] = var.time_apply(x, y1, args.reps)
    alg = 32.0 * n * N + 20.0 * M.nnz + 4.0 * (n + 1)
    out["default_frac"] = alg / out["default_ms"] / 1e6 / 6548.8
    out["variant_frac"] = alg / out["variant_ms"] / 1e6 / 6548.8
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
