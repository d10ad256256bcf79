"""N > 1: the row partition and halo plan (host logic, world_size-2 gloo on CPU) and, when
at least two GPUs are present, the NCCL path of the CUDA library against the oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from control_b200 import partition
from synthetic import fem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ownership_is_petsc_split():
    for n, w in ((10, 3), (121, 2), (1050625, 8), (7, 8)):
        ranges = [partition.ownership_range(n, w, r) for r in range(w)]
        assert ranges[0][0] == 0 and sum(c for _, c in ranges) == n
        for (b0, c0), (b1, _) in zip(ranges, ranges[1:]):
            assert b0 + c0 == b1
        counts = [c for _, c in ranges]
        assert max(counts) - min(counts) <= 1 and counts == sorted(counts, reverse=True)


def test_halo_plan_is_consistent_between_ranks():
    M, K, _, _ = fem.assemble_p1_2d(9, 7)
    for world in (2, 3, 4):
        plans = [partition.halo_plan(M.indptr, M.indices, world, r) for r in range(world)]
        for r, pr in enumerate(plans):
            for p, (off, cnt) in pr["recv"].items():
                wanted = pr["ghosts"][off:off + cnt]
                sent = plans[p]["send"][r] + plans[p]["row_begin"]
                assert np.array_equal(wanted, sent)       # same rows, same order, no negotiation


def _gloo_worker(rank, world, port, n_t, out):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    M, K, _, bd = fem.assemble_p1_2d(12, 9, 2.0, 1.0)
    n = M.shape[0]
    plan = partition.halo_plan(M.indptr, M.indices, world, rank)
    rb, nl = plan["row_begin"], plan["n_local"]
    rng = np.random.default_rng(0)
    X = rng.standard_normal((n_t, n))                   # global, same on every rank
    x_loc = partition.local_blocks(X, n, world, rank)   # (n_t, nl)
    ghost = np.zeros((n_t, plan["ghosts"].size))
    reqs = []
    for p, rows in plan["send"].items():
        reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(x_loc[:, rows])), p))
    bufs = {}
    for p, (off, cnt) in plan["recv"].items():
        bufs[p] = torch.zeros((n_t, cnt), dtype=torch.float64)
        reqs.append(dist.irecv(bufs[p], p))
    for r in reqs:
        r.wait()
    for p, (off, cnt) in plan["recv"].items():
        ghost[:, off:off + cnt] = bufs[p].numpy()
    # local rows of K in local column numbering: owned first, ghosts after
    Kl = K[rb:rb + nl].tocsr()
    colmap = -np.ones(n, dtype=np.int64)
    colmap[rb:rb + nl] = np.arange(nl)
    colmap[plan["ghosts"]] = nl + np.arange(plan["ghosts"].size)
    Kl_local = Kl.copy()
    Kl_local.indices = colmap[Kl.indices].astype(np.int32)
    Kl_local._shape = (nl, nl + plan["ghosts"].size)
    y_loc = (Kl_local @ np.concatenate([x_loc, ghost], axis=1).T).T
    y_ref = (K @ X.T).T[:, rb:rb + nl]
    # a global dot product through an all-reduce of local partial sums
    part = torch.tensor([float((x_loc * x_loc).sum())], dtype=torch.float64)
    dist.all_reduce(part)
    ok = np.abs(y_loc - y_ref).max() < 1e-13 * np.abs(y_ref).max() and abs(part.item() - (X * X).sum()) < 1e-9
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_partitioned_product_with_gloo_world_size_2():
    import torch.multiprocessing as mp
    world = 2
    with mp.Manager() as m:
        out = m.dict()
        port = 29500 + (os.getpid() % 2000)
        mp.spawn(_gloo_worker, args=(world, port, 5, out), nprocs=world, join=True)
        assert all(out[r] for r in range(world))


@pytest.mark.gpu
def test_multi_gpu_solve_matches_oracle():
    n_gpu = torch.cuda.device_count()
    if n_gpu < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + os.getpid() % 1000),
           os.path.join(ROOT, "tests", "mp_gpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert "MP_GPU_CHECK_OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
