"""Row partition of the spatial dofs over ranks (host-side mirror of the plan the CUDA
library derives in csrc/api.cu + csrc/comm.cu).

The reference distributes its matrices and vectors the PETSc way: every rank owns a
contiguous block of rows of every time block (preconditioner/preconditioner.py:706-722 sizes
the shell matrix ((n_local, N_global), ...)).  ``ownership_range`` is PETSc's
``PetscSplitOwnership`` rule; ``halo_plan`` lists, for one rank, the ghost columns it reads
(sorted by global index = grouped by owner) and the owned rows every peer reads from it.
No communication is needed to build the plan: every rank holds the global pattern.
"""
import numpy as np

__all__ = ["ownership_range", "halo_plan", "local_blocks", "gather_blocks"]


def ownership_range(n, world, rank):
    """(first owned row, number of owned rows)."""
    base, rem = divmod(n, world)
    return rank * base + min(rank, rem), base + (1 if rank < rem else 0)


def halo_plan(indptr, indices, world, rank):
    """dict(ghosts=global ids of this rank's ghost columns (sorted),
            recv={peer: (offset into ghosts, count)},
            send={peer: local owned rows the peer needs, in the peer's ghost order})."""
    n = len(indptr) - 1
    rb, nl = ownership_range(n, world, rank)
    cols = np.unique(indices[indptr[rb]:indptr[rb + nl]])
    ghosts = cols[(cols < rb) | (cols >= rb + nl)]
    recv, send = {}, {}
    for p in range(world):
        if p == rank:
            continue
        pb, pc = ownership_range(n, world, p)
        lo, hi = np.searchsorted(ghosts, [pb, pb + pc])
        pcols = np.unique(indices[indptr[pb]:indptr[pb + pc]])
        need = pcols[(pcols >= rb) & (pcols < rb + nl)]
        if hi > lo or need.size:
            recv[p] = (int(lo), int(hi - lo))
            send[p] = (need - rb).astype(np.int64)
    return dict(ghosts=ghosts.astype(np.int64), recv=recv, send=send, row_begin=rb, n_local=nl)


def local_blocks(x, n, world, rank):
    """This rank's columns of a global block array (N, n)."""
    rb, nl = ownership_range(n, world, rank)
    return np.ascontiguousarray(np.asarray(x)[:, rb:rb + nl])


def gather_blocks(parts):
    """Inverse of ``local_blocks`` over all ranks (list ordered by rank)."""
    return np.concatenate(parts, axis=1)
