"""Cell-sharded multi-GPU plumbing: one process per GPU, torch.distributed for rendezvous only.

The data path never goes through torch: each rank owns a `Context`, joins the library's NCCL
communicator (sb_comm_init) and runs the same call sequence on its own contiguous cell range.
A^T.Y stays shard-local; A.X partials, gene moments, gene totals and the Gram matrix are all-reduced
inside the library; gene-sized results come back replicated, cell-sized ones per shard."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def shard_bounds(n_cells: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, near-equal cell ranges: [lo, hi) of `rank`."""
    return n_cells * rank // world, n_cells * (rank + 1) // world


def shard_bounds_by_nnz(cell_nnz: Sequence[int], world: int) -> List[Tuple[int, int]]:
    """Contiguous ranges with near-equal non-zero counts (matters for skewed inputs)."""
    csum = np.concatenate([[0], np.cumsum(np.asarray(cell_nnz, dtype=np.int64))])
    total = int(csum[-1])
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(csum, total * r / world, side="left")))
    cuts.append(len(cell_nnz))
    cuts = np.maximum.accumulate(cuts)
    return [(int(cuts[r]), int(cuts[r + 1])) for r in range(world)]


def init_context(local_rank: int):
    """Context for this rank, joined to the library communicator when torch.distributed is initialised."""
    import torch.distributed as dist
    from .sqz import Context
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    ctx = Context(local_rank)
    if world > 1:
        obj = [Context.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        ctx.comm_init(world, rank, obj[0])
    return ctx


def gather_rows(local: np.ndarray) -> np.ndarray | None:
    """Concatenate per-rank row blocks (e.g. the V shards) on rank 0."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    parts = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(local, parts, dst=0)
    return np.concatenate(parts, axis=0) if dist.get_rank() == 0 else None
