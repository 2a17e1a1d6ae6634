"""Multi-GPU plumbing: one process per GPU, the DataSet sharded by residue count, ONE all-reduce of the
sufficient statistics per EM iteration.

The reference's only parallel construct is Compute.oneIteration/estimateFromStatistics
(HMM/models/HigherOrderHMM.java:1250-1288): per-worker statistics, joinStatistics (sum over workers, copied back to
every worker) and setParameters (broadcast).  Sum + copy-back is an all-reduce; because every rank then runs the
same deterministic M-step on the same summed statistics no broadcast is needed.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(offsets: np.ndarray, rank: int, world: int):
    """Sequence range [lo, hi) of `rank`: contiguous, balanced by residues — boundary r is the first sequence whose
    start offset reaches r/world of all residues.  (The reference splits by sequence count, i*N/T,
    HigherOrderHMM.java:1241-1248, which imbalances ragged data.)"""
    if not 0 <= rank < world:
        raise ValueError("rank outside the world")
    n = offsets.size - 1
    total = int(offsets[-1])
    cuts = [int(np.searchsorted(offsets[:-1], (total * r) // world, side="left")) for r in range(world)] + [n]
    cuts[0] = 0
    for r in range(1, world + 1):
        cuts[r] = max(cuts[r], cuts[r - 1])
    return cuts[rank], cuts[rank + 1]


def shard_lpt(offsets: np.ndarray, rank: int, world: int) -> np.ndarray:
    """Sequence indices of `rank` under longest-processing-time bin packing by residue count (chromosome-scale
    data — cfg5's 24 sequences of 47..248 Mb — cannot be balanced by a contiguous cut).  Deterministic: sequences by
    decreasing length (ties by index) go to the currently lightest rank (ties by rank); indices are returned ascending."""
    if not 0 <= rank < world:
        raise ValueError("rank outside the world")
    lens = np.diff(offsets)
    order = np.lexsort((np.arange(lens.size), -lens))
    load = [0] * world
    mine = []
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        load[r] += int(lens[i])
        if r == rank:
            mine.append(int(i))
    return np.array(sorted(mine), dtype=np.int64)


def allreduce_stats_(tensor, group=None):
    """In-place sum of the [transition | emission | score] statistics block over all ranks (torch.distributed;
    NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
    return tensor


def broadcast_comm_id(make_id, device=None):
    """Rank 0 creates the 128-byte ncclUniqueId (jh_comm_unique_id); everybody receives it through torch.distributed."""
    import torch
    import torch.distributed as dist

    buf = torch.zeros(128, dtype=torch.uint8, device=device)
    if dist.get_rank() == 0:
        buf.copy_(torch.from_numpy(np.frombuffer(make_id(), dtype=np.uint8).copy()))
    dist.broadcast(buf, src=0)
    return bytes(buf.cpu().numpy().tobytes())
