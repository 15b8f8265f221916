"""Stream sharding across GPUs: one process per GPU, no collective on the data path.

A stream (one file = all of its channels) shares nothing with any other stream, so the batch is partitioned by stream
index.  Equal-length batches use a contiguous block partition; ragged batches are sorted by length and dealt
round-robin so every rank gets a similar amount of audio.  The only communication is the host-side gather of results
(torch.distributed object gather; NCCL or gloo), which is not on the hot path.
"""
from __future__ import annotations

from typing import Callable, List, Sequence

import numpy as np


def block_partition(n_streams: int, world: int, rank: int) -> range:
    """Contiguous, balanced block of stream indices for `rank`."""
    base, extra = divmod(n_streams, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def balanced_partition(lengths: Sequence[int], world: int) -> List[np.ndarray]:
    """Ragged batches: longest first, dealt round-robin (boustrophedon) -> per-rank index arrays (sorted)."""
    order = np.argsort(-np.asarray(lengths, dtype=np.int64), kind="stable")
    shards: List[list] = [[] for _ in range(world)]
    for i, idx in enumerate(order):
        lap, pos = divmod(i, world)
        shards[pos if lap % 2 == 0 else world - 1 - pos].append(int(idx))
    return [np.array(sorted(s), dtype=np.int64) for s in shards]


def library_partition(lengths: Sequence[int], world: int) -> List[np.ndarray]:
    """The partition libpvgpu.so itself uses for pvgpu_mbatch (pvgpu_shard_streams): per-rank index arrays (sorted)."""
    from .phasevocoder import shard_streams
    owner = shard_streams(lengths, world) if len(lengths) else np.zeros(0, np.int32)
    return [np.flatnonzero(owner == r).astype(np.int64) for r in range(world)]


def run_sharded(streams: Sequence[np.ndarray], process_local: Callable[[List[np.ndarray]], List[np.ndarray]], group=None):
    """Every rank holds the full list of host streams (or at least its own shard's entries), processes its shard with
    `process_local` (the GPU batch on that rank) and rank 0 receives all outputs in the original order."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lengths = [int(x.shape[1]) for x in streams]
    shards = library_partition(lengths, world)
    mine = shards[rank]
    local_out = process_local([streams[int(i)] for i in mine]) if len(mine) else []
    if world == 1:
        return list(local_out)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((mine.tolist(), local_out), gathered, dst=0, group=group)
    if rank != 0:
        return None
    out: List = [None] * len(streams)
    for idxs, outs in gathered:
        for i, y in zip(idxs, outs):
            out[i] = y
    return out
