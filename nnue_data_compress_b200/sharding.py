"""Multi-GPU plumbing of the conversion path: one process per GPU, every rank converts its own
contiguous shard, and the only exchange is an all-gather of the per-shard output byte counts,
which gives each shard its offset in the concatenated output. The payload never
crosses NVLink.

Sharded output semantics: shard outputs are concatenated. For .binpack that is exactly what the
reference produces when it is run once per shard file with ``-a`` (BINP chunks are
self-delimiting, compress_file.cpp:449-522, :1663-1666); for .bin and .plain outputs the
concatenation equals the single-run output byte for byte because records are independent.
"""
from __future__ import annotations

from typing import List, Tuple


def shard_bounds(n_units: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) range of units (records or chunks) for `rank`."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(n_units, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def offsets_from_sizes(sizes: List[int]) -> List[int]:
    out, run = [], 0
    for s in sizes:
        out.append(run)
        run += int(s)
    return out


def exchange_offsets(local_bytes: int, device=None, group=None) -> Tuple[int, int, List[int]]:
    """All-gathers the per-rank byte counts (NCCL on GPUs, gloo on CPU) and returns
    (this rank's output offset, total bytes, all sizes)."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized():
        return 0, int(local_bytes), [int(local_bytes)]
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    mine = torch.tensor([int(local_bytes)], dtype=torch.int64, device=device)
    gathered = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    sizes = [int(t.item()) for t in gathered]
    offs = offsets_from_sizes(sizes)
    return offs[rank], sum(sizes), sizes


def chunk_bounds_binpack(data: bytes) -> List[Tuple[int, int]]:
    """(offset, total length incl. header) of every BINP chunk: the unit decompression shards by.
    Header walk only (compress_file.cpp:500-521); raises on a bad magic."""
    out, pos, n = [], 0, len(data)
    while pos < n:
        if n - pos < 8 or data[pos:pos + 4] != b"BINP":
            raise ValueError("Invalid binpack file or chunk.")
        size = int.from_bytes(data[pos + 4:pos + 8], "little")
        out.append((pos, 8 + size))
        pos += 8 + size
    return out
