"""Multi-GPU plumbing of the conversion path: one process per GPU, every rank converts its own
contiguous shard and only a few integers per rank are exchanged. The payload never
crosses NVLink.

Two ways to shard .bin -> .binpack:

* independent shards (``exchange_offsets``): every rank compresses its records as a file of its own
  and the outputs are concatenated. That is what the reference produces when it is run once per
  shard file with ``-a`` (BINP chunks are self-delimiting, compress_file.cpp:449-522, :1663-1666);
  chains that cross a shard boundary are cut and chunks restart per shard.
* one file (``compress_sharded``): byte-identical to ONE reference run over all records. Chains
  belong to the rank that holds their head (the rank before sees the rest of the chain through an
  overlap window), and the chunk-flush rule is replayed in rank order with a 16-byte carry.

.binpack -> .bin / .plain shards ONE file by chunk ranges (``decompress_sharded``): rank r decodes the
contiguous chunk range shard_bounds(chunks, world, r) (nnp_shard_decompress_dev), the ranks all-gather
their position counts and rank r's records belong at byte 40 * (positions before it); the
concatenation equals the single-run output byte for byte because chunks are independent
(compress_file.cpp:468-480, :1128-1214).
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
    sizes = _gather_ints(int(local_bytes), world, device, group)
    offs = offsets_from_sizes(sizes)
    return offs[rank], sum(sizes), sizes


def _gather_rows(values: List[int], world: int, device=None, group=None) -> List[List[int]]:
    """One all-gather of a few int64 per rank, read back with a single device-to-host copy."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return [[int(v) for v in values]]
    mine = torch.tensor([int(v) for v in values], dtype=torch.int64, device=device)
    out = torch.empty(world * len(values), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, mine, group=group)
    flat = [int(v) for v in out.tolist()]
    return [flat[r * len(values):(r + 1) * len(values)] for r in range(world)]


def _gather_ints(value: int, world: int, device=None, group=None) -> List[int]:
    return [row[0] for row in _gather_rows([value], world, device, group)]


class ShardStatusError(RuntimeError):
    """A rank's local step of a sharded conversion failed; every rank raises this together, before any
    further collective, so that no rank is left waiting in one."""

    def __init__(self, statuses: List[int]):
        super().__init__(f"sharded conversion: per-rank status {statuses}")
        self.statuses = statuses


def decompress_sharded(decode, device=None, group=None) -> Tuple[int, int, int]:
    """ONE .binpack decoded by all ranks (BASELINE configs[2]): ``decode(world, rank) ->
    bytes produced`` wraps nnp_shard_decompress_dev (the rank's contiguous chunk range into the rank's
    own buffer); one all-gather of the byte counts gives every rank the offset of its records in the
    .bin file. Returns (bytes produced, file offset of this rank's records, total file bytes)."""
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized():
        world, rank = 1, 0
    else:
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    produced = int(decode(world, rank))
    sizes = _gather_ints(produced, world, device, group)
    return produced, offsets_from_sizes(sizes)[rank], sum(sizes)


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


# ---------------------------------------------------------------------------------------------
# one .binpack from record ranges spread over ranks (include/nnuepack.h, "sharded" entry points)

NO_CARRY = (1 << 64) - 1


def shard_window(n_records: int, world: int, rank: int, overlap: int):
    """Record range a rank must hold for ``compress_sharded``: its balanced share plus one halo record
    in front and ``overlap`` records behind. Returns (g0, g1, own_lo, own_hi, reaches_eof): [g0, g1)
    are file record indices, own_lo / own_hi index into that buffer."""
    lo, hi = shard_bounds(n_records, world, rank)
    g0 = lo - 1 if lo > 0 else 0
    g1 = min(n_records, hi + overlap)
    return g0, g1, lo - g0, hi - g0, g1 == n_records


ORBIT_TABLE_ENTRIES = 30848  # NNP_ORBIT_TABLE_ENTRIES (include/nnuepack.h)


def compress_sharded(payload_bytes: int, orbit, emit, device=None, group=None, table=None, resolve=None, status: int = 0):
    """The exchange steps 2-4 of the sharded compressor around a rank's local calls.

    ``payload_bytes``: from nnp_shard_compress_begin_dev (step 1, done by the caller).
    ``orbit(payload_base, carry_in) -> (n_chunk_starts, first_start, carry_out)`` and
    ``emit(next_start) -> result`` wrap nnp_shard_compress_orbit / nnp_shard_compress_emit_dev.
    With ``table() -> int64 tensor of 3 * ORBIT_TABLE_ENTRIES`` and ``resolve(tables, sizes, world, rank) ->
    (carry_in, chunks_before, next_start, total_chunks)`` (nnp_shard_compress_table_dev / _resolve_dev) the
    rank-order carry is replaced by one all-gather of the ranks' orbit tables; without them the carry
    travels from rank to rank (16 bytes per hop).
    ``status``: the nnp_status of the rank's nnp_shard_compress_begin_dev. NNP_ERR_WINDOW and
    NNP_ERR_BAD_SFEN are rank-local outcomes, so the statuses travel with the payload sizes in the first
    all-gather and every rank raises ShardStatusError together when one of them is not 0 (the caller
    widens the overlap window on every rank, or gives up on every rank).
    Returns (emit's result, file offset of this rank's slice, total file bytes)."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized():
        world, rank = 1, 0
    else:
        world, rank = dist.get_world_size(group), dist.get_rank(group)

    def gather(value: int):
        return _gather_ints(value, world, device, group)

    rows = _gather_rows([payload_bytes, status], world, device, group)
    if any(row[1] != 0 for row in rows):
        raise ShardStatusError([row[1] for row in rows])
    sizes = [row[0] for row in rows]
    bases = offsets_from_sizes(sizes)
    total_payload = sum(sizes)

    if table is not None and resolve is not None:
        mine = table()
        if world == 1:
            tables = mine
        else:
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine, group=group)
            tables = torch.cat(parts)
        carry_in, chunks_before, next_start, total_chunks = resolve(tables, sizes, world, rank)
        orbit(bases[rank], carry_in)
        result = emit(next_start)
        return result, bases[rank] + 8 * chunks_before, total_payload + 8 * total_chunks

    def p2p(op, tensor, peer):
        for req in dist.batch_isend_irecv([dist.P2POp(op, tensor, peer, group=group)]):
            req.wait()

    # the chunk-flush rule in rank order: (offset of the last chunk start or -1, chunks so far)
    carry, chunks_before = -1, 0
    if rank > 0:
        buf = torch.zeros(2, dtype=torch.int64, device=device)
        p2p(dist.irecv, buf, rank - 1)
        carry, chunks_before = int(buf[0].item()), int(buf[1].item())
    n_starts, first_start, carry_out = orbit(bases[rank], NO_CARRY if carry < 0 else carry)
    if rank < world - 1:
        buf = torch.tensor([-1 if carry_out == NO_CARRY else carry_out, chunks_before + n_starts], dtype=torch.int64,
                           device=device)
        p2p(dist.isend, buf, rank + 1)

    firsts = gather(-1 if first_start == NO_CARRY else first_start)
    later = [f for f in firsts[rank + 1:] if f >= 0]
    next_start = later[0] if later else total_payload
    result = emit(next_start)
    all_chunks = gather(n_starts)
    total_file = total_payload + 8 * sum(all_chunks)
    return result, bases[rank] + 8 * chunks_before, total_file
