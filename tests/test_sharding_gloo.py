"""CPU suite: the multi-rank path (world_size 2, gloo). Each rank converts its own shard --
with the oracle standing in for the CUDA converter -- exchanges byte counts with the same
all-gather the GPU path uses, and the shard outputs assembled at the gathered offsets must equal
what the reference semantics prescribe (one run per shard, appended)."""
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, tmpdir):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ctypes

    import nnue_data_compress_b200 as pkg
    from nnue_data_compress_b200.sharding import chunk_bounds_binpack, decompress_sharded, exchange_offsets, shard_bounds
    from refutil import BIN_TO_BINPACK, BINPACK_TO_BIN, golden, oracle_convert

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b = golden("twochunks.bin")
        n = len(b) // 40
        lo, hi = shard_bounds(n, world, rank)
        rc, mine = oracle_convert(BIN_TO_BINPACK, b[lo * 40:hi * 40])
        assert rc == 0
        off, total, sizes = exchange_offsets(len(mine))
        assert sizes[rank] == len(mine) and total == sum(sizes)
        with open(os.path.join(tmpdir, f"pack{rank}"), "wb") as f:
            f.write(off.to_bytes(8, "little") + mine)
        # decompression: ONE file, chunk ranges per rank (sharding.decompress_sharded). The range comes
        # from the library's host-side header walk (nnp_binpack_chunk_range, no GPU needed); the oracle
        # stands in for nnp_binpack_to_bin_dev on the rank's bytes.
        full = golden("twochunks.binpack")
        out = {}

        def decode(w, r):
            rng = pkg.ChunkRange()
            assert pkg.lib().nnp_binpack_chunk_range(full, len(full), w, r, ctypes.byref(rng)) == 0
            chunks = chunk_bounds_binpack(full)
            clo, chi = shard_bounds(len(chunks), w, r)
            assert (rng.chunks_total, rng.chunk_lo, rng.chunk_hi) == (len(chunks), clo, chi)
            assert rng.byte_lo == (chunks[clo][0] if clo < len(chunks) else len(full))
            rc, rec = oracle_convert(BINPACK_TO_BIN, full[rng.byte_lo:rng.byte_hi])
            assert rc == 0
            out["rec"] = rec
            return len(rec)

        got, off2, total2 = decompress_sharded(decode)
        assert got == len(out["rec"]) and total2 == len(golden("twochunks.rt.bin"))
        with open(os.path.join(tmpdir, f"bin{rank}"), "wb") as f:
            f.write(off2.to_bytes(8, "little") + out["rec"])
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_shards_assemble(tmp_path):
    import torch.multiprocessing as mp

    from refutil import BIN_TO_BINPACK, golden, oracle_convert
    from nnue_data_compress_b200.sharding import shard_bounds

    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)

    b = golden("twochunks.bin")
    n = len(b) // 40
    expect = b""
    for r in range(world):
        lo, hi = shard_bounds(n, world, r)
        expect += oracle_convert(BIN_TO_BINPACK, b[lo * 40:hi * 40])[1]  # reference: one run per shard, -a
    out = bytearray(len(expect))
    for r in range(world):
        raw = (tmp_path / f"pack{r}").read_bytes()
        off = int.from_bytes(raw[:8], "little")
        out[off:off + len(raw) - 8] = raw[8:]
    assert bytes(out) == expect

    full_bin = golden("twochunks.rt.bin")
    out = bytearray(len(full_bin))
    for r in range(world):
        raw = (tmp_path / f"bin{r}").read_bytes()
        off = int.from_bytes(raw[:8], "little")
        out[off:off + len(raw) - 8] = raw[8:]
    assert bytes(out) == full_bin  # chunk-sharded decompression is exact


def test_shard_bounds_cover_everything():
    from nnue_data_compress_b200.sharding import offsets_from_sizes, shard_bounds

    for n in (0, 1, 7, 8, 9, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    assert offsets_from_sizes([3, 0, 5]) == [0, 3, 3]


# ---------------------------------------------------------------------------------------------
# one .binpack from several ranks, byte-identical to a single run


def _worker_one_file(rank, world, port, tmpdir, name, overlap, small_threshold, use_table):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import shardsim
    from nnue_data_compress_b200.sharding import compress_sharded, shard_window
    from refutil import golden

    if small_threshold:
        shardsim.THRESHOLD = small_threshold
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b = golden(name)
        n = len(b) // 40
        g0, g1, lo, hi, eof = shard_window(n, world, rank, overlap)
        sh = shardsim.OracleShard(b[g0 * 40:g1 * 40], lo, hi, eof)
        extra = dict(table=sh.table, resolve=sh.resolve) if use_table else {}
        data, off, total = compress_sharded(sh.payload_bytes, sh.orbit, sh.emit, **extra)
        with open(os.path.join(tmpdir, f"slice{rank}"), "wb") as f:
            f.write(off.to_bytes(8, "little") + total.to_bytes(8, "little") + data)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("use_table", [False, True])
@pytest.mark.parametrize("name,world,overlap", [("twochunks.bin", 2, 64), ("twochunks.bin", 3, 8), ("games100.bin", 3, 128),
                                               ("long400.bin", 3, 500), ("restart.bin", 2, 200)])
def test_ranks_write_one_file(tmp_path, name, world, overlap, use_table):
    """The rank orchestration (all-gather of payload sizes, carry of the chunk-flush rule in rank
    order, all-gather of first chunk starts) with the oracle standing in for the CUDA entry points:
    the slices assembled at their offsets are the single-run .binpack of the whole input."""
    import torch.multiprocessing as mp

    from refutil import BIN_TO_BINPACK, golden, oracle_convert

    port = _free_port()
    mp.spawn(_worker_one_file, args=(world, port, str(tmp_path), name, overlap, 0, use_table), nprocs=world, join=True)
    rc, expect = oracle_convert(BIN_TO_BINPACK, golden(name))
    assert rc == 0
    out = bytearray(len(expect))
    for r in range(world):
        raw = (tmp_path / f"slice{r}").read_bytes()
        off, total = int.from_bytes(raw[:8], "little"), int.from_bytes(raw[8:16], "little")
        assert total == len(expect)
        out[off:off + len(raw) - 16] = raw[16:]
    assert bytes(out) == expect
