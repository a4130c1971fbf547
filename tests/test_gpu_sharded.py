"""GPU suite: sharded .bin -> .binpack through the C ABI (nnp_shard_compress_*). One process plays
all ranks in turn (the library holds one shard at a time, so the orbit pass and the emit pass each
redo `begin`); the slices written at their offsets must be the single-run .binpack byte for byte."""
import ctypes
import os

import pytest

from refutil import BIN_TO_BINPACK, golden, oracle_convert

pytestmark = pytest.mark.gpu

NO_CARRY = (1 << 64) - 1


def _sharded(nnp, b, world, overlap):
    import numpy as np
    import torch

    from nnue_data_compress_b200.sharding import offsets_from_sizes, shard_window

    L = nnp.lib()
    n = len(b) // 40
    d_all = torch.from_numpy(np.frombuffer(b, dtype=np.uint8).copy()).cuda()

    def begin(r):
        g0, g1, lo, hi, eof = shard_window(n, world, r, overlap)
        info = nnp.ShardInfo()
        rc = L.nnp_shard_compress_begin_dev(ctypes.c_void_p(d_all.data_ptr() + g0 * 40), g1 - g0, lo, hi, int(eof),
                                            ctypes.byref(info))
        assert rc == 0, (r, rc)
        return info

    def orbit(base, carry):
        a, f, c = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
        assert L.nnp_shard_compress_orbit(base, carry, ctypes.byref(a), ctypes.byref(f), ctypes.byref(c)) == 0
        return a.value, f.value, c.value

    sizes = [begin(r).payload_bytes for r in range(world)]
    bases = offsets_from_sizes(sizes)
    total_payload = sum(sizes)
    carry, chunks = NO_CARRY, 0
    carry_in, chunks_before, firsts = [], [], []
    for r in range(world):
        begin(r)
        carry_in.append(carry)
        chunks_before.append(chunks)
        ns, first, carry = orbit(bases[r], carry)
        chunks += ns
        firsts.append(first)
    # the table form of the carry chain must agree with the rank-order chain for every rank
    import torch as _torch

    from nnue_data_compress_b200.sharding import ORBIT_TABLE_ENTRIES

    tables = []
    for r in range(world):
        begin(r)
        t = _torch.empty(3 * ORBIT_TABLE_ENTRIES, dtype=_torch.int64, device="cuda")
        assert L.nnp_shard_compress_table_dev(ctypes.c_void_p(t.data_ptr())) == 0
        tables.append(t)
    all_tables = _torch.cat(tables)
    arr = (ctypes.c_uint64 * world)(*sizes)
    for r in range(world):
        c, b2, nx, tot = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
        assert L.nnp_shard_compress_resolve_dev(ctypes.c_void_p(all_tables.data_ptr()), arr, world, r, ctypes.byref(c),
                                                ctypes.byref(b2), ctypes.byref(nx), ctypes.byref(tot)) == 0
        later = [f for f in firsts[r + 1:] if f != NO_CARRY]
        assert (c.value, b2.value, tot.value) == (carry_in[r], chunks_before[r], chunks), r
        assert nx.value == (later[0] if later else total_payload), r
    out = bytearray(total_payload + 8 * chunks)
    for r in range(world):
        begin(r)
        ns, _, _ = orbit(bases[r], carry_in[r])
        later = [f for f in firsts[r + 1:] if f != NO_CARRY]
        next_start = later[0] if later else total_payload
        need = ctypes.c_size_t(0)
        assert L.nnp_shard_compress_emit_dev(next_start, None, 0, ctypes.byref(need)) == 0
        assert need.value == sizes[r] + 8 * ns
        d_out = torch.empty(max(need.value, 8), dtype=torch.uint8, device="cuda")
        got = ctypes.c_size_t(0)
        assert L.nnp_shard_compress_emit_dev(next_start, ctypes.c_void_p(d_out.data_ptr()), need.value, ctypes.byref(got)) == 0
        off = bases[r] + 8 * chunks_before[r]
        out[off:off + got.value] = d_out[: got.value].cpu().numpy().tobytes()
    return bytes(out)


@pytest.mark.parametrize("name,world,overlap", [("twochunks", 2, 64), ("twochunks", 5, 8), ("games100", 3, 128),
                                               ("long400", 4, 500), ("restart", 2, 200), ("heads", 3, 4),
                                               ("shuffled", 7, 4)])
def test_sharded_golden(nnp, name, world, overlap):
    assert _sharded(nnp, golden(name + ".bin"), world, overlap) == golden(name + ".binpack")


@pytest.mark.parametrize("n,plies,world,overlap", [(600_000, 100, 4, 1024), (2_000_000, 3, 8, 16), (300_000, 400, 8, 4096),
                                                  (500_000, 100, 1, 0)])
def test_sharded_synthetic(nnp, n, plies, world, overlap):
    b = nnp.generate_bin(n, plies, 17)
    rc, want = oracle_convert(BIN_TO_BINPACK, b)
    assert rc == 0
    assert _sharded(nnp, b, world, overlap) == want
    assert nnp.bin_to_binpack(b) == want


def test_sharded_reference_16m(nnp):
    """The 8-rank one-file compressor against ONE run of the compiled reference on 16 M records."""
    from refutil import have_ref, ref_convert

    if not have_ref():
        pytest.skip("oracle/_ref not built")
    b = nnp.generate_bin(16_000_000, 100, 4242)
    want = ref_convert(BIN_TO_BINPACK, b)
    assert len(want) > 30 << 20
    assert _sharded(nnp, b, 8, 65536) == want


def test_sharded_window_too_small(nnp):
    import numpy as np
    import torch

    b = golden("long400.bin")
    d = torch.from_numpy(np.frombuffer(b, dtype=np.uint8).copy()).cuda()
    info = nnp.ShardInfo()
    rc = nnp.lib().nnp_shard_compress_begin_dev(ctypes.c_void_p(d.data_ptr()), 60, 0, 50, 0, ctypes.byref(info))
    assert rc == -12  # NNP_ERR_WINDOW: the 400-ply chain does not end within 10 records of the boundary


def test_shard_sequence_ends_when_another_compressor_call_reuses_its_buffers(nnp):
    """begin / orbit / emit keep pointers into the context's workspace; a whole-buffer conversion in between
    reuses those slots, so the rest of the sequence must be refused (NNP_ERR_BAD_ARG), not run on its memory."""
    import numpy as np
    import torch

    L = nnp.lib()
    b = golden("games100.bin")
    d = torch.from_numpy(np.frombuffer(b, dtype=np.uint8).copy()).cuda()
    info = nnp.ShardInfo()
    n = len(b) // 40
    assert L.nnp_shard_compress_begin_dev(ctypes.c_void_p(d.data_ptr()), n, 0, n, 1, ctypes.byref(info)) == 0
    assert nnp.bin_to_binpack(golden("heads.bin")) == golden("heads.binpack")
    a, f, c = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
    assert L.nnp_shard_compress_orbit(0, NO_CARRY, ctypes.byref(a), ctypes.byref(f), ctypes.byref(c)) == -6  # NNP_ERR_BAD_ARG
    assert L.nnp_shard_compress_begin_dev(ctypes.c_void_p(d.data_ptr()), n, 0, n, 1, ctypes.byref(info)) == 0
    assert L.nnp_shard_compress_orbit(0, NO_CARRY, ctypes.byref(a), ctypes.byref(f), ctypes.byref(c)) == 0


# ---------------------------------------------------------------------------------------------
# ONE .binpack decoded by several ranks (nnp_shard_decompress_dev; BASELINE configs[2])


def _shard_decompress(nnp, pack: bytes, world: int):
    """One process plays all ranks in turn: every rank decodes its chunk range of the whole file into a
    buffer of its own; the pieces placed at 40 * (positions of the ranks before) are the .bin file."""
    import numpy as np
    import torch

    d_all = torch.from_numpy(np.frombuffer(pack, dtype=np.uint8).copy()).cuda() if pack else torch.empty(0, dtype=torch.uint8,
                                                                                                       device="cuda")
    pieces, ranges = [], []
    for r in range(world):
        count, rng = nnp.shard_decompress(d_all, world, r, None)
        d_out = torch.empty(max(count, 8), dtype=torch.uint8, device="cuda")
        got, rng2 = nnp.shard_decompress(d_all, world, r, d_out)
        assert got == count and got == 40 * rng2.positions
        assert (rng.chunk_lo, rng.chunk_hi, rng.byte_lo, rng.byte_hi) == (rng2.chunk_lo, rng2.chunk_hi, rng2.byte_lo, rng2.byte_hi)
        # the host-side header walk names the same range
        host = nnp.ChunkRange()
        assert nnp.lib().nnp_binpack_chunk_range(pack, len(pack), world, r, ctypes.byref(host)) == 0
        assert (host.chunks_total, host.chunk_lo, host.chunk_hi, host.byte_lo, host.byte_hi) == (
            rng.chunks_total, rng.chunk_lo, rng.chunk_hi, rng.byte_lo, rng.byte_hi)
        pieces.append(d_out[:got].cpu().numpy().tobytes())
        ranges.append((rng.chunk_lo, rng.chunk_hi))
    from nnue_data_compress_b200.sharding import shard_bounds

    assert ranges == [shard_bounds(rng.chunks_total, world, r) for r in range(world)]
    return b"".join(pieces)


@pytest.mark.parametrize("name,world", [("twochunks", 2), ("twochunks", 3), ("games100", 2), ("long400", 4), ("heads", 1)])
def test_shard_decompress_golden(nnp, name, world):
    assert _shard_decompress(nnp, golden(name + ".binpack"), world) == golden(name + ".rt.bin")


@pytest.mark.parametrize("n,plies,world", [(3_000_000, 100, 8), (2_500_000, 100, 3), (1_200_000, 2, 7), (700_000, 400, 2)])
def test_shard_decompress_synthetic(nnp, n, plies, world):
    from refutil import BINPACK_TO_BIN

    b = nnp.generate_bin(n, plies, 23)
    pack = nnp.bin_to_binpack(b)
    rc, want = oracle_convert(BINPACK_TO_BIN, pack)
    assert rc == 0
    assert pack.count(b"BINP") >= min(world, 2)  # several chunks: the ranks really split the file
    assert _shard_decompress(nnp, pack, world) == want
    assert nnp.binpack_to_bin(pack) == want


def test_shard_decompress_more_ranks_than_chunks(nnp):
    pack = golden("games100.binpack")  # one chunk
    out = _shard_decompress(nnp, pack, 4)
    assert out == golden("games100.rt.bin")
    assert _shard_decompress(nnp, b"", 2) == b""


def test_shard_decompress_bad_header_is_reported_by_every_rank(nnp):
    import numpy as np
    import torch

    pack = bytearray(golden("twochunks.binpack"))
    second = pack.index(b"BINP", 8)
    pack[second] = ord("X")
    d_all = torch.from_numpy(np.frombuffer(bytes(pack), dtype=np.uint8).copy()).cuda()
    for r in range(2):
        rng = nnp.ChunkRange()
        got = ctypes.c_size_t(0)
        d_out = torch.empty(len(golden("twochunks.rt.bin")), dtype=torch.uint8, device="cuda")
        rc = nnp.lib().nnp_shard_decompress_dev(ctypes.c_void_p(d_all.data_ptr()), d_all.numel(), 2, r,
                                                ctypes.c_void_p(d_out.data_ptr()), d_out.numel(), ctypes.byref(got), ctypes.byref(rng))
        assert rc == -1 and rng.chunks_total == 1  # NNP_ERR_BAD_MAGIC; the chunk in front of it is still decoded
        if r == 0:
            first = oracle_convert(1, bytes(pack[:second]))[1]
            assert d_out[: got.value].cpu().numpy().tobytes() == first
        else:
            assert got.value == 0


def test_two_devices_in_one_process_if_present(nnp):
    """Per-device contexts: a second GPU (when the box has one) converts while the first stays bound."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU")
    L = nnp.lib()
    assert L.nnp_init(1) == 0 and L.nnp_device_count() >= 2
    try:
        b = golden("games100.bin")
        assert nnp.bin_to_binpack(b) == golden("games100.binpack")  # host buffers, device 1
    finally:
        assert L.nnp_bind_device(int(os.environ.get("LOCAL_RANK", "0"))) == 0
    assert nnp.bin_to_binpack(golden("games100.bin")) == golden("games100.binpack")
