"""Device functions on the CPU (tests/host_sim): chess.cuh / stream.cuh / walk.cuh are compiled with
plain-C++ stand-ins for the CUDA intrinsics and checked against each other and against data written
by the reference. This is test infrastructure: the product library is never built this way."""
import ctypes
import os
import subprocess

import pytest

from refutil import GOLDEN_SETS, ROOT, golden, have_ref, host_sim, ref_generate

SIM_DIR = os.path.join(ROOT, "tests", "host_sim")
SIM_SO = os.path.join(SIM_DIR, "libsim.so")
CSRC = os.path.join(ROOT, "nnue_data_compress_b200", "csrc")


@pytest.fixture(scope="module")
def sim():
    return host_sim()


def _inputs():
    data = [(s, golden(s + ".bin")) for s in GOLDEN_SETS] + [(s + ".rt", golden(s + ".rt.bin")) for s in GOLDEN_SETS]
    if have_ref():
        data += [("gen100", ref_generate(120_000, 100, 42, 0)), ("gen400", ref_generate(40_000, 400, 9, 0)),
                 ("shuffled", ref_generate(30_000, 100, 5, 1)), ("restart", ref_generate(30_000, 60, 11, 2)),
                 ("heads", ref_generate(20_000, 1, 3, 0)), ("gen8", ref_generate(30_000, 8, 3, 0))]
    return data


def test_stream_splice_equals_from_scratch_encoding(sim):
    """stream_from_pos + stream_with_tail == SfenPacker::pack restatement, and the stream spliced by
    the record's own move == the encoding of the position after the move."""
    for name, b in _inputs():
        a = [ctypes.c_uint64() for _ in range(4)]
        sim.sim_stream_check(b, len(b) // 40, *[ctypes.byref(x) for x in a])
        assert a[0].value == 0 and a[1].value == 0, (name, [x.value for x in a])
        assert a[2].value == len(b) // 40, name  # every move of real data is inside the splice domain


def test_stream_splice_fuzz(sim):
    """Pseudo-random moves of all four types (mostly illegal): whatever the splice accepts must
    equal pos_do_move + from-scratch encoding."""
    b = golden("long400.bin") + golden("games100.bin")
    mm = ctypes.c_uint64()
    accepted = sim.sim_stream_fuzz(b, len(b) // 40, 256, 7, ctypes.byref(mm))
    assert accepted > 50_000 and mm.value == 0


@pytest.mark.parametrize("run", [1, 5, 16, 64, -7, -128, -1000])
def test_chain_walk_equals_record_parallel_step(sim, run):
    """walk_item (chain-walking K1) produces the codes and stems of link_code (record-parallel K1); run > 0:
    runs with parked heads (k_walk_runs), run < 0: threads owning the chains whose heads lie in their range
    of -run records (k_walk_chains)."""
    for name, b in _inputs():
        parked, err = ctypes.c_uint64(), ctypes.c_uint64()
        bad = sim.sim_walk_check(b, len(b) // 40, run, ctypes.byref(parked), ctypes.byref(err))
        assert bad == 0 and err.value == 2**64 - 1, (name, bad, err.value)


def test_chain_walk_reports_bad_sfen(sim):
    good = golden("games100.bin")[:4000]
    bits = [0] + [0] * 6 + [1, 0, 0, 0, 0, 0] + [1, 0, 0, 0, 0] * 49
    sfen = bytearray(32)
    for i, v in enumerate(bits[:256]):
        sfen[i // 8] |= v << (i & 7)
    bad_rec = bytes(sfen) + bytes(8)
    for cut, run in ((100, 16), (37, 5), (16, 16), (15, 16), (1, 3)):
        data = good[: 40 * cut] + bad_rec + good
        parked, err = ctypes.c_uint64(), ctypes.c_uint64()
        bad = sim.sim_walk_check(data, len(data) // 40, run, ctypes.byref(parked), ctypes.byref(err))
        assert bad == 0 and err.value == cut


def test_chain_decoder_reproduces_reference_bin(sim):
    """chain.cuh (bit reader window, decode_ply, doMove, spliced stream) over whole .binpack files:
    the records must be the reference's own decompression output."""
    from refutil import BIN_TO_BINPACK, BINPACK_TO_BIN, oracle_convert

    cases = [(golden(s + ".binpack"), golden(s + ".rt.bin")) for s in GOLDEN_SETS]
    if have_ref():
        for args in ((150_000, 100, 42, 0), (40_000, 400, 9, 0), (30_000, 8, 3, 0)):
            b = ref_generate(*args)
            rc, bp = oracle_convert(BIN_TO_BINPACK, b)
            assert rc == 0
            rc, rt = oracle_convert(BINPACK_TO_BIN, bp)
            assert rc == 0
            cases.append((bp, rt))
    for bp, want in cases:
        out = ctypes.create_string_buffer(len(want) + 40)
        n = sim.sim_decode_binpack(bp, len(bp), out, len(want) // 40)
        assert n == len(want) // 40
        assert out.raw[: len(want)] == want


def test_halfkp_row_updates_equal_rebuilt_rows(sim):
    """halfkp_apply_move (the in-place row update of the HalfKP chain kernels) against rows rebuilt
    from the position after every ply; on data written from legal games every ply is incremental."""
    packs = [(s, golden(s + ".binpack")) for s in GOLDEN_SETS]
    if have_ref():
        from refutil import BIN_TO_BINPACK, oracle_convert
        for name, b in (("gen100", ref_generate(120_000, 100, 42, 0)), ("gen400", ref_generate(40_000, 400, 9, 0))):
            packs.append((name, oracle_convert(BIN_TO_BINPACK, b)[1]))
    total = 0
    for name, bp in packs:
        upd, reb = ctypes.c_uint64(), ctypes.c_uint64()
        bad = sim.sim_halfkp_chains(bp, len(bp), ctypes.byref(upd), ctypes.byref(reb))
        assert bad == 0 and reb.value == 0, (name, bad, upd.value, reb.value)
        total += upd.value
    assert total > 100_000


def test_halfkp_row_update_fuzz(sim):
    b = golden("long400.bin") + golden("games100.bin")
    mm = ctypes.c_uint64()
    accepted = sim.sim_halfkp_fuzz(b, len(b) // 40, 256, 11, ctypes.byref(mm))
    assert accepted > 50_000 and mm.value == 0


def test_halfkp_rows_of_the_walker_equal_the_oracle(sim):
    """sim_halfkp_rows (rows rebuilt from the chain walker's positions; the GPU suite's reference for
    corrupted movetext) against the oracle's rows of the decoded .bin on well-formed files."""
    import numpy as np
    from refutil import BINPACK_TO_BIN, oracle_convert, oracle_halfkp
    for name in GOLDEN_SETS:
        bp = golden(name + ".binpack")
        rc, want_bin = oracle_convert(BINPACK_TO_BIN, bp)
        _, white, black, _meta, _ = oracle_halfkp(want_bin)
        n = len(want_bin) // 40
        w = np.empty((n, 32), dtype=np.int32)
        k = np.empty((n, 32), dtype=np.int32)
        assert sim.sim_halfkp_rows(bp, len(bp), w.ctypes.data, k.ctypes.data, n) == n
        assert np.array_equal(w, white) and np.array_equal(k, black), name


def test_halfkp_row_listed_from_sfen_tokens(sim):
    """k_bin_halfkp lists a record's row from the piece tokens sfen_decode hands out; it must hold the
    entries of the row rebuilt from the decoded position."""
    for name, b in _inputs():
        assert sim.sim_halfkp_tokens(b, len(b) // 40) == 0, name


def test_stem_transcode_equals_the_general_route(sim):
    """stem_to_record (single-position chains: stem nibbles -> PackedSfen tokens, no position in between) gives
    the bytes of stem_unpack + stream_from_pos + stream_with_tail on every stem of the golden files, at every
    byte alignment, and on randomly damaged stems whenever it accepts them."""
    total = 0
    for name in GOLDEN_SETS:
        b = golden(name + ".binpack")
        mm = ctypes.c_uint64()
        total += sim.sim_stem_transcode_fuzz(b, len(b), 12 if name in ("heads", "shuffled", "restart") else 8, 5, ctypes.byref(mm))
        assert mm.value == 0, name
    assert total > 100_000


def test_heads_transcode_equals_the_general_route(sim):
    """record_to_stem (chain heads: PackedSfen tokens -> stem nibbles, no position in between) gives the bytes of
    sfen_decode + stem_pack on every record of the golden files and on randomly damaged records, and reports
    exactly the streams the decoder reports as malformed."""
    total = 0
    for name in GOLDEN_SETS:
        b = golden(name + ".bin")
        mm = ctypes.c_uint64()
        total += sim.sim_heads_transcode_fuzz(b, len(b) // 40, 4, 11, ctypes.byref(mm))
        assert mm.value == 0, name
    assert total > 100_000


def test_flat_sfen_decode_equals_the_token_loop(sim):
    """sfen_decode_flat (the decoder every kernel without a piece callback uses) gives the verdict and the
    position of the token-driven loop on every golden record and on randomly damaged copies."""
    total = 0
    for name in GOLDEN_SETS:
        b = golden(name + ".bin")
        mm = ctypes.c_uint64()
        total += sim.sim_sfen_decode_fuzz(b, len(b) // 40, 6, 3, ctypes.byref(mm))
        assert mm.value == 0, name
    assert total > 150_000
