"""CPU suite: the oracle (oracle/oracle.c) against the golden vectors the compiled reference
produced (tests/golden/make_golden.py). These pin the oracle; the GPU parity tests then
compare the CUDA path with the oracle and with the same vectors."""
import json
import os

import pytest

from refutil import (BIN_TO_BINPACK, BIN_TO_PLAIN, BINPACK_TO_BIN, BINPACK_TO_PLAIN, GOLDEN, GOLDEN_PLAIN_SETS,
                     GOLDEN_SETS, PLAIN_TO_BIN, PLAIN_TO_BINPACK, golden, oracle, oracle_convert)


@pytest.mark.parametrize("name", GOLDEN_SETS)
def test_bin_to_binpack(name):
    rc, out = oracle_convert(BIN_TO_BINPACK, golden(name + ".bin"))
    assert rc == 0
    assert out == golden(name + ".binpack")


@pytest.mark.parametrize("name", GOLDEN_SETS)
def test_binpack_to_bin(name):
    rc, out = oracle_convert(BINPACK_TO_BIN, golden(name + ".binpack"))
    assert rc == 0
    assert out == golden(name + ".rt.bin")


@pytest.mark.parametrize("name", GOLDEN_PLAIN_SETS)
def test_plain_directions(name):
    bp = golden(name + ".binpack")
    rc, pl = oracle_convert(BINPACK_TO_PLAIN, bp)
    assert rc == 0 and pl == golden(name + ".plain")
    rc, bp2 = oracle_convert(PLAIN_TO_BINPACK, pl)
    assert rc == 0 and bp2 == golden(name + ".p.binpack")
    rc, bpl = oracle_convert(BIN_TO_PLAIN, golden(name + ".bin"))
    assert rc == 0 and bpl == golden(name + ".b.plain")
    rc, b2 = oracle_convert(PLAIN_TO_BIN, bpl)
    assert rc == 0 and b2 == golden(name + ".p.bin")


def test_survey_kat():
    """three hand-written records, byte values hand-decoded in the survey."""
    with open(os.path.join(GOLDEN, "manifest.json")) as f:
        manifest = json.load(f)
    expect = bytes.fromhex(
        "42494e5026000000" "ffff00000000ffff" "2d844ad200000000111111113e955be3" "0c70" "0032" "8000" "0000" "0002" "4299a040"
    )
    assert bytes.fromhex(manifest["kat"]["binpack_hex"]) == expect
    rc, bp = oracle_convert(PLAIN_TO_BINPACK, golden("kat.plain"))
    assert rc == 0 and bp == expect == golden("kat.binpack")
    rc, b = oracle_convert(BINPACK_TO_BIN, bp)
    assert rc == 0 and b == golden("kat.bin")
    assert b[:32].hex() == "08fece9aebbc31c618638c000000002184104208679454c67900000000000000"
    assert b[32:40].hex() == "19001c03000001ff"
    rc, bp_from_bin = oracle_convert(BIN_TO_BINPACK, b)
    assert rc == 0 and bp_from_bin == expect
    rc, pl = oracle_convert(BINPACK_TO_PLAIN, bp)
    assert rc == 0 and pl == golden("kat.rt.plain")
    # ep squares e3/e6 are dropped (no capturing pawn) and fullmove = (ply + 1) / 2
    fens = [l for l in pl.decode().splitlines() if l.startswith("fen ")]
    assert [f.split()[-3:] for f in fens] == [["-", "0", "0"], ["-", "0", "1"], ["-", "0", "1"]]


def test_empty_and_ragged_inputs():
    for mode in range(6):
        rc, out = oracle_convert(mode, b"")
        assert rc == 0 and out == b""
    b = golden("games100.bin")
    # a short trailing record is dropped (compress_file.cpp:1360)
    rc, a = oracle_convert(BIN_TO_BINPACK, b[: 40 * 10 + 17])
    rc2, c = oracle_convert(BIN_TO_BINPACK, b[: 40 * 10])
    assert rc == 0 and rc2 == 0 and a == c and len(a) > 8


def test_error_statuses():
    bp = golden("games100.binpack")
    rc, out = oracle_convert(BINPACK_TO_BIN, b"XINP" + bp[4:])
    assert rc == -1 and out == b""
    rc, out = oracle_convert(BINPACK_TO_BIN, bp[:4] + (200 * 1024 * 1024).to_bytes(4, "little") + bp[8:])
    assert rc == -2 and out == b""
    # sfen whose Huffman stream overruns 256 bits: 32 pieces need 13 + 30*5 + 32 + ... bits, all-ones overflows
    bad = bytes([0xFF] * 32) + bytes(8)
    good = golden("games100.bin")[:400]
    rc, out = oracle_convert(BIN_TO_BINPACK, good + bad + good)
    assert rc == -3
    rc0, ref = oracle_convert(BIN_TO_BINPACK, good)
    assert out == ref  # the writer still flushes what it gathered before the bad record


def test_files_of_single_positions_have_an_arithmetic_layout():
    """The premise of the one-kernel route for files of chain heads (k_heads_direct), pinned on the
    oracle: when every chain is 34 bytes the writer's flush rule (compress_file.cpp:1076-1080) closes a chunk after
    exactly ceil(2^20 / 34) = 30 841 chains, so record r's stem lies at
    (r / 30841) * (8 + 34 * 30841) + 8 + 34 * (r mod 30841) of the file."""
    from refutil import golden

    heads = golden("heads.bin")
    n_unit = len(heads) // 40
    reps = 70_000 // n_unit + 1
    data = heads * reps  # (a record never continues the copy of the file's last record in front of it: ply 0 follows)
    n = len(data) // 40
    rc, pack = oracle_convert(BIN_TO_BINPACK, data)
    assert rc == 0
    per_chunk = ((1 << 20) + 33) // 34
    chunks = (n + per_chunk - 1) // per_chunk
    if len(pack) != 34 * n + 8 * chunks:
        pytest.skip("the golden heads file repeats into a continuation at its seam")
    pos = 0
    for c in range(chunks):
        assert pack[pos:pos + 4] == b"BINP"
        size = int.from_bytes(pack[pos + 4:pos + 8], "little")
        assert size == 34 * min(per_chunk, n - c * per_chunk)
        pos += 8 + size
    assert pos == len(pack)
    rc, one = oracle_convert(BIN_TO_BINPACK, data[40 * 40_000:40 * 40_001])
    r = 40_000
    off = (r // per_chunk) * (8 + 34 * per_chunk) + 8 + 34 * (r % per_chunk)
    assert pack[off:off + 34] == one[8:42]
