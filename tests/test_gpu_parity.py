"""GPU suite (-m gpu): the CUDA path, called through the C ABI (libnnuepack.so), against
(1) the golden vectors written by the compiled reference, (2) the CPU oracle on seeded
synthetic inputs, (3) the compiled reference itself when oracle/_ref travelled to the box.
Everything is integer/byte work: equality is exact."""
import os

import pytest

from refutil import (BIN_TO_BINPACK, BINPACK_TO_BIN, GOLDEN_SETS, golden, have_ref, oracle_convert, ref_convert,
                     ref_generate)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", GOLDEN_SETS)
def test_golden_bin_to_binpack(nnp, name):
    assert nnp.bin_to_binpack(golden(name + ".bin")) == golden(name + ".binpack")


@pytest.mark.parametrize("name", GOLDEN_SETS)
def test_golden_binpack_to_bin(nnp, name):
    assert nnp.binpack_to_bin(golden(name + ".binpack")) == golden(name + ".rt.bin")


def test_kat(nnp):
    expect = bytes.fromhex(
        "42494e5026000000ffff00000000ffff2d844ad200000000111111113e955be30c7000328000000000024299a040"
    )
    assert nnp.bin_to_binpack(golden("kat.bin")) == expect
    assert nnp.binpack_to_bin(expect) == golden("kat.bin")


def test_empty_and_ragged(nnp):
    assert nnp.bin_to_binpack(b"") == b""
    assert nnp.binpack_to_bin(b"") == b""
    b = golden("games100.bin")
    assert nnp.bin_to_binpack(b[:39]) == b""
    assert nnp.bin_to_binpack(b[: 40 * 10 + 17]) == nnp.bin_to_binpack(b[: 40 * 10])
    for n in (1, 2, 3, 254, 255, 256, 257, 511, 2047, 2048, 2049):
        rc, want = oracle_convert(BIN_TO_BINPACK, b[: 40 * n])
        assert rc == 0
        got = nnp.bin_to_binpack(b[: 40 * n])
        assert got == want, n
        assert nnp.binpack_to_bin(got) == oracle_convert(BINPACK_TO_BIN, want)[1], n


def _synthetic(n, plies, seed, mode=0):
    if not have_ref():
        pytest.skip("oracle/_ref (compiled reference + generator) did not travel to this box")
    return ref_generate(n, plies, seed, mode)


@pytest.mark.parametrize("n,plies,seed,mode", [
    (200_000, 100, 42, 0),
    (100_000, 1, 7, 0),
    (100_000, 8, 7, 0),
    (100_000, 64, 7, 0),
    (150_000, 400, 9, 0),
    (60_000, 100, 5, 1),
    (60_000, 60, 11, 2),
])
def test_oracle_parity_synthetic(nnp, n, plies, seed, mode):
    b = _synthetic(n, plies, seed, mode)
    rc, want = oracle_convert(BIN_TO_BINPACK, b)
    assert rc == 0
    got = nnp.bin_to_binpack(b)
    assert got == want
    rc, want_bin = oracle_convert(BINPACK_TO_BIN, want)
    assert rc == 0
    assert nnp.binpack_to_bin(want) == want_bin


@pytest.mark.parametrize("n,plies", [(1_500_000, 1), (2_500_000, 3)])
def test_many_chunks(nnp, n, plies):
    """Dozens of chunks: the chunk-flush orbit (compress) and the chunk table (decompress)."""
    b = nnp.generate_bin(n, plies, 8)
    rc, want = oracle_convert(BIN_TO_BINPACK, b)
    assert rc == 0 and want.count(b"BINP") >= 30
    got = nnp.bin_to_binpack(b)
    assert got == want
    assert nnp.binpack_to_bin(got) == oracle_convert(BINPACK_TO_BIN, want)[1]


HEADS_PER_CHUNK = ((1 << 20) + 33) // 34  # a chunk of single positions closes after this many chains


@pytest.mark.parametrize("n", [4096, HEADS_PER_CHUNK - 1, HEADS_PER_CHUNK, HEADS_PER_CHUNK + 1, 2 * HEADS_PER_CHUNK,
                               2 * HEADS_PER_CHUNK + 127, 400_003])
def test_files_of_single_positions_take_one_kernel(nnp, n):
    """A .bin in which every record starts a chain is converted by one kernel (k_heads_direct: the stem is the
    record's token stream in another order, and every byte's place in the file follows from its record index);
    the bytes are the oracle's, whole chunks and a ragged last one alike."""
    b = nnp.generate_bin(n, 1, 77)
    rc, want = oracle_convert(BIN_TO_BINPACK, b)
    assert rc == 0
    assert nnp.bin_to_binpack(b) == want
    assert nnp.lib().nnp_last_dominant_kernel() == b"k_heads_direct"
    nnp.lib().nnp_debug_config(b"k1_direct", 2)
    try:
        assert nnp.bin_to_binpack(b) == want
        assert nnp.lib().nnp_last_dominant_kernel() != b"k_heads_direct"
    finally:
        nnp.lib().nnp_debug_config(b"k1_direct", 0)
    # and back: chunks of nothing but 34-byte chains are read without looking for chain starts
    rc, back = oracle_convert(BINPACK_TO_BIN, want)
    assert rc == 0
    assert nnp.binpack_to_bin(want) == back
    assert nnp.lib().nnp_last_dominant_kernel() == b"k_emit_heads_only"
    nnp.lib().nnp_debug_config(b"dec_direct", 2)
    try:
        assert nnp.binpack_to_bin(want) == back
        assert nnp.lib().nnp_last_dominant_kernel() != b"k_emit_heads_only"
    finally:
        nnp.lib().nnp_debug_config(b"dec_direct", 0)


def test_single_position_chunks_among_game_chunks(nnp):
    """Chunks of nothing but single positions between ordinary chunks (shuffled data in which a few records
    continue their predecessor; files concatenated from both kinds): the decoder lists such a chunk as one entry
    and writes its records without looking for chain starts, the rest takes the verified walk; the records and
    the strategy's outcome are the oracle's. With a damaged numPlies in one of them the file still decodes."""
    import numpy as np

    games = np.frombuffer(nnp.generate_bin(300_000, 100, 9), dtype=np.uint8).reshape(-1, 40)
    shuffled = games[np.random.default_rng(4).permutation(len(games))].tobytes()
    singles = nnp.generate_bin(100_000, 1, 5)
    for data in (shuffled, singles + games.tobytes()[:40 * 120_000] + singles, games.tobytes()[:40 * 50_000] + singles):
        rc, pack = oracle_convert(BIN_TO_BINPACK, data)
        assert rc == 0
        assert nnp.bin_to_binpack(data) == pack
        rc, want = oracle_convert(BINPACK_TO_BIN, pack)
        assert rc == 0
        before = nnp.decode_stats()["optimistic_hits"]
        assert nnp.binpack_to_bin(pack) == want
        if data is shuffled:  # (every chunk may be of the single-position kind: then nothing is verified)
            assert nnp.lib().nnp_last_dominant_kernel() in (b"k_emit_chains_verify", b"k_emit_heads_only")
        else:
            assert nnp.decode_stats()["optimistic_hits"] == before + 1
            assert nnp.lib().nnp_last_dominant_kernel() == b"k_emit_chains_verify"
        bad = bytearray(pack)
        bad[8 + 34 * 1000 + 33] ^= 1  # numPlies of a chain in the first chunk
        rc, want = oracle_convert(BINPACK_TO_BIN, bytes(bad))
        if rc == 0:
            try:
                assert nnp.binpack_to_bin(bytes(bad)) == want
            except nnp.NnpError as e:
                assert e.status == -4, e.status


def test_single_position_chunks_with_damaged_stems(nnp):
    """The candidate-free reader takes every 34 bytes as a stem whatever they hold, as the reference does: stems
    with flipped bits (irregular nibbles, stray special codes) decode to the oracle's records; a flipped
    numPlies byte sends the file back through the general reader."""
    import random

    rng = random.Random(5)
    pack = bytearray(nnp.bin_to_binpack(nnp.generate_bin(70_000, 1, 12)))
    starts = [8]  # payload starts of the chunks
    while starts[-1] - 8 + 8 + int.from_bytes(pack[starts[-1] - 4:starts[-1]], "little") < len(pack):
        starts.append(starts[-1] + int.from_bytes(pack[starts[-1] - 4:starts[-1]], "little") + 8)
    for trial in range(6):
        b = bytearray(pack)
        for _ in range(400):
            c = rng.choice(starts)
            size = int.from_bytes(b[c - 4:c], "little")
            k = rng.randrange(size // 34)
            byte = rng.randrange(34 if trial == 5 else 32)
            b[c + 34 * k + byte] ^= 1 << rng.randrange(8)
        rc, want = oracle_convert(BINPACK_TO_BIN, bytes(b))
        if rc != 0:
            continue
        try:
            got = nnp.binpack_to_bin(bytes(b))
        except nnp.NnpError as e:
            assert e.status == -4, e.status
            continue
        assert got == want
        if trial < 5:
            assert nnp.lib().nnp_last_dominant_kernel() == b"k_emit_heads_only"


def test_one_kernel_route_gives_way(nnp):
    """The one-kernel route checks its premise while it runs: shuffled game positions (chain heads, but some with
    ply / result fields that happen to link), real chains, and a malformed record all end in the general
    pipeline's bytes and status when the route is forced on them."""
    import numpy as np

    L = nnp.lib()
    games = np.frombuffer(nnp.generate_bin(300_000, 100, 9), dtype=np.uint8).reshape(-1, 40)
    shuffled = games[np.random.default_rng(4).permutation(len(games))].tobytes()
    bad = bytearray(nnp.generate_bin(50_000, 1, 3))
    bad[40 * 31_000 + 2:40 * 31_000 + 32] = b"\xff" * 30  # every square a 5-bit token of type 7
    cases = [shuffled, games.tobytes(), golden("heads.bin"), golden("shuffled.bin"), golden("restart.bin"), bytes(bad)]
    for data in cases:
        rc, want = oracle_convert(BIN_TO_BINPACK, data)
        L.nnp_debug_config(b"k1_direct", 1)
        try:
            if rc == 0:
                assert nnp.bin_to_binpack(data) == want
            else:
                with pytest.raises(nnp.NnpError) as ei:
                    nnp.bin_to_binpack(data)
                assert ei.value.status == rc and ei.value.partial == want
        finally:
            L.nnp_debug_config(b"k1_direct", 0)
    # left to itself the shuffled file is sampled, found to be nearly all heads and transcoded by K1
    assert nnp.bin_to_binpack(shuffled) == oracle_convert(BIN_TO_BINPACK, shuffled)[1]
    assert L.nnp_last_dominant_kernel() in (b"k_heads_transcode", b"k_heads_direct")


@pytest.mark.parametrize("force", ["k1_walk", "k1_per_record", "k1_heads"])
@pytest.mark.parametrize("n,plies", [(400_000, 100), (400_000, 1), (300_000, 5)])
def test_both_forms_of_k1(nnp, force, n, plies):
    """The compressor's first kernel exists in a chain-walking and a record-parallel form, and as a transcoder
    for files of chain heads (a sample of the chain-head density picks one); each must give the oracle's
    bytes on long and short chains."""
    b = nnp.generate_bin(n, plies, 31)
    rc, want = oracle_convert(BIN_TO_BINPACK, b)
    assert rc == 0
    nnp.lib().nnp_debug_config(force.encode(), 1)
    try:
        assert nnp.bin_to_binpack(b) == want
    finally:
        nnp.lib().nnp_debug_config(force.encode(), 0)
    assert nnp.bin_to_binpack(b) == want


def test_reference_parity_1m(nnp):
    """BASELINE config 1: 1M positions, ~100 plies per chain, against the reference binary."""
    b = _synthetic(1_000_000, 100, 42)
    want = ref_convert(BIN_TO_BINPACK, b)
    got = nnp.bin_to_binpack(b)
    assert len(got) == len(want) == 2182401
    assert got == want
    assert nnp.binpack_to_bin(want) == ref_convert(BINPACK_TO_BIN, want)


def test_bad_sfen_stops_like_the_reference(nnp):
    good = golden("games100.bin")[:4000]
    # 62 white-pawn tokens need 13 + 310 bits: the cursor passes 256 ("Improperly encoded bin sfen")
    bits = [0] + [0] * 6 + [1, 0, 0, 0, 0, 0] + [1, 0, 0, 0, 0] * 49
    sfen = bytearray(32)
    for i, v in enumerate(bits[:256]):
        sfen[i // 8] |= v << (i & 7)
    bad = bytes(sfen) + bytes(8)
    rc, want = oracle_convert(BIN_TO_BINPACK, good + bad + good)
    assert rc == -3
    with pytest.raises(nnp.NnpError) as ei:
        nnp.bin_to_binpack(good + bad + good)
    assert ei.value.status == -3
    assert ei.value.partial == want == oracle_convert(BIN_TO_BINPACK, good)[1]


def test_bad_chunk_headers(nnp):
    bp = golden("twochunks.binpack")
    second = 8 + int.from_bytes(bp[4:8], "little")
    for mutate, status in (
        (lambda d: b"XINP" + d[4:], -1),
        (lambda d: d[:second] + b"BINX" + d[second + 4:], -1),
        (lambda d: d[:second + 4] + (101 * 1024 * 1024).to_bytes(4, "little") + d[second + 8:], -2),
    ):
        data = mutate(bp)
        rc, want = oracle_convert(BINPACK_TO_BIN, data)
        assert rc == status
        with pytest.raises(nnp.NnpError) as ei:
            nnp.binpack_to_bin(data)
        assert ei.value.status == status
        assert ei.value.partial == want


def test_roundtrip_property_large(nnp):
    """decode(encode(x)) keeps the record count, and re-encoding it is again byte-identical to the
    oracle (encode o decode is NOT the identity on bytes: quirk Q1 of SURVEY.md 8a drops some
    ep squares, so the comparison is against the oracle, not against x)."""
    b = _synthetic(400_000, 100, 1234)
    y = nnp.bin_to_binpack(b)
    x2 = nnp.binpack_to_bin(y)
    assert len(x2) == len(b)
    rc, y2 = oracle_convert(BIN_TO_BINPACK, x2)
    assert rc == 0
    assert nnp.bin_to_binpack(x2) == y2


@pytest.mark.parametrize("n,plies,seed", [(300_000, 100, 3), (50_000, 1, 4), (200_000, 400, 5), (100_000, 13, 6)])
def test_device_generator_parity(nnp, n, plies, seed):
    """Inputs from the device-side generator (what bench.py uses): deterministic, accepted by the
    oracle, and converted identically by the CUDA path."""
    b = nnp.generate_bin(n, plies, seed)
    assert len(b) == n * 40
    assert b == nnp.generate_bin(n, plies, seed)
    assert b[: 40 * 1000] == nnp.generate_bin(1000, plies, seed)
    rc, want = oracle_convert(BIN_TO_BINPACK, b)
    assert rc == 0
    got = nnp.bin_to_binpack(b)
    assert got == want
    rc, want_bin = oracle_convert(BINPACK_TO_BIN, want)
    assert rc == 0 and len(want_bin) == len(b)
    assert nnp.binpack_to_bin(got) == want_bin
    # games are played by the rules: except for the rare ep quirk the round trip
    # reproduces the generated records, in particular every move is decoded back
    diff = sum(1 for i in range(0, len(b), 40) if b[i:i + 40] != want_bin[i:i + 40])
    assert diff <= max(5, n // 100_000 * 5), diff
    if plies >= 100:
        assert len(got) < len(b) // 10  # chains really link


def test_device_generator_vs_reference_binary(nnp):
    if not have_ref():
        pytest.skip("oracle/_ref did not travel to this box")
    b = nnp.generate_bin(250_000, 100, 77)
    want = ref_convert(BIN_TO_BINPACK, b)
    assert nnp.bin_to_binpack(b) == want
    assert nnp.binpack_to_bin(want) == ref_convert(BINPACK_TO_BIN, want)


@pytest.mark.parametrize("env", [{"NNP_DEBUG_REJECT_MOD": "5"}, {"NNP_DEBUG_EXHAUSTIVE": "1"}, {"NNP_DEBUG_SINGLES": "0"}])
def test_decode_fallbacks_are_exact(nnp, env):
    """binpack -> bin has three strategies: the optimistic single walk (default), the exhaustive
    probe/resolve walk (NNP_DEBUG_EXHAUSTIVE=1, or whenever the optimistic walk sees a violation)
    and the sequential per-chunk decoder. NNP_DEBUG_REJECT_MOD drops a pseudo-random subset of
    chain-start candidates, which fails the optimistic walk and forces the affected chunks through
    the sequential decoder. The output must not change."""
    import subprocess
    import sys

    code = (
        "import sys; sys.path.insert(0, 'tests'); sys.path.insert(0, '.');"
        "import nnue_data_compress_b200 as n; from refutil import golden, GOLDEN_SETS;"
        "n.init(0);"
        "ok = all(n.binpack_to_bin(golden(s + '.binpack')) == golden(s + '.rt.bin') for s in GOLDEN_SETS);"
        "b = n.generate_bin(300000, 100, 99); p = n.bin_to_binpack(b); r = n.binpack_to_bin(p);"
        "ok = ok and len(r) == len(b) and n.bin_to_binpack(r) == p;"
        "print('FALLBACK_OK' if ok else 'FALLBACK_MISMATCH')"
    )
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], cwd=root, env=dict(os.environ, **env), capture_output=True, text=True,
                         timeout=600)
    assert "FALLBACK_OK" in out.stdout, out.stdout + out.stderr


def test_corrupted_binpack_follows_the_oracle(nnp):
    """Random bit flips in valid .binpack files: whenever both the oracle (= the reference's reading of
    malformed movetext) and the CUDA path decode the file, the records are identical; the CUDA path may
    instead refuse movetext that runs off its chunk (NNP_ERR_TRUNCATED), never anything else."""
    import random

    rng = random.Random(20260)
    packs = [golden(n + ".binpack") for n in ("games100", "long400", "restart", "heads", "shuffled")]
    compared = 0
    for _ in range(250):
        b = bytearray(rng.choice(packs))
        for _ in range(rng.randrange(1, 5)):
            b[rng.randrange(8, len(b))] ^= 1 << rng.randrange(8)
        rc, want = oracle_convert(BINPACK_TO_BIN, bytes(b))
        if rc != 0:
            continue
        try:
            got = nnp.binpack_to_bin(bytes(b))
        except nnp.NnpError as e:
            assert e.status == -4, e.status
            continue
        assert got == want
        compared += 1
    assert compared > 100


# ---------------------------------------------------------------------------------------------
# .plain path (compressPlain / decompressPlain / convertBinToPlain / convertPlainToBin)

from refutil import BIN_TO_PLAIN, BINPACK_TO_PLAIN, GOLDEN_PLAIN_SETS, PLAIN_TO_BIN, PLAIN_TO_BINPACK  # noqa: E402


@pytest.mark.parametrize("name", GOLDEN_PLAIN_SETS)
def test_golden_plain_directions(nnp, name):
    assert nnp.binpack_to_plain(golden(name + ".binpack")) == golden(name + ".plain")
    assert nnp.plain_to_binpack(golden(name + ".plain")) == golden(name + ".p.binpack")
    assert nnp.bin_to_plain(golden(name + ".bin")) == golden(name + ".b.plain")
    assert nnp.plain_to_bin(golden(name + ".b.plain")) == golden(name + ".p.bin")


def test_kat_plain(nnp):
    expect = bytes.fromhex(
        "42494e5026000000ffff00000000ffff2d844ad200000000111111113e955be30c7000328000000000024299a040"
    )
    assert nnp.plain_to_binpack(golden("kat.plain")) == expect
    assert nnp.binpack_to_plain(expect) == golden("kat.rt.plain")
    assert nnp.plain_to_bin(golden("kat.plain")) == oracle_convert(PLAIN_TO_BIN, golden("kat.plain"))[1]


@pytest.mark.parametrize("n,plies,seed", [(120_000, 100, 21), (40_000, 1, 22), (60_000, 400, 23)])
def test_oracle_parity_plain(nnp, n, plies, seed):
    b = nnp.generate_bin(n, plies, seed)
    rc, bp = oracle_convert(BIN_TO_BINPACK, b)
    assert rc == 0
    rc, pl = oracle_convert(BINPACK_TO_PLAIN, bp)
    assert rc == 0
    assert nnp.binpack_to_plain(bp) == pl
    rc, bp2 = oracle_convert(PLAIN_TO_BINPACK, pl)
    assert rc == 0
    assert nnp.plain_to_binpack(pl) == bp2
    rc, bpl = oracle_convert(BIN_TO_PLAIN, b)
    assert rc == 0
    assert nnp.bin_to_plain(b) == bpl
    rc, b2 = oracle_convert(PLAIN_TO_BIN, bpl)
    assert rc == 0
    assert nnp.plain_to_bin(bpl) == b2


def test_plain_layout_variants(nnp):
    """Key order, unknown keys, inherited fields, blank lines, CR before LF on the terminator, no
    trailing newline: everything the reference tokeniser reads line by line."""
    base = golden("kat.plain").decode()
    recs = base.strip().split("e\n")
    recs = [r for r in recs if r.strip()]
    shuffled = ""
    for r in recs:
        lines = r.strip().split("\n")
        shuffled += "\n".join(reversed(lines)) + "\ncomment ignored by the parser\n\n  e  \n"
    inherited = recs[0] + "e\n" + "move e2e3\ne\n" + "fen " + recs[1].split("fen ")[1] + "e"
    for text in (shuffled, inherited):
        data = text.encode()
        for mode, fn in ((PLAIN_TO_BINPACK, nnp.plain_to_binpack), (PLAIN_TO_BIN, nnp.plain_to_bin)):
            rc, want = oracle_convert(mode, data)
            assert rc == 0
            assert fn(data) == want
    assert nnp.plain_to_bin(b"") == b"" and nnp.plain_to_binpack(b"\n\n") == b""


def test_plain_inherited_fields_in_linear_time(nnp):
    """A key that (almost) never appears: every record inherits it from far back (fields persist,
    compress_file.cpp:1254). The parser resolves that with a running maximum over the records, not by
    walking back through the file; 300 000 records with no `result` line at all and `ply` only in the
    first record take milliseconds and match the oracle."""
    import time

    b = nnp.generate_bin(300_000, 100, 77)
    rc, text = oracle_convert(BIN_TO_PLAIN, b)
    assert rc == 0
    lines = text.split(b"\n")
    seen_ply = False
    kept = []
    for ln in lines:
        if ln.startswith(b"result "):
            continue
        if ln.startswith(b"ply "):
            if seen_ply:
                continue
            seen_ply = True
        kept.append(ln)
    data = b"\n".join(kept)
    assert len(data) < len(text) - 300_000 * 8
    for mode, fn in ((PLAIN_TO_BIN, nnp.plain_to_bin), (PLAIN_TO_BINPACK, nnp.plain_to_binpack)):
        rc, want = oracle_convert(mode, data)
        assert rc == 0
        t0 = time.time()
        got = fn(data)
        assert time.time() - t0 < 20
        assert got == want
    # every second record drops its fen line: the position persists, the move is read in it
    kept, k = [], 0
    for ln in lines:
        if ln.startswith(b"fen "):
            k += 1
            if k % 2 == 0:
                continue
        kept.append(ln)
    data = b"\n".join(kept)
    rc, want = oracle_convert(PLAIN_TO_BIN, data)
    if rc == 0:  # (the oracle refuses what the reference would crash on)
        assert nnp.plain_to_bin(data) == want


def test_plain_rejected_layouts(nnp):
    for text in (b"fen\nrnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1\nmove e2e4\nscore 1\nply 0\nresult 0\ne\n",
                 b"fen rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1\nmove e2e4\nscore x\nply 0\nresult 0\ne\n",
                 b"fen rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1\nmove e2e4\nscore 1\nply 0\nresult 0\ne fen x\n"):
        with pytest.raises(nnp.NnpError) as ei:
            nnp.plain_to_binpack(text)
        assert ei.value.status == -7


def test_plain_error_partial_outputs(nnp):
    """bad sfen in .bin -> .plain and bad chunk in .binpack -> .plain: only what the reference had
    flushed (1 MiB buffer rule) survives."""
    b = nnp.generate_bin(30_000, 100, 31)
    bits = [0] + [0] * 6 + [1, 0, 0, 0, 0, 0] + [1, 0, 0, 0, 0] * 49
    sfen = bytearray(32)
    for i, v in enumerate(bits[:256]):
        sfen[i // 8] |= v << (i & 7)
    bad = bytes(sfen) + bytes(8)
    data = b[: 40 * 25_000] + bad + b[40 * 25_000:]
    rc, want = oracle_convert(BIN_TO_PLAIN, data)
    assert rc == -3 and len(want) > 1 << 20
    with pytest.raises(nnp.NnpError) as ei:
        nnp.bin_to_plain(data)
    assert ei.value.status == -3 and ei.value.partial == want
    bp = golden("twochunks.binpack")
    second = 8 + int.from_bytes(bp[4:8], "little")
    data = bp[:second] + b"BINX" + bp[second + 4:]
    rc, want = oracle_convert(BINPACK_TO_PLAIN, data)
    assert rc == -1 and len(want) > 1 << 20
    with pytest.raises(nnp.NnpError) as ei:
        nnp.binpack_to_plain(data)
    assert ei.value.status == -1 and ei.value.partial == want


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_fuzz_differential_against_the_oracle(nnp, seed):
    """Random bit flips in .bin and .binpack inputs: wherever the oracle (= the reference's semantics)
    gives a result -- OK or one of the three reference errors -- status and bytes agree. This includes
    stored moves that are not pseudo-legal, whose ids the reference does not mask (addBitsLE8
    :840-862): the payload writer adds the same stray bits (k_write_payload<true>)."""
    import random

    rng = random.Random(seed)

    def ours(fn, data):
        try:
            return 0, fn(data)
        except nnp.NnpError as e:
            return e.status, (e.partial or b"")

    packs = [golden(n + ".binpack") for n in ("games100", "long400", "restart", "heads")]
    bins = [golden(n + ".bin") for n in ("games100", "long400", "restart")]
    agree = skipped = 0
    for it in range(300):
        if it & 1:
            b = bytearray(rng.choice(packs))
            for _ in range(rng.randrange(1, 4)):
                b[rng.randrange(8, len(b))] ^= 1 << rng.randrange(8)
            mode, fn = BINPACK_TO_BIN, nnp.binpack_to_bin
        else:
            b = bytearray(rng.choice(bins))
            for _ in range(rng.randrange(1, 4)):
                b[rng.randrange(len(b))] ^= 1 << rng.randrange(8)
            mode, fn = BIN_TO_BINPACK, nnp.bin_to_binpack
        rc_o, out_o = oracle_convert(mode, bytes(b))
        if rc_o not in (0, -1, -2, -3):
            skipped += 1
            continue
        rc, out = ours(fn, bytes(b))
        assert (rc, out) == (rc_o, out_o), (seed, it, mode, rc_o, rc, len(out_o), len(out))
        agree += 1
    assert agree > 150


def test_illegal_stored_moves_bleed_like_the_reference(nnp):
    """Every stored move replaced by a random one (mostly not pseudo-legal): ids overflow their fields in
    many plies; bytes must still be the oracle's, through both forms of K1 and through the .plain path."""
    import random

    rng = random.Random(11)
    b = bytearray(golden("games100.bin"))
    for r in range(len(b) // 40):
        if rng.random() < 0.5:
            mv = rng.randrange(1 << 16)
            b[40 * r + 34] = mv & 0xFF
            b[40 * r + 35] = mv >> 8
    data = bytes(b)
    rc, want = oracle_convert(BIN_TO_BINPACK, data)
    assert rc == 0
    L = nnp.lib()
    for key in (b"k1_walk", b"k1_per_record", b"k1_heads"):
        L.nnp_debug_config(key, 1)
        try:
            assert nnp.bin_to_binpack(data) == want
        finally:
            L.nnp_debug_config(key, 0)


def test_segmented_header_walk(nnp):
    """The chunk headers of large files are walked in parallel segments (k_seg_find / k_seg_walk) and only
    believed when every segment's walk ends on the next segment's start; tiny segments force that path on
    small files, and a broken header or a "BINP" inside a payload must leave the file to the sequential walk."""
    import numpy as np
    import torch

    L = nnp.lib()
    b = nnp.generate_bin(1_500_000, 2, 41)  # ~25 chunks
    pack = nnp.bin_to_binpack(b)
    rc, want = oracle_convert(BINPACK_TO_BIN, pack)
    assert rc == 0 and pack.count(b"BINP") >= 20
    L.nnp_debug_config(b"walk_seg_bytes", 1 << 20)
    try:
        assert nnp.binpack_to_bin(pack) == want
        d = torch.from_numpy(np.frombuffer(pack, dtype=np.uint8).copy()).cuda()
        for world in (1, 3, 8):
            pieces = []
            for r in range(world):
                out = torch.empty(len(want) + 64, dtype=torch.uint8, device="cuda")
                got, rng = nnp.shard_decompress(d, world, r, out)
                pieces.append(out[:got].cpu().numpy().tobytes())
            assert b"".join(pieces) == want
        # a false header inside a payload, behind a segment boundary: the walk in front of it steps over it
        fake = bytearray(pack)
        at = (len(pack) // 2 // 16) * 16 + 64
        fake[at:at + 8] = b"BINP" + (40).to_bytes(4, "little")
        rc, want2 = oracle_convert(BINPACK_TO_BIN, bytes(fake))
        if rc == 0:
            assert nnp.binpack_to_bin(bytes(fake)) == want2
        # a broken header: the reference's error and partial output
        bad = bytearray(pack)
        third = [i for i in range(len(pack) - 4) if pack[i:i + 4] == b"BINP"][12]
        bad[third] = ord("X")
        rc, want3 = oracle_convert(BINPACK_TO_BIN, bytes(bad))
        assert rc == -1
        with pytest.raises(nnp.NnpError) as ei:
            nnp.binpack_to_bin(bytes(bad))
        assert ei.value.status == -1 and ei.value.partial == want3
    finally:
        L.nnp_debug_config(b"walk_seg_bytes", 0)
