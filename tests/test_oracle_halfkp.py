"""CPU suite: the oracle's HalfKP rows (orc_bin_to_halfkp) against an independent pure-Python statement
of the published index formula applied to the FEN of the same record, and against hand-computed
values for the start position. The reference has no feature code (parity unpinned for the formula;
the decoded positions it is applied to are pinned elsewhere)."""
import ctypes
import struct

import numpy as np

from refutil import golden, halfkp_from_fen, oracle, oracle_halfkp


def _fen(rec):
    buf = ctypes.create_string_buffer(128)
    assert oracle().orc_sfen_to_fen(bytes(rec[:32]), buf, 128) == 0
    return buf.value.decode()


def test_start_position_by_hand():
    rc, white, black, meta, _ = oracle_halfkp(golden("kat.bin")[:40])
    assert rc == 0
    # white: king e1 = 4; the row starts with the white pawns a2..h2 (kind 0), then the black pawns (kind 1)
    assert white[0][0] == 1 + 8 + 641 * 4
    assert white[0][8] == 1 + 48 + 64 * 1 + 641 * 4
    # black: king e8 = 60 -> 60 ^ 63 = 3; the white pawn a2 seen from black: sq 8 ^ 63 = 55, kind 1
    assert black[0][0] == 1 + 55 + 64 * 1 + 641 * 3
    assert sorted(white[0][:30]) == list(white[0][:30])
    assert list(meta[0]) == [25, 0, 0, 0, 1, 0, 30, 0]
    assert (white[0][30:] == -1).all() and (black[0][30:] == -1).all()
    assert white.max() < 41024 and black.max() < 41024


def test_rows_match_the_formula_applied_to_the_fen():
    b = golden("games100.bin") + golden("long400.bin")[: 40 * 3000]
    n = len(b) // 40
    rc, white, black, meta, _ = oracle_halfkp(b)
    assert rc == 0
    for r in range(0, n, 7):
        rec = b[r * 40: r * 40 + 40]
        w, k = halfkp_from_fen(_fen(rec))
        assert list(white[r][: len(w)]) == w and (white[r][len(w):] == -1).all(), r
        assert list(black[r][: len(k)]) == k and (black[r][len(k):] == -1).all(), r
        score, move, ply, result = struct.unpack_from("<hHHb", rec, 32)
        m = bytes(meta[r])
        assert struct.unpack("<hHbBBB", m) == (score, ply, result, _fen(rec).split()[1] == "b", len(w), 0), r


def test_malformed_record_is_reported():
    b = bytearray(golden("games100.bin")[: 40 * 5])
    b[40 * 3: 40 * 3 + 32] = b"\xff" * 32
    rc, *_rest, bad = oracle_halfkp(bytes(b))
    assert rc == -3 and bad == 3
