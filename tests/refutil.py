"""Shared helpers for the test-suite: ctypes access to the CPU oracle (oracle/liboracle.so)
and, when present, the compiled reference (oracle/_ref). TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "nnue_data_compression")
REF_GEN = os.path.join(ORACLE_DIR, "_ref", "gen_ref")
GOLDEN = os.path.join(ROOT, "tests", "golden")

BIN_TO_BINPACK, BINPACK_TO_BIN, PLAIN_TO_BINPACK, BINPACK_TO_PLAIN, BIN_TO_PLAIN, PLAIN_TO_BIN = range(6)
MODE_EXT = {
    BIN_TO_BINPACK: (".bin", ".binpack"),
    BINPACK_TO_BIN: (".binpack", ".bin"),
    PLAIN_TO_BINPACK: (".plain", ".binpack"),
    BINPACK_TO_PLAIN: (".binpack", ".plain"),
    BIN_TO_PLAIN: (".bin", ".plain"),
    PLAIN_TO_BIN: (".plain", ".bin"),
}

_oracle = None


def build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "oracle"], check=True)


def oracle():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(
            os.path.join(ORACLE_DIR, "oracle.c")
        ):
            build_oracle()
        lib = ctypes.CDLL(ORACLE_SO)
        lib.orc_convert.restype = ctypes.c_int
        lib.orc_convert.argtypes = [
            ctypes.c_int,
            ctypes.c_char_p,
            ctypes.c_size_t,
            ctypes.POINTER(ctypes.c_void_p),
            ctypes.POINTER(ctypes.c_size_t),
        ]
        lib.orc_free.argtypes = [ctypes.c_void_p]
        lib.orc_sfen_to_fen.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
        lib.orc_fen_to_sfen.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
        lib.orc_is_continuation.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
        lib.orc_binpack_count.restype = ctypes.c_int64
        lib.orc_binpack_count.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
        lib.orc_bin_to_halfkp.restype = ctypes.c_int
        lib.orc_bin_to_halfkp.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.POINTER(ctypes.c_size_t)]
        _oracle = lib
    return _oracle


def oracle_halfkp(bin_data):
    """HalfKP rows of .bin records by the oracle: (rc, white [n,32] int32, black [n,32] int32, meta [n,8] uint8,
    index of the first malformed record)."""
    import numpy as np

    n = len(bin_data) // 40
    white = np.empty((n, 32), dtype=np.int32)
    black = np.empty((n, 32), dtype=np.int32)
    meta = np.empty((n, 8), dtype=np.uint8)
    bad = ctypes.c_size_t(n)
    rc = oracle().orc_bin_to_halfkp(bytes(bin_data), n, white.ctypes.data, black.ctypes.data, meta.ctypes.data,
                                    ctypes.byref(bad))
    return rc, white, black, meta, bad.value


def halfkp_from_fen(fen):
    """The published HalfKP index (nnue-pytorch halfkp_idx) from a FEN string, written independently
    of the oracle's C code: returns (white row, black row) as lists ordered by (kind as white sees it, square)."""
    board = fen.split()[0]
    pieces = {}
    for r, row in enumerate(board.split("/")):
        f = 0
        for ch in row:
            if ch.isdigit():
                f += int(ch)
            else:
                pieces[(7 - r) * 8 + f] = ch
                f += 1
    ksq = {True: next(s for s, c in sorted(pieces.items()) if c == "K"),
           False: next(s for s, c in sorted(pieces.items()) if c == "k")}
    rows = {}
    for white_pov in (True, False):
        def orient(sq):
            return sq if white_pov else sq ^ 63
        row = []
        for sq, ch in sorted(pieces.items(), key=lambda it: ("pnbrq".find(it[1].lower()) * 2 + it[1].islower(), it[0])):
            if ch in "Kk":
                continue
            p_idx = "pnbrq".index(ch.lower()) * 2 + (ch.isupper() != white_pov)
            row.append(1 + orient(sq) + p_idx * 64 + orient(ksq[white_pov]) * 641)
        rows[white_pov] = row
    return rows[True], rows[False]


def oracle_convert(mode, data):
    """Returns (rc, bytes)."""
    lib = oracle()
    out = ctypes.c_void_p()
    n = ctypes.c_size_t()
    rc = lib.orc_convert(mode, bytes(data), len(data), ctypes.byref(out), ctypes.byref(n))
    res = ctypes.string_at(out, n.value) if n.value else b""
    lib.orc_free(out)
    return rc, res


def have_ref():
    return os.access(REF_BIN, os.X_OK) and os.access(REF_GEN, os.X_OK)


def ref_convert(mode, data, append_to=None):
    """Runs the compiled reference CLI on `data`; returns the output file's bytes."""
    ein, eout = MODE_EXT[mode]
    with tempfile.TemporaryDirectory() as d:
        pin, pout = os.path.join(d, "in" + ein), os.path.join(d, "out" + eout)
        with open(pin, "wb") as f:
            f.write(data)
        args = [REF_BIN]
        if append_to is not None:
            with open(pout, "wb") as f:
                f.write(append_to)
            args.append("-a")
        subprocess.run(args + [pin, pout], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        if not os.path.exists(pout):
            return b""
        with open(pout, "rb") as f:
            return f.read()


def ref_generate(n, plies, seed, mode=0):
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "g.bin")
        subprocess.run([REF_GEN, p, str(n), str(plies), str(seed), str(mode)], check=True)
        with open(p, "rb") as f:
            return f.read()


def golden(name):
    """Reads tests/golden/<name>.gz (written by tests/golden/make_golden.py from the reference)."""
    import gzip

    with gzip.open(os.path.join(GOLDEN, name + ".gz"), "rb") as f:
        return f.read()


GOLDEN_SETS = ["games100", "heads", "long400", "shuffled", "restart", "twochunks"]
GOLDEN_PLAIN_SETS = ["games100", "heads", "long400"]


_sim = None


def host_sim():
    """tests/host_sim: the device headers compiled for the CPU (TEST INFRASTRUCTURE, see host_sim.h)."""
    global _sim
    if _sim is None:
        sim_dir = os.path.join(ROOT, "tests", "host_sim")
        so = os.path.join(sim_dir, "libsim.so")
        csrc = os.path.join(ROOT, "nnue_data_compress_b200", "csrc")
        srcs = [os.path.join(sim_dir, "sim.cpp"), os.path.join(sim_dir, "host_sim.h")] + [
            os.path.join(csrc, f) for f in ("chess.cuh", "stream.cuh", "walk.cuh", "link.cuh", "chain.cuh", "heads.cuh", "halfkp.cuh")]
        if not os.path.exists(so) or any(os.path.getmtime(x) > os.path.getmtime(so) for x in srcs):
            subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wno-unknown-pragmas", "-DNNP_HOST_SIM", "-I" + sim_dir,
                            "-shared", "-fPIC", "-o", so, srcs[0]], check=True)
        L = ctypes.CDLL(so)
        u64p = ctypes.POINTER(ctypes.c_uint64)
        L.sim_stream_check.argtypes = [ctypes.c_char_p, ctypes.c_size_t, u64p, u64p, u64p, u64p]
        L.sim_stream_fuzz.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint64, u64p]
        L.sim_stream_fuzz.restype = ctypes.c_uint64
        L.sim_walk_check.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, u64p, u64p]
        L.sim_walk_check.restype = ctypes.c_uint64
        L.sim_decode_binpack.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t]
        L.sim_decode_binpack.restype = ctypes.c_longlong
        L.sim_halfkp_chains.argtypes = [ctypes.c_char_p, ctypes.c_size_t, u64p, u64p]
        L.sim_halfkp_chains.restype = ctypes.c_longlong
        L.sim_halfkp_fuzz.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint64, u64p]
        L.sim_halfkp_fuzz.restype = ctypes.c_uint64
        L.sim_halfkp_rows.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
        L.sim_halfkp_rows.restype = ctypes.c_longlong
        L.sim_stem_transcode_fuzz.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint64, u64p]
        L.sim_stem_transcode_fuzz.restype = ctypes.c_uint64
        L.sim_heads_transcode_fuzz.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint64, u64p]
        L.sim_heads_transcode_fuzz.restype = ctypes.c_uint64
        L.sim_sfen_decode_fuzz.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint64, u64p]
        L.sim_sfen_decode_fuzz.restype = ctypes.c_uint64
        L.sim_halfkp_tokens.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
        L.sim_halfkp_tokens.restype = ctypes.c_uint64
        _sim = L
    return _sim
