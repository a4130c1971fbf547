"""Regenerates the golden fixtures in this directory from the *compiled reference*
(oracle/_ref, built by oracle/Makefile from /root/reference). Run in the build container:

    python tests/golden/make_golden.py

Every file is an output of the unmodified reference tool (or of gen_ref, which drives the
reference's own chess library); nothing here was produced by the code under test. Files are
gzip-compressed to keep the repository small; tests/refutil.py::golden() reads them.
"""
import gzip
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from refutil import (BIN_TO_BINPACK, BIN_TO_PLAIN, BINPACK_TO_BIN, BINPACK_TO_PLAIN, PLAIN_TO_BIN, PLAIN_TO_BINPACK,
                     ref_convert, ref_generate)

HERE = os.path.dirname(os.path.abspath(__file__))

# name -> (positions, max plies per chain, seed, generator mode, with .plain outputs)
SETS = {
    "games100": (1000, 100, 1, 0, True),      # ordinary games, chains of ~100 plies
    "heads": (300, 1, 2, 0, True),            # every record is a chain head
    "long400": (500, 400, 3, 0, True),        # promotions, castling, endgames
    "shuffled": (400, 100, 5, 1, False),      # constant gamePly: no record continues its predecessor
    "restart": (600, 60, 11, 2, False),       # game restarts that pass the result/ply test of isContinuation
    "twochunks": (40000, 1, 4, 0, False),     # > 1 MiB of payload: exercises the chunk-flush rule
}

# SURVEY.md section 8c known-answer test (three hand-written records)
KAT_PLAIN = (
    b"fen rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1\nmove e2e4\nscore 25\nply 0\nresult 1\ne\n"
    b"fen rnbqkbnr/pppppppp/8/8/4P3/8/PPPP1PPP/RNBQKBNR b KQkq e3 0 1\nmove e7e5\nscore -20\nply 1\nresult -1\ne\n"
    b"fen rnbqkbnr/pppp1ppp/8/4p3/4P3/8/PPPP1PPP/RNBQKBNR w KQkq e6 0 2\nmove g1f3\nscore 30\nply 2\nresult 1\ne\n"
)


def put(name, data):
    with gzip.GzipFile(os.path.join(HERE, name + ".gz"), "wb", mtime=0) as f:
        f.write(data)
    return len(data)


def main():
    manifest = {}
    for name, (n, plies, seed, mode, with_plain) in SETS.items():
        b = ref_generate(n, plies, seed, mode)
        bp = ref_convert(BIN_TO_BINPACK, b)
        rt = ref_convert(BINPACK_TO_BIN, bp)
        entry = {"positions": n, "plies": plies, "seed": seed, "mode": mode}
        entry["bin"] = put(f"{name}.bin", b)
        entry["binpack"] = put(f"{name}.binpack", bp)
        entry["roundtrip_bin"] = put(f"{name}.rt.bin", rt)
        if with_plain:
            pl = ref_convert(BINPACK_TO_PLAIN, bp)
            entry["plain"] = put(f"{name}.plain", pl)
            entry["plain_binpack"] = put(f"{name}.p.binpack", ref_convert(PLAIN_TO_BINPACK, pl))
            bpl = ref_convert(BIN_TO_PLAIN, b)
            entry["bin_plain"] = put(f"{name}.b.plain", bpl)
            entry["plain_bin"] = put(f"{name}.p.bin", ref_convert(PLAIN_TO_BIN, bpl))
        manifest[name] = entry
    kat_bp = ref_convert(PLAIN_TO_BINPACK, KAT_PLAIN)
    put("kat.plain", KAT_PLAIN)
    put("kat.binpack", kat_bp)
    put("kat.bin", ref_convert(BINPACK_TO_BIN, kat_bp))
    put("kat.rt.plain", ref_convert(BINPACK_TO_PLAIN, kat_bp))
    manifest["kat"] = {"binpack_hex": kat_bp.hex()}
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
