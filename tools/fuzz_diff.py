"""Diagnostic (GPU box): differential fuzz against the CPU oracle. Corrupted .binpack / .bin inputs go
through the CUDA path and through the oracle; wherever the oracle (= the reference's semantics) gives a
result (OK or one of the three reference errors), status and bytes must agree."""
import random
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import nnue_data_compress_b200 as nnp
from refutil import BIN_TO_BINPACK, BINPACK_TO_BIN, golden, oracle_convert

nnp.init(0)
rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 400
REF = (0, -1, -2, -3)


def ours(fn, data):
    try:
        return 0, fn(data)
    except nnp.NnpError as e:
        return e.status, (e.partial or b"")


stats = {"agree": 0, "skipped": 0, "differ": 0}
examples = []
packs = [golden(n + ".binpack") for n in ("games100", "long400", "restart", "heads")]
bins = [golden(n + ".bin") for n in ("games100", "long400", "restart")]
for it in range(rounds):
    if it & 1:
        b = bytearray(rng.choice(packs))
        for _ in range(rng.randrange(1, 4)):
            i = rng.randrange(8, len(b))
            b[i] ^= 1 << rng.randrange(8)
        mode, fn = BINPACK_TO_BIN, nnp.binpack_to_bin
    else:
        b = bytearray(rng.choice(bins))
        for _ in range(rng.randrange(1, 4)):
            i = rng.randrange(len(b))
            b[i] ^= 1 << rng.randrange(8)
        mode, fn = BIN_TO_BINPACK, nnp.bin_to_binpack
    rc_o, out_o = oracle_convert(mode, bytes(b))
    if rc_o not in REF:
        stats["skipped"] += 1
        continue
    rc, out = ours(fn, bytes(b))
    if rc == rc_o and out == out_o:
        stats["agree"] += 1
    else:
        stats["differ"] += 1
        if len(examples) < 5:
            examples.append((it, mode, rc_o, rc, len(out_o), len(out)))
print("FUZZ_DIFF", stats, examples)
