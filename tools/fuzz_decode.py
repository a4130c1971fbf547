"""Diagnostic (GPU box): corrupted .binpack inputs must come back with a status, never hang or fault.
Run under `timeout`. Random byte flips / truncations / header edits of golden files."""
import random
import sys
import time

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import nnue_data_compress_b200 as nnp
from refutil import golden

nnp.init(0)
rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 300
bases = [golden(n + ".binpack") for n in ("games100", "long400", "heads", "restart")] + [golden("twochunks.binpack")[:200_000]]
stat = {}
t0 = time.time()
for it in range(rounds):
    b = bytearray(rng.choice(bases))
    kind = rng.randrange(5)
    if kind == 0:
        for _ in range(rng.randrange(1, 8)):
            b[rng.randrange(len(b))] ^= 1 << rng.randrange(8)
    elif kind == 1:
        del b[rng.randrange(8, len(b)):]
    elif kind == 2:
        i = rng.randrange(len(b) - 8)
        b[i:i + 8] = bytes(rng.randrange(256) for _ in range(8))
    elif kind == 3:
        b[4:8] = rng.randrange(1 << 32).to_bytes(4, "little")
    else:
        i = rng.randrange(8, len(b))
        b[i:] = bytes([0xFF]) * (len(b) - i)
    for fn in (nnp.binpack_to_bin, nnp.binpack_to_plain):
        try:
            fn(bytes(b))
            key = "ok"
        except nnp.NnpError as e:
            key = str(e.status)
        stat[key] = stat.get(key, 0) + 1
# corrupted .bin and .plain inputs through the other drivers
bin_bases = [golden(n + ".bin") for n in ("games100", "long400", "restart")]
plain_bases = [golden(n + ".plain") for n in ("games100", "long400")]
for it in range(rounds):
    b = bytearray(rng.choice(bin_bases))
    for _ in range(rng.randrange(1, 30)):
        b[rng.randrange(len(b))] ^= 1 << rng.randrange(8)
    if rng.randrange(4) == 0:
        del b[rng.randrange(1, len(b)):]
    t = bytearray(rng.choice(plain_bases))
    for _ in range(rng.randrange(1, 10)):
        t[rng.randrange(len(t))] = rng.choice(b"0123456789abcdefghKQRBNPkqrbnp/- \nwe")
    for fn, data in ((nnp.bin_to_binpack, b), (nnp.bin_to_plain, b), (nnp.plain_to_bin, t), (nnp.plain_to_binpack, t)):
        try:
            fn(bytes(data))
            key = "ok"
        except nnp.NnpError as e:
            key = str(e.status)
        stat[key] = stat.get(key, 0) + 1
# the library is still healthy
assert nnp.binpack_to_bin(golden("games100.binpack")) == golden("games100.rt.bin")
print("FUZZ_DONE", rounds, "inputs in", round(time.time() - t0, 1), "s; statuses", stat)
