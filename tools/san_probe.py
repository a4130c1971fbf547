"""Diagnostic (GPU box): a small pass over every compressor / decoder route with consistency checks (all K1 forms,
the single-position routes and their fallbacks, collapsed chunks, .plain). Small enough to run under
compute-sanitizer where a pool allows it (this round's pool does not)."""
import sys

import numpy as np

sys.path.insert(0, ".")
import nnue_data_compress_b200 as nnp

nnp.init(0)
L = nnp.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 70_000
games = nnp.generate_bin(n, 100, 3)
singles = nnp.generate_bin(n, 1, 4)
rows = np.frombuffer(games, dtype=np.uint8).reshape(-1, 40)
shuffled = rows[np.random.default_rng(1).permutation(len(rows))].tobytes()
mixed = singles + games[: 40 * (n // 2)] + singles
for name, data in (("games", games), ("singles", singles), ("shuffled", shuffled), ("mixed", mixed)):
    for key in (None, b"k1_walk", b"k1_runs", b"k1_per_record", b"k1_heads"):
        if key:
            L.nnp_debug_config(key, 1)
        pack = nnp.bin_to_binpack(data)
        kernel = L.nnp_last_dominant_kernel().decode()
        if key:
            L.nnp_debug_config(key, 0)
        back = nnp.binpack_to_bin(pack)
        assert len(back) == len(data), (name, key)
        print(name, key, kernel, L.nnp_last_dominant_kernel().decode(), len(pack))
    L.nnp_debug_config(b"k1_direct", 1)
    assert nnp.bin_to_binpack(data) == pack
    L.nnp_debug_config(b"k1_direct", 0)
    text = nnp.binpack_to_plain(pack)
    again = nnp.plain_to_binpack(text)  # (not the same file: .plain does not carry everything a stem does)
    assert nnp.binpack_to_plain(again) == text
print("ok")
