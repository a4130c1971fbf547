"""One-off large parity check against the compiled reference binary (oracle/_ref): device-generated
games of several chain lengths are converted by the reference CLI on the host and by the CUDA path."""
import sys
import time

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import nnue_data_compress_b200 as nnp
from refutil import BIN_TO_BINPACK, BINPACK_TO_BIN, BINPACK_TO_PLAIN, have_ref, ref_convert

assert have_ref(), "oracle/_ref missing"
nnp.init(0)
ok = True
for n, plies, seed in ((8_000_000, 100, 101), (2_000_000, 1, 102), (4_000_000, 8, 103), (5_000_000, 400, 104)):
    b = nnp.generate_bin(n, plies, seed)
    t0 = time.time()
    want = ref_convert(BIN_TO_BINPACK, b)
    t1 = time.time()
    got = nnp.bin_to_binpack(b)
    same_c = got == want
    want_bin = ref_convert(BINPACK_TO_BIN, want)
    same_d = nnp.binpack_to_bin(want) == want_bin
    same_p = True
    if n <= 4_000_000:
        same_p = nnp.binpack_to_plain(want) == ref_convert(BINPACK_TO_PLAIN, want)
    print(f"REF_CHECK n={n} plies={plies}: binpack {len(want)} B, reference compress {t1 - t0:.1f} s; "
          f"bin->binpack {'IDENTICAL' if same_c else 'DIFFERS'}, binpack->bin {'IDENTICAL' if same_d else 'DIFFERS'}, "
          f"binpack->plain {'IDENTICAL' if same_p else 'DIFFERS'}", flush=True)
    ok = ok and same_c and same_d and same_p
print("REF_CHECK", "ALL IDENTICAL" if ok else "MISMATCH")
