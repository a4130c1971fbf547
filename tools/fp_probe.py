"""Diagnostic (GPU box): how many chain-start candidates of a large binpack are false, and where."""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
import nnue_data_compress_b200 as nnp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
plies = int(sys.argv[2]) if len(sys.argv) > 2 else 100
nnp.init(0)
L = nnp.lib()
d_bin = torch.empty(n * 40, dtype=torch.uint8, device="cuda")
assert L.nnp_generate_bin_dev(ctypes.c_void_p(d_bin.data_ptr()), n, plies, 42) == 0
cap = n * 40 // 8 + (1 << 20) if plies > 20 else n * 36 + (1 << 20)
d_pack = torch.empty(cap, dtype=torch.uint8, device="cuda")
sz = ctypes.c_size_t(0)
rc = L.nnp_bin_to_binpack_dev(ctypes.c_void_p(d_bin.data_ptr()), n * 40, ctypes.c_void_p(d_pack.data_ptr()), cap, ctypes.byref(sz))
assert rc == 0, rc
d_out = torch.empty(n * 40, dtype=torch.uint8, device="cuda")
o = ctypes.c_size_t(0)
rc = L.nnp_binpack_to_bin_dev(ctypes.c_void_p(d_pack.data_ptr()), sz.value, ctypes.c_void_p(d_out.data_ptr()), n * 40, ctypes.byref(o))
st = nnp.decode_stats()
print("rc", rc, "binpack", sz.value, "out", o.value, st)
pack = d_pack[: sz.value].cpu().numpy()
# chunk table on the host
starts = []
pos = 0
while pos < len(pack):
    size = int.from_bytes(pack[pos + 4:pos + 8].tobytes(), "little")
    starts.append(pos + 8)
    pos += 8 + size
for c, off in st["false_sample"]:
    if c == 0 and off == 0:
        continue
    a = starts[c] + off
    print("false candidate chunk", c, "off", off, pack[a:a + 34].tobytes().hex())
    print("   before:", pack[a - 40:a].tobytes().hex())
