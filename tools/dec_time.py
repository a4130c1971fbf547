"""Diagnostic (GPU box): times binpack->bin at N positions (argv[1], default 100M) of at most argv[2] plies per
chain (default 100) for the library named by NNP_LIB."""
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
assert L.nnp_bin_to_binpack_dev(ctypes.c_void_p(d_bin.data_ptr()), n * 40, ctypes.c_void_p(d_pack.data_ptr()), cap, ctypes.byref(sz)) == 0
a = ctypes.c_float(); b = ctypes.c_float()
L.nnp_last_timing(ctypes.byref(a), ctypes.byref(b))
print("compress ms", a.value, b.value)
d_out = torch.empty(n * 40, dtype=torch.uint8, device="cuda")
o = ctypes.c_size_t(0)
for it in range(3):
    rc = L.nnp_binpack_to_bin_dev(ctypes.c_void_p(d_pack.data_ptr()), sz.value, ctypes.c_void_p(d_out.data_ptr()), n * 40, ctypes.byref(o))
    L.nnp_last_timing(ctypes.byref(a), ctypes.byref(b))
print("decompress rc", rc, "ms", a.value, b.value, "hits", nnp.decode_stats()["optimistic_hits"])
sz2 = ctypes.c_size_t(0)
d_pack2 = torch.empty(cap, dtype=torch.uint8, device="cuda")
assert L.nnp_bin_to_binpack_dev(ctypes.c_void_p(d_out.data_ptr()), n * 40, ctypes.c_void_p(d_pack2.data_ptr()), cap, ctypes.byref(sz2)) == 0
print("re-encode identical:", sz2.value == sz.value and torch.equal(d_pack[: sz.value], d_pack2[: sz.value]))
