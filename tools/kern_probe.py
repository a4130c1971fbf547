"""Diagnostic (GPU box): one compress + one decompress of N positions for the library named by NNP_LIB;
meant to run under `ncu --metrics ...` for instruction and pipe counts of the two dominant kernels."""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
import nnue_data_compress_b200 as nnp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
shuffled = len(sys.argv) > 2 and sys.argv[2] == "shuffled"  # positions of 100-ply games in random order
plies = 100 if shuffled or len(sys.argv) <= 2 else int(sys.argv[2])
nnp.init(0)
L = nnp.lib()
d_bin = torch.empty(n * 40, dtype=torch.uint8, device="cuda")
assert L.nnp_generate_bin_dev(ctypes.c_void_p(d_bin.data_ptr()), n, plies, 42) == 0
if shuffled:
    rows = d_bin.view(torch.int64).view(n, 5)
    rows.copy_(rows[torch.randperm(n, device="cuda")])
    plies = 1
cap = n * 40 // 8 + (1 << 20) if plies > 20 else n * 36 + (1 << 20)
d_pack = torch.empty(cap, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n * 40, dtype=torch.uint8, device="cuda")
sz, o = ctypes.c_size_t(0), ctypes.c_size_t(0)
if len(sys.argv) <= 3 or sys.argv[3] != "auto":
    L.nnp_debug_config(b"k1_walk", 1)  # third argument "auto": let the density sample pick K1
for _ in range(2):
    assert L.nnp_bin_to_binpack_dev(ctypes.c_void_p(d_bin.data_ptr()), n * 40, ctypes.c_void_p(d_pack.data_ptr()), cap, ctypes.byref(sz)) == 0
    assert L.nnp_binpack_to_bin_dev(ctypes.c_void_p(d_pack.data_ptr()), sz.value, ctypes.c_void_p(d_out.data_ptr()), n * 40, ctypes.byref(o)) == 0
print("ok", sz.value, o.value)
