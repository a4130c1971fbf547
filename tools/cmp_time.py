"""Diagnostic (GPU box): times bin->binpack at N positions for the library named by NNP_LIB and checks
the chain-walking K1 against the record-parallel K1 on the same input."""
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
outs = []
for mode in (0, 1, 2, 3):
    L.nnp_debug_config(b"k1_walk", int(mode == 0))
    L.nnp_debug_config(b"k1_per_record", int(mode == 1))
    L.nnp_debug_config(b"k1_runs", int(mode == 3))
    d_pack = torch.empty(cap, dtype=torch.uint8, device="cuda")
    sz = ctypes.c_size_t(0)
    a = ctypes.c_float(); b = ctypes.c_float()
    for it in range(3):
        assert L.nnp_bin_to_binpack_dev(ctypes.c_void_p(d_bin.data_ptr()), n * 40, ctypes.c_void_p(d_pack.data_ptr()), cap, ctypes.byref(sz)) == 0
        L.nnp_last_timing(ctypes.byref(a), ctypes.byref(b))
    print(("chain-walk K1", "per-record K1", "auto K1", "run-walk K1")[mode], "compress ms", round(a.value, 3), "K1 ms", round(b.value, 3), "bytes", sz.value)
    outs.append(d_pack[: sz.value])
print("identical:", all(o.shape == outs[0].shape and torch.equal(o, outs[0]) for o in outs[1:]))
