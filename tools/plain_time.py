"""Diagnostic (GPU box): the four .plain directions on device buffers, timed, with consistency checks."""
import ctypes
import sys
import time

import torch

sys.path.insert(0, ".")
import nnue_data_compress_b200 as nnp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
nnp.init(0)
nnp.use_torch_stream()
L = nnp.lib()


def call(name, src, src_bytes, cap):
    out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    got = ctypes.c_size_t(0)
    fn = getattr(L, name)
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rc = fn(ctypes.c_void_p(src.data_ptr()), src_bytes, ctypes.c_void_p(out.data_ptr()), cap, ctypes.byref(got))
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
        assert rc == 0, (name, rc)
    print(f"{name:28s} {best * 1e3:8.2f} ms  {n / best / 1e6:9.1f} Mpos/s  in {src_bytes / 1e6:.1f} MB out {got.value / 1e6:.1f} MB", flush=True)
    return out, got.value


d_bin = torch.empty(n * 40, dtype=torch.uint8, device="cuda")
assert L.nnp_generate_bin_dev(ctypes.c_void_p(d_bin.data_ptr()), n, 100, 42) == 0
pack, pack_n = call("nnp_bin_to_binpack_dev", d_bin, n * 40, n * 5 + (1 << 20))
plain, plain_n = call("nnp_bin_to_plain_dev", d_bin, n * 40, n * 130)
bin2, bin2_n = call("nnp_plain_to_bin_dev", plain, plain_n, n * 40)
plain2, plain2_n = call("nnp_binpack_to_plain_dev", pack, pack_n, n * 130)
pack2, pack2_n = call("nnp_plain_to_binpack_dev", plain2, plain2_n, n * 5 + (1 << 20))
rt, rt_n = call("nnp_binpack_to_bin_dev", pack, pack_n, n * 40)
# Note: binpack -> plain -> binpack and binpack -> bin -> plain are NOT round trips in the reference
# either (full-move counters and the UCI en-passant rule); parity of every
# direction is tested against the oracle in tests/test_gpu_parity.py.
print("binpack -> plain -> binpack sizes:", pack_n, pack2_n)
