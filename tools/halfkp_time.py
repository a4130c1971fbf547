"""Diagnostic (GPU box): times .binpack -> HalfKP rows at N positions against the two-step route
(.binpack -> .bin, then .bin -> rows) and checks both give the same rows."""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
import nnue_data_compress_b200 as nnp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
nnp.init(0)
nnp.use_torch_stream()
L = nnp.lib()
P = lambda t: ctypes.c_void_p(t.data_ptr())
d_bin = torch.empty(n * 40, dtype=torch.uint8, device="cuda")
assert L.nnp_generate_bin_dev(P(d_bin), n, 100, 42) == 0
cap = n * 40 // 8 + (1 << 20)
d_pack = torch.empty(cap, dtype=torch.uint8, device="cuda")
sz = ctypes.c_size_t(0)
assert L.nnp_bin_to_binpack_dev(P(d_bin), n * 40, P(d_pack), cap, ctypes.byref(sz)) == 0
a = ctypes.c_float(); b = ctypes.c_float()
white = torch.empty((n, 32), dtype=torch.int32, device="cuda")
black = torch.empty((n, 32), dtype=torch.int32, device="cuda")
meta = torch.empty((n, 8), dtype=torch.uint8, device="cuda")
cnt = ctypes.c_size_t(0)
for it in range(3):
    rc = L.nnp_binpack_to_halfkp_dev(P(d_pack), sz.value, P(white), P(black), P(meta), n, ctypes.byref(cnt))
    L.nnp_last_timing(ctypes.byref(a), ctypes.byref(b))
fused = a.value
print(f"binpack->halfkp rc {rc} rows {cnt.value} ms {a.value:.3f} (emit {b.value:.3f})  {n / a.value / 1e3:.0f} Mpos/s  "
      f"out {(264 * n) / b.value / 1e6:.0f} GB/s  hits {nnp.decode_stats()['optimistic_hits']}")
o = ctypes.c_size_t(0)
w2 = torch.empty_like(white); k2 = torch.empty_like(black); m2 = torch.empty_like(meta)
for it in range(3):
    rc = L.nnp_binpack_to_bin_dev(P(d_pack), sz.value, P(d_bin), n * 40, ctypes.byref(o))
    L.nnp_last_timing(ctypes.byref(a), ctypes.byref(b))
    t1 = a.value
    rc2 = L.nnp_bin_to_halfkp_dev(P(d_bin), n * 40, P(w2), P(k2), P(m2), n, ctypes.byref(cnt))
    L.nnp_last_timing(ctypes.byref(a), ctypes.byref(b))
print(f"two-step: binpack->bin {t1:.3f} ms + bin->halfkp {a.value:.3f} ms = {t1 + a.value:.3f} ms (fused {fused:.3f})")
torch.cuda.synchronize()
big = torch.iinfo(torch.int32).max
M = 2_000_000


def sorted_pairs(w, k):
    order = torch.where(w < 0, big, w).argsort(dim=1, stable=True)
    return w.gather(1, order), k.gather(1, order)


a_w, a_k = sorted_pairs(white[:M], black[:M])
b_w, b_k = sorted_pairs(w2[:M], k2[:M])
print("same rows (first 2M, pairs sorted by white index):", torch.equal(a_w, b_w) and torch.equal(a_k, b_k) and torch.equal(meta, m2))
