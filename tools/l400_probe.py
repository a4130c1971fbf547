import ctypes, sys, torch
sys.path.insert(0, ".")
import nnue_data_compress_b200 as nnp
n = 125_000_000
nnp.init(0); L = nnp.lib()
for seed in (42, 1042, 2042, 3042):
    d_bin = torch.empty(n * 40, dtype=torch.uint8, device="cuda")
    assert L.nnp_generate_bin_dev(ctypes.c_void_p(d_bin.data_ptr()), n, 400, seed) == 0
    cap = n * 40 // 8 + (1 << 20)
    d_pack = torch.empty(cap, dtype=torch.uint8, device="cuda")
    sz = ctypes.c_size_t(0)
    assert L.nnp_bin_to_binpack_dev(ctypes.c_void_p(d_bin.data_ptr()), n * 40, ctypes.c_void_p(d_pack.data_ptr()), cap, ctypes.byref(sz)) == 0
    d_out = d_bin  # reuse
    o = ctypes.c_size_t(0)
    a = ctypes.c_float(); b = ctypes.c_float()
    for it in range(2):
        rc = L.nnp_binpack_to_bin_dev(ctypes.c_void_p(d_pack.data_ptr()), sz.value, ctypes.c_void_p(d_out.data_ptr()), n * 40, ctypes.byref(o))
        L.nnp_last_timing(ctypes.byref(a), ctypes.byref(b))
    st = nnp.decode_stats()
    print("seed", seed, "rc", rc, "ms", round(a.value, 2), round(b.value, 2), L.nnp_last_dominant_kernel().decode(), {k: st[k] for k in ("optimistic_hits", "optimistic_misses", "candidates", "violations", "false_candidates")}, st["false_sample"][:2])
    del d_pack
