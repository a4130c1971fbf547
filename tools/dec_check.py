import sys, ctypes, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import nnue_data_compress_b200 as nnp
nnp.init(0)
for n, pl in ((100_000_000, 100),):
    import torch
    L = nnp.lib()
    d_bin = torch.empty(n * 40, dtype=torch.uint8, device="cuda")
    assert L.nnp_generate_bin_dev(ctypes.c_void_p(d_bin.data_ptr()), n, pl, 42) == 0
    cap = n * 40 // 8 + (1 << 20)
    d_pack = torch.empty(cap, dtype=torch.uint8, device="cuda")
    sz = ctypes.c_size_t(0)
    assert L.nnp_bin_to_binpack_dev(ctypes.c_void_p(d_bin.data_ptr()), n * 40, ctypes.c_void_p(d_pack.data_ptr()), cap, ctypes.byref(sz)) == 0
    d_out = torch.empty(n * 40, dtype=torch.uint8, device="cuda")
    o = ctypes.c_size_t(0)
    for it in range(3):
        rc = L.nnp_binpack_to_bin_dev(ctypes.c_void_p(d_pack.data_ptr()), sz.value, ctypes.c_void_p(d_out.data_ptr()), n * 40, ctypes.byref(o))
        a = ctypes.c_float(); b = ctypes.c_float(); L.nnp_last_timing(ctypes.byref(a), ctypes.byref(b))
        print("rc", rc, o.value, "ms", a.value, b.value, nnp.decode_stats())
    # compare with the exhaustive strategy
    ref = d_out.clone()
    L.nnp_debug_config(b"exhaustive", 1)
    rc = L.nnp_binpack_to_bin_dev(ctypes.c_void_p(d_pack.data_ptr()), sz.value, ctypes.c_void_p(d_out.data_ptr()), n * 40, ctypes.byref(o))
    a = ctypes.c_float(); b = ctypes.c_float(); L.nnp_last_timing(ctypes.byref(a), ctypes.byref(b))
    print("exhaustive rc", rc, o.value, "ms", a.value, b.value, nnp.decode_stats())
    print("optimistic == exhaustive:", torch.equal(ref, d_out))
    L.nnp_debug_config(b"exhaustive", 0)
    # and the compressor reproduces the binpack from the decoded records
    d_pack2 = torch.empty(cap, dtype=torch.uint8, device="cuda")
    sz2 = ctypes.c_size_t(0)
    assert L.nnp_bin_to_binpack_dev(ctypes.c_void_p(ref.data_ptr()), n * 40, ctypes.c_void_p(d_pack2.data_ptr()), cap, ctypes.byref(sz2)) == 0
    print("re-encode identical:", sz2.value == sz.value and torch.equal(d_pack[: sz.value], d_pack2[: sz.value]))
