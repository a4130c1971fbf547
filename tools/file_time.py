"""Diagnostic (GPU box): file-to-file conversion times from tmpfs, every repetition printed.
    python tools/file_time.py [positions] [reps] [devices: 0 = all]"""
import ctypes
import os
import shutil
import sys
import tempfile
import time

import torch

sys.path.insert(0, ".")
import nnue_data_compress_b200 as nnp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
devs = int(sys.argv[3]) if len(sys.argv) > 3 else 1
slab = int(os.environ.get("SLAB", "0"))
if devs == 1:
    nnp.init(0)
else:
    print("devices", nnp.init_all(devs))
L = nnp.lib()
d_bin = torch.empty(n * 40, dtype=torch.uint8, device="cuda")
assert L.nnp_generate_bin_dev(ctypes.c_void_p(d_bin.data_ptr()), n, 100, 42) == 0
work = tempfile.mkdtemp(prefix="nnp_ft_", dir="/dev/shm")
try:
    p_in, p_pack, p_back = (os.path.join(work, x).encode() for x in ("in.bin", "out.binpack", "back.bin"))
    with open(p_in, "wb") as f:
        for off in range(0, n * 40, 256 << 20):
            f.write(d_bin[off:off + (256 << 20)].cpu().numpy().tobytes())
    del d_bin
    pos = ctypes.c_uint64(0)
    for r in range(reps):
        t0 = time.perf_counter()
        assert L.nnp_bin_to_binpack_file(p_in, p_pack, 0, slab, ctypes.byref(pos)) == 0
        t1 = time.perf_counter()
        assert L.nnp_binpack_to_bin_file(p_pack, p_back, 0, slab // 8, ctypes.byref(pos)) == 0
        t2 = time.perf_counter()
        a, b, sec = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_double()
        L.nnp_internal_pool_stats(ctypes.byref(a), ctypes.byref(b), ctypes.byref(sec))
        print(f"   pinned allocations so far: {a.value}, {b.value / 1e9:.2f} GB, {sec.value:.2f} s")
        print(f"rep {r}: bin->binpack {t1 - t0:.3f} s = {n * 40 / (t1 - t0) / 1e9:.2f} GB/s in; binpack->bin {t2 - t1:.3f} s = "
              f"{n * 40 / (t2 - t1) / 1e9:.2f} GB/s out; positions {pos.value}", flush=True)
finally:
    shutil.rmtree(work, ignore_errors=True)
