"""GPU suite: byte parity with the compiled reference at the headline scale (BASELINE configs[1]:
100 M positions, byte-identical) and past it: 112 M records = 4.48e9 bytes of .bin, so every byte
offset of the record array crosses 2^32. The reference binary (oracle/_ref, built from
/root/reference by oracle/Makefile) runs single-process on the host copy of the same records --
compressBin then decompressBin, compress_file.cpp:1338-1412 -- and both of its files are compared
byte for byte with the output of the CUDA path. Takes about four minutes, nearly all of it the
reference."""
import ctypes
import os
import shutil
import subprocess
import tempfile
import time

import pytest

from refutil import REF_BIN, have_ref

pytestmark = pytest.mark.gpu

N_RECORDS = 112_000_000
SLAB = 256 << 20


def _scratch_dir(need_bytes):
    for base in ("/dev/shm", tempfile.gettempdir()):
        if os.path.isdir(base) and shutil.disk_usage(base).free > need_bytes:
            return tempfile.mkdtemp(prefix="nnp_parity_", dir=base)
    return None


def _file_equals_device(path, d_buf, n_bytes):
    """Compares a file with the first n_bytes of a CUDA uint8 tensor, slab by slab."""
    import numpy as np

    if os.path.getsize(path) != n_bytes:
        return False, f"size {os.path.getsize(path)} != {n_bytes}"
    with open(path, "rb") as f:
        for off in range(0, n_bytes, SLAB):
            want = np.frombuffer(f.read(SLAB), dtype=np.uint8)
            got = d_buf[off:off + len(want)].cpu().numpy()
            if not np.array_equal(want, got):
                first = int(np.nonzero(want != got)[0][0])
                return False, f"first difference at byte {off + first}"
    return True, ""


@pytest.mark.timeout(1500)
def test_reference_parity_112m(nnp):
    import torch

    if not have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    n = N_RECORDS
    bin_bytes = n * 40
    assert bin_bytes > 1 << 32
    work = _scratch_dir(2 * bin_bytes + (1 << 30))
    if work is None:
        pytest.skip("no scratch space for two 4.5 GB files")
    L = nnp.lib()
    try:
        d_bin = torch.empty(bin_bytes, dtype=torch.uint8, device="cuda")
        assert L.nnp_generate_bin_dev(ctypes.c_void_p(d_bin.data_ptr()), n, 100, 20260) == 0
        cap = bin_bytes // 8 + (1 << 20)
        d_pack = torch.empty(cap, dtype=torch.uint8, device="cuda")
        n_pack = ctypes.c_size_t(0)
        assert L.nnp_bin_to_binpack_dev(ctypes.c_void_p(d_bin.data_ptr()), bin_bytes, ctypes.c_void_p(d_pack.data_ptr()), cap,
                                        ctypes.byref(n_pack)) == 0
        d_rt = torch.empty(bin_bytes, dtype=torch.uint8, device="cuda")
        n_rt = ctypes.c_size_t(0)
        assert L.nnp_binpack_to_bin_dev(ctypes.c_void_p(d_pack.data_ptr()), n_pack.value, ctypes.c_void_p(d_rt.data_ptr()),
                                        bin_bytes, ctypes.byref(n_rt)) == 0
        assert n_rt.value == bin_bytes
        assert nnp.decode_stats()["violations"] == 0  # the optimistic single walk stood

        p_bin, p_pack, p_rt = (os.path.join(work, x) for x in ("in.bin", "out.binpack", "rt.bin"))
        with open(p_bin, "wb") as f:
            for off in range(0, bin_bytes, SLAB):
                f.write(d_bin[off:off + SLAB].cpu().numpy().tobytes())
        del d_bin
        t0 = time.time()
        subprocess.run([REF_BIN, p_bin, p_pack], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        t1 = time.time()
        os.remove(p_bin)
        ok, why = _file_equals_device(p_pack, d_pack, n_pack.value)
        assert ok, ".bin -> .binpack differs from the reference: " + why
        subprocess.run([REF_BIN, p_pack, p_rt], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        t2 = time.time()
        ok, why = _file_equals_device(p_rt, d_rt, bin_bytes)
        assert ok, ".binpack -> .bin differs from the reference: " + why
        print(f"\n112M records: reference compress {t1 - t0:.0f} s, decompress {t2 - t1:.0f} s; "
              f"binpack {n_pack.value} bytes identical, .bin {bin_bytes} bytes identical")
    finally:
        shutil.rmtree(work, ignore_errors=True)


@pytest.mark.timeout(900)
@pytest.mark.parametrize("kind,n", [("shuffled", 24_000_000), ("startpos", 12_000_000)])
def test_reference_parity_single_positions(nnp, kind, n):
    """Files of single positions at a scale with hundreds of chunks, against the reference binary: the
    positions of 100-ply games in random order (what a shuffled training set looks like: nearly every record starts a
    chain, a few happen to continue their predecessor -- K1 transcodes, the decoder collapses the chunks that hold
    nothing but 34-byte chains) and the generator's own chain-length-1 file (every record a head: one kernel each
    way). Both files of the reference are compared byte for byte."""
    import torch

    if not have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    bin_bytes = n * 40
    work = _scratch_dir(3 * bin_bytes)
    if work is None:
        pytest.skip("no scratch space")
    L = nnp.lib()
    try:
        d_bin = torch.empty(bin_bytes, dtype=torch.uint8, device="cuda")
        assert L.nnp_generate_bin_dev(ctypes.c_void_p(d_bin.data_ptr()), n, 100 if kind == "shuffled" else 1, 77) == 0
        if kind == "shuffled":
            g = torch.Generator(device="cuda")
            g.manual_seed(5)
            rows = d_bin.view(torch.int64).view(n, 5)
            rows.copy_(rows[torch.randperm(n, device="cuda", generator=g)])
            torch.cuda.synchronize()
        cap = n * 34 + (1 << 20)
        d_pack = torch.empty(cap, dtype=torch.uint8, device="cuda")
        n_pack = ctypes.c_size_t(0)
        assert L.nnp_bin_to_binpack_dev(ctypes.c_void_p(d_bin.data_ptr()), bin_bytes, ctypes.c_void_p(d_pack.data_ptr()), cap,
                                        ctypes.byref(n_pack)) == 0
        k_compress = L.nnp_last_dominant_kernel()
        d_rt = torch.empty(bin_bytes, dtype=torch.uint8, device="cuda")
        n_rt = ctypes.c_size_t(0)
        assert L.nnp_binpack_to_bin_dev(ctypes.c_void_p(d_pack.data_ptr()), n_pack.value, ctypes.c_void_p(d_rt.data_ptr()),
                                        bin_bytes, ctypes.byref(n_rt)) == 0
        k_decompress = L.nnp_last_dominant_kernel()
        assert n_rt.value == bin_bytes
        if kind == "startpos":
            assert (k_compress, k_decompress) == (b"k_heads_direct", b"k_emit_heads_only")
        else:
            assert k_compress in (b"k_heads_transcode", b"k_heads_direct")
            assert k_decompress in (b"k_emit_chains_verify", b"k_emit_heads_only")

        p_bin, p_pack, p_rt = (os.path.join(work, x) for x in ("in.bin", "out.binpack", "rt.bin"))
        with open(p_bin, "wb") as f:
            for off in range(0, bin_bytes, SLAB):
                f.write(d_bin[off:off + SLAB].cpu().numpy().tobytes())
        subprocess.run([REF_BIN, p_bin, p_pack], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        ok, why = _file_equals_device(p_pack, d_pack, n_pack.value)
        assert ok, ".bin -> .binpack differs from the reference: " + why
        subprocess.run([REF_BIN, p_pack, p_rt], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        ok, why = _file_equals_device(p_rt, d_rt, bin_bytes)
        assert ok, ".binpack -> .bin differs from the reference: " + why
    finally:
        shutil.rmtree(work, ignore_errors=True)
