"""CPU suite: the C-ABI library loads and exports every symbol include/nnuepack.h declares, it
refuses to run without a GPU (no CPU fallback), and the host-side mirrors of the reference CLI
dispatch exactly like compress_file.cpp:1593-1709. No compute calls are made here."""
import ctypes
import io
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nnuepack.h")
LIB = os.path.join(ROOT, "nnue_data_compress_b200", "libnnuepack.so")
CLI = os.path.join(ROOT, "nnue_data_compress_b200", "nnue_data_compression")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nnp_[a-z0-9_]+)\s*\(", text)))


def have_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB), "build with python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(LIB)
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), name
    import nnue_data_compress_b200 as pkg

    assert sorted(pkg.EXPORTS) == names


def test_header_is_plain_c_and_library_links_from_c(tmp_path):
    """The drop-in boundary is a C ABI: the header compiles as C99 with -pedantic and a C program links
    against the library (and is refused without a device)."""
    exe = str(tmp_path / "abi_c_check")
    libdir = os.path.dirname(LIB)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "abi_c_check.c"), "-o", exe, "-L", libdir, "-lnnuepack",
                    "-Wl,-rpath," + libdir], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "ABI_C_OK" in r.stdout, r.stdout + r.stderr


def test_error_strings_are_the_reference_messages():
    lib = ctypes.CDLL(LIB)
    lib.nnp_strerror.restype = ctypes.c_char_p
    assert lib.nnp_strerror(-1) == b"Invalid binpack file or chunk."  # compress_file.cpp:506
    assert lib.nnp_strerror(-2) == b"Chunks size larger than supported. Malformed file?"  # :517
    assert lib.nnp_strerror(-3) == b"Improperly encoded bin sfen"  # :408, :442


def test_host_header_walk_names_the_chunk_ranges():
    """nnp_binpack_chunk_range is the header walk of compress_file.cpp:500-521 on host memory (no
    conversion, no device): balanced contiguous chunk ranges that cover the file; bad headers end it."""
    import sys

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import nnue_data_compress_b200 as pkg
    from nnue_data_compress_b200.sharding import chunk_bounds_binpack, shard_bounds
    from refutil import golden

    full = golden("twochunks.binpack") * 3  # appended files are valid files (:1663-1666): six chunks
    chunks = chunk_bounds_binpack(full)
    assert len(chunks) == 6
    L = pkg.lib()
    for world in (1, 2, 4, 6, 7):
        prev_hi = 0
        for r in range(world):
            rng = pkg.ChunkRange()
            assert L.nnp_binpack_chunk_range(full, len(full), world, r, ctypes.byref(rng)) == 0
            assert (rng.chunk_lo, rng.chunk_hi) == shard_bounds(6, world, r) and rng.chunks_total == 6
            assert rng.byte_lo == prev_hi
            assert rng.byte_hi == (chunks[rng.chunk_hi - 1][0] + chunks[rng.chunk_hi - 1][1] if rng.chunk_hi > rng.chunk_lo
                                   else rng.byte_lo)
            prev_hi = rng.byte_hi
        assert prev_hi == len(full)
    rng = pkg.ChunkRange()
    bad = bytearray(full)
    bad[chunks[2][0]] = ord("X")
    assert L.nnp_binpack_chunk_range(bytes(bad), len(bad), 2, 1, ctypes.byref(rng)) == -1  # NNP_ERR_BAD_MAGIC
    assert (rng.chunks_total, rng.chunk_lo, rng.chunk_hi, rng.byte_hi) == (2, 1, 2, chunks[2][0])
    huge = bytearray(full)
    huge[chunks[1][0] + 4:chunks[1][0] + 8] = (101 << 20).to_bytes(4, "little")
    assert L.nnp_binpack_chunk_range(bytes(huge), len(huge), 1, 0, ctypes.byref(rng)) == -2  # NNP_ERR_CHUNK_TOO_LARGE
    assert L.nnp_binpack_chunk_range(full[:-5], len(full) - 5, 1, 0, ctypes.byref(rng)) == -4  # NNP_ERR_TRUNCATED
    assert L.nnp_binpack_chunk_range(None, 0, 3, 2, ctypes.byref(rng)) == 0 and rng.chunks_total == 0


@pytest.mark.skipif(have_gpu(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    lib = ctypes.CDLL(LIB)
    n = ctypes.c_size_t(0)
    buf = ctypes.create_string_buffer(64)
    # not initialised: every driver refuses
    for name in ("nnp_bin_to_binpack", "nnp_binpack_to_bin", "nnp_plain_to_binpack", "nnp_binpack_to_plain",
                 "nnp_bin_to_plain", "nnp_plain_to_bin"):
        assert getattr(lib, name)(buf, 40, buf, 64, ctypes.byref(n)) == -10
    for name in ("nnp_bin_to_binpack_file", "nnp_binpack_to_bin_file"):
        fn = getattr(lib, name)
        fn.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p]
        assert fn(b"/nonexistent/in", b"/nonexistent/out", 0, 0, None) == -10
    for name in ("nnp_binpack_to_halfkp_dev", "nnp_bin_to_halfkp_dev"):
        fn = getattr(lib, name)
        fn.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                       ctypes.POINTER(ctypes.c_size_t)]
        assert fn(None, 0, None, None, None, 0, ctypes.byref(n)) == -10
    assert lib.nnp_init(0) == -9  # NNP_ERR_NO_DEVICE
    assert lib.nnp_init_all(0) == -9 and lib.nnp_bind_device(0) == -10 and lib.nnp_device_count() == 0
    lib.nnp_shard_decompress_dev.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                             ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t), ctypes.c_void_p]
    assert lib.nnp_shard_decompress_dev(None, 0, 1, 0, None, 0, ctypes.byref(n), buf) == -10
    import nnue_data_compress_b200 as pkg

    with pytest.raises(pkg.NnpError) as ei:
        pkg.bin_to_binpack(b"\0" * 40)
    assert ei.value.status == -9


def _run_main(pkg, argv, monkeypatch, calls):
    for name in ("bin_to_binpack", "binpack_to_bin", "plain_to_binpack", "binpack_to_plain", "bin_to_plain", "plain_to_bin"):
        monkeypatch.setattr(pkg, name, (lambda nm: (lambda data: calls.append(nm) or nm.encode()))(name))
    out, err = io.StringIO(), io.StringIO()
    rc = pkg.main(argv, out=out, err=err)
    return rc, out.getvalue(), err.getvalue()


def test_cli_mirror_dispatch(tmp_path, monkeypatch):
    import nnue_data_compress_b200 as pkg

    calls = []
    rc, out, err = _run_main(pkg, [], monkeypatch, calls)
    assert rc == 0 and out.startswith("Usage:")
    rc, out, err = _run_main(pkg, ["a", "b", "c"], monkeypatch, calls)
    assert rc == 1 and err == "Invalid arguments.\n"
    rc, out, err = _run_main(pkg, [str(tmp_path / "missing.bin"), "x"], monkeypatch, calls)
    assert rc == 0 and err == "Input file doesn't exist.\n"

    table = [("a.bin", "o", "bin_to_binpack", "o.binpack"), ("a.plain", "o.binpack", "plain_to_binpack", "o.binpack"),
             ("a.binpack", "o.bin", "binpack_to_bin", "o.bin"), ("a.binpack", "o.plain", "binpack_to_plain", "o.plain"),
             ("a.bin", "o.plain", "bin_to_plain", "o.plain"), ("a.plain", "o.bin", "plain_to_bin", "o.bin")]
    for src, dst, fn, produced in table:
        (tmp_path / src).write_bytes(b"x")
        calls.clear()
        rc, out, err = _run_main(pkg, [str(tmp_path / src), str(tmp_path / dst)], monkeypatch, calls)
        assert rc == 0 and calls == [fn]
        assert (tmp_path / produced).read_bytes() == fn.encode()
    # -a appends, --append is silently ignored (stored as "-append", compress_file.cpp:1684-1687)
    (tmp_path / "o.binpack").write_bytes(b"bin_to_binpack")
    rc, _, _ = _run_main(pkg, ["-a", str(tmp_path / "a.bin"), str(tmp_path / "o")], monkeypatch, calls)
    assert (tmp_path / "o.binpack").read_bytes() == b"bin_to_binpack" * 2
    rc, _, _ = _run_main(pkg, ["--append", str(tmp_path / "a.bin"), str(tmp_path / "o")], monkeypatch, calls)
    assert (tmp_path / "o.binpack").read_bytes() == b"bin_to_binpack"
    (tmp_path / "a.txt").write_bytes(b"x")
    rc, out, err = _run_main(pkg, [str(tmp_path / "a.txt"), "o"], monkeypatch, calls)
    assert rc == 0 and err == "Unsupported extension."


def test_cli_mirror_reference_error(tmp_path, monkeypatch):
    """A reference-style error writes the partial output, prints the message and 'Exiting...',
    and still exits 0 (compress_file.cpp:1697-1709)."""
    import nnue_data_compress_b200 as pkg

    def boom(data):
        raise pkg.NnpError(-1, "Invalid binpack file or chunk.", partial=b"partial")

    monkeypatch.setattr(pkg, "binpack_to_bin", boom)
    (tmp_path / "a.binpack").write_bytes(b"x")
    out, err = io.StringIO(), io.StringIO()
    rc = pkg.main([str(tmp_path / "a.binpack"), str(tmp_path / "o.bin")], out=out, err=err)
    assert rc == 0
    assert err.getvalue() == "Invalid binpack file or chunk.\nExiting...\n"
    assert (tmp_path / "o.bin").read_bytes() == b"partial"


def test_native_cli_argument_handling():
    assert os.access(CLI, os.X_OK)
    r = subprocess.run([CLI], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("Usage:")
    r = subprocess.run([CLI, "-h", "a", "b"], capture_output=True, text=True)
    assert r.returncode == 0 and "nnue_data_compression [-h] [-a] input_path output_path" in r.stdout
    r = subprocess.run([CLI, "a", "b", "c"], capture_output=True, text=True)
    assert r.returncode == 1 and r.stderr.startswith("Invalid arguments.")
    if not have_gpu():
        r = subprocess.run([CLI, "a.bin", "b"], capture_output=True, text=True)
        assert r.returncode == 2 and "no usable sm_100a CUDA device" in r.stderr
