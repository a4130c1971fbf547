"""nnue_data_compress_b200 -- B200-native (sm_100a) conversion path for NNUE training data.

Host-side mirror of the reference tool's interface (Sopel97/nnue_data_compress,
``src/compress_file.cpp:1535-1709``): the six file drivers, ``convert()`` with its
extension dispatch, and ``main()``/``run()`` with the same flags and messages. Everything
that touches the data runs in ``libnnuepack.so`` (hand-written CUDA behind the C ABI of
``include/nnuepack.h``); this module is ctypes plumbing only and raises if the library or a
GPU is missing -- there is no CPU path.
"""
from __future__ import annotations

import ctypes
import os
import sys

__all__ = [
    "NnpError",
    "lib",
    "init",
    "shutdown",
    "bin_to_binpack",
    "binpack_to_bin",
    "plain_to_binpack",
    "binpack_to_plain",
    "bin_to_plain",
    "plain_to_bin",
    "convert",
    "main",
    "STATUS",
]

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NNP_LIB", os.path.join(_HERE, "libnnuepack.so"))  # NNP_LIB: experiment builds

STATUS = {
    0: "NNP_OK",
    -1: "NNP_ERR_BAD_MAGIC",
    -2: "NNP_ERR_CHUNK_TOO_LARGE",
    -3: "NNP_ERR_BAD_SFEN",
    -4: "NNP_ERR_TRUNCATED",
    -5: "NNP_ERR_NOMEM",
    -6: "NNP_ERR_BAD_ARG",
    -7: "NNP_ERR_BAD_TEXT",
    -8: "NNP_ERR_CAPACITY",
    -9: "NNP_ERR_NO_DEVICE",
    -10: "NNP_ERR_NOT_INITIALISED",
    -11: "NNP_ERR_CUDA",
    -12: "NNP_ERR_WINDOW",
}
# statuses after which the reference prints a message and still leaves partial output behind
REFERENCE_ERRORS = (-1, -2, -3)

EXPORTS = [
    "nnp_init", "nnp_shutdown", "nnp_strerror", "nnp_set_stream", "nnp_last_cuda_error", "nnp_kernel_launches",
    "nnp_host_alloc", "nnp_host_free",
    "nnp_bin_to_binpack", "nnp_binpack_to_bin", "nnp_plain_to_binpack", "nnp_binpack_to_plain",
    "nnp_bin_to_plain", "nnp_plain_to_bin",
    "nnp_bin_to_binpack_dev", "nnp_binpack_to_bin_dev", "nnp_plain_to_binpack_dev",
    "nnp_binpack_to_plain_dev", "nnp_bin_to_plain_dev", "nnp_plain_to_bin_dev",
    "nnp_binpack_count_dev", "nnp_generate_bin_dev", "nnp_last_timing", "nnp_decode_stats",
    "nnp_debug_config",
    "nnp_shard_compress_begin_dev", "nnp_shard_compress_orbit", "nnp_shard_compress_emit_dev",
    "nnp_shard_compress_table_dev", "nnp_shard_compress_resolve_dev",
    "nnp_bin_to_binpack_file", "nnp_binpack_to_bin_file",
    "nnp_binpack_to_halfkp_dev", "nnp_bin_to_halfkp_dev",
    "nnp_init_all", "nnp_bind_device", "nnp_device_count", "nnp_device_at",
    "nnp_bin_to_binpack_multi", "nnp_binpack_to_bin_multi",
    "nnp_binpack_chunk_range", "nnp_binpack_chunk_range_dev", "nnp_shard_decompress_dev",
    "nnp_last_dominant_kernel", "nnp_last_positions",
]


class ShardInfo(ctypes.Structure):
    """nnp_shard_info (include/nnuepack.h)."""
    _fields_ = [("first_owned_record", ctypes.c_uint64), ("end_owned_record", ctypes.c_uint64),
                ("payload_bytes", ctypes.c_uint64), ("chains", ctypes.c_uint64), ("first_bad_record", ctypes.c_uint64)]


class ChunkRange(ctypes.Structure):
    """nnp_chunk_range (include/nnuepack.h)."""
    _fields_ = [("chunks_total", ctypes.c_uint64), ("chunk_lo", ctypes.c_uint64), ("chunk_hi", ctypes.c_uint64),
                ("byte_lo", ctypes.c_uint64), ("byte_hi", ctypes.c_uint64), ("positions", ctypes.c_uint64)]


NO_CARRY = (1 << 64) - 1


class NnpError(RuntimeError):
    def __init__(self, status: int, message: str, partial: bytes | None = None):
        super().__init__(f"{STATUS.get(status, status)}: {message}")
        self.status = status
        self.message = message
        self.partial = partial


_lib = None


def lib() -> ctypes.CDLL:
    """Loads libnnuepack.so (built in-tree by ``__graft_entry__.build()`` / ``csrc/Makefile``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). nnue_data_compress_b200 has no CPU fallback."
            )
        L = ctypes.CDLL(LIB_PATH)
        conv = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]
        for name in EXPORTS:
            fn = getattr(L, name)
            if name.endswith(("_to_binpack", "_to_bin", "_to_plain", "_dev", "_multi")) and not name.endswith("_file") and name not in (
                "nnp_binpack_count_dev",
                "nnp_generate_bin_dev",
                "nnp_shard_compress_begin_dev",
                "nnp_shard_compress_emit_dev",
                "nnp_shard_compress_table_dev",
                "nnp_shard_compress_resolve_dev",
                "nnp_binpack_to_halfkp_dev",
                "nnp_bin_to_halfkp_dev",
                "nnp_binpack_chunk_range_dev",
                "nnp_shard_decompress_dev",
            ):
                fn.argtypes = conv
                fn.restype = ctypes.c_int
        L.nnp_init.argtypes = [ctypes.c_int]
        L.nnp_init.restype = ctypes.c_int
        L.nnp_shutdown.restype = None
        L.nnp_set_stream.argtypes = [ctypes.c_void_p]
        L.nnp_set_stream.restype = ctypes.c_int
        L.nnp_strerror.argtypes = [ctypes.c_int]
        L.nnp_strerror.restype = ctypes.c_char_p
        L.nnp_last_cuda_error.restype = ctypes.c_char_p
        L.nnp_kernel_launches.restype = ctypes.c_uint64
        L.nnp_host_alloc.argtypes = [ctypes.c_size_t]
        L.nnp_host_alloc.restype = ctypes.c_void_p
        L.nnp_host_free.argtypes = [ctypes.c_void_p]
        L.nnp_host_free.restype = None
        L.nnp_binpack_count_dev.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint64)]
        L.nnp_binpack_count_dev.restype = ctypes.c_int
        L.nnp_generate_bin_dev.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_uint64]
        L.nnp_generate_bin_dev.restype = ctypes.c_int
        u64p = ctypes.POINTER(ctypes.c_uint64)
        L.nnp_shard_compress_begin_dev.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t,
                                                   ctypes.c_int, ctypes.POINTER(ShardInfo)]
        L.nnp_shard_compress_begin_dev.restype = ctypes.c_int
        L.nnp_shard_compress_orbit.argtypes = [ctypes.c_uint64, ctypes.c_uint64, u64p, u64p, u64p]
        L.nnp_shard_compress_orbit.restype = ctypes.c_int
        L.nnp_shard_compress_emit_dev.argtypes = [ctypes.c_uint64, ctypes.c_void_p, ctypes.c_size_t,
                                                  ctypes.POINTER(ctypes.c_size_t)]
        L.nnp_shard_compress_emit_dev.restype = ctypes.c_int
        for name in ("nnp_bin_to_binpack_file", "nnp_binpack_to_bin_file"):
            getattr(L, name).argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_size_t, u64p]
            getattr(L, name).restype = ctypes.c_int
        L.nnp_shard_compress_table_dev.argtypes = [ctypes.c_void_p]
        L.nnp_shard_compress_table_dev.restype = ctypes.c_int
        L.nnp_shard_compress_resolve_dev.argtypes = [ctypes.c_void_p, u64p, ctypes.c_int, ctypes.c_int, u64p, u64p, u64p, u64p]
        L.nnp_shard_compress_resolve_dev.restype = ctypes.c_int
        for name in ("nnp_binpack_to_halfkp_dev", "nnp_bin_to_halfkp_dev"):
            getattr(L, name).argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]
            getattr(L, name).restype = ctypes.c_int
        for name in ("nnp_init_all", "nnp_bind_device", "nnp_device_at"):
            getattr(L, name).argtypes = [ctypes.c_int]
            getattr(L, name).restype = ctypes.c_int
        L.nnp_device_count.restype = ctypes.c_int
        for name in ("nnp_binpack_chunk_range", "nnp_binpack_chunk_range_dev"):
            getattr(L, name).argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ChunkRange)]
            getattr(L, name).restype = ctypes.c_int
        L.nnp_shard_decompress_dev.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                               ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ChunkRange)]
        L.nnp_shard_decompress_dev.restype = ctypes.c_int
        L.nnp_debug_config.argtypes = [ctypes.c_char_p, ctypes.c_uint64]
        L.nnp_debug_config.restype = ctypes.c_int
        L.nnp_decode_stats.argtypes = [ctypes.POINTER(ctypes.c_uint64)]
        L.nnp_decode_stats.restype = ctypes.c_int
        L.nnp_last_timing.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]
        L.nnp_last_timing.restype = ctypes.c_int
        L.nnp_last_dominant_kernel.restype = ctypes.c_char_p
        L.nnp_last_positions.restype = ctypes.c_uint64
        _lib = L
    return _lib


_initialised = False


def init(device: int | None = None) -> None:
    """Binds the process to one GPU (default: LOCAL_RANK, else 0). Raises without a usable B200."""
    global _initialised
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    rc = lib().nnp_init(device)
    if rc != 0:
        raise NnpError(rc, lib().nnp_strerror(rc).decode() + " / " + lib().nnp_last_cuda_error().decode())
    _initialised = True


def init_all(n_devices: int = 0) -> int:
    """nnp_init_all: every visible GPU (or the first `n_devices`) in this one process; the slab-wise file
    drivers and the *_multi drivers then spread their slabs over all of them. Returns the device count."""
    global _initialised
    rc = lib().nnp_init_all(n_devices)
    if rc <= 0:
        raise NnpError(rc, lib().nnp_strerror(rc).decode() + " / " + lib().nnp_last_cuda_error().decode())
    _initialised = True
    return rc


def bin_to_binpack_multi(data: bytes) -> bytes:
    """compressBin over every initialised device (nnp_bin_to_binpack_multi)."""
    return _convert_host("nnp_bin_to_binpack_multi", data)


def binpack_to_bin_multi(data: bytes) -> bytes:
    """decompressBin over every initialised device (nnp_binpack_to_bin_multi)."""
    return _convert_host("nnp_binpack_to_bin_multi", data)


def convert_file(direction: str, input_path: str, output_path: str, append: bool = False, slab_bytes: int = 0) -> int:
    """File-to-file conversion in slabs (nnp_bin_to_binpack_file / nnp_binpack_to_bin_file): inputs of
    any size. `direction` is "bin_to_binpack" or "binpack_to_bin". Returns the number of positions; a
    reference error raises NnpError after the file holds what the reference would have left."""
    _ensure_init()
    fn = getattr(lib(), f"nnp_{direction}_file")
    n = ctypes.c_uint64(0)
    rc = fn(os.fsencode(input_path), os.fsencode(output_path), int(append), slab_bytes, ctypes.byref(n))
    if rc != 0:
        raise NnpError(rc, _strerror(rc))
    return int(n.value)


class ShardCalls:
    """The local calls of the sharded compressor (include/nnuepack.h) in the shape
    sharding.compress_sharded() expects: orbit / emit / table / resolve as bound methods. `d_out` is the
    device tensor the rank's slice is written to."""

    def __init__(self, d_out, device=None):
        self.d_out = d_out
        self.device = device

    @staticmethod
    def _check(rc):
        if rc != 0:
            raise NnpError(rc, _strerror(rc))

    def begin(self, d_records, n_records: int, own_lo: int, own_hi: int, reaches_eof: bool) -> ShardInfo:
        """nnp_shard_compress_begin_dev. Does not raise on a rank-local failure: `last_status` goes into
        sharding.compress_sharded(status=...), which lets all ranks fail together."""
        info = ShardInfo()
        self.last_status = lib().nnp_shard_compress_begin_dev(ctypes.c_void_p(d_records.data_ptr()), n_records, own_lo, own_hi,
                                                              int(reaches_eof), ctypes.byref(info))
        self.last_payload_bytes = int(info.payload_bytes)
        return info

    def orbit(self, base: int, carry: int):
        a, f, c = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
        self._check(lib().nnp_shard_compress_orbit(base, carry, ctypes.byref(a), ctypes.byref(f), ctypes.byref(c)))
        return a.value, f.value, c.value

    def emit(self, next_start: int) -> int:
        got = ctypes.c_size_t(0)
        self._check(lib().nnp_shard_compress_emit_dev(next_start, ctypes.c_void_p(self.d_out.data_ptr()), self.d_out.numel(),
                                                     ctypes.byref(got)))
        return got.value

    def table(self):
        import torch

        from .sharding import ORBIT_TABLE_ENTRIES

        t = torch.empty(3 * ORBIT_TABLE_ENTRIES, dtype=torch.int64, device=self.device or self.d_out.device)
        self._check(lib().nnp_shard_compress_table_dev(ctypes.c_void_p(t.data_ptr())))
        return t

    def resolve(self, tables, sizes, world: int, rank: int):
        arr = (ctypes.c_uint64 * world)(*[int(v) for v in sizes])
        c, b, nx, tot = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
        self._check(lib().nnp_shard_compress_resolve_dev(ctypes.c_void_p(tables.data_ptr()), arr, world, rank, ctypes.byref(c),
                                                        ctypes.byref(b), ctypes.byref(nx), ctypes.byref(tot)))
        return c.value, b.value, nx.value, tot.value


def shard_decompress(d_binpack, world: int, rank: int, d_out=None):
    """nnp_shard_decompress_dev: decodes rank `rank`'s chunk range of the ONE .binpack held in the CUDA
    uint8 tensor `d_binpack` into `d_out` (None: count only). Returns (bytes written, ChunkRange)."""
    _ensure_init()
    rng = ChunkRange()
    got = ctypes.c_size_t(0)
    rc = lib().nnp_shard_decompress_dev(ctypes.c_void_p(d_binpack.data_ptr() if d_binpack.numel() else 0), d_binpack.numel(),
                                        world, rank, ctypes.c_void_p(d_out.data_ptr()) if d_out is not None else None,
                                        d_out.numel() if d_out is not None else 0, ctypes.byref(got), ctypes.byref(rng))
    if rc != 0:
        raise NnpError(rc, _strerror(rc))
    return got.value, rng


def use_torch_stream() -> None:
    """Runs the library's kernels on torch's current CUDA stream, so that device tensors produced by
    torch ops are ordered before the library reads them (the library's own stream is non-blocking
    and does not wait for torch's). torch's default stream has the handle 0, which nnp_set_stream
    takes as "back to the library's stream"; it is passed as cudaStreamLegacy (0x1) instead."""
    import torch

    _ensure_init()
    handle = torch.cuda.current_stream().cuda_stream or 1
    rc = lib().nnp_set_stream(ctypes.c_void_p(handle))
    if rc != 0:
        raise NnpError(rc, _strerror(rc))


def shutdown() -> None:
    global _initialised
    if _lib is not None:
        _lib.nnp_shutdown()
    _initialised = False


def _ensure_init():
    if not _initialised:
        init()


def _strerror(rc: int) -> str:
    return lib().nnp_strerror(rc).decode()


def _convert_host(fn_name: str, data: bytes, allow_partial: bool = False) -> bytes:
    """Calls one of the six host-buffer drivers: capacity query, then the conversion."""
    _ensure_init()
    L = lib()
    fn = getattr(L, fn_name)
    src = (ctypes.c_char * len(data)).from_buffer_copy(data) if len(data) else (ctypes.c_char * 1)()
    need = ctypes.c_size_t(0)
    rc = fn(src, len(data), None, 0, ctypes.byref(need))
    if rc != 0:
        raise NnpError(rc, _strerror(rc))
    cap = max(int(need.value), 1)
    dst = (ctypes.c_char * cap)()
    out_n = ctypes.c_size_t(0)
    rc = fn(src, len(data), dst, cap, ctypes.byref(out_n))
    result = bytes(memoryview(dst)[: out_n.value]) if rc in (0,) + REFERENCE_ERRORS else b""
    if rc != 0:
        if allow_partial and rc in REFERENCE_ERRORS:
            raise NnpError(rc, _strerror(rc), partial=result)
        raise NnpError(rc, _strerror(rc), partial=result if rc in REFERENCE_ERRORS else None)
    return result


# -- the six drivers (compress_file.cpp:1246-1533), in-memory -------------------------------------

def bin_to_binpack(data: bytes) -> bytes:
    """compressBin (compress_file.cpp:1338-1374)."""
    return _convert_host("nnp_bin_to_binpack", data)


def binpack_to_bin(data: bytes) -> bytes:
    """decompressBin (compress_file.cpp:1376-1412)."""
    return _convert_host("nnp_binpack_to_bin", data)


def plain_to_binpack(data: bytes) -> bytes:
    """compressPlain (compress_file.cpp:1246-1297)."""
    return _convert_host("nnp_plain_to_binpack", data)


def binpack_to_plain(data: bytes) -> bytes:
    """decompressPlain (compress_file.cpp:1299-1335)."""
    return _convert_host("nnp_binpack_to_plain", data)


def bin_to_plain(data: bytes) -> bytes:
    """convertBinToPlain (compress_file.cpp:1414-1465)."""
    return _convert_host("nnp_bin_to_plain", data)


def plain_to_bin(data: bytes) -> bytes:
    """convertPlainToBin (compress_file.cpp:1467-1533)."""
    return _convert_host("nnp_plain_to_bin", data)


# -- convert() / run() / main(): the reference CLI (compress_file.cpp:1535-1709) ------------------

PLAIN_EXT, BIN_EXT, BINPACK_EXT = ".plain", ".bin", ".binpack"

_HELP = """Usage:
    nnue_data_compression [-h] [-a] input_path output_path

-h, --help                show help
-a, --append              append to the output file instead of truncating it

input_path                the path to the file to process
output_path               the path to the file to create/append to

Behaviour depends on file extensions. If the input
file has extension either .bin or .plain
it will be compressed. The output file has then an implied
extension of .binpack and it doesn't have to be specified.
If the input file's extension is .binpack then it will be decompressed
to either a .bin or .plain file, depending on the extension.

Example usage:
1. convert from plain to binpack in append mode:
    nnue_data_compression -a data.plain data
2. convert from binpack to plain in truncate/replace mode:
    nnue_data_compression data.binpack data.plain
"""


_FILE_DRIVERS = {"bin_to_binpack", "binpack_to_bin"}


def _run_driver(driver, label: str, input_path: str, output_path: str, append: bool, out=sys.stdout) -> None:
    out.write(f"{label} {input_path} to {output_path}\n")
    # like the native CLI: the headline directions go file to file, slab by slab, once the larger of
    # input and expected output (a .binpack expands about 24 times) exceeds NNP_STREAM_THRESHOLD bytes
    # (default 256 MiB), so that neither host nor device memory has to hold the whole file
    threshold = int(os.environ.get("NNP_STREAM_THRESHOLD", str(256 << 20)))
    weight = 24 if driver.__name__ == "binpack_to_bin" else 1
    if driver.__name__ in _FILE_DRIVERS and os.path.getsize(input_path) * weight > threshold:
        convert_file(driver.__name__, input_path, output_path, append, int(os.environ.get("NNP_SLAB_BYTES", "0")))
        return
    with open(input_path, "rb") as f:
        data = f.read()
    err = None
    try:
        result = driver(data)
    except NnpError as e:
        if e.status not in REFERENCE_ERRORS:
            raise
        result, err = e.partial or b"", e
    with open(output_path, "ab" if append else "wb") as f:
        f.write(result)
    if err is not None:
        raise err


def convert(input_path: str, output_path: str, append: bool = False, out=sys.stdout, err=sys.stderr) -> None:
    """convert() (compress_file.cpp:1593-1621): dispatch on the file extensions."""
    if not os.path.exists(input_path):
        err.write("Input file doesn't exist.\n")
        return
    if input_path.endswith(BIN_EXT) and output_path.endswith(PLAIN_EXT):
        _run_driver(bin_to_plain, "Converting", input_path, output_path, append, out)
    elif input_path.endswith(PLAIN_EXT) and output_path.endswith(BIN_EXT):
        _run_driver(plain_to_bin, "Compressing", input_path, output_path, append, out)
    elif input_path.endswith(PLAIN_EXT) or input_path.endswith(BIN_EXT):
        if not output_path.endswith(BINPACK_EXT):
            output_path += BINPACK_EXT
        drv = bin_to_binpack if input_path.endswith(BIN_EXT) else plain_to_binpack
        _run_driver(drv, "Compressing", input_path, output_path, append, out)
    elif input_path.endswith(BINPACK_EXT):
        if output_path.endswith(BIN_EXT):
            _run_driver(binpack_to_bin, "Decompressing", input_path, output_path, append, out)
        elif output_path.endswith(PLAIN_EXT):
            _run_driver(binpack_to_plain, "Decompressing", input_path, output_path, append, out)
        else:
            err.write("Unrecognized file format. Only .bin and .plain are supported for decompression.")
    else:
        err.write("Unsupported extension.")


def main(argv: list[str] | None = None, out=sys.stdout, err=sys.stderr) -> int:
    """main()/run()/readArgs() (compress_file.cpp:1648-1709), including its quirks: any argument
    starting with '-' is a flag named by the text after the FIRST dash, so ``--append`` is stored
    as ``-append`` and ignored; reference errors print their message and ``Exiting...`` with exit
    code 0."""
    argv = list(sys.argv[1:] if argv is None else argv)
    flags = {a[1:] for a in argv if a.startswith("-")}
    pos = [a for a in argv if not a.startswith("-")]
    if len(pos) == 0 or "help" in flags or "h" in flags:
        out.write(_HELP)
        return 0
    if len(pos) == 2:
        append = "a" in flags or "append" in flags
        try:
            convert(pos[0], pos[1], append, out, err)
        except NnpError as e:
            if e.status not in REFERENCE_ERRORS:
                raise
            err.write(e.message + "\nExiting...\n")
        return 0
    err.write("Invalid arguments.\n")
    out.write(_HELP)
    return 1


def decode_stats() -> dict:
    """Diagnostics of the .binpack decoder (nnp_decode_stats)."""
    a = (ctypes.c_uint64 * 14)()
    lib().nnp_decode_stats(a)
    return {"optimistic_hits": a[0], "optimistic_misses": a[1], "candidates": a[2], "tentative_positions": a[3],
            "violations": a[4], "false_candidates": a[5], "false_sample": [(v >> 32, v & 0xFFFFFFFF) for v in a[6:14]]}


HALFKP_ROW = 32
HALFKP_FEATURES = 41024


def halfkp_rows(d_data, kind: str = "binpack", positions: int | None = None):
    """HalfKP feature rows of a .binpack (or .bin) held in a CUDA uint8 tensor, produced on the device
    (nnp_binpack_to_halfkp_dev / nnp_bin_to_halfkp_dev). Returns
    (white [n, 32] int32, black [n, 32] int32, meta [n, 8] uint8 = nnp_halfkp_meta); torch only provides
    the device buffers. The library's stream must be ordered with the producer of `d_data`
    (use_torch_stream())."""
    import torch

    _ensure_init()
    fn = {"binpack": lib().nnp_binpack_to_halfkp_dev, "bin": lib().nnp_bin_to_halfkp_dev}[kind]
    src = ctypes.c_void_p(d_data.data_ptr() if d_data.numel() else 0)
    n = ctypes.c_size_t(0)
    if positions is None:
        rc = fn(src, d_data.numel(), None, None, None, 0, ctypes.byref(n))
        if rc != 0:
            raise NnpError(rc, _strerror(rc))
        positions = n.value
    rows = max(positions, 1)
    white = torch.empty((rows, HALFKP_ROW), dtype=torch.int32, device=d_data.device)
    black = torch.empty((rows, HALFKP_ROW), dtype=torch.int32, device=d_data.device)
    meta = torch.empty((rows, 8), dtype=torch.uint8, device=d_data.device)
    rc = fn(src, d_data.numel(), ctypes.c_void_p(white.data_ptr()), ctypes.c_void_p(black.data_ptr()),
            ctypes.c_void_p(meta.data_ptr()), positions, ctypes.byref(n))
    if rc != 0:
        raise NnpError(rc, _strerror(rc))
    return white[: n.value], black[: n.value], meta[: n.value]


def generate_bin(n_positions: int, max_plies: int = 100, seed: int = 42) -> bytes:
    """Synthetic .bin input generated on the device (SURVEY.md 8d recipe); torch only provides the
    device buffer and the copy back to the host."""
    import torch

    _ensure_init()
    buf = torch.empty(max(n_positions * 40, 8), dtype=torch.uint8, device="cuda")
    rc = lib().nnp_generate_bin_dev(ctypes.c_void_p(buf.data_ptr()), n_positions, max_plies, seed)
    if rc != 0:
        raise NnpError(rc, _strerror(rc))
    return buf[: n_positions * 40].cpu().numpy().tobytes()
