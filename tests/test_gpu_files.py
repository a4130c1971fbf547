"""GPU suite: file-to-file conversion in slabs (nnp_bin_to_binpack_file / nnp_binpack_to_bin_file).
Tiny slabs force many slabs on small inputs; the files must be what a single whole-file run writes."""
import pytest

from refutil import BIN_TO_BINPACK, BINPACK_TO_BIN, golden, oracle_convert

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,slab", [("games100", 40 * 64), ("long400", 40 * 100), ("heads", 40 * 16), ("restart", 40 * 37),
                                       ("twochunks", 40 * 3000), ("twochunks", 0)])
def test_bin_to_binpack_file_in_slabs(nnp, tmp_path, name, slab):
    src, dst = tmp_path / "in.bin", tmp_path / "out.binpack"
    src.write_bytes(golden(name + ".bin") + b"\x01\x02\x03")  # a short trailing record is dropped
    n = nnp.convert_file("bin_to_binpack", str(src), str(dst), slab_bytes=slab)
    assert n == len(golden(name + ".bin")) // 40
    assert dst.read_bytes() == golden(name + ".binpack")


def test_files_many_chunks_and_append(nnp, tmp_path):
    b = nnp.generate_bin(1_200_000, 2, 5)  # ~20 chunks
    rc, want = oracle_convert(BIN_TO_BINPACK, b)
    assert rc == 0 and want.count(b"BINP") >= 15
    src, dst, back = tmp_path / "in.bin", tmp_path / "out.binpack", tmp_path / "back.bin"
    src.write_bytes(b)
    dst.write_bytes(golden("heads.binpack"))
    nnp.convert_file("bin_to_binpack", str(src), str(dst), append=True, slab_bytes=40 * 100_000)
    assert dst.read_bytes() == golden("heads.binpack") + want
    dst.write_bytes(want)
    n = nnp.convert_file("binpack_to_bin", str(dst), str(back), slab_bytes=3 << 20)
    rc, want_bin = oracle_convert(BINPACK_TO_BIN, want)
    assert rc == 0 and n == len(want_bin) // 40
    assert back.read_bytes() == want_bin
    nnp.convert_file("binpack_to_bin", str(dst), str(back), append=True)
    assert back.read_bytes() == want_bin + want_bin


def test_files_reference_errors(nnp, tmp_path):
    good = golden("games100.bin")[:4000]
    bits = [0] + [0] * 6 + [1, 0, 0, 0, 0, 0] + [1, 0, 0, 0, 0] * 49
    sfen = bytearray(32)
    for i, v in enumerate(bits[:256]):
        sfen[i // 8] |= v << (i & 7)
    bad = bytes(sfen) + bytes(8)
    src, dst = tmp_path / "in.bin", tmp_path / "out.binpack"
    src.write_bytes(good + bad + good)
    with pytest.raises(nnp.NnpError) as ei:
        nnp.convert_file("bin_to_binpack", str(src), str(dst), slab_bytes=40 * 30)
    assert ei.value.status == -3
    assert dst.read_bytes() == oracle_convert(BIN_TO_BINPACK, good)[1]
    bp = golden("twochunks.binpack")
    second = 8 + int.from_bytes(bp[4:8], "little")
    data = bp[:second] + b"BINX" + bp[second + 4:]
    rc, want = oracle_convert(BINPACK_TO_BIN, data)
    assert rc == -1
    (tmp_path / "bad.binpack").write_bytes(data)
    with pytest.raises(nnp.NnpError) as ei:
        nnp.convert_file("binpack_to_bin", str(tmp_path / "bad.binpack"), str(tmp_path / "bad.bin"), slab_bytes=1 << 19)
    assert ei.value.status == -1
    assert (tmp_path / "bad.bin").read_bytes() == want


@pytest.mark.parametrize("name", ["games100", "twochunks", "restart", "heads"])
def test_multi_drivers_on_host_buffers(nnp, name):
    """nnp_*_multi: the slab pipeline between host buffers (all initialised devices; one here unless the
    box has more) gives the bytes of the whole-buffer drivers."""
    assert nnp.bin_to_binpack_multi(golden(name + ".bin")) == golden(name + ".binpack")
    assert nnp.binpack_to_bin_multi(golden(name + ".binpack")) == golden(name + ".rt.bin")
    assert nnp.bin_to_binpack_multi(b"") == b"" and nnp.binpack_to_bin_multi(b"") == b""


def test_multi_drivers_reference_errors(nnp):
    bp = golden("twochunks.binpack")
    second = 8 + int.from_bytes(bp[4:8], "little")
    data = bp[:second] + b"BINX" + bp[second + 4:]
    rc, want = oracle_convert(BINPACK_TO_BIN, data)
    assert rc == -1
    with pytest.raises(nnp.NnpError) as ei:
        nnp.binpack_to_bin_multi(data)
    assert ei.value.status == -1 and ei.value.partial == want


def test_file_pipeline_on_all_devices(tmp_path):
    """One process, every GPU of the box (nnp_init_all): slabs are dealt to the devices, the chunk-flush
    rule is replayed in file order in host memory; the files are the single-run files. A separate
    process, because the suite's own process is bound to one device."""
    import subprocess
    import sys

    from refutil import ROOT

    code = """
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r + "/tests")
import nnue_data_compress_b200 as nnp
from refutil import BIN_TO_BINPACK, BINPACK_TO_BIN, oracle_convert
n_dev = nnp.init_all(0)
b = nnp.generate_bin(1_500_000, 100, 99)
rc, want = oracle_convert(BIN_TO_BINPACK, b)
assert rc == 0
d = %r
open(d + "/in.bin", "wb").write(b)
n = nnp.convert_file("bin_to_binpack", d + "/in.bin", d + "/out.binpack", slab_bytes=40 * 130_000)
assert n == 1_500_000 and open(d + "/out.binpack", "rb").read() == want
rc, want_bin = oracle_convert(BINPACK_TO_BIN, want)
n = nnp.convert_file("binpack_to_bin", d + "/out.binpack", d + "/back.bin", slab_bytes=1 << 20)
assert n == len(want_bin) // 40 and open(d + "/back.bin", "rb").read() == want_bin
assert nnp.bin_to_binpack_multi(b) == want
print("devices", n_dev, "ok")
""" % (ROOT, ROOT, str(tmp_path))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr
