"""GPU suite: the native command line (nnue_data_compress_b200/nnue_data_compression, csrc/cli.cpp)
run as a process on files, as a user of the reference tool would: same dispatch by extension, same
bytes, `-a` appends, reference errors print their message and `Exiting...` with exit code 0."""
import os
import subprocess

import pytest

from refutil import BIN_TO_BINPACK, ROOT, golden, have_ref, ref_convert

pytestmark = pytest.mark.gpu

CLI = os.path.join(ROOT, "nnue_data_compress_b200", "nnue_data_compression")


def _run(*args):
    return subprocess.run([CLI, *args], capture_output=True, text=True, timeout=300)


def test_cli_all_six_directions(tmp_path):
    d = tmp_path
    (d / "g.bin").write_bytes(golden("games100.bin"))
    r = _run(str(d / "g.bin"), str(d / "g"))  # implied .binpack extension
    assert r.returncode == 0 and "Compressing" in r.stdout
    assert (d / "g.binpack").read_bytes() == golden("games100.binpack")
    assert _run(str(d / "g.binpack"), str(d / "rt.bin")).returncode == 0
    assert (d / "rt.bin").read_bytes() == golden("games100.rt.bin")
    assert _run(str(d / "g.binpack"), str(d / "g.plain")).returncode == 0
    assert (d / "g.plain").read_bytes() == golden("games100.plain")
    assert _run(str(d / "g.plain"), str(d / "p.binpack")).returncode == 0
    assert (d / "p.binpack").read_bytes() == golden("games100.p.binpack")
    assert _run(str(d / "g.bin"), str(d / "b.plain")).returncode == 0
    assert (d / "b.plain").read_bytes() == golden("games100.b.plain")
    assert _run(str(d / "b.plain"), str(d / "p.bin")).returncode == 0
    assert (d / "p.bin").read_bytes() == golden("games100.p.bin")


def test_cli_append_and_errors(tmp_path):
    d = tmp_path
    (d / "a.bin").write_bytes(golden("heads.bin"))
    (d / "b.bin").write_bytes(golden("long400.bin"))
    assert _run(str(d / "a.bin"), str(d / "out.binpack")).returncode == 0
    assert _run("-a", str(d / "b.bin"), str(d / "out.binpack")).returncode == 0
    assert (d / "out.binpack").read_bytes() == golden("heads.binpack") + golden("long400.binpack")
    # "--append" is stored as "-append" by the reference's readArgs and ignored: the file is replaced
    assert _run("--append", str(d / "a.bin"), str(d / "out.binpack")).returncode == 0
    assert (d / "out.binpack").read_bytes() == golden("heads.binpack")
    bp = bytearray(golden("twochunks.binpack"))
    second = 8 + int.from_bytes(bp[4:8], "little")
    bp[second:second + 4] = b"BINX"
    (d / "bad.binpack").write_bytes(bytes(bp))
    r = _run(str(d / "bad.binpack"), str(d / "bad.bin"))
    assert r.returncode == 0 and "Invalid binpack file or chunk." in r.stderr and "Exiting..." in r.stderr
    if have_ref():
        from refutil import BINPACK_TO_BIN

        assert (d / "bad.bin").read_bytes() == ref_convert(BINPACK_TO_BIN, bytes(bp))
    r = _run(str(d / "missing.bin"), str(d / "x"))
    assert "Input file doesn't exist." in r.stderr and r.returncode == 0
    assert _run("only_one_argument").returncode == 1


def test_cli_streams_large_files(tmp_path, monkeypatch):
    """Above NNP_STREAM_THRESHOLD the CLI converts slab by slab (nnp_*_file); same bytes."""
    d = tmp_path
    (d / "g.bin").write_bytes(golden("twochunks.bin"))
    env = dict(os.environ, NNP_STREAM_THRESHOLD="1000", NNP_SLAB_BYTES=str(40 * 7000))
    r = subprocess.run([CLI, str(d / "g.bin"), str(d / "g.binpack")], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and r.stdout.startswith("Compressing")
    assert (d / "g.binpack").read_bytes() == golden("twochunks.binpack")
    env["NNP_SLAB_BYTES"] = str(700_000)
    r = subprocess.run([CLI, str(d / "g.binpack"), str(d / "rt.bin")], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and "positions" in r.stdout
    assert (d / "rt.bin").read_bytes() == golden("twochunks.rt.bin")


def test_cli_prints_the_reference_start_and_end_lines(tmp_path, nnp):
    """stdout: the reference's first line and the LAST of its progress lines (compress_file.cpp:1369-1372,
    :1395-1410); the lines in between are not reproduced. Whole-buffer and slab-wise drivers alike."""
    from refutil import REF_BIN

    if not have_ref():
        pytest.skip("oracle/_ref not built")
    d = tmp_path
    (d / "g.bin").write_bytes(nnp.generate_bin(250_000, 100, 5))
    (d / "s.bin").write_bytes(golden("games100.bin"))
    for stream in ("1", str(1 << 40)):
        env = dict(os.environ, NNP_STREAM_THRESHOLD=stream)
        for stem in ("g", "s"):
            for args in ((str(d / f"{stem}.bin"), str(d / f"{stem}.binpack")), (str(d / f"{stem}.binpack"), str(d / f"{stem}.rt.bin"))):
                ours = subprocess.run([CLI, *args], capture_output=True, text=True, timeout=300, env=env).stdout.strip().split("\n")
                ref = subprocess.run([REF_BIN, *args], capture_output=True, text=True, timeout=300).stdout.strip().split("\n")
                assert ours == ref[:1] + (ref[-1:] if len(ref) > 1 else []), (ours, ref[:1], ref[-1:])
