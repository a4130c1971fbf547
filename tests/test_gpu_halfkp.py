"""GPU suite (-m gpu): HalfKP feature rows produced on the device (nnp_binpack_to_halfkp_dev,
nnp_bin_to_halfkp_dev) against the oracle: rows(binpack) must equal the oracle's rows
of the .bin file the oracle decodes from the same binpack. Integer work: equality is exact. The order
of a row's entries follows the source (include/nnuepack.h: stream order for .bin records, slot-stable
along a chain), the oracle's is (kind, square): the (white, black) pairs are sorted by the white index
before comparing -- which also checks that slot j of both rows describes the same piece."""
import numpy as np
import pytest
import torch

from refutil import BINPACK_TO_BIN, GOLDEN_SETS, golden, halfkp_from_fen, have_ref, oracle_convert, oracle_halfkp, ref_generate

pytestmark = pytest.mark.gpu


def _dev(data):
    return torch.frombuffer(bytearray(data) if data else bytearray(8), dtype=torch.uint8)[: len(data)].cuda()


def _sorted_pairs(w, k):
    key = np.where(w < 0, np.iinfo(np.int32).max, w)
    order = np.argsort(key, axis=1, kind="stable")
    return np.take_along_axis(w, order, 1), np.take_along_axis(k, order, 1)


def _check(nnp, binpack):
    rc, want_bin = oracle_convert(BINPACK_TO_BIN, binpack)
    assert rc == 0
    rc, white, black, meta, _ = oracle_halfkp(want_bin)
    assert rc == 0
    nnp.use_torch_stream()
    for kind, data in (("binpack", binpack), ("bin", want_bin)):
        w, k, m = nnp.halfkp_rows(_dev(data), kind)
        torch.cuda.synchronize()
        assert w.shape == (len(want_bin) // 40, 32)
        w, k = w.cpu().numpy(), k.cpu().numpy()
        assert ((w >= 0).sum(axis=1) == meta[:, 6]).all() and ((w >= 0) == (k >= 0)).all(), kind
        assert (np.diff((w >= 0).astype(np.int8), axis=1) <= 0).all(), kind  # padding only at the end
        w, k = _sorted_pairs(w, k)
        assert np.array_equal(w, white), kind
        assert np.array_equal(k, black), kind
        assert np.array_equal(m.cpu().numpy(), meta), kind


@pytest.mark.parametrize("name", GOLDEN_SETS)
def test_golden_rows(nnp, name):
    _check(nnp, golden(name + ".binpack"))


def test_kat_by_hand(nnp):
    nnp.use_torch_stream()
    w, k, m = nnp.halfkp_rows(_dev(golden("kat.binpack")))
    assert w.shape[0] == 3
    start = "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1"
    ww, kk = halfkp_from_fen(start)
    assert w[0].tolist() == ww + [-1, -1] and k[0].tolist() == kk + [-1, -1]
    assert m[0].tolist() == [25, 0, 0, 0, 1, 0, 30, 0]
    assert m[1].tolist()[4:7] == [255, 1, 30]


@pytest.mark.parametrize("n,plies,seed", [(200_000, 100, 42), (60_000, 1, 7), (100_000, 400, 9)])
def test_synthetic_rows(nnp, n, plies, seed):
    if not have_ref():
        pytest.skip("oracle/_ref did not travel to this box")
    _check(nnp, nnp.bin_to_binpack(ref_generate(n, plies, seed)))


def test_both_decode_strategies_and_count(nnp):
    import ctypes

    binpack = golden("twochunks.binpack")
    rc, want_bin = oracle_convert(BINPACK_TO_BIN, binpack)
    _, white, black, meta, _ = oracle_halfkp(want_bin)
    nnp.use_torch_stream()
    d = _dev(binpack)
    n = ctypes.c_size_t(0)
    assert nnp.lib().nnp_binpack_to_halfkp_dev(ctypes.c_void_p(d.data_ptr()), d.numel(), None, None, None, 0,
                                               ctypes.byref(n)) == 0
    assert n.value == len(want_bin) // 40
    for exhaustive in (1, 0):
        assert nnp.lib().nnp_debug_config(b"exhaustive", exhaustive) == 0
        try:
            w, k, m = nnp.halfkp_rows(d)
        finally:
            nnp.lib().nnp_debug_config(b"exhaustive", 0)
        w, k = _sorted_pairs(w.cpu().numpy(), k.cpu().numpy())
        assert np.array_equal(w, white) and np.array_equal(k, black)
        assert np.array_equal(m.cpu().numpy(), meta)
    # capacity too small: the needed count comes back
    w = torch.empty((4, 32), dtype=torch.int32, device="cuda")
    k = torch.empty((4, 32), dtype=torch.int32, device="cuda")
    m = torch.empty((4, 8), dtype=torch.uint8, device="cuda")
    rc = nnp.lib().nnp_binpack_to_halfkp_dev(ctypes.c_void_p(d.data_ptr()), d.numel(), ctypes.c_void_p(w.data_ptr()),
                                             ctypes.c_void_p(k.data_ptr()), ctypes.c_void_p(m.data_ptr()), 4, ctypes.byref(n))
    assert nnp.STATUS.get(rc) == "NNP_ERR_CAPACITY", rc
    assert n.value == len(want_bin) // 40


def test_errors_and_empty(nnp):
    nnp.use_torch_stream()
    w, k, m = nnp.halfkp_rows(_dev(b""))
    assert w.shape[0] == 0
    w, k, m = nnp.halfkp_rows(_dev(b""), "bin")
    assert w.shape[0] == 0
    bad = bytearray(golden("games100.binpack"))
    bad[0] = 0x43
    with pytest.raises(nnp.NnpError) as e:
        nnp.halfkp_rows(_dev(bytes(bad)))
    assert e.value.status == -1
    b = bytearray(golden("games100.bin")[: 40 * 50])
    b[40 * 20: 40 * 20 + 32] = b"\xff" * 32
    with pytest.raises(nnp.NnpError) as e:
        nnp.halfkp_rows(_dev(bytes(b)), "bin", positions=50)
    assert e.value.status == -3


def test_sequential_fallback_rows(nnp):
    """reject_mod drops a pseudo-random subset of chain-start candidates: the optimistic walk fails and
    the affected chunks go through k_slow_emit_halfkp; the rows must not change."""
    binpack = golden("twochunks.binpack")
    rc, want_bin = oracle_convert(BINPACK_TO_BIN, binpack)
    _, white, black, meta, _ = oracle_halfkp(want_bin)
    nnp.use_torch_stream()
    assert nnp.lib().nnp_debug_config(b"reject_mod", 5) == 0
    try:
        w, k, m = nnp.halfkp_rows(_dev(binpack))
    finally:
        nnp.lib().nnp_debug_config(b"reject_mod", 0)
    # the sequential kernel rebuilds every row: (kind, square) order, no sorting needed for its chunks
    w, k = _sorted_pairs(w.cpu().numpy(), k.cpu().numpy())
    assert np.array_equal(w, white) and np.array_equal(k, black) and np.array_equal(m.cpu().numpy(), meta)


def test_corrupted_movetext_rows_follow_the_walker(nnp):
    """Bit flips in the movetext produce moves no legal game contains (pieces appearing from nothing,
    kings captured, castling through pieces): the in-place row update must fall back to rebuilding the
    row exactly there. The reference here is tests/host_sim (rows rebuilt per position from the same
    chain walker on the CPU), because a position without a king has no faithful .bin record for the
    oracle to start from; the walker itself is pinned by the .bin parity tests on the same inputs."""
    import random

    from refutil import host_sim

    sim = host_sim()
    rng = random.Random(414)
    packs = [golden(n + ".binpack") for n in ("games100", "long400", "restart")]
    nnp.use_torch_stream()
    compared = 0
    for _ in range(300):
        b = bytearray(rng.choice(packs))
        for _ in range(rng.randrange(1, 5)):
            b[rng.randrange(8, len(b))] ^= 1 << rng.randrange(8)
        b = bytes(b)
        try:
            w, k, m = nnp.halfkp_rows(_dev(b))
        except nnp.NnpError as e:
            assert e.status == -4, e.status
            continue
        n = w.shape[0]
        white = np.empty((n + 1, 32), dtype=np.int32)
        black = np.empty((n + 1, 32), dtype=np.int32)
        assert sim.sim_halfkp_rows(b, len(b), white.ctypes.data, black.ctypes.data, n + 1) == n
        w, k = _sorted_pairs(w.cpu().numpy(), k.cpu().numpy())
        # the walker's rows are in (kind, square) order = ascending white index, like the sorted pairs
        assert np.array_equal(w, white[:n]) and np.array_equal(k, black[:n])
        compared += 1
    assert compared > 50
