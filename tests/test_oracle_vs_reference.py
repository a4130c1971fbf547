"""CPU suite, only where oracle/_ref exists (the build container, and GPU boxes the prebuilt
binaries travelled to): the oracle against the compiled reference itself, all six directions,
on freshly generated inputs. This is the live version of the golden-vector pinning."""
import pytest

from refutil import (BIN_TO_BINPACK, BIN_TO_PLAIN, BINPACK_TO_BIN, BINPACK_TO_PLAIN, PLAIN_TO_BIN, PLAIN_TO_BINPACK,
                     have_ref, oracle_convert, ref_convert, ref_generate)

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")


@pytest.mark.parametrize("n,plies,seed,mode", [(60_000, 100, 101, 0), (30_000, 1, 102, 0), (40_000, 400, 103, 0),
                                               (20_000, 100, 104, 1), (20_000, 50, 105, 2)])
def test_six_directions(n, plies, seed, mode):
    b = ref_generate(n, plies, seed, mode)
    bp = ref_convert(BIN_TO_BINPACK, b)
    assert oracle_convert(BIN_TO_BINPACK, b) == (0, bp)
    assert oracle_convert(BINPACK_TO_BIN, bp) == (0, ref_convert(BINPACK_TO_BIN, bp))
    pl = ref_convert(BINPACK_TO_PLAIN, bp)
    assert oracle_convert(BINPACK_TO_PLAIN, bp) == (0, pl)
    assert oracle_convert(PLAIN_TO_BINPACK, pl) == (0, ref_convert(PLAIN_TO_BINPACK, pl))
    bpl = ref_convert(BIN_TO_PLAIN, b)
    assert oracle_convert(BIN_TO_PLAIN, b) == (0, bpl)
    assert oracle_convert(PLAIN_TO_BIN, bpl) == (0, ref_convert(PLAIN_TO_BIN, bpl))


def test_append_mode_is_concatenation():
    """-a: BINP chunks are self-delimiting, so appending shard outputs is a valid binpack that
    decodes to the concatenated records (the multi-GPU output model)."""
    a = ref_generate(5_000, 100, 201)
    b = ref_generate(7_000, 100, 202)
    pa = ref_convert(BIN_TO_BINPACK, a)
    both = ref_convert(BIN_TO_BINPACK, b, append_to=pa)
    assert both == pa + ref_convert(BIN_TO_BINPACK, b)
    assert oracle_convert(BINPACK_TO_BIN, both)[1] == ref_convert(BINPACK_TO_BIN, pa) + ref_convert(BINPACK_TO_BIN, both[len(pa):])
