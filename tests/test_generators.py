"""CPU-only: the workload generators (the reference has no encoders) round-trip through the oracle."""
import numpy as np
import pytest

from libmspack_b200 import gen
from libmspack_b200.units import CODEC_LZX, CODEC_MSZIP, CODEC_QUANTUM


@pytest.mark.parametrize("codec,kw", [(CODEC_MSZIP, dict()), (CODEC_MSZIP, dict(unit_bytes=100000, data="binary")), (CODEC_QUANTUM, dict(window_bits=10)),
                                      (CODEC_QUANTUM, dict(unit_bytes=65536)), (CODEC_LZX, dict()), (CODEC_LZX, dict(block_mode=4, split=4, unit_bytes=98304)),
                                      (CODEC_LZX, dict(block_mode=3)), (CODEC_LZX, dict(window_bits=16, unit_bytes=65536, block_frames=2))],
                         ids=lambda x: str(x))
def test_roundtrip(oracle_ref, codec, kw):
    b = gen.make_batch(codec, 32, keep_raw=True, **kw)
    out, st, _ = oracle_ref.decode_batch(b.units, b.comp, b.out_bytes, threads=4)
    assert (st == 0).all()
    ub = int(b.units["out_len"][0]); stride = (ub + 15) & ~15
    assert np.array_equal(out.reshape(-1, stride)[:, :ub].reshape(-1), b.raw)


@pytest.mark.parametrize("kw", [dict(window_bits=17), dict(window_bits=19, ref_bytes=30000, unit_bytes=98304 + 5, block_mode=4, block_frames=2),
                                dict(window_bits=24, ref_bytes=65536, data="zeros"), dict(window_bits=17, ref_bytes=131072, unit_bytes=3 * 131072, block_mode=3)],
                         ids=lambda x: str(x))
def test_roundtrip_lzx_delta(oracle_ref, kw):
    """The LZX DELTA encoder (chunk sizes, long matches, matches into the reference data) through the reference decoder."""
    b = gen.make_batch(CODEC_LZX, 8, keep_raw=True, delta=1, **kw)
    out, st, _ = oracle_ref.decode_batch(b.units, b.comp, b.out_bytes, threads=4, out_init=b.out_init)
    assert (st == 0).all()
    ub = int(b.units["out_len"][0])
    for i in range(b.n):
        assert np.array_equal(b.unit_output(out, i), b.raw[i * ub:(i + 1) * ub])
    if kw.get("ref_bytes") and kw.get("data") != "zeros":
        # the reference data is really used: without it the same streams do not decode to the same bytes
        out2, st2, _ = oracle_ref.decode_batch(b.units, b.comp, b.out_bytes, threads=4)
        assert not np.array_equal(out2, out)


def test_corpus_is_deterministic_and_block_independent():
    a = gen.raw_units(8, 32768, first_unit=3)
    b = gen.raw_units(4, 32768, first_unit=5)
    assert np.array_equal(a[2 * 32768:6 * 32768], b)
