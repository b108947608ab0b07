"""GPU parity tests of the rows added after the first round-1 measurements (they run after the core parity tests):
LZX DELTA units and MSZIP folders with short blocks in the middle (the reference's ring window),
through the C-ABI against the reference's own decoders."""
import numpy as np
import pytest

from libmspack_b200 import gen
from libmspack_b200.units import CODEC_LZX, CODEC_MSZIP, CODEC_QUANTUM
from util import RING_CASES, ring_batch
from test_gpu_parity import _check_batch

pytestmark = pytest.mark.gpu


def test_mszip_short_blocks_ring_window(decoder, oracle_ref):
    """MSZIP folders with short blocks in the middle (k_p2_ring): bit-exact with the reference's ring window; mixed with ordinary
    MSZIP units in the same batch."""
    b, raws = ring_batch(RING_CASES[:-1])
    m = gen.concat_batches([b, gen.make_batch(CODEC_MSZIP, 64, unit_bytes=65536), gen.make_batch(CODEC_MSZIP, 64, first_unit=500)])
    out, st = _check_batch(decoder, oracle_ref, m, "mszip ring")
    assert (st == 0).all()
    for i, r in enumerate(raws):
        assert b.unit_output(out, i).tobytes() == r


LZX_DELTA_CASES = [dict(window_bits=17), dict(window_bits=17, ref_bytes=20000), dict(window_bits=22, ref_bytes=100000, unit_bytes=100000, block_mode=4, split=2),
                   dict(window_bits=25, unit_bytes=70000, ref_bytes=50000, data="binary", intel=1), dict(window_bits=18, data="zeros", unit_bytes=65536),
                   dict(window_bits=17, unit_bytes=196608, ref_bytes=131072, block_mode=4, block_frames=2), dict(window_bits=17, unit_bytes=196685, ref_bytes=1000, block_mode=3),
                   dict(window_bits=17, unit_bytes=327680, ref_bytes=70000, block_mode=4), dict(window_bits=17, unit_bytes=131072, ref_bytes=1000, reset_interval=1, block_mode=4, slack=4)]


@pytest.mark.parametrize("case", LZX_DELTA_CASES, ids=lambda c: ",".join(f"{k}={v}" for k, v in c.items()))
def test_lzx_delta_batches(decoder, oracle_ref, case):
    """LZX DELTA units (lzxd_init(is_delta=1) + lzxd_set_reference_data): chunk sizes, matches longer than 257 bytes, matches into
    the reference data, windows up to 2^25, units longer than the window; intact and corrupted; and mixed into a batch of plain units."""
    b = gen.make_batch(CODEC_LZX, 48, delta=1, **case)
    _check_batch(decoder, oracle_ref, b, f"lzx delta {case}")
    rng = np.random.default_rng(17)
    comp, units = b.comp.copy(), b.units.copy()
    for i, u in enumerate(units):
        lo, n = int(u["in_off"]), int(u["in_len"])
        if i % 2 == 0:
            comp[lo + int(rng.integers(0, n))] ^= 1 << int(rng.integers(0, 8))
        else:
            units["in_len"][i] = max(1, n - int(rng.integers(1, 40)))
    c = gen.Batch(units, comp, None, b.out_bytes, b.out_init)
    _check_batch(decoder, oracle_ref, c, f"corrupt lzx delta {case}")


def test_lzx_delta_mixed_with_plain_units(decoder, oracle_ref):
    parts = [gen.make_batch(CODEC_LZX, 40, delta=1, window_bits=20, ref_bytes=40000, unit_bytes=50000), gen.make_batch(CODEC_LZX, 40, block_mode=4, split=2, first_unit=100),
             gen.make_batch(CODEC_MSZIP, 40, first_unit=200), gen.make_batch(CODEC_QUANTUM, 40, first_unit=300)]
    m = gen.concat_batches(parts)
    perm = np.random.default_rng(5).permutation(m.n)
    m.units = m.units[perm].copy()
    _check_batch(decoder, oracle_ref, m, "delta mixed")
