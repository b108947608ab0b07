"""The experimental kernel shapes (MSGPU_LZX_VARIANT 31-53, MSGPU_ZIP_VARIANT 15-19, MSGPU_QTM_VARIANT 1-4, 7,
MSGPU_P2_VARIANT 1) against the reference's decoders on the GPU: intact, damaged and truncated units, unaligned inputs.
They are not defaults and have not run on a B200 yet, so this file only runs when MSGPU_TEST_EXPERIMENTAL=1 is set
(tools/r2_first_call.sh does, under its own timeout) - the default gpu tier never launches an unmeasured kernel."""
import os

import numpy as np
import pytest

from libmspack_b200 import gen
from libmspack_b200.units import CODEC_LZX, CODEC_MSZIP, CODEC_QUANTUM
from util import assert_same

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("MSGPU_TEST_EXPERIMENTAL") != "1", reason="set MSGPU_TEST_EXPERIMENTAL=1")]

SHAPES = [("MSGPU_LZX_VARIANT", v, CODEC_LZX) for v in list(range(31, 40)) + list(range(40, 54))] + [("MSGPU_ZIP_VARIANT", v, CODEC_MSZIP) for v in (15, 16, 17, 18, 19)] + \
         [("MSGPU_QTM_VARIANT", v, CODEC_QUANTUM) for v in (1, 2, 3, 4, 7)] + [("MSGPU_P2_VARIANT", 1, CODEC_LZX), ("MSGPU_P2_VARIANT", 1, CODEC_MSZIP)]
CASES = {CODEC_LZX: [dict(), dict(block_mode=4, split=3), dict(unit_bytes=65536, reset_interval=2, block_mode=4), dict(intel=1, data="binary", unit_bytes=70000, block_mode=2),
                     dict(window_bits=15, unit_bytes=100000, block_mode=4), dict(data="random"), dict(data="zeros")],
         CODEC_MSZIP: [dict(), dict(data="random", unit_bytes=40000), dict(unit_bytes=65536, level=1), dict(data="zeros"), dict(unit_bytes=100000, data="binary")],
         CODEC_QUANTUM: [dict(), dict(window_bits=10, unit_bytes=100000), dict(window_bits=12, unit_bytes=65536, data="binary"), dict(window_bits=21, unit_bytes=200000)]}


@pytest.mark.parametrize("env,value,codec", SHAPES, ids=lambda x: str(x))
def test_experimental_shape(oracle_ref, env, value, codec):
    from libmspack_b200.codec import BatchDecoder
    old = os.environ.get(env)
    os.environ[env] = str(value)
    try:
        dec = BatchDecoder(0)            # the variant is read when the context is created
    finally:
        if old is None:
            del os.environ[env]
        else:
            os.environ[env] = old
    try:
        rng = np.random.default_rng(value)
        for kw in CASES[codec]:
            b = gen.make_batch(codec, 96, **kw)
            for shift in (0, 1):
                comp = np.concatenate([np.zeros(shift, np.uint8), b.comp]) if shift else b.comp
                units = b.units.copy()
                units["in_off"] += shift
                for damaged in (False, True):
                    c2, u2 = comp.copy(), units.copy()
                    if damaged:
                        for i, u in enumerate(u2):
                            lo, n = int(u["in_off"]), int(u["in_len"])
                            if i % 2 == 0:
                                c2[lo + int(rng.integers(0, n))] ^= 1 << int(rng.integers(0, 8))
                            else:
                                u2["in_len"][i] = max(1, n - int(rng.integers(1, 40)))
                    og, sg = dec.decode_host(u2, c2, b.out_bytes)
                    oo, so, _ = oracle_ref.decode_batch(u2, c2, b.out_bytes, threads=8)
                    assert_same(u2, oo, so, og, sg, f"{env}={value} {kw} shift {shift} damaged {damaged}")
    finally:
        dec.close()
