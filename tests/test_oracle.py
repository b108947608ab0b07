"""CPU-only: the oracles themselves.  The reference oracle (oracle/_ref/libmspack_ref.so = the reference's own
lzxd.c / qtmd.c / mszipd.c) is pinned against the MD5s the reference's tests assert; the plain-C port is pinned
against the reference oracle and against the committed golden vectors."""
import hashlib
import os

import numpy as np
import pytest

from libmspack_b200 import gen
from libmspack_b200.units import CODEC_LZX, CODEC_MSZIP, CODEC_QUANTUM
from util import assert_same, golden_expected, golden_manifest, golden_unit

HAVE_REF = os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libmspack_ref.so"))


def _one(ora, entry):
    u, comp = golden_unit(entry)
    out, st, _ = ora.decode_batch(u, comp, entry["out_len"])
    return out, int(st[0])


@pytest.mark.parametrize("entry", golden_manifest(), ids=lambda e: e["name"])
def test_port_matches_golden(oracle_port, entry):
    out, err = _one(oracle_port, entry)
    assert err == entry["err"]
    if entry["err"] == 0:
        assert hashlib.md5(out.tobytes()).hexdigest() == entry["md5"]
        exp = golden_expected(entry)
        if exp is not None:
            assert out.tobytes() == exp


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref/libmspack_ref.so not built (needs /root/reference)")
@pytest.mark.parametrize("entry", golden_manifest(), ids=lambda e: e["name"])
def test_reference_matches_golden(oracle_ref, entry):
    """Includes the MD5s asserted by libmspack/test/cabd_test.c:472-478 and cabextract/test/large-files.test."""
    assert oracle_ref.kind == "reference"
    out, err = _one(oracle_ref, entry)
    assert err == entry["err"]
    if entry["err"] == 0:
        assert hashlib.md5(out.tobytes()).hexdigest() == entry["md5"]
        if "asserted_md5" in entry:
            assert hashlib.md5(out.tobytes()).hexdigest() == entry["asserted_md5"]


CASES = [(CODEC_MSZIP, dict()), (CODEC_MSZIP, dict(data="random", unit_bytes=70000)), (CODEC_QUANTUM, dict(window_bits=14)),
         (CODEC_QUANTUM, dict(unit_bytes=65536)), (CODEC_LZX, dict(block_mode=4, split=3)),
         (CODEC_LZX, dict(unit_bytes=131072, reset_interval=2, block_mode=4, intel=1, data="binary")),
         (CODEC_LZX, dict(window_bits=15, unit_bytes=100000, block_mode=4))]


@pytest.mark.skipif(not HAVE_REF, reason="needs the reference oracle")
@pytest.mark.parametrize("codec,kw", CASES, ids=lambda x: str(x))
def test_port_matches_reference(oracle_ref, oracle_port, codec, kw):
    b = gen.make_batch(codec, 48, **kw)
    o1, s1, _ = oracle_ref.decode_batch(b.units, b.comp, b.out_bytes, threads=4)
    o2, s2, _ = oracle_port.decode_batch(b.units, b.comp, b.out_bytes, threads=4)
    assert_same(b.units, o1, s1, o2, s2, f"port vs reference {codec} {kw}")


@pytest.mark.skipif(not HAVE_REF, reason="needs the reference oracle")
def test_port_matches_reference_on_corrupt_streams(oracle_ref, oracle_port):
    rng = np.random.default_rng(11)
    for codec in (CODEC_MSZIP, CODEC_LZX, CODEC_QUANTUM):
        b = gen.make_batch(codec, 64)
        comp = b.comp.copy()
        for i, u in enumerate(b.units):
            lo, n = int(u["in_off"]), int(u["in_len"])
            if i % 2 == 0:
                comp[lo + int(rng.integers(0, n))] ^= 1 << int(rng.integers(0, 8))
            else:
                b.units["in_len"][i] = max(1, n - int(rng.integers(1, 40)))
        o1, s1, _ = oracle_ref.decode_batch(b.units, comp, b.out_bytes)
        o2, s2, _ = oracle_port.decode_batch(b.units, comp, b.out_bytes)
        assert_same(b.units, o1, s1, o2, s2, f"corrupt {codec}")


DELTA_CASES = [dict(window_bits=17, ref_bytes=20000), dict(window_bits=25, unit_bytes=70000, ref_bytes=50000, data="binary", intel=1),
               dict(window_bits=18, data="zeros", unit_bytes=65536), dict(window_bits=17, unit_bytes=327680, ref_bytes=70000, block_mode=4, block_frames=2),
               dict(window_bits=17, unit_bytes=131072, ref_bytes=1000, reset_interval=1, block_mode=4)]


@pytest.mark.skipif(not HAVE_REF, reason="needs the reference oracle")
@pytest.mark.parametrize("kw", DELTA_CASES, ids=lambda c: ",".join(f"{k}={v}" for k, v in c.items()))
def test_port_matches_reference_lzx_delta(oracle_ref, oracle_port, kw):
    """LZX DELTA (is_delta=1 + lzxd_set_reference_data): the port against the reference, intact and corrupted."""
    b = gen.make_batch(CODEC_LZX, 12, delta=1, **kw)
    o1, s1, _ = oracle_ref.decode_batch(b.units, b.comp, b.out_bytes, threads=4, out_init=b.out_init)
    o2, s2, _ = oracle_port.decode_batch(b.units, b.comp, b.out_bytes, threads=4, out_init=b.out_init)
    assert_same(b.units, o1, s1, o2, s2, f"delta {kw}")
    rng = np.random.default_rng(13)
    comp, units = b.comp.copy(), b.units.copy()
    for i, u in enumerate(units):
        lo, n = int(u["in_off"]), int(u["in_len"])
        if i % 2 == 0:
            comp[lo + int(rng.integers(0, n))] ^= 1 << int(rng.integers(0, 8))
        else:
            units["in_len"][i] = max(1, n - int(rng.integers(1, 40)))
    o1, s1, _ = oracle_ref.decode_batch(units, comp, b.out_bytes, threads=4, out_init=b.out_init)
    o2, s2, _ = oracle_port.decode_batch(units, comp, b.out_bytes, threads=4, out_init=b.out_init)
    assert_same(units, o1, s1, o2, s2, f"corrupt delta {kw}")


@pytest.mark.skipif(not HAVE_REF, reason="needs the reference oracle")
@pytest.mark.parametrize("seed", range(6))
def test_port_matches_reference_mszip_repair_mode(oracle_ref, oracle_port, seed):
    """Repair mode (mszipd.c:420-433): where the reference goes on after a block it gave up depends on its stale bit state and on
    refills of its input buffer (oracle/port/mspack_port.c zip_repair_restart) - pinned for buffer sizes from 2 bytes to 4 KiB."""
    from util import damaged_mszip_batch
    units, comp, out_bytes = damaged_mszip_batch(200 + seed, level=(6, 1, 0)[seed % 3])
    o1, s1, _ = oracle_ref.decode_batch(units, comp, out_bytes, threads=4)
    o2, s2, _ = oracle_port.decode_batch(units, comp, out_bytes, threads=4)
    assert_same(units, o1, s1, o2, s2, f"repair seed {seed}")
