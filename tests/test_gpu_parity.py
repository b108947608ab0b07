"""GPU parity tests: the CUDA path, called through the C-ABI (libmsgpu.so), against the oracle
(the reference's own decoders when oracle/_ref/libmspack_ref.so is present) - bit-exact output
and identical MSPACK_ERR_* per unit."""
import hashlib

import numpy as np
import pytest

from libmspack_b200 import gen
from libmspack_b200.units import CODEC_LZX, CODEC_MSZIP, CODEC_QUANTUM
from util import assert_same, golden_expected, golden_manifest, golden_unit

pytestmark = pytest.mark.gpu


def _check_batch(decoder, oracle_ref, b, what):
    out_g, st_g = decoder.decode_host(b.units, b.comp, b.out_bytes, out_init=b.out_init)
    out_o, st_o, _ = oracle_ref.decode_batch(b.units, b.comp, b.out_bytes, threads=8, out_init=b.out_init)
    assert_same(b.units, out_o, st_o, out_g, st_g, what)
    return out_g, st_g


@pytest.mark.parametrize("entry", golden_manifest(), ids=lambda e: e["name"])
def test_golden_vectors(decoder, oracle_ref, entry):
    """The reference's own fixture cabinets (tests/golden/manifest.json): same MD5 / same error code."""
    u, comp = golden_unit(entry)
    out, st = decoder.decode_host(u, comp, entry["out_len"])
    assert int(st[0]) == entry["err"]
    if entry["err"] == 0:
        assert hashlib.md5(out.tobytes()).hexdigest() == entry["md5"]
        exp = golden_expected(entry)
        if exp is not None:
            assert out.tobytes() == exp


@pytest.mark.parametrize("data", ["text", "binary", "random", "zeros"])
@pytest.mark.parametrize("unit_bytes", [32768, 65536, 100000, 4097, 1])
def test_mszip_batches(decoder, oracle_ref, data, unit_bytes):
    b = gen.make_batch(CODEC_MSZIP, 96, unit_bytes=unit_bytes, data=data, keep_raw=True)
    out, st = _check_batch(decoder, oracle_ref, b, f"mszip {data} {unit_bytes}")
    assert (st == 0).all()
    stride = (unit_bytes + 15) & ~15
    got = out.reshape(-1, stride)[:, :unit_bytes].reshape(-1)
    assert np.array_equal(got, b.raw)


LZX_CASES = [
    dict(), dict(block_mode=1), dict(block_mode=2), dict(block_mode=3), dict(block_mode=4, split=3),
    dict(window_bits=15, block_mode=4), dict(window_bits=17), dict(intel=1, data="binary"), dict(intel=1, data="binary", block_mode=4),
    dict(intel=1, intel_filesize=40000, data="binary"), dict(data="zeros"), dict(data="random"),
    dict(unit_bytes=65536, reset_interval=2), dict(unit_bytes=65536, reset_interval=2, slack=4), dict(unit_bytes=65536, reset_interval=2, block_frames=2, block_mode=4),
    dict(unit_bytes=65536, block_frames=2), dict(unit_bytes=131072, reset_interval=1, block_mode=4, intel=1, data="binary"),
    dict(unit_bytes=100000, block_mode=4, split=2), dict(unit_bytes=5000), dict(unit_bytes=1), dict(unit_bytes=32769),
    dict(unit_bytes=163840, window_bits=15, block_mode=4),
]


@pytest.mark.parametrize("case", LZX_CASES, ids=lambda c: ",".join(f"{k}={v}" for k, v in c.items()) or "default")
def test_lzx_batches(decoder, oracle_ref, case):
    b = gen.make_batch(CODEC_LZX, 64, **case)
    _check_batch(decoder, oracle_ref, b, f"lzx {case}")


QTM_CASES = [dict(), dict(window_bits=10), dict(window_bits=12, data="binary"), dict(window_bits=16, unit_bytes=100000),
             dict(data="zeros", unit_bytes=65536), dict(data="random"), dict(unit_bytes=3), dict(unit_bytes=32769)]


@pytest.mark.parametrize("case", QTM_CASES, ids=lambda c: ",".join(f"{k}={v}" for k, v in c.items()) or "default")
def test_quantum_batches(decoder, oracle_ref, case):
    b = gen.make_batch(CODEC_QUANTUM, 64, **case)
    out, st = _check_batch(decoder, oracle_ref, b, f"quantum {case}")
    assert (st == 0).all()


def test_mixed_codec_batch(decoder, oracle_ref):
    """Config 5 in small: units of all three codecs interleaved, per-unit dispatch inside one call."""
    parts = [gen.make_batch(c, 40, first_unit=100 * c) for c in (CODEC_MSZIP, CODEC_LZX, CODEC_QUANTUM)]
    b = gen.concat_batches(parts)
    rng = np.random.default_rng(0x51544D31)
    perm = rng.permutation(b.n)
    b.units = b.units[perm].copy()
    _check_batch(decoder, oracle_ref, b, "mixed")


def test_corrupt_streams_same_error_class(decoder, oracle_ref):
    """Bit flips and truncation: the GPU path must report exactly the reference's error code per unit
    and the same bytes wherever the reference still decodes."""
    rng = np.random.default_rng(7)
    for codec in (CODEC_MSZIP, CODEC_LZX, CODEC_QUANTUM):
        b = gen.make_batch(codec, 96)
        comp = b.comp.copy()
        for i, u in enumerate(b.units):
            lo, n = int(u["in_off"]), int(u["in_len"])
            if i % 3 == 0:
                pos = lo + int(rng.integers(0, n))
                comp[pos] ^= 1 << int(rng.integers(0, 8))
            elif i % 3 == 1:
                b.units["in_len"][i] = max(1, n - int(rng.integers(1, 64)))
        b.comp = comp
        out_g, st_g = decoder.decode_host(b.units, b.comp, b.out_bytes)
        out_o, st_o, _ = oracle_ref.decode_batch(b.units, b.comp, b.out_bytes, threads=8)
        assert np.array_equal(st_g, st_o), (codec, st_g[st_g != st_o][:8], st_o[st_g != st_o][:8])
        assert_same(b.units, out_o, st_o, out_g, st_g, f"corrupt codec {codec}")


@pytest.mark.parametrize("shift", [1, 2, 3])
def test_unaligned_input(decoder, oracle_ref, shift):
    """Unit inputs at any byte offset (CFDATA payloads inside a cabinet are not aligned)."""
    for codec, kw in ((CODEC_LZX, dict(block_mode=4, split=2)), (CODEC_LZX, dict(block_mode=3)), (CODEC_MSZIP, dict(data="random", unit_bytes=40000)),
                      (CODEC_MSZIP, dict()), (CODEC_QUANTUM, dict())):
        b = gen.make_batch(codec, 64, **kw)
        b.comp = np.concatenate([np.zeros(shift, np.uint8), b.comp])
        b.units["in_off"] += np.uint64(shift)
        _check_batch(decoder, oracle_ref, b, f"unaligned {codec} {kw} shift {shift}")


def test_host_pipeline_many_subwaves(decoder, oracle_ref):
    """Batches large enough for the host-buffer path to cut them into several sub-waves (copy queues + compute streams,
    msgpu.cu run_wave): one codec (many small sub-waves) and a mixed batch with multi-block MSZIP units (straggler rounds)."""
    b = gen.make_batch(CODEC_LZX, 6000, unit_bytes=8192, block_mode=4, split=2)
    _check_batch(decoder, oracle_ref, b, "host pipeline lzx")
    parts = [gen.make_batch(CODEC_MSZIP, 11000, unit_bytes=3000), gen.make_batch(CODEC_MSZIP, 200, unit_bytes=70000, first_unit=20000),
             gen.make_batch(CODEC_LZX, 11000, unit_bytes=3000, first_unit=40000), gen.make_batch(CODEC_QUANTUM, 11000, unit_bytes=2000, first_unit=60000)]
    m = gen.concat_batches(parts)
    perm = np.random.default_rng(11).permutation(m.n)
    m.units = m.units[perm].copy()
    _check_batch(decoder, oracle_ref, m, "host pipeline mixed")


def test_host_pipeline_codec_streams(decoder, oracle_ref):
    """Host buffers, several codecs, contiguous units (per-range copies, no primed span): every sub-wave's codecs run on streams of
    their own with Quantum's output queued last (msgpu.cu run_wave `qsplit`), MSZIP output goes home before the straggler check and
    the sub-waves that hold folders of short blocks (more rounds than their size suggests) are copied once more."""
    from util import RING_CASES, ring_batch
    rb, raws = ring_batch(RING_CASES[:-1])
    parts = [gen.make_batch(CODEC_MSZIP, 4600), rb, gen.make_batch(CODEC_MSZIP, 900, first_unit=7000),
             gen.make_batch(CODEC_LZX, 5000, first_unit=10000), gen.make_batch(CODEC_QUANTUM, 4700, unit_bytes=4096, first_unit=20000)]
    m = gen.concat_batches(parts)
    out, st = _check_batch(decoder, oracle_ref, m, "host pipeline, a stream per codec")
    assert (st == 0).all()
    for i, r in enumerate(raws):
        assert m.unit_output(out, 4600 + i).tobytes() == r
    z = gen.concat_batches([gen.make_batch(CODEC_MSZIP, 4600), rb, gen.make_batch(CODEC_MSZIP, 4600, first_unit=7000)])
    _check_batch(decoder, oracle_ref, z, "host pipeline, MSZIP with stragglers")


def test_full_size_lzx_properties(decoder):
    """BASELINE config 3 at a quarter of full size (16 384 LZX wb21 units, 512 MiB): round trip against the
    generator's raw data - the size-independent property decode(encode(x)) == x; bench.py checks the
    full 65 536-unit batch the same way."""
    n = 16384
    b = gen.make_batch(CODEC_LZX, n, keep_raw=True)
    out, st = decoder.decode_host(b.units, b.comp, b.out_bytes)
    assert (st == 0).all()
    assert np.array_equal(out, b.raw)


@pytest.mark.parametrize("n,unit_bytes", [(40, 5003), (3000, 1001)], ids=["few-ranges", "primed-span"])
def test_host_call_writes_only_unit_bytes(decoder, oracle_ref, n, unit_bytes):
    """msgpu_decode_batch_host returns exactly the bytes units own: the gaps between ragged units (out_off is 16-byte aligned) and
    the bytes of failed units' neighbours keep what the caller had there - per merged range for a few ranges, through the primed
    span for many (msgpu.cu run_wave out_ranges)."""
    parts = [gen.make_batch(CODEC_MSZIP, n, unit_bytes=unit_bytes), gen.make_batch(CODEC_LZX, n, unit_bytes=unit_bytes + 2, first_unit=5000),
             gen.make_batch(CODEC_QUANTUM, n // 4 + 1, unit_bytes=unit_bytes + 5, first_unit=9000)]
    m = gen.concat_batches(parts)
    perm = np.random.default_rng(3).permutation(m.n)
    m.units = m.units[perm].copy()
    init = np.full(m.out_bytes, 0xA5, dtype=np.uint8)
    out_g, st_g = decoder.decode_host(m.units, m.comp, m.out_bytes, out_init=init)
    out_o, st_o, _ = oracle_ref.decode_batch(m.units, m.comp, m.out_bytes, threads=8)
    assert_same(m.units, out_o, st_o, out_g, st_g, "ragged mixed batch")
    owned = np.zeros(m.out_bytes, dtype=bool)
    for u in m.units:
        owned[int(u["out_off"]):int(u["out_off"]) + int(u["out_len"])] = True
    assert (~owned).any()
    assert (out_g[~owned] == 0xA5).all(), "bytes outside every unit were overwritten"


def test_multi_device_host_call(oracle_ref):
    """msgpu_decode_batch_host_multi: one batch, host buffers, every visible device (two contexts on the one device of a 1-GPU box
    - the sharding, re-basing and threading are the same): mixed codecs, ragged units, LZX DELTA reference data, an MSZIP chain
    across the shard boundary; output and status equal the oracle's, bytes outside the units untouched."""
    import torch
    from libmspack_b200.sharding import MultiDecoder
    from util import chain_batch
    ndev = torch.cuda.device_count()
    devices = list(range(ndev)) if ndev >= 2 else [0, 0]
    chain, plain, raws = chain_batch([32768 * 9 + 5])
    parts = [gen.make_batch(CODEC_LZX, 300, unit_bytes=40001), chain, gen.make_batch(CODEC_QUANTUM, 100, unit_bytes=9000, first_unit=700),
             gen.make_batch(CODEC_LZX, 40, window_bits=17, unit_bytes=50000, delta=1, ref_bytes=20000, first_unit=900), gen.make_batch(CODEC_MSZIP, 301, unit_bytes=5003)]
    m = gen.concat_batches(parts)
    md = MultiDecoder(devices)
    try:
        init = m.out_init.copy() if m.out_init is not None else np.zeros(m.out_bytes, np.uint8)
        out_g, st_g = md.decode_host(m.units, m.comp, m.out_bytes, out_init=init)
        assert md.launches > 0
    finally:
        md.close()
    pm = gen.concat_batches([parts[0], plain, parts[2], parts[3], parts[4]])       # the chain as ONE plain unit: what the reference decodes
    out_o, st_o, _ = oracle_ref.decode_batch(pm.units, pm.comp, pm.out_bytes, threads=8, out_init=pm.out_init)
    assert (st_g == 0).all() and (st_o == 0).all()
    for u in pm.units:
        lo, n = int(u["out_off"]), int(u["out_len"])
        assert np.array_equal(out_g[lo:lo + n], out_o[lo:lo + n])


def test_device_digest_sinks(decoder, oracle_ref):
    """MD5 / CRC-32 of every unit's output computed on the device (msgpu_decode_batch_host_digest: no output copy)
    equal hashlib.md5 / zlib.crc32 of the bytes the reference decodes - ragged sizes around MD5's 55 / 56 / 64-byte padding cases,
    all three codecs, and a failing unit (all-zero digest, the reference's error code)."""
    import zlib
    for codec, sizes in ((CODEC_MSZIP, (1, 55, 56, 63, 64, 65, 119, 120, 4097, 32768, 40000)), (CODEC_LZX, (3, 57, 128, 9999, 32768, 70001)), (CODEC_QUANTUM, (5, 64, 5000))):
        parts = [gen.make_batch(codec, 24, unit_bytes=s, first_unit=100 * k) for k, s in enumerate(sizes)]
        m = gen.concat_batches(parts)
        lo, nbad = int(m.units["in_off"][5]), int(m.units["in_len"][5])
        m.comp = m.comp.copy(); m.comp[lo + nbad // 2:lo + nbad // 2 + 8] ^= 0xFF        # one damaged unit
        out_o, st_o, _ = oracle_ref.decode_batch(m.units, m.comp, m.out_bytes, threads=8)
        md5, st = decoder.decode_host_digest(m.units, m.comp, m.out_bytes, 1)
        crc, st2 = decoder.decode_host_digest(m.units, m.comp, m.out_bytes, 2)
        assert np.array_equal(st, st_o) and np.array_equal(st2, st_o)
        for i, u in enumerate(m.units):
            data = out_o[int(u["out_off"]):int(u["out_off"]) + int(u["out_len"])].tobytes()
            if st_o[i]:
                assert not md5[i].any() and crc[i] == 0
            else:
                assert md5[i].tobytes() == hashlib.md5(data).digest(), (codec, i)
                assert int(crc[i]) == zlib.crc32(data), (codec, i)


def test_device_digest_of_the_golden_vectors(decoder):
    """The MD5s the reference's own test asserts (libmspack/test/cabd_test.c:472-478, tests/golden/manifest.json), computed on the device."""
    for entry in golden_manifest():
        if entry["err"]:
            continue
        u, comp = golden_unit(entry)
        md5, st = decoder.decode_host_digest(u, comp, entry["out_len"], 1)
        assert int(st[0]) == 0 and md5[0].tobytes().hex() == entry["md5"], entry["name"]
