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


def test_mszip_block_chains(decoder, oracle_ref):
    """MSZIP folders handed over as block chains (one unit per CK block; entropy stage per block, k_p2_chain in
    chain order) decode to what the reference decodes from each folder as ONE stream; chains and ordinary units in one batch."""
    from util import chain_batch
    sizes = [32768 * 4 + 1000, 32768 * 2, 32768 + 1, 100000, 32768 * 40 + 7] + [32768 * 3 + 11 * k for k in range(60)]
    chain, plain, raws = chain_batch(sizes)
    o1, s1, _ = oracle_ref.decode_batch(plain.units, plain.comp, plain.out_bytes, threads=8)
    assert (s1 == 0).all()
    extra = gen.make_batch(CODEC_MSZIP, 100, unit_bytes=50000, first_unit=900)
    m = gen.concat_batches([chain, extra])
    out, st = decoder.decode_host(m.units, m.comp, m.out_bytes)
    assert (st == 0).all()
    for k, r in enumerate(raws):
        assert np.array_equal(plain.unit_output(out, k), plain.unit_output(o1, k)), k
        assert plain.unit_output(out, k).tobytes() == r
    # the whole buffer: the bytes between two folders (alignment gaps) belong to nobody and keep the caller's contents
    init = np.full(m.out_bytes, 0x5A, dtype=np.uint8)
    out2, st2 = decoder.decode_host(m.units, m.comp, m.out_bytes, out_init=init)
    ref = init.copy()
    for k in range(plain.n):
        lo = int(plain.units["out_off"][k]); ref[lo:lo + int(plain.units["out_len"][k])] = plain.unit_output(o1, k)
    for k in range(extra.n):
        lo = chain.out_bytes + int(extra.units["out_off"][k]); ref[lo:lo + int(extra.units["out_len"][k])] = extra.unit_output(out[chain.out_bytes:], k)
    assert (st2 == 0).all() and np.array_equal(out2, ref)
    oe, se, _ = oracle_ref.decode_batch(extra.units, extra.comp, extra.out_bytes, threads=8)
    for k in range(extra.n):
        assert np.array_equal(extra.unit_output(out[chain.out_bytes:], k), extra.unit_output(oe, k)), k


def test_chain_units_that_are_not_one_ck_block(decoder):
    """MSGPU_ERR_CHAIN (100) for a chain unit with trailing bytes / a stream that runs out / two CK blocks; malformed chains are
    refused by the call itself."""
    from util import chain_batch

    def damage(k, b, piece):
        if k == 0 and b == 1:
            return piece + b"\0"
        if k == 1 and b == 0:
            return piece[:-3]
        if k == 2 and b == 1:
            return piece + piece
        return piece
    chain, plain, raws = chain_batch([32768 * 3, 32768 * 2 + 5, 32768 * 3, 32768 * 2], damage=damage)
    out, st = decoder.decode_host(chain.units, chain.comp, chain.out_bytes)
    first = np.nonzero(chain.units["flags"] == 4)[0]
    assert st[first[0] + 1] == 100 and st[first[0]] == 0
    assert st[first[1]] != 0
    assert st[first[2] + 1] == 100
    assert (st[first[3]:] == 0).all()
    assert plain.unit_output(out, 3).tobytes() == raws[3]
    bad = chain.units.copy()
    bad["flags"][0] = 0x8                                 # a NEXT unit without a predecessor
    with pytest.raises(RuntimeError):
        decoder.decode_host(bad, chain.comp, chain.out_bytes)
    bad = chain.units.copy()
    bad["out_off"][1] += 16                               # not contiguous with its predecessor
    with pytest.raises(RuntimeError):
        decoder.decode_host(bad, chain.comp, chain.out_bytes)


def test_cabinet_with_long_mszip_folders(decoder, monkeypatch):
    """Cabinet front end: multi-block MSZIP folders go through the block chains; same bytes and statuses as with
    MSGPU_CAB_NOCHAIN=1 (every folder one stream), including a folder with a corrupt block (falls back to one stream)."""
    import zlib
    from libmspack_b200 import cab
    from cabfile import build_cab
    folders, raws = [], []
    for k, n in enumerate([32768 * 50 + 123, 32768 * 3, 40000, 32768 * 8]):
        raw = gen.raw_units(1, n, data="text" if k != 1 else "binary", first_unit=50 * k).tobytes()
        blocks = []
        for off in range(0, n, 32768):
            kw = {"zdict": raw[off - 32768:off]} if off else {}
            c = zlib.compressobj(6, zlib.DEFLATED, -15, **kw)
            blocks.append((b"CK" + c.compress(raw[off:off + 32768]) + c.flush(), len(raw[off:off + 32768])))
        folders.append(dict(comp_type=1, blocks=blocks, files=[(f"f{k}.bin", 0, n)]))
        raws.append(raw)
    # folder 3: a payload with a stray byte behind its CK block - not "exactly one CK block", so the chain is refused and the folder
    # decoded as one stream, which skips the byte while looking for the next CK (mszipd.c:405-413); folder 1: a damaged block
    p3, u3 = folders[3]["blocks"][4]
    folders[3]["blocks"][4] = (p3 + b"\x00", u3)
    p1, u1 = folders[1]["blocks"][1]
    folders[1]["blocks"][1] = (p1[:300] + bytes(64) + p1[364:], u1)
    img = build_cab(folders, with_checksums=False)
    plan = cab.scan(img)
    out_c, st_c = plan.decode(decoder)
    monkeypatch.setenv("MSGPU_CAB_NOCHAIN", "1")
    out_s, st_s = plan.decode(decoder)
    assert list(st_c) == list(st_s)
    assert st_c[0] == 0 and st_c[2] == 0 and st_c[3] == 0
    for k in (0, 2, 3):
        base, n = int(plan.folders["out_off"][k]), int(plan.folders["out_len"][k])
        assert out_c[base:base + n].tobytes() == raws[k]
        assert out_s[base:base + n].tobytes() == raws[k]
