"""CPU-only: the SAME per-lane P1 state machines and per-lane P2 functions the CUDA kernels run
(libmspack_b200/csrc/*.cuh compiled as plain C++, tests/emul/emul.cpp) against the oracle.  This checks the
decoder logic without a GPU; the GPU parity tests (-m gpu) check the real kernels through the C-ABI."""
import ctypes
import hashlib
import os

import numpy as np
import pytest

from libmspack_b200 import gen
from libmspack_b200.units import CODEC_LZX, CODEC_MSZIP, CODEC_QUANTUM
from util import RING_CASES, assert_same, chain_batch, golden_manifest, golden_unit, ring_batch

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emul():
    from libmspack_b200 import build
    so = build.build_emul()
    lib = ctypes.CDLL(so)
    lib.emul_decode_batch.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]

    def run(units, comp, out_bytes, frames_per_round=1, out_shift=0, out_init=None):
        units = np.ascontiguousarray(units)
        comp = np.concatenate([np.ascontiguousarray(comp, dtype=np.uint8), np.zeros(64, np.uint8)])
        out = np.zeros(out_bytes + 128, np.uint8)
        base = (-out.ctypes.data) % 16 + out_shift               # 16-byte aligned + the requested misalignment
        if out_init is not None:
            out[base:base + len(out_init)] = out_init
        st = np.full(len(units), -1, np.int32)
        lib.emul_decode_batch(units.ctypes.data, len(units), comp.ctypes.data, out.ctypes.data + base, st.ctypes.data, frames_per_round)
        return out[base:base + out_bytes], st
    return run


@pytest.mark.parametrize("entry", golden_manifest(), ids=lambda e: e["name"])
@pytest.mark.parametrize("frames_per_round", [1, 2, 7, 64])
def test_device_logic_on_golden_vectors(emul, entry, frames_per_round):
    u, comp = golden_unit(entry)
    out, st = emul(u, comp, entry["out_len"], frames_per_round)
    assert int(st[0]) == entry["err"]
    if entry["err"] == 0:
        assert hashlib.md5(out.tobytes()).hexdigest() == entry["md5"]


CASES = [(CODEC_MSZIP, dict()), (CODEC_MSZIP, dict(data="random", unit_bytes=40000)), (CODEC_MSZIP, dict(unit_bytes=65536, level=1)),
         (CODEC_QUANTUM, dict()), (CODEC_QUANTUM, dict(window_bits=12, unit_bytes=65536, data="binary")),
         (CODEC_LZX, dict()), (CODEC_LZX, dict(block_mode=4, split=3)), (CODEC_LZX, dict(unit_bytes=65536, reset_interval=2, block_mode=4)),
         (CODEC_LZX, dict(intel=1, data="binary", unit_bytes=70000)), (CODEC_LZX, dict(window_bits=15, unit_bytes=100000, block_mode=4))]


@pytest.mark.parametrize("codec,kw", CASES, ids=lambda x: str(x))
def test_device_logic_matches_oracle(emul, oracle_ref, codec, kw):
    b = gen.make_batch(codec, 20, **kw)
    o1, s1, _ = oracle_ref.decode_batch(b.units, b.comp, b.out_bytes, threads=4)
    for fpr in (1, 2):
        o2, s2 = emul(b.units, b.comp, b.out_bytes, fpr)
        assert_same(b.units, o1, s1, o2, s2, f"emulation {codec} {kw} F={fpr}")


def test_device_logic_on_corrupt_streams(emul, oracle_ref):
    rng = np.random.default_rng(5)
    for codec in (CODEC_MSZIP, CODEC_LZX, CODEC_QUANTUM):
        b = gen.make_batch(codec, 48)
        comp = b.comp.copy()
        for i, u in enumerate(b.units):
            lo, n = int(u["in_off"]), int(u["in_len"])
            if i % 2 == 0:
                comp[lo + int(rng.integers(0, n))] ^= 1 << int(rng.integers(0, 8))
            else:
                b.units["in_len"][i] = max(1, n - int(rng.integers(1, 40)))
        o1, s1, _ = oracle_ref.decode_batch(b.units, comp, b.out_bytes)
        o2, s2 = emul(b.units, comp, b.out_bytes, 1)
        assert_same(b.units, o1, s1, o2, s2, f"corrupt {codec}")


def _shifted(b, shift):
    """The same batch with every unit's input moved `shift` bytes (unit inputs no longer 4-byte aligned)."""
    comp = np.concatenate([np.zeros(shift, np.uint8), b.comp])
    units = b.units.copy()
    units["in_off"] += np.uint64(shift)
    return units, comp


@pytest.mark.parametrize("shift", [1, 2, 3])
def test_device_logic_unaligned_input(emul, oracle_ref, shift):
    """Unit inputs at any byte offset: the word loads of the bit readers fall back to byte loads."""
    for codec, kw in ((CODEC_LZX, dict(block_mode=4, split=2)), (CODEC_MSZIP, dict(data="random", unit_bytes=40000)), (CODEC_QUANTUM, dict())):
        b = gen.make_batch(codec, 12, **kw)
        units, comp = _shifted(b, shift)
        o1, s1, _ = oracle_ref.decode_batch(units, comp, b.out_bytes, threads=4)
        o2, s2 = emul(units, comp, b.out_bytes, 1)
        assert_same(units, o1, s1, o2, s2, f"unaligned {codec} shift {shift}")


@pytest.mark.parametrize("shift", [1, 2])
def test_device_logic_unaligned_output(emul, oracle_ref, shift):
    """An output base that is not 4/16-byte aligned: literal words and resolved 16-byte groups fall back to byte stores."""
    for codec, kw in ((CODEC_LZX, dict(block_mode=4, split=2, unit_bytes=40000)), (CODEC_MSZIP, dict(data="random", unit_bytes=40000)),
                      (CODEC_MSZIP, dict()), (CODEC_QUANTUM, dict())):
        b = gen.make_batch(codec, 8, **kw)
        o1, s1, _ = oracle_ref.decode_batch(b.units, b.comp, b.out_bytes, threads=4)
        o2, s2 = emul(b.units, b.comp, b.out_bytes, 1, out_shift=shift)
        assert_same(b.units, o1, s1, o2, s2, f"unaligned output {codec} shift {shift}")


DELTA_CASES = [dict(window_bits=17), dict(window_bits=17, ref_bytes=20000), dict(window_bits=22, ref_bytes=100000, unit_bytes=100000, block_mode=4, split=2),
               dict(window_bits=25, unit_bytes=70000, ref_bytes=50000, data="binary", intel=1), dict(window_bits=18, data="zeros", unit_bytes=65536),
               dict(window_bits=17, unit_bytes=196608, ref_bytes=131072, block_mode=4, block_frames=2), dict(window_bits=17, unit_bytes=196685, ref_bytes=1000, block_mode=3),
               dict(window_bits=17, unit_bytes=327680, ref_bytes=70000, block_mode=4), dict(window_bits=17, unit_bytes=131072, ref_bytes=1000, reset_interval=1, block_mode=4),
               dict(window_bits=17, unit_bytes=131072, ref_bytes=1000, reset_interval=1, block_mode=4, slack=4)]


@pytest.mark.parametrize("kw", DELTA_CASES, ids=lambda c: ",".join(f"{k}={v}" for k, v in c.items()))
def test_device_logic_lzx_delta(emul, oracle_ref, kw):
    """LZX DELTA units (lzxd_init(is_delta=1), lzxd.c:441-444 chunk sizes, :589-611 long matches, :348-382 + :622-628
    reference data; windows up to 2^25, units longer than the window): the DELTA / WIDE instantiations against the
    reference, intact and corrupted."""
    b = gen.make_batch(CODEC_LZX, 10, delta=1, **kw)
    o1, s1, _ = oracle_ref.decode_batch(b.units, b.comp, b.out_bytes, threads=4, out_init=b.out_init)
    for fpr in (1, 2):
        o2, s2 = emul(b.units, b.comp, b.out_bytes, fpr, out_init=b.out_init)
        assert_same(b.units, o1, s1, o2, s2, f"delta {kw} F={fpr}")
    rng = np.random.default_rng(3)
    comp, units = b.comp.copy(), b.units.copy()
    for i, u in enumerate(units):
        lo, n = int(u["in_off"]), int(u["in_len"])
        if i % 2 == 0:
            comp[lo + int(rng.integers(0, n))] ^= 1 << int(rng.integers(0, 8))
        else:
            units["in_len"][i] = max(1, n - int(rng.integers(1, 40)))
    o1, s1, _ = oracle_ref.decode_batch(units, comp, b.out_bytes, threads=4, out_init=b.out_init)
    o2, s2 = emul(units, comp, b.out_bytes, 1, out_init=b.out_init)
    assert_same(units, o1, s1, o2, s2, f"corrupt delta {kw}")


@pytest.mark.parametrize("kw", [dict(), dict(block_mode=4, split=3), dict(unit_bytes=65536, reset_interval=2, block_mode=4), dict(window_bits=15, unit_bytes=100000, block_mode=4)],
                         ids=lambda c: ",".join(f"{k}={v}" for k, v in c.items()) or "default")
def test_plain_lzx_through_the_delta_instantiation(emul, oracle_ref, kw):
    """A wave that holds LZX DELTA units runs every LZX unit through the DELTA kernels (msgpu.cu run_wave)."""
    b = gen.make_batch(CODEC_LZX, 10, **kw)
    o1, s1, _ = oracle_ref.decode_batch(b.units, b.comp, b.out_bytes, threads=4)
    o2, s2 = emul(b.units, b.comp, b.out_bytes, 0x101)
    assert_same(b.units, o1, s1, o2, s2, f"forced wide {kw}")


def test_delta_argument_errors(emul, oracle_ref):
    """lzxd_init refuses window_bits outside 17..25 for DELTA (NULL -> NOMEMORY), lzxd_set_reference_data refuses reference data
    on a plain stream or longer than the window (ARGS)."""
    b = gen.make_batch(CODEC_LZX, 4, delta=1, window_bits=17, ref_bytes=4096)
    u = b.units.copy()
    u["window_bits"][0] = 16
    u["window_bits"][1] = 26
    u["flags"][2] = (4096 << 6)                      # reference data without the DELTA flag
    u["flags"][3] = 0x2 | ((1 << 17) + 16 << 6)      # longer than the window
    u["out_off"][3] += (1 << 17) + 16
    o1, s1, _ = oracle_ref.decode_batch(u, b.comp, b.out_bytes + (1 << 18), threads=1, out_init=b.out_init)
    o2, s2 = emul(u, b.comp, b.out_bytes + (1 << 18), 1, out_init=b.out_init)
    assert list(s1) == [6, 6, 1, 1]
    assert list(s2) == list(s1)


def test_device_logic_mszip_short_blocks(emul, oracle_ref):
    """MSZIP folders with blocks shorter than 32 KiB in the middle: matches that reach in front of their block see the
    reference's 32 KiB ring (mszipd.c:267-268), not the linear output (msgpu_p2.cuh "MSZIP ring history", k_p2_ring)."""
    b, raws = ring_batch(RING_CASES)
    o1, s1, _ = oracle_ref.decode_batch(b.units, b.comp, b.out_bytes, threads=4)
    assert (s1 == 0).all()
    for i, r in enumerate(raws):
        assert b.unit_output(o1, i).tobytes() == r                  # the construction is what the reference decodes
    for fpr in (1, 2):
        o2, s2 = emul(b.units, b.comp, b.out_bytes, fpr)
        deep = len(RING_CASES) - 1                                  # beyond the history depth the device refuses loudly
        assert int(s2[deep]) == 11
        s2[deep] = 0
        assert_same(b.units[:deep], o1, s1[:deep], o2, s2[:deep], f"ring F={fpr}")


def test_device_logic_mszip_block_chains(emul, oracle_ref):
    """the CK blocks of an MSZIP folder as a chain of units (entropy stage per block, resolve stage in chain
    order) produce what the reference produces for the folder as one stream."""
    chain, plain, raws = chain_batch([32768 * 4 + 1000, 32768 * 2, 32768 + 1, 100000])
    o1, s1, _ = oracle_ref.decode_batch(plain.units, plain.comp, plain.out_bytes, threads=4)
    assert (s1 == 0).all()
    o2, s2 = emul(chain.units, chain.comp, chain.out_bytes, 1)
    assert (s2 == 0).all()
    assert np.array_equal(o1, o2)
    for k, r in enumerate(raws):
        assert plain.unit_output(o2, k).tobytes() == r


def test_device_logic_chain_blocks_that_are_not_one_ck_block(emul):
    """A chain unit that is not exactly one CK block using exactly its input reports MSGPU_ERR_CHAIN (100): the caller then
    decodes the folder as one stream (msgpu_cab.cu) - trailing bytes, a short or long block, a stream that runs out."""
    def damage(k, b, piece):
        if k == 0 and b == 1:
            return piece + b"\0"                  # trailing byte: the one-stream decoder would scan it for the next CK
        if k == 1 and b == 0:
            return piece[:-3]                       # runs into the next block's bytes
        if k == 2 and b == 1:
            return piece + piece                    # two CK blocks in one unit
        return piece
    chain, plain, raws = chain_batch([32768 * 3, 32768 * 2 + 5, 32768 * 3, 32768 * 2], damage=damage)
    o2, s2 = emul(chain.units, chain.comp, chain.out_bytes, 1)
    first = np.nonzero(chain.units["flags"] == 4)[0]
    assert s2[first[0] + 1] == 100 and s2[first[0]] == 0
    assert s2[first[1]] != 0
    assert s2[first[2] + 1] == 100
    assert (s2[first[3]:] == 0).all()


PACKED_CASES = [dict(), dict(block_mode=1), dict(block_mode=2), dict(block_mode=4, split=3), dict(unit_bytes=65536, reset_interval=2, block_mode=4),
                dict(intel=1, data="binary", unit_bytes=70000, block_mode=2), dict(window_bits=15, unit_bytes=100000, block_mode=4), dict(data="random"),
                dict(data="zeros"), dict(unit_bytes=131072, block_frames=2, block_mode=2)]


@pytest.mark.parametrize("layout", [0x400], ids=["Q"])
@pytest.mark.parametrize("kw", PACKED_CASES, ids=lambda c: ",".join(f"{k}={v}" for k, v in c.items()) or "default")
def test_device_logic_lzx_packed_layout(emul, oracle_ref, kw, layout):
    """The packed shared-memory layout of the plain LZX kernel (LzxSharedQ: byte + two-bit head entries, four-word aligned-offset
    tree, counters sharing the LENGTH limits' array, 16-bit per-length bases and a byte head of the LENGTH tree) decodes exactly
    like the 16-bit one the DELTA instantiation keeps - intact and corrupted streams."""
    b = gen.make_batch(CODEC_LZX, 12, **kw)
    o1, s1, _ = oracle_ref.decode_batch(b.units, b.comp, b.out_bytes, threads=4)
    for fpr in (1, 2):
        o2, s2 = emul(b.units, b.comp, b.out_bytes, layout | fpr)
        assert_same(b.units, o1, s1, o2, s2, f"packed {kw} F={fpr}")
    rng = np.random.default_rng(23)
    comp, units = b.comp.copy(), b.units.copy()
    for i, u in enumerate(units):
        lo, n = int(u["in_off"]), int(u["in_len"])
        if i % 3 != 2:
            comp[lo + int(rng.integers(0, n))] ^= 1 << int(rng.integers(0, 8))
        else:
            units["in_len"][i] = max(1, n - int(rng.integers(1, 40)))
    o1, s1, _ = oracle_ref.decode_batch(units, comp, b.out_bytes, threads=4)
    o2, s2 = emul(units, comp, b.out_bytes, layout | 1)
    assert_same(units, o1, s1, o2, s2, f"corrupt packed {kw}")


def test_device_logic_packed_layout_on_golden_vectors(emul):
    import hashlib
    for entry in golden_manifest():
        if entry["codec"] != CODEC_LZX:
            continue
        u, comp = golden_unit(entry)
        for layout in (0x400,):
            out, st = emul(u, comp, entry["out_len"], layout | 2)
            assert int(st[0]) == entry["err"], entry["name"]
            if entry["err"] == 0:
                assert hashlib.md5(out.tobytes()).hexdigest() == entry["md5"], entry["name"]


def test_device_logic_long_codes_all_layouts(emul, oracle_ref):
    """Main-tree codes up to the 16-bit limit (skewed literals): every storage form of the per-length bases (one word, or 16 bits
    with the split K[16] of MsBoK) and every head layout decodes them."""
    from libmspack_b200.units import UNIT_DTYPE
    rng = np.random.default_rng(9)
    p = 0.5 ** np.arange(1, 41)
    p /= p.sum()
    for trial in range(3):
        n = 32768
        data = rng.choice(40, size=n, p=p).astype(np.uint8)
        idx = rng.integers(0, n, 300)
        data[idx] = rng.integers(40, 256, 300)
        comp = gen.lzx_encode(data.tobytes(), window_bits=16, block_mode=1, chain=1)
        u = np.zeros(1, dtype=UNIT_DTYPE)
        u["codec"], u["window_bits"], u["in_len"], u["out_len"] = CODEC_LZX, 16, len(comp), n
        buf = np.frombuffer(comp + b"\0" * 16, dtype=np.uint8)
        o1, s1, _ = oracle_ref.decode_batch(u, buf, n)
        assert s1[0] == 0 and o1.tobytes() == data.tobytes()
        for mode in (1, 0x201, 0x401, 0xC01, 0x2401, 0x101):
            o2, s2 = emul(u, buf, n, mode)
            assert s2[0] == 0 and np.array_equal(o2, o1), (trial, hex(mode))


def _compare(emul, oracle_ref, units, comp, out_bytes, what, modes=(1,)):
    o1, s1, _ = oracle_ref.decode_batch(units, comp, out_bytes, threads=4)
    for m in modes:
        o2, s2 = emul(units, comp, out_bytes, m)
        assert_same(units, o1, s1, o2, s2, f"{what} mode {m:#x}")
    return s1


def test_device_logic_lzx_crafted_repeat_offsets(emul, oracle_ref):
    """R0-R2 taken from an uncompressed block's header can be anything (lzxd.c:510-515): zero, the window size, beyond it,
    0xFFFFFFFF.  Units that start with an uncompressed block get such values; the verbatim / aligned blocks after it use them
    through the repeated-offset slots (eff == 0, off > window, source in front of the unit ...) - same bytes or same error."""
    import struct
    rng = np.random.default_rng(7)
    specials = [0, 1, 2, 3, 100, 32767, 32768, 32769, 65535, 65536, 65537, (1 << 21) - 3, 1 << 21, (1 << 21) + 1, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFF]
    mutated = 0
    for wb in (15, 16, 21):
        b = gen.make_batch(CODEC_LZX, 96, unit_bytes=50000, window_bits=wb, block_mode=4, split=4, first_unit=wb)
        comp = b.comp.copy()
        for u in b.units:
            lo = int(u["in_off"])
            if ((int(comp[lo]) | (int(comp[lo + 1]) << 8)) >> 12) == 0b0011:        # intel bit 0, block type 3: R0-R2 are bytes 4..15
                vals = [int(rng.choice(specials)) for _ in range(3)]
                comp[lo + 4:lo + 16] = np.frombuffer(struct.pack("<III", *vals), dtype=np.uint8)
                mutated += 1
        _compare(emul, oracle_ref, b.units, comp, b.out_bytes, f"crafted R wb{wb}", (1, 0x401, 0x4401, 0x6401))
    assert mutated > 30


def test_device_logic_mszip_structural_cases(emul, oracle_ref):
    """Deflate shapes zlib level 6 never produces by itself: fixed-Huffman and stored blocks, several deflate blocks per CK block
    (with empty stored ones), Huffman-only and RLE strategies, junk in front of / between CK blocks, empty CK blocks, a block
    longer than 32 KiB (error), requests that end inside a block, streams a byte short."""
    import zlib
    from libmspack_b200.units import UNIT_DTYPE
    raw = gen.raw_units(1, 400000, data="text").tobytes()
    rnd = bytes(np.random.default_rng(3).integers(0, 256, 40000, dtype=np.uint8))

    def ck(data, zdict=None, strategy=zlib.Z_DEFAULT_STRATEGY, level=6, flushes=0):
        kw = {"zdict": zdict} if zdict else {}
        c = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy, **kw)
        out = b""
        if flushes:
            step = max(1, len(data) // (flushes + 1))
            for k in range(0, len(data), step):
                out += c.compress(data[k:k + step]) + c.flush(zlib.Z_FULL_FLUSH if (k // step) % 2 else zlib.Z_SYNC_FLUSH)
        else:
            out += c.compress(data)
        return b"CK" + out + c.flush()

    def blk(k, n=32768):
        return raw[k * 32768:k * 32768 + n]
    streams = [(ck(blk(0), strategy=zlib.Z_FIXED), 32768), (ck(blk(1), flushes=5), 32768), (ck(blk(2), level=0), 32768),
               (ck(blk(3)) + b"garbage!!C" + ck(blk(4), zdict=blk(3)), 65536), (b"xxCxK" + ck(blk(5)), 32768),
               (ck(blk(9)), 20000), (ck(blk(9)) + ck(blk(10), zdict=blk(9)), 40000), (ck(blk(0) + blk(1)[:100]), 32868),
               (ck(b""), 10), (ck(b"") + ck(blk(2)), 32768), (ck(rnd[:32768], strategy=zlib.Z_HUFFMAN_ONLY), 32768),
               (ck(blk(3), strategy=zlib.Z_RLE, flushes=2), 32768), (ck(bytes(32768)), 32768), (ck(blk(4))[:-1], 32768),
               (ck(blk(4)) + b"\0\0\0", 32768)]
    units = np.zeros(len(streams), dtype=UNIT_DTYPE)
    comps, ioff, ooff = [], 0, 0
    for i, (c, n) in enumerate(streams):
        units[i] = (CODEC_MSZIP, 0, 0, 0, ioff, len(c), n, ooff)
        pad = (-len(c)) % 4
        comps.append(c + b"\0" * pad)
        ioff += len(c) + pad
        ooff += (n + 15) & ~15
    comp = np.frombuffer(b"".join(comps) + b"\0" * 16, dtype=np.uint8).copy()
    s1 = _compare(emul, oracle_ref, units, comp, ooff, "mszip structural", (1, 2))
    assert list(s1[:7]) == [0] * 7 and s1[7] == 11 and s1[8] == 3          # the long block fails, the lone empty block runs out of input


def test_device_logic_quantum_two_level_scan(emul, oracle_ref):
    """The Quantum lanes' two-level model scan (group sums over every 8 model entries) and loop-free renormalisation against the
    reference - every window size's model sizes, long units (rescales and re-sorts of every model), state reloads between launches
    (F = 1 and 2), damaged streams."""
    rng = np.random.default_rng(41)
    for kw in (dict(), dict(window_bits=10, unit_bytes=100000), dict(window_bits=12, unit_bytes=65536, data="binary"), dict(window_bits=15, unit_bytes=200000),
               dict(window_bits=17, data="random", unit_bytes=40000), dict(window_bits=21, unit_bytes=300000, data="binary"), dict(data="zeros", unit_bytes=70000)):
        b = gen.make_batch(CODEC_QUANTUM, 8, **kw)
        _compare(emul, oracle_ref, b.units, b.comp, b.out_bytes, f"quantum two-level {kw}", (1, 2))
        comp, units = b.comp.copy(), b.units.copy()
        for i, u in enumerate(units):
            lo, n = int(u["in_off"]), int(u["in_len"])
            if i % 2 == 0:
                comp[lo + int(rng.integers(0, n))] ^= 1 << int(rng.integers(0, 8))
            else:
                units["in_len"][i] = max(1, n - int(rng.integers(1, 40)))
        _compare(emul, oracle_ref, units, comp, b.out_bytes, f"quantum two-level corrupt {kw}", (1,))


def test_device_logic_quantum_many_window_laps(emul, oracle_ref):
    """Quantum units much longer than their window (the copy paths at the window's end, qtmd.c:358-416), intact and damaged."""
    rng = np.random.default_rng(5)
    for wb, ub in ((10, 100000), (11, 70000), (12, 40000)):
        b = gen.make_batch(CODEC_QUANTUM, 12, unit_bytes=ub, window_bits=wb)
        _compare(emul, oracle_ref, b.units, b.comp, b.out_bytes, f"quantum wb{wb}", (1, 2))
        comp = b.comp.copy()
        for u in b.units:
            lo, n = int(u["in_off"]), int(u["in_len"])
            comp[lo + int(rng.integers(0, n))] ^= 1 << int(rng.integers(0, 8))
        _compare(emul, oracle_ref, b.units, comp, b.out_bytes, f"quantum corrupt wb{wb}")


def test_device_logic_mszip_kwaj_framing(emul, oracle_ref):
    """mszipd_decompress_kwaj (mszipd.c:462-495): blocks behind 16-bit lengths until a zero length; out_len is only the capacity
    of the output area.  Same bytes, same produced size, same status as the reference - including a missing terminator (the
    reference's two free zero bytes at EOF end the stream), a bad signature (DATAFORMAT), a damaged block, a too small area."""
    import ctypes
    from libmspack_b200.units import UNIT_DTYPE, FLAG_MSZIP_KWAJ
    from util import kwaj_mszip_stream
    cases = [([32768, 32768, 5000], {}), ([100, 32768, 7, 20000], {}), ([32768], dict(terminator=False)), ([3000] * 12, {}), ([], {})]
    streams = [kwaj_mszip_stream(l, seed=i, **kw) for i, (l, kw) in enumerate(cases)]
    bad_sig = bytearray(streams[0][0]); bad_sig[2 + 32768 // 8] ^= 0; bad_sig[2] = ord("X")
    streams.append((bytes(bad_sig), b""))
    dmg = bytearray(streams[1][0]); dmg[len(dmg) // 2] ^= 0x10
    streams.append((bytes(dmg), None))
    units = np.zeros(len(streams) + 1, dtype=UNIT_DTYPE)
    comps, ioff, ooff = [], 0, 0
    for i, (c, out) in enumerate(streams + [streams[0]]):
        cap = 200000 if i < len(streams) else 40000                 # the last one: the area is too small
        units[i] = (CODEC_MSZIP, 0, 0, FLAG_MSZIP_KWAJ, ioff, len(c), cap, ooff)
        pad = (-len(c)) % 4
        comps.append(c + b"\0" * pad)
        ioff += len(c) + pad
        ooff += (cap + 15) & ~15
    comp = np.frombuffer(b"".join(comps) + b"\0" * 16, dtype=np.uint8).copy()
    ref_decode = oracle_ref._decode
    for i in range(len(units)):
        out_ref = np.zeros(ooff, np.uint8)
        prod = ctypes.c_uint32(0)
        st_ref = ref_decode(units[i:i + 1].ctypes.data, comp.ctypes.data, out_ref.ctypes.data, ctypes.byref(prod))
        for fpr in (1, 2):
            o2, s2 = emul(units[i:i + 1], comp, ooff, fpr)
            lo = int(units[i]["out_off"])
            if i == len(units) - 1:
                assert int(s2[0]) == 101                             # MSGPU_ERR_CAPACITY; the reference's writer just drops the surplus
                continue
            assert int(s2[0]) == st_ref, (i, int(s2[0]), st_ref)
            if st_ref == 0:
                n = prod.value
                assert np.array_equal(o2[lo:lo + n], out_ref[lo:lo + n]), i
                if streams[i][1] is not None:
                    assert o2[lo:lo + n].tobytes() == streams[i][1]
    assert ref_decode is not None


@pytest.mark.parametrize("seed", range(6))
def test_device_logic_mszip_repair_mode(emul, oracle_ref, seed):
    """mszipd_init(repair_mode = 1) (mszipd.c:420-433, cabd's fix_mszip): a block that does not inflate is zero-filled and decoding
    goes on with the reference's STALE bit state (last STORE_BITS, input-buffer refills - ZipLaneC::repair_block); a block that
    overflows 32 KiB overwrites the start of its own window image (ZipLaneC::qbase, k_p2_ring<true>).  Same bytes, same status."""
    from util import damaged_mszip_batch
    units, comp, out_bytes = damaged_mszip_batch(100 + seed, level=(6, 1, 0)[seed % 3], data=("text", "binary")[seed % 2])
    o1, s1, _ = oracle_ref.decode_batch(units, comp, out_bytes, threads=4)
    for fpr in (1, 2):
        o2, s2 = emul(units, comp, out_bytes, fpr)
        assert_same(units, o1, s1, o2, s2, f"repair seed {seed} F={fpr}")





@pytest.mark.parametrize("fpr", [2, 5, 64])
def test_long_units_frames_per_round(emul, oracle_ref, fpr):
    """SURVEY.md 8 f3 (LZX / Quantum folders): long units decoded 2 / 5 / 64 frames per launch round (msgpu.cu frame_slots) give
    the reference's bytes and status, intact and damaged (the GPU twin is tests/test_w_long_units_gpu.py)"""
    parts = [gen.make_batch(CODEC_LZX, 2, unit_bytes=40 * 32768 + 777, block_mode=4, split=2, intel=1, data="binary"),
             gen.make_batch(CODEC_LZX, 2, unit_bytes=33 * 32768, reset_interval=4, window_bits=16, first_unit=10),
             gen.make_batch(CODEC_QUANTUM, 2, unit_bytes=20 * 32768 + 5, window_bits=17, first_unit=20),
             gen.make_batch(CODEC_MSZIP, 2, unit_bytes=30 * 32768 + 100, first_unit=30)]
    m = gen.concat_batches(parts)
    o1, s1, _ = oracle_ref.decode_batch(m.units, m.comp, m.out_bytes, threads=4)
    assert (s1 == 0).all()
    o2, s2 = emul(m.units, m.comp, m.out_bytes, fpr)
    assert_same(m.units, o1, s1, o2, s2, f"long units F={fpr}")
    rng = np.random.default_rng(3)
    comp, units = m.comp.copy(), m.units.copy()
    for i, u in enumerate(units):
        lo, n = int(u["in_off"]), int(u["in_len"])
        if i % 2 == 0:
            comp[lo + n // 2 + int(rng.integers(0, n // 4))] ^= 1 << int(rng.integers(0, 8))
        else:
            units["in_len"][i] = n - n // 3
    o1, s1, _ = oracle_ref.decode_batch(units, comp, m.out_bytes, threads=4)
    o2, s2 = emul(units, comp, m.out_bytes, fpr)
    assert (s1 != 0).any()
    assert_same(units, o1, s1, o2, s2, f"damaged long units F={fpr}")


QTM_CONV = 0x800          # tests/emul/emul.cpp: QtmLane<1, true> (the converged eight-wide scans, msgpu_p1_qtm.cuh scan8)


@pytest.mark.parametrize("kw", [dict(), dict(window_bits=10), dict(window_bits=12, data="binary", unit_bytes=65536), dict(window_bits=16, unit_bytes=100000),
                                dict(data="zeros", unit_bytes=65536), dict(data="random"), dict(unit_bytes=3), dict(unit_bytes=32769)],
                         ids=lambda c: ",".join(f"{k}={v}" for k, v in c.items()) or "default")
def test_quantum_converged_scans(emul, oracle_ref, kw):
    """GET_SYMBOL's two scan levels as branch-free eight-wide walks (QtmLane CONV): the same bytes and status as the reference on
    intact, bit-flipped and truncated units, 1 and 2 frames per launch round."""
    b = gen.make_batch(CODEC_QUANTUM, 12, **kw)
    o1, s1, _ = oracle_ref.decode_batch(b.units, b.comp, b.out_bytes, threads=4)
    assert (s1 == 0).all()
    for fpr in (1, 2):
        o2, s2 = emul(b.units, b.comp, b.out_bytes, fpr | QTM_CONV)
        assert_same(b.units, o1, s1, o2, s2, f"quantum converged {kw} F={fpr}")
    rng = np.random.default_rng(9)
    comp, units = b.comp.copy(), b.units.copy()
    for i, u in enumerate(units):
        lo, n = int(u["in_off"]), int(u["in_len"])
        if i % 2 == 0:
            comp[lo + int(rng.integers(0, n))] ^= 1 << int(rng.integers(0, 8))
        else:
            units["in_len"][i] = max(1, n - int(rng.integers(1, 40)))
    o1, s1, _ = oracle_ref.decode_batch(units, comp, b.out_bytes, threads=4)
    o2, s2 = emul(units, comp, b.out_bytes, 1 | QTM_CONV)
    assert_same(units, o1, s1, o2, s2, f"damaged quantum converged {kw}")


def test_quantum_converged_scans_goldens(emul):
    for entry in golden_manifest():
        if entry["codec"] != CODEC_QUANTUM:
            continue
        u, comp = golden_unit(entry)
        out, st = emul(u, comp, entry["out_len"], 2 | QTM_CONV)
        assert int(st[0]) == entry["err"], entry["name"]
        if entry["err"] == 0:
            assert hashlib.md5(out.tobytes()).hexdigest() == entry["md5"], entry["name"]
