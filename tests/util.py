"""Shared helpers for the tests."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def golden_manifest():
    return json.load(open(os.path.join(GOLDEN, "manifest.json")))


def golden_unit(entry):
    """(units ndarray[1], comp uint8[]) for one golden manifest entry."""
    from libmspack_b200.units import UNIT_DTYPE
    data = open(os.path.join(GOLDEN, "units", entry["name"] + ".in"), "rb").read()
    u = np.zeros(1, dtype=UNIT_DTYPE)
    u["codec"], u["window_bits"], u["reset_interval"] = entry["codec"], entry["window_bits"], entry["reset_interval"]
    u["in_len"], u["out_len"] = len(data), entry["out_len"]
    comp = np.frombuffer(data + b"\0" * 16, dtype=np.uint8).copy()
    return u, comp


def golden_expected(entry):
    p = os.path.join(GOLDEN, "units", entry["name"] + ".out")
    return open(p, "rb").read() if os.path.exists(p) else None


def assert_same(units, out_a, st_a, out_b, st_b, what=""):
    """Bit-exact on every unit both decode without error; identical status everywhere."""
    assert np.array_equal(st_a, st_b), (f"{what}: status differs at {np.nonzero(st_a != st_b)[0][:8]}: "
                                        f"{st_a[st_a != st_b][:8]} vs {st_b[st_a != st_b][:8]}")
    if np.array_equal(out_a, out_b):
        return
    for i, u in enumerate(units):
        if st_a[i] != 0:
            continue
        lo, n = int(u["out_off"]), int(u["out_len"])
        if not np.array_equal(out_a[lo:lo + n], out_b[lo:lo + n]):
            d = np.nonzero(out_a[lo:lo + n] != out_b[lo:lo + n])[0]
            raise AssertionError(f"{what}: unit {i} differs at byte {d[0]} ({len(d)} bytes differ)")


def mszip_ring_folder(lens, seed=0):
    """One MSZIP folder whose CK blocks inflate to the given lengths (<= 32768 each) and reference each other across block
    boundaries the way the reference's window allows: a block starts writing the 32 KiB ring at index 0 (mszipd.c:416-417)
    and a match reaching in front of the block reads window[32768 + posn - dist] (:267-268), i.e. after a SHORT block the
    bytes older blocks left further up the ring, not the end of the previous block.  Returns (compressed, expected output)."""
    import zlib
    from libmspack_b200 import gen
    raw = gen.raw_units(1, int(sum(lens)) + 1, data="text", first_unit=seed).tobytes()
    win, comp, out, pos = bytearray(32768), b"", b"", 0
    for k, n in enumerate(lens):
        data = raw[pos:pos + n]
        pos += n
        kw = {"zdict": bytes(win)} if k else {}      # distance d at position 0 reads win[32768 - d]: the ring, read linearly
        c = zlib.compressobj(6, zlib.DEFLATED, -15, **kw)
        comp += b"CK" + c.compress(data) + c.flush()
        out += data
        win[0:n] = data
    return comp, out


def ring_batch(cases):
    """A batch of mszip_ring_folder() units."""
    from libmspack_b200.units import UNIT_DTYPE
    from libmspack_b200 import gen
    units = np.zeros(len(cases), dtype=UNIT_DTYPE)
    comps, raws, ioff, ooff = [], [], 0, 0
    for i, lens in enumerate(cases):
        comp, out = mszip_ring_folder([int(x) for x in lens], seed=i)
        units[i] = (1, 0, 0, 0, ioff, len(comp), len(out), ooff)
        pad = (-len(comp)) % 4
        comps.append(comp + b"\0" * pad)
        raws.append(out)
        ioff += len(comp) + pad
        ooff += (len(out) + 15) & ~15
    comp = np.frombuffer(b"".join(comps) + b"\0" * 16, dtype=np.uint8).copy()
    return gen.Batch(units, comp, None, ooff), raws


RING_CASES = [[32768, 100, 32768], [32768, 100, 50, 20000, 32768, 7], [500, 400, 300, 200, 100, 50, 25, 12, 6, 3, 32768, 1000], [32768, 32768, 1, 32768],
              [15506, 16772, 24746, 31145, 1143, 4724, 9000, 32768, 32768, 31, 4000], [1, 2, 3, 4, 5, 32768, 5, 4, 3, 2, 1, 30000],
              [3000 - 100 * k for k in range(20)] + [32768]]       # the last one: 20 blocks, each shorter than the one before (deeper than P2_HIST_K)


def chain_batch(folder_bytes, data="text", damage=None):
    """MSZIP folders handed over as block chains (include/msgpu.h MSGPU_FLAG_CHAIN_*): one unit per CK block.  Returns
    (chain batch, the same folders as one plain unit each, raw data per folder).  damage(k, b, block) may alter block b of
    folder k (bytes -> bytes)."""
    import zlib
    from libmspack_b200.units import UNIT_DTYPE
    from libmspack_b200 import gen
    cu, pu, comps, raws, ioff, ooff = [], [], [], [], 0, 0
    for k, n in enumerate(folder_bytes):
        raw = gen.raw_units(1, n, data=data, first_unit=10 * k).tobytes()
        raws.append(raw)
        fin, first = ioff, True
        for b, off in enumerate(range(0, n, 32768)):
            blk = raw[off:off + 32768]
            kw = {"zdict": raw[off - 32768:off]} if off else {}
            c = zlib.compressobj(6, zlib.DEFLATED, -15, **kw)
            piece = b"CK" + c.compress(blk) + c.flush()
            if damage:
                piece = damage(k, b, piece)
            cu.append((1, 0, 0, 0x4 if first else 0x8, ioff, len(piece), len(blk), ooff + off))
            comps.append(piece)
            ioff += len(piece)
            first = False
        pu.append((1, 0, 0, 0, fin, ioff - fin, n, ooff))
        pad = (-ioff) % 4
        comps.append(b"\0" * pad)
        ioff += pad
        ooff += (n + 15) & ~15
    comp = np.frombuffer(b"".join(comps) + b"\0" * 16, dtype=np.uint8).copy()
    chain = gen.Batch(np.array(cu, dtype=UNIT_DTYPE), comp, None, ooff)
    plain = gen.Batch(np.array(pu, dtype=UNIT_DTYPE), comp, None, ooff)
    return chain, plain, raws


def kwaj_mszip_stream(lens, seed=0, terminator=True):
    """The MSZIP payload of a KWAJ file (mszipd.c:462-495): per block a 16-bit length, 'CK', a deflate block; a zero length ends it.
    Blocks may have any length <= 32768 and see the 32 KiB ring like a CAB folder's.  Returns (stream, expected output)."""
    import zlib
    from libmspack_b200 import gen
    raw = gen.raw_units(1, int(sum(lens)) + 1, data="text", first_unit=seed + 77).tobytes()
    win, comp, out, pos = bytearray(32768), b"", b"", 0
    for k, n in enumerate(lens):
        data = raw[pos:pos + n]
        pos += n
        kw = {"zdict": bytes(win)} if k else {}
        c = zlib.compressobj(6, zlib.DEFLATED, -15, **kw)
        blk = b"CK" + c.compress(data) + c.flush()
        comp += len(blk).to_bytes(2, "little") + blk
        out += data
        win[0:n] = data
    return comp + (b"\0\0" if terminator else b""), out


def damaged_mszip_batch(seed, n=16, level=6, data="text"):
    """MSZIP folders of several blocks with 1-3 damages each (bit flips, overwritten / zeroed stretches, truncation), repair mode on,
    a random size of the decoder's input buffer (the reference's repair behaviour depends on it): units, compressed bytes, out size."""
    from libmspack_b200 import gen
    rng = np.random.default_rng(seed)
    nblk = int(rng.integers(2, 7))
    ub = 32768 * nblk         # whole frames: a block that OVERFLOWS inside a frame cut short by out_len is the one stated deviation
    b = gen.make_batch(1, n, unit_bytes=ub, first_unit=int(rng.integers(0, 1 << 20)), data=data, level=level)
    comp, units = b.comp.copy(), b.units.copy()
    bufsize = int(rng.choice([4096, 4096, 2048, 512, 64, 16, 6, 2]))
    units["flags"] = 0x1 | (bufsize << 6)                 # MSGPU_FLAG_MSZIP_REPAIR | input buffer size << MSGPU_FLAG_REF_SHIFT
    for i, u in enumerate(units):
        lo, ln = int(u["in_off"]), int(u["in_len"])
        for _ in range(int(rng.integers(1, 4))):
            k, pos = int(rng.integers(0, 4)), lo + int(rng.integers(2, ln))
            if k == 0:
                comp[pos] ^= 1 << int(rng.integers(0, 8))
            elif k == 1:
                comp[pos:pos + 16] = rng.integers(0, 256, min(16, lo + ln - pos), dtype=np.uint8)
            elif k == 2:
                comp[pos:min(pos + 300, lo + ln)] = 0
            else:
                units["in_len"][i] = max(4, ln - int(rng.integers(1, 3000)))
    return units, comp, b.out_bytes
