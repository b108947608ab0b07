"""Writers for the two Offline Address Book containers libmspack's oabd.c reads (test helper; the reference ships no OAB
fixtures and no encoder):

* full file   (oabd.c:103-234, oab.h:34-44): header {3, 1, block_max, target_size}, then per block
  {flags, comp_size, uncomp_size, crc} + data; flags 1 = an LZX DELTA stream (window_bits from the block size, :191-194),
  flags 0 = stored.
* patch file  (oabd.c:236-400, oab.h:46-59): header {3, 2, block_max, source_size, target_size, source_crc, target_crc},
  then per block {patch_size, target_size, source_size, crc} + an LZX DELTA stream whose reference data are the next
  source_size bytes of the base file (lzxd_set_reference_data, :349); window_bits from roundup32k(source_size) + target_size.

The block CRC is the raw table-driven CRC-32 state started at 0xFFFFFFFF without the final inversion (oabd.c:97-98, crc32.h).
"""
import struct
import zlib

from libmspack_b200 import gen


def _crc(data: bytes) -> int:
    return (zlib.crc32(data) ^ 0xFFFFFFFF) & 0xFFFFFFFF


def _wbits(size: int) -> int:
    wb = 17
    while wb < 25 and (1 << wb) < size:
        wb += 1
    return wb


def full_oab(target: bytes, block: int = 200000, stored_every: int = 0, **enc) -> bytes:
    out = [struct.pack("<IIII", 3, 1, block, len(target))]
    for k, off in enumerate(range(0, len(target), block)):
        blk = target[off:off + block]
        if stored_every and k % stored_every == stored_every - 1:
            out.append(struct.pack("<IIII", 0, len(blk), len(blk), _crc(blk)) + blk)
            continue
        comp = gen.lzx_encode(blk, window_bits=_wbits(len(blk)), delta=1, seed=k + 1, **enc)
        out.append(struct.pack("<IIII", 1, len(comp), len(blk), _crc(blk)) + comp)
    return b"".join(out)


def patch_oab(base: bytes, target: bytes, tblock: int = 150000, sblock: int = 150000, **enc) -> bytes:
    """Block k turns base[k*sblock : (k+1)*sblock] (+ nothing else) into target[k*tblock : (k+1)*tblock]."""
    nblk = (len(target) + tblock - 1) // tblock
    out = [struct.pack("<IIIIIII", 3, 2, max(tblock, sblock), len(base), len(target), _crc(base), _crc(target))]
    for k in range(nblk):
        blk = target[k * tblock:(k + 1) * tblock]
        src = base[k * sblock:(k + 1) * sblock]
        wsize = ((len(src) + 32767) & ~32767) + len(blk)
        comp = gen.lzx_encode(blk, window_bits=_wbits(wsize), delta=1, ref=src, seed=k + 1, **enc)
        out.append(struct.pack("<IIII", len(comp), len(blk), len(src), _crc(blk)) + comp)
    return b"".join(out)
