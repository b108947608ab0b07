"""Unit descriptor shared by the host API and the generators.

Mirrors `msgpu_unit` in include/msgpu.h (32 bytes, little-endian, no padding).  Codec ids are the
CAB folder compression types (libmspack/mspack/cab.h:50-52).
"""
import numpy as np

UNIT_DTYPE = np.dtype([
    ("codec", "u1"), ("window_bits", "u1"), ("reset_interval", "<u2"), ("flags", "<u4"),
    ("in_off", "<u8"), ("in_len", "<u4"), ("out_len", "<u4"), ("out_off", "<u8"),
])
assert UNIT_DTYPE.itemsize == 32

CODEC_MSZIP, CODEC_QUANTUM, CODEC_LZX = 1, 2, 3
FLAG_MSZIP_REPAIR = 0x1
FLAG_LZX_DELTA = 0x2          # include/msgpu.h MSGPU_FLAG_LZX_DELTA
FLAG_CHAIN_FIRST, FLAG_CHAIN_NEXT = 0x4, 0x8   # MSZIP block chains
FLAG_MSZIP_KWAJ = 0x10        # MSZIP inside a KWAJ file: out_len is a capacity, the stream ends at a zero block length
ERR_CHAIN, ERR_CAPACITY = 100, 101
FLAG_LZX_STREAM_BASE = 0x20  # include/msgpu.h MSGPU_FLAG_LZX_STREAM_BASE: flags >> 6 = index of the unit's first frame in its stream
FLAG_REF_SHIFT = 6            # flags >> 6 = LZX DELTA reference bytes stored in front of the unit's output

# MSPACK_ERR_* (libmspack/mspack/mspack.h:485-507)
ERR_OK, ERR_ARGS, ERR_OPEN, ERR_READ, ERR_WRITE, ERR_SEEK, ERR_NOMEMORY = 0, 1, 2, 3, 4, 5, 6
ERR_SIGNATURE, ERR_DATAFORMAT, ERR_CHECKSUM, ERR_CRUNCH, ERR_DECRUNCH = 7, 8, 9, 10, 11
