"""Synthetic batches of compressed units for the benchmarks and tests.

The reference ships decoders only, so the workloads BASELINE.json names are produced here:
MSZIP units with Python's zlib (raw deflate + the two-byte "CK" signature, mszipd.c:406-413),
LZX and Quantum units with the C encoders in msgen.c.  Nothing in this package is on the decode
path; it only feeds it.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from ..units import UNIT_DTYPE, CODEC_MSZIP, CODEC_QUANTUM, CODEC_LZX

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
CORPUS_SEED = 0x4D534346
FRAME = 32768
DATA_KINDS = {"text": 0, "binary": 1, "zeros": 2, "random": 3}


class _LzxParams(ctypes.Structure):
    _fields_ = [("window_bits", ctypes.c_int), ("reset_interval", ctypes.c_int), ("block_frames", ctypes.c_int),
                ("split", ctypes.c_int), ("block_mode", ctypes.c_int), ("intel", ctypes.c_int),
                ("intel_filesize", ctypes.c_uint32), ("chain", ctypes.c_int), ("seed", ctypes.c_uint32),
                ("delta", ctypes.c_int), ("ref_len", ctypes.c_uint32)]


class _QtmParams(ctypes.Structure):
    _fields_ = [("window_bits", ctypes.c_int), ("chain", ctypes.c_int)]


class _Batch(ctypes.Structure):
    _fields_ = [("codec", ctypes.c_int), ("data_kind", ctypes.c_int), ("seed", ctypes.c_uint64),
                ("first_block", ctypes.c_uint64), ("unit_bytes", ctypes.c_uint32), ("slot_bytes", ctypes.c_uint32),
                ("lzx", _LzxParams), ("qtm", _QtmParams)]


def build_lib(force: bool = False) -> str:
    src = os.path.join(HERE, "msgen.c")
    so = os.path.join(HERE, "libmsgen.so")
    if force or not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(so) < os.path.getmtime(src)):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", so, src, "-lpthread"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_lib())
        _LIB.msgen_generate.restype = ctypes.c_int
        _LIB.msgen_generate.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_int]
        _LIB.msgen_generate_ref.restype = ctypes.c_int
        _LIB.msgen_generate_ref.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_int]
        _LIB.msgen_fill_raw.restype = ctypes.c_int
        _LIB.msgen_fill_raw.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int]
        _LIB.msgen_lzx_encode.restype = ctypes.c_size_t
        _LIB.msgen_lzx_encode.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
        _LIB.msgen_qtm_encode.restype = ctypes.c_size_t
        _LIB.msgen_qtm_encode.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    return _LIB


def _threads(threads):
    return threads or min(32, os.cpu_count() or 1)


def raw_units(n: int, unit_bytes: int = FRAME, data: str = "text", seed: int = CORPUS_SEED,
              first_unit: int = 0, threads: int = 0) -> np.ndarray:
    """n * unit_bytes of synthetic uncompressed data (unit i = corpus blocks of unit first_unit+i)."""
    nfr = (unit_bytes + FRAME - 1) // FRAME
    b = _Batch(codec=0, data_kind=DATA_KINDS[data], seed=seed, first_block=first_unit * nfr,
               unit_bytes=unit_bytes, slot_bytes=0)
    raw = np.empty(n * unit_bytes, dtype=np.uint8)
    _lib().msgen_fill_raw(ctypes.byref(b), n, raw.ctypes.data, _threads(threads))
    return raw


def lzx_encode(data: bytes, window_bits: int = 21, reset_interval: int = 0, block_frames: int = 1, split: int = 1,
               block_mode: int = 0, intel: int = 0, intel_filesize: int = 12000000, chain: int = 24, seed: int = 1,
               delta: int = 0, ref: bytes = b"") -> bytes:
    """One LZX stream.  delta=1: LZX DELTA (16-bit chunk size per frame, long matches); `ref` = its reference data."""
    src = np.frombuffer(bytes(ref) + bytes(data), dtype=np.uint8)
    cap = len(data) + len(data) // 8 + 4096
    dst = np.empty(cap, dtype=np.uint8)
    p = _LzxParams(window_bits, reset_interval, block_frames, split, block_mode, intel, intel_filesize, chain, seed, delta, len(ref))
    r = _lib().msgen_lzx_encode(src.ctypes.data + len(ref), len(data), dst.ctypes.data, cap, ctypes.byref(p))
    if r == 0 and len(data):
        raise RuntimeError("lzx_encode overflow")
    return dst[:r].tobytes()


def qtm_encode(data: bytes, window_bits: int = 21, chain: int = 24) -> bytes:
    src = np.frombuffer(bytes(data), dtype=np.uint8)
    cap = len(data) * 2 + 4096
    dst = np.empty(cap, dtype=np.uint8)
    p = _QtmParams(window_bits, chain)
    r = _lib().msgen_qtm_encode(src.ctypes.data, len(data), dst.ctypes.data, cap, ctypes.byref(p))
    if r == 0 and len(data):
        raise RuntimeError("qtm_encode overflow")
    return dst[:r].tobytes()


def mszip_encode(data: bytes, level: int = 6, history: bool = True) -> bytes:
    """One MSZIP folder: a 'CK' + raw-deflate block per 32 KiB, each block's dictionary being the
    previous block's 32 KiB as the decoder keeps it (mszipd.c:267-268, mszip.h:70)."""
    out = []
    for off in range(0, max(len(data), 1), FRAME):
        blk = data[off:off + FRAME]
        kw = {}
        if history and off:
            kw["zdict"] = data[off - FRAME:off]
        c = zlib.compressobj(level, zlib.DEFLATED, -15, **kw)
        out.append(b"CK" + c.compress(blk) + c.flush())
    return b"".join(out)


class Batch:
    """A batch of independent units: descriptors, packed compressed bytes, optional raw data."""

    def __init__(self, units: np.ndarray, comp: np.ndarray, raw: np.ndarray | None, out_bytes: int,
                 out_init: np.ndarray | None = None):
        self.units, self.comp, self.raw, self.out_bytes = units, comp, raw, out_bytes
        # LZX DELTA: initial contents of the output buffer - each unit's reference data sits in front of its output
        # (include/msgpu.h MSGPU_FLAG_REF_SHIFT); None for every other kind of batch
        self.out_init = out_init

    def unit_output(self, out: np.ndarray, i: int) -> np.ndarray:
        lo, n = int(self.units["out_off"][i]), int(self.units["out_len"][i])
        return out[lo:lo + n]

    @property
    def n(self):
        return len(self.units)

    @property
    def in_bytes(self):
        return int(self.units["in_len"].astype(np.int64).sum())


def _pack(codec, window_bits, reset_interval, pieces, unit_bytes, raw):
    n = len(pieces)
    units = np.zeros(n, dtype=UNIT_DTYPE)
    lens = np.array([len(p) for p in pieces], dtype=np.int64)
    # keep every unit's input 4-byte aligned so device-side word loads are aligned
    padded = (lens + 3) & ~3
    offs = np.concatenate([[0], np.cumsum(padded)[:-1]]) if n else np.zeros(0, dtype=np.int64)
    comp = np.zeros(int(padded.sum()) + 16, dtype=np.uint8)
    for i, p in enumerate(pieces):
        comp[offs[i]:offs[i] + lens[i]] = np.frombuffer(p, dtype=np.uint8)
    units["codec"], units["window_bits"], units["reset_interval"] = codec, window_bits, reset_interval
    units["in_off"], units["in_len"] = offs, lens
    units["out_len"] = unit_bytes
    units["out_off"] = np.arange(n, dtype=np.uint64) * np.uint64((unit_bytes + 15) & ~15)
    return Batch(units, comp, raw, n * ((unit_bytes + 15) & ~15))


def make_batch(codec: int, n: int, unit_bytes: int = FRAME, window_bits: int = 21, data: str = "text",
               seed: int = CORPUS_SEED, first_unit: int = 0, threads: int = 0, keep_raw: bool = False,
               reset_interval: int = 0, block_frames: int = 1, split: int = 1, block_mode: int = 0, intel: int = 0,
               intel_filesize: int = 12000000, chain: int = 24, level: int = 6, slack: int = 0,
               delta: int = 0, ref_bytes: int = 0) -> Batch:
    """n independent units of `unit_bytes` uncompressed bytes each, all of one codec.

    slack: extra bytes added to every unit's in_len (they are the next unit's first bytes, as in a CHM content stream).
    The reference runs one more, empty, frame pass when a request ends on a frame boundary and at a reset point that
    pass reads ahead (lzxd.c:419-453, :696-697): a reset-interval unit cut exactly at its last byte decodes completely
    but returns MSPACK_ERR_READ; 4 bytes of slack avoid that.
    delta / ref_bytes: LZX DELTA units (window_bits 17..25), each with ref_bytes of reference data - an older version of
    the unit's own data - stored in front of its output (Batch.out_init holds the output buffer's initial contents)."""
    threads = _threads(threads)
    nfr = (unit_bytes + FRAME - 1) // FRAME
    if codec == CODEC_MSZIP:
        raw = raw_units(n, unit_bytes, data, seed, first_unit, threads)
        mv = memoryview(raw)

        def one(i):
            return mszip_encode(bytes(mv[i * unit_bytes:(i + 1) * unit_bytes]), level)
        with ThreadPoolExecutor(threads) as ex:
            pieces = list(ex.map(one, range(n), chunksize=max(1, n // (threads * 8) or 1)))
        return _pack(codec, 0, 0, pieces, unit_bytes, raw if keep_raw else None)

    slot = unit_bytes + unit_bytes // 4 + 2048
    if codec == CODEC_QUANTUM:
        slot = unit_bytes * 2 + 2048
    b = _Batch(codec=codec, data_kind=DATA_KINDS[data], seed=seed, first_block=first_unit * nfr,
               unit_bytes=unit_bytes, slot_bytes=slot)
    if codec != CODEC_LZX or not delta:
        delta, ref_bytes = 0, 0
    b.lzx = _LzxParams(window_bits, reset_interval, block_frames, split, block_mode, intel, intel_filesize, chain, seed & 0xFFFFFFFF,
                       delta, ref_bytes)
    b.qtm = _QtmParams(window_bits, chain)
    raw = np.empty(n * unit_bytes, dtype=np.uint8) if keep_raw else None
    ref = np.empty(n * ref_bytes, dtype=np.uint8) if ref_bytes else None
    comp_slots = np.empty(n * slot, dtype=np.uint8)
    lens = np.zeros(n, dtype=np.uint32)
    _lib().msgen_generate_ref(ctypes.byref(b), n, raw.ctypes.data if keep_raw else None, ref.ctypes.data if ref_bytes else None,
                              comp_slots.ctypes.data, lens.ctypes.data, threads)
    if n and int(lens.min()) == 0:
        raise RuntimeError("encoder overflowed its slot")
    units = np.zeros(n, dtype=UNIT_DTYPE)
    l64 = lens.astype(np.int64)
    padded = (l64 + 3) & ~3
    offs = np.concatenate([[0], np.cumsum(padded)[:-1]]) if n else np.zeros(0, dtype=np.int64)
    comp = np.zeros(int(padded.sum()) + 16, dtype=np.uint8)
    # compact the slots (vectorised gather, a few thousand units at a time: the index arrays are 24 bytes per compressed byte)
    for c0 in range(0, n, 4096):
        c1 = min(n, c0 + 4096)
        lc = l64[c0:c1]
        idx_unit = np.repeat(np.arange(c0, c1, dtype=np.int64), lc)
        within = np.arange(int(lc.sum()), dtype=np.int64) - np.repeat(np.cumsum(lc) - lc, lc)
        comp[np.repeat(offs[c0:c1], lc) + within] = comp_slots[idx_unit * slot + within]
    units["codec"], units["window_bits"] = codec, window_bits
    units["reset_interval"] = reset_interval if codec == CODEC_LZX else 0
    units["in_off"], units["in_len"], units["out_len"] = offs, lens + np.uint32(slack), unit_bytes
    stride = (unit_bytes + 15) & ~15
    if delta:
        units["flags"] = 0x2 | (ref_bytes << 6)                       # MSGPU_FLAG_LZX_DELTA | reference bytes << MSGPU_FLAG_REF_SHIFT
        rpad = (ref_bytes + 15) & ~15
        stride += rpad
        units["out_off"] = np.arange(n, dtype=np.uint64) * np.uint64(stride) + np.uint64(rpad)
        out_init = None
        if ref_bytes:
            out_init = np.zeros(n * stride, dtype=np.uint8)
            out_init.reshape(n, stride)[:, rpad - ref_bytes:rpad] = ref.reshape(n, ref_bytes)
        return Batch(units, comp, raw, n * stride, out_init)
    units["out_off"] = np.arange(n, dtype=np.uint64) * np.uint64(stride)
    return Batch(units, comp, raw, n * stride)


def concat_batches(batches) -> Batch:
    """Concatenate batches (e.g. a mixed-codec batch) re-basing the offsets."""
    units, comps, in_base, out_base = [], [], 0, 0
    any_init = any(b.out_init is not None for b in batches)
    inits = []
    for b in batches:
        if any_init:
            inits.append(b.out_init if b.out_init is not None else np.zeros(b.out_bytes, dtype=np.uint8))
        u = b.units.copy()
        u["in_off"] += np.uint64(in_base)
        u["out_off"] += np.uint64(out_base)
        units.append(u)
        comps.append(b.comp)
        in_base += len(b.comp)
        out_base += b.out_bytes
    return Batch(np.concatenate(units), np.concatenate(comps), None, out_base, np.concatenate(inits) if any_init else None)
