"""CHM section front end (include/msgpu_chm.h, SURVEY 8 f2 / BASELINE config 4): ControlData + ResetTable -> one unit per LZX
reset interval.  The reference has no CHM with a multi-interval section among its fixtures, so the
pinning is: (1) the unit table rebuilt from the tables equals the generator's, (2) decoding those units - with the device
logic (CPU emulation) here, with the kernels in the gpu test - gives the bytes the UNMODIFIED reference lzxd produces for the
whole section as ONE continuous stream with that reset interval (chmd.c's own way of reading it), (3) malformed tables are
refused with the codes chmd.c:1096-1149,1213-1258 names."""
import ctypes
import struct

import numpy as np
import pytest

from libmspack_b200 import gen
from libmspack_b200.chm import chm_units
from libmspack_b200.units import CODEC_LZX, UNIT_DTYPE

DATAFORMAT, SIGNATURE, ARGS = 8, 7, 1


def control_data(version=2, reset_interval=65536, window=1 << 21, sig=b"LZXC"):
    div = 32768 if version == 2 else 1
    return struct.pack("<I4sIIIII", 6, sig, version, reset_interval // div, window // div, 0, 0)


def reset_table(offsets_per_frame, uncomp_len, comp_len, entry_size=8, frame_len=32768):
    fmt = "<Q" if entry_size == 8 else "<I"
    body = b"".join(struct.pack(fmt, o) for o in offsets_per_frame)
    return struct.pack("<IIIIQQQ", 2, len(offsets_per_frame), entry_size, 0x28, uncomp_len, comp_len, frame_len) + body


def section(n_units=24, frames=2, window_bits=21, tail=True, **kw):
    """A CHM-style LZX content section: n_units reset intervals of `frames` frames each, back to back."""
    b = gen.make_batch(CODEC_LZX, n_units, unit_bytes=frames * 32768, window_bits=window_bits, reset_interval=frames, keep_raw=True, **kw)
    lens = b.units["in_len"].astype(np.int64)
    assert (lens % 2 == 0).all()                       # LZX streams are whole 16-bit words
    content = b"".join(bytes(b.comp[int(u["in_off"]):int(u["in_off"]) + int(u["in_len"])]) for u in b.units)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]])
    if tail:
        content += b"\0" * (8 if kw.get("intel") else 4)      # something behind the last interval (see msgpu_chm.h on look-ahead)
    return b, content, offs


@pytest.mark.parametrize("entry_size", [4, 8])
@pytest.mark.parametrize("version", [1, 2])
def test_units_from_tables(entry_size, version):
    b, content, offs = section()
    per_frame = []
    for o in offs:
        per_frame += [int(o), 0xDEAD]                  # the mid-interval frame offsets are never used by chmd.c
    total = b.n * 65536
    rc, info, units = chm_units(control_data(version), reset_table(per_frame, total - 1234, len(content), entry_size), len(content))
    assert rc == 0
    assert (info.window_bits, info.reset_interval, info.uncomp_len, info.padded_len, info.num_units) == (21, 65536, total - 1234, total, b.n)
    assert (units["codec"] == CODEC_LZX).all() and (units["window_bits"] == 21).all() and (units["reset_interval"] == 2).all()
    assert np.array_equal(units["in_off"], offs.astype(np.uint64))
    assert np.array_equal(units["flags"], np.where(np.arange(b.n) > 0, 0x20 | (np.arange(b.n) * 2 << 6), 0).astype(np.uint32))      # MSGPU_FLAG_LZX_STREAM_BASE
    assert np.array_equal(units["in_len"].astype(np.int64)[:-1], b.units["in_len"].astype(np.int64)[:-1] + 8)      # look-ahead into the next interval
    assert int(units["in_len"][-1]) == int(b.units["in_len"][-1]) + 4
    assert np.array_equal(units["out_off"], np.arange(b.n, dtype=np.uint64) * 65536) and (units["out_len"] == 65536).all()


def test_malformed_tables_are_refused():
    b, content, offs = section(n_units=4)
    pf = [int(o) for o in offs for _ in range(2)]
    good_rt = reset_table(pf, 4 * 65536, len(content))
    assert chm_units(control_data(), good_rt, len(content))[0] == 0
    assert chm_units(control_data(sig=b"LZXD"), good_rt, len(content))[0] == SIGNATURE
    assert chm_units(control_data(version=3), good_rt, len(content))[0] == DATAFORMAT
    assert chm_units(control_data()[:-1], good_rt, len(content))[0] == DATAFORMAT
    assert chm_units(control_data(window=3 << 20), good_rt, len(content))[0] == DATAFORMAT
    assert chm_units(control_data(version=1, reset_interval=65536 + 512), good_rt, len(content))[0] == DATAFORMAT
    assert chm_units(control_data(version=1, reset_interval=0, window=1 << 21) , good_rt, len(content))[0] == DATAFORMAT
    assert chm_units(control_data(), reset_table(pf, 4 * 65536, len(content), frame_len=16384), len(content))[0] == DATAFORMAT
    assert chm_units(control_data(), reset_table(pf, 4 * 65536, len(content), entry_size=8)[:-9], len(content))[0] == DATAFORMAT   # last entry cut
    assert chm_units(control_data(), reset_table(pf[:4], 4 * 65536, len(content)), len(content))[0] == DATAFORMAT                    # too few entries
    assert chm_units(control_data(), good_rt[:0x20], len(content))[0] == DATAFORMAT


def _whole_stream_reference(oracle_ref, content, total, frames, window_bits):
    u = np.zeros(1, UNIT_DTYPE)
    u["codec"], u["window_bits"], u["reset_interval"], u["in_len"], u["out_len"] = CODEC_LZX, window_bits, frames, len(content), total
    out, st, _ = oracle_ref.decode_batch(u, np.frombuffer(content + b"\0" * 16, np.uint8), total)
    return out, int(st[0])


def _tables(b, content, offs, frames):
    pf = [int(o) for o in offs for _ in range(frames)]
    return control_data(reset_interval=frames * 32768, window=1 << int(b.units["window_bits"][0])), reset_table(pf, b.n * frames * 32768, len(content))


@pytest.mark.parametrize("case", [dict(), dict(frames=1, window_bits=16, block_mode=4), dict(frames=3, window_bits=17, block_mode=2, n_units=10),
                                  dict(intel=1, data="binary", n_units=12), dict(intel=1, data="binary", frames=1, window_bits=17, block_mode=4, n_units=40, intel_filesize=200000)],
                         ids=lambda c: ",".join(f"{k}={v}" for k, v in c.items()) or "config4")
def test_interval_batch_equals_continuous_reference_stream(oracle_ref, case):
    """fresh-state-per-interval units (what this project decodes) == the reference decoding the section as one stream"""
    frames = case.get("frames", 2)
    b, content, offs = section(**case)
    total = b.n * frames * 32768
    want, st = _whole_stream_reference(oracle_ref, content, total, frames, int(b.units["window_bits"][0]))
    assert st == 0
    if not case.get("intel"):      # (with E8 translation on, the continuous stream's call offsets count from the SECTION start, the generator's from each unit's)
        assert np.array_equal(want, b.raw)
    rc, info, units = chm_units(*_tables(b, content, offs, frames), len(content))
    assert rc == 0 and info.num_units == b.n
    # device logic on the CPU (tests/emul): same lanes / resolve code as the kernels
    from libmspack_b200 import build
    lib = ctypes.CDLL(build.build_emul())
    lib.emul_decode_batch.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    comp = np.frombuffer(content + b"\0" * 64, np.uint8)
    out = np.zeros(total + 64, np.uint8)
    stt = np.full(len(units), -1, np.int32)
    lib.emul_decode_batch(units.ctypes.data, len(units), comp.ctypes.data, out.ctypes.data, stt.ctypes.data, 2)
    assert (stt == 0).all()
    assert np.array_equal(out[:total], want)


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(), dict(intel=1, data="binary")], ids=["text", "e8"])
def test_interval_batch_on_gpu(decoder, oracle_ref, kw):
    b, content, offs = section(n_units=512, **kw)
    total = b.n * 65536
    want, st = _whole_stream_reference(oracle_ref, content, total, 2, 21)
    assert st == 0
    rc, info, units = chm_units(*_tables(b, content, offs, 2), len(content))
    assert rc == 0
    out, stt = decoder.decode_host(units, np.frombuffer(content + b"\0" * 16, np.uint8), total)
    assert (stt == 0).all() and np.array_equal(out, want)
