"""CPU-only: the C-ABI library loads, exports every symbol include/msgpu.h declares, and refuses to run
without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from libmspack_b200 import build, codec
    build.build_msgpu()
    lib = ctypes.CDLL(codec.LIB_PATH)
    hdr = open(os.path.join(ROOT, "include", "msgpu.h")).read()
    declared = set(re.findall(r"\b(msgpu_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(codec.ABI_SYMBOLS), declared ^ set(codec.ABI_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in codec.load_library().msgpu_version()


def test_library_exports_cabinet_front_end():
    """include/msgpu_cab.h: every declared entry point is exported; struct layouts match the binding."""
    from libmspack_b200 import cab, codec
    lib = ctypes.CDLL(codec.LIB_PATH)
    hdr = open(os.path.join(ROOT, "include", "msgpu_cab.h")).read()
    declared = set(re.findall(r"\b(msgpu_cab_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(cab.CAB_SYMBOLS), declared ^ set(cab.CAB_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert (cab.FOLDER_DTYPE.itemsize, cab.BLOCK_DTYPE.itemsize, cab.FILE_DTYPE.itemsize) == (56, 32, 16)
    from libmspack_b200 import chm
    hdr = open(os.path.join(ROOT, "include", "msgpu_chm.h")).read()
    declared = set(re.findall(r"\b(msgpu_chm_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(chm.CHM_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def test_unit_descriptor_layout_matches_header():
    from libmspack_b200.units import UNIT_DTYPE
    assert UNIT_DTYPE.itemsize == 32
    assert [UNIT_DTYPE.fields[k][1] for k in ("codec", "window_bits", "reset_interval", "flags", "in_off", "in_len", "out_len", "out_off")] == \
        [0, 1, 2, 4, 8, 16, 20, 24]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from libmspack_b200.codec import BatchDecoder
    with pytest.raises(RuntimeError):
        BatchDecoder(0)


def test_dropin_exports_reference_entry_points():
    """The streaming entry points cabd.c / chmd.c bind (lzx.h:146-214, qtm.h:92-122, mszip.h:85-120)."""
    from libmspack_b200 import build
    so = build.build_dropin()
    if so is None:
        pytest.skip("drop-in not built")
    lib = ctypes.CDLL(so)
    for name in ("lzxd_init", "lzxd_set_output_length", "lzxd_set_reference_data", "lzxd_decompress", "lzxd_free",
                 "qtmd_init", "qtmd_decompress", "qtmd_free", "mszipd_init", "mszipd_decompress", "mszipd_decompress_kwaj", "mszipd_free"):
        assert hasattr(lib, name), name


def test_batch_plan_frame_slots_and_waves(monkeypatch):
    """Host logic of msgpu.cu (pick_fmax / frame_slots / wave_end) through msgpu_plan_batch, no GPU: big batches of short units keep
    two frames per launch round and one frame slot per one-frame unit (the headline batch's scratch footprint), batches of few long
    units get up to 64 (DESIGN.md section 7 "Long units"), MSGPU_FMAX overrides, a scratch budget cuts a batch into waves without
    splitting an MSZIP block chain."""
    from libmspack_b200.codec import plan_batch
    from libmspack_b200.units import UNIT_DTYPE, CODEC_LZX, CODEC_MSZIP, CODEC_QUANTUM
    monkeypatch.delenv("MSGPU_FMAX", raising=False)
    GIB = 1 << 30
    slot_bytes = 16400 * 8

    def units(spec):
        u = np.zeros(len(spec), UNIT_DTYPE)
        for i, (codec, out_len, flags) in enumerate(spec):
            u["codec"][i], u["out_len"][i], u["flags"][i], u["window_bits"][i], u["in_len"][i] = codec, out_len, flags, 21, 100
        return u

    # BASELINE configs[2]: 65 536 one-frame LZX units: one slot each, one round, one wave within the B200's budget
    p = plan_batch(units([(CODEC_LZX, 32768, 0)] * 65536), 80 * GIB)
    assert p == dict(fmax=2, frame_slots=65536, waves=1, rounds=1)
    # BASELINE configs[3]: reset intervals of two frames: two slots each
    p = plan_batch(units([(CODEC_LZX, 65536, 0)] * 4096), 80 * GIB)
    assert (p["fmax"], p["frame_slots"], p["rounds"]) == (2, 8192, 1)
    # cabextract's large-files.cab: an MSZIP chain of 65 535 blocks next to two LZX folders of 65 535 frames
    chain = [(CODEC_MSZIP, 32768, 0x4)] + [(CODEC_MSZIP, 32768, 0x8)] * 65534
    long_units = [(CODEC_LZX, 65535 * 32768, 0)] * 2
    p = plan_batch(units(chain + long_units), 80 * GIB)
    assert p == dict(fmax=64, frame_slots=65535 + 2 * 64, waves=1, rounds=1024)
    # many medium-long units: the frame-slot ceiling (131 072 slots) holds fmax down
    p = plan_batch(units([(CODEC_QUANTUM, 100 * 32768, 0)] * 10000), 80 * GIB)
    assert p["fmax"] == 8 and p["frame_slots"] == 80000 and p["rounds"] == 13
    # a small scratch budget: fewer slots per unit, several waves of at least 1 024 units
    p = plan_batch(units([(CODEC_LZX, 40 * 32768, 0)] * 5000), GIB // 2)
    assert p["fmax"] == 2 and p["frame_slots"] == 10000 and p["rounds"] == 20
    assert p["waves"] == -(-5000 // ((GIB // 2) // (2 * slot_bytes + 10368)))       # 1 968 units per wave: 3 waves
    # a KWAJ / repair-mode MSZIP unit always has two slots (a repaired block may need both)
    p = plan_batch(units([(CODEC_MSZIP, 1000, 0x10), (CODEC_MSZIP, 1000, 0)]), GIB)
    assert p["frame_slots"] == 3
    # the environment override, as the GPU tests use it
    monkeypatch.setenv("MSGPU_FMAX", "5")
    p = plan_batch(units(long_units), 80 * GIB)
    assert (p["fmax"], p["frame_slots"], p["rounds"]) == (5, 10, 13107)
    # a chain is never cut by a wave boundary: 3 000 one-block units of which the last 2 500 are one chain, budget for ~1 500 units
    monkeypatch.delenv("MSGPU_FMAX", raising=False)
    spec = [(CODEC_MSZIP, 32768, 0)] * 500 + [(CODEC_MSZIP, 32768, 0x4)] + [(CODEC_MSZIP, 32768, 0x8)] * 2499
    with pytest.raises(RuntimeError):
        plan_batch(units(spec), 1500 * (slot_bytes + 10368))          # the chain alone does not fit: MSGPU_ERR_NOMEMORY, as the decode call says
    p = plan_batch(units(spec), 2600 * (slot_bytes + 10368))
    assert p["waves"] == 2
