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
