"""Python binding of the CHM section front end (include/msgpu_chm.h) - plumbing for tests."""
from __future__ import annotations

import ctypes

import numpy as np

from .codec import load_library
from .units import UNIT_DTYPE

CHM_SYMBOLS = ["msgpu_chm_units"]


class ChmInfo(ctypes.Structure):
    _fields_ = [("window_bits", ctypes.c_uint32), ("reset_interval", ctypes.c_uint32), ("uncomp_len", ctypes.c_uint64),
                ("padded_len", ctypes.c_uint64), ("num_units", ctypes.c_uint64)]


def chm_units(control_data: bytes, reset_table: bytes, content_bytes: int):
    """-> (rc, info, units).  rc != 0: the MSGPU_ERR_* chmd.c would report for these system files."""
    lib = load_library()
    lib.msgpu_chm_units.restype = ctypes.c_int
    lib.msgpu_chm_units.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_uint64,
                                    ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ChmInfo)]
    info = ChmInfo()
    rc = lib.msgpu_chm_units(control_data, len(control_data), reset_table, len(reset_table), content_bytes, None, 0, ctypes.byref(info))
    if rc:
        return rc, info, None
    units = np.zeros(int(info.num_units), dtype=UNIT_DTYPE)
    rc = lib.msgpu_chm_units(control_data, len(control_data), reset_table, len(reset_table), content_bytes,
                             units.ctypes.data if len(units) else None, len(units), ctypes.byref(info))
    return rc, info, units
