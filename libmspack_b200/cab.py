"""Python binding of the cabinet front end (include/msgpu_cab.h) - plumbing for tests.

`scan()` parses the headers of a .cab image on the host (no GPU needed); `CabPlan.decode()` runs the CFDATA framing and the
checksums on the device and decodes every folder as one batch.
"""
from __future__ import annotations

import ctypes

import numpy as np

from .codec import load_library

CAB_SYMBOLS = ["msgpu_cab_scan", "msgpu_cab_free", "msgpu_cab_num_folders", "msgpu_cab_num_blocks", "msgpu_cab_num_files",
               "msgpu_cab_folders", "msgpu_cab_blocks", "msgpu_cab_files", "msgpu_cab_out_bytes", "msgpu_cab_packed_bytes",
               "msgpu_cab_decode_host", "msgpu_cab_scan_set", "msgpu_cab_decode_host_set"]
SALVAGE = 1          # MSGPU_CAB_SALVAGE

FOLDER_DTYPE = np.dtype([("comp_type", "<u2"), ("codec", "u1"), ("window_bits", "u1"), ("num_blocks", "<u4"), ("first_block", "<u4"),
                         ("scan_status", "<i4"), ("bad_block", "<u4"), ("_pad", "<u4"), ("out_off", "<u8"), ("out_len", "<u8"),
                         ("in_off", "<u8"), ("in_len", "<u8")])
BLOCK_DTYPE = np.dtype([("payload_off", "<u8"), ("dst_off", "<u8"), ("checksum", "<u4"), ("comp_len", "<u2"), ("uncomp_len", "<u2"),
                        ("folder", "<u4"), ("flags", "<u4")])
FILE_DTYPE = np.dtype([("folder", "<u4"), ("offset", "<u4"), ("length", "<u4"), ("name_off", "<u4")])

_declared = False


def _lib():
    global _declared
    lib = load_library()
    if not _declared:
        vp, sz = ctypes.c_void_p, ctypes.c_size_t
        lib.msgpu_cab_scan.restype = vp
        lib.msgpu_cab_scan.argtypes = [vp, sz, ctypes.POINTER(ctypes.c_int)]
        lib.msgpu_cab_free.restype = None
        lib.msgpu_cab_free.argtypes = [vp]
        for name in ("msgpu_cab_num_folders", "msgpu_cab_num_blocks", "msgpu_cab_num_files", "msgpu_cab_out_bytes", "msgpu_cab_packed_bytes"):
            getattr(lib, name).restype = sz
            getattr(lib, name).argtypes = [vp]
        for name in ("msgpu_cab_folders", "msgpu_cab_blocks", "msgpu_cab_files"):
            getattr(lib, name).restype = vp
            getattr(lib, name).argtypes = [vp]
        lib.msgpu_cab_decode_host.restype = ctypes.c_int
        lib.msgpu_cab_decode_host.argtypes = [vp, vp, vp, sz, vp, sz, vp]
        lib.msgpu_cab_scan_set.restype = vp
        lib.msgpu_cab_scan_set.argtypes = [vp, vp, sz, ctypes.c_uint32, ctypes.POINTER(ctypes.c_int)]
        lib.msgpu_cab_decode_host_set.restype = ctypes.c_int
        lib.msgpu_cab_decode_host_set.argtypes = [vp, vp, vp, vp, sz, vp, sz, vp]
        _declared = True
    return lib


class CabError(Exception):
    def __init__(self, code):
        super().__init__(f"msgpu_cab_scan failed with MSGPU_ERR {code}")
        self.code = code


def _table(ptr, n, dtype):
    if not n:
        return np.zeros(0, dtype)
    buf = (ctypes.c_uint8 * (n * dtype.itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).copy()


class CabPlan:
    """One cabinet (image: bytes) or a set of cabinets in order (image: list of bytes); flags: SALVAGE."""

    def __init__(self, image, flags: int = 0):
        self.lib = _lib()
        images = [image] if isinstance(image, (bytes, bytearray, memoryview)) else list(image)
        self.images = [np.frombuffer(bytes(im), dtype=np.uint8) for im in images]
        self.image = self.images[0]
        self._ptrs = (ctypes.c_void_p * len(self.images))(*[im.ctypes.data for im in self.images])
        self._sizes = (ctypes.c_size_t * len(self.images))(*[im.size for im in self.images])
        err = ctypes.c_int(0)
        if len(self.images) == 1 and not flags:
            self.ptr = self.lib.msgpu_cab_scan(self.image.ctypes.data, self.image.size, ctypes.byref(err))
        else:
            self.ptr = self.lib.msgpu_cab_scan_set(self._ptrs, self._sizes, len(self.images), flags, ctypes.byref(err))
        if not self.ptr:
            raise CabError(err.value)
        assert FOLDER_DTYPE.itemsize == 56 and BLOCK_DTYPE.itemsize == 32 and FILE_DTYPE.itemsize == 16
        self.folders = _table(self.lib.msgpu_cab_folders(self.ptr), self.lib.msgpu_cab_num_folders(self.ptr), FOLDER_DTYPE)
        self.blocks = _table(self.lib.msgpu_cab_blocks(self.ptr), self.lib.msgpu_cab_num_blocks(self.ptr), BLOCK_DTYPE)
        self.files = _table(self.lib.msgpu_cab_files(self.ptr), self.lib.msgpu_cab_num_files(self.ptr), FILE_DTYPE)
        self.out_bytes = self.lib.msgpu_cab_out_bytes(self.ptr)
        self.packed_bytes = self.lib.msgpu_cab_packed_bytes(self.ptr)

    def file_name(self, i: int) -> bytes:
        off = int(self.files["name_off"][i])
        return bytes(self.image[off:off + 256]).split(b"\0", 1)[0]

    def decode(self, decoder):
        """-> (out uint8[out_bytes], status int32[num_folders]); folder f's bytes are out[out_off : out_off + out_len]."""
        out = np.zeros(max(self.out_bytes, 1), np.uint8)
        st = np.full(len(self.folders), -1, np.int32)
        rc = self.lib.msgpu_cab_decode_host_set(decoder.ctx, self.ptr, self._ptrs, self._sizes, len(self.images), out.ctypes.data, out.size,
                                                st.ctypes.data)
        if rc:
            raise RuntimeError(f"msgpu_cab_decode_host failed: {rc}")
        return out[:self.out_bytes], st

    def close(self):
        if getattr(self, "ptr", None):
            self.lib.msgpu_cab_free(self.ptr)
            self.ptr = None

    def __del__(self):
        self.close()


def scan(image, flags: int = 0) -> CabPlan:
    return CabPlan(image, flags)
