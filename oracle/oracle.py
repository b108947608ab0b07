"""oracle.py - TEST INFRASTRUCTURE ONLY: ctypes loader for the CPU oracles.

Two libraries can sit in oracle/_ref/ (built by oracle/Makefile):

* ``libmspack_ref.so``  - the reference's own lzxd.c / qtmd.c / mszipd.c compiled where they lie
  under /root/reference, driven through oracle/ref_harness.c  (kind "reference").
* ``libmspack_port.so`` - oracle/port/mspack_port.c, this repository's plain-C restatement of the
  same algorithms (kind "port").

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.  The product (libmspack_b200) never does.
"""
from __future__ import annotations

import ctypes
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# must match include/msgpu.h `msgpu_unit` (32 bytes, no padding)
UNIT_DTYPE = np.dtype([
    ("codec", "u1"), ("window_bits", "u1"), ("reset_interval", "<u2"), ("flags", "<u4"),
    ("in_off", "<u8"), ("in_len", "<u4"), ("out_len", "<u4"), ("out_off", "<u8"),
])
assert UNIT_DTYPE.itemsize == 32

CODEC_MSZIP, CODEC_QUANTUM, CODEC_LZX = 1, 2, 3


class Oracle:
    """One loaded oracle library (kind = 'reference' or 'port')."""

    def __init__(self, kind: str):
        name = {"reference": "libmspack_ref.so", "port": "libmspack_port.so"}[kind]
        path = os.path.join(HERE, "_ref", name)
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing - run `make -C oracle` (or __graft_entry__.build())")
        self.kind = kind
        self.path = path
        self.lib = ctypes.CDLL(path)
        pre = "oracle_ref" if kind == "reference" else "oracle_port"
        self._decode = getattr(self.lib, pre + "_decode")
        self._decode.restype = ctypes.c_int
        self._decode.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        self._batch = getattr(self.lib, pre + "_decode_batch")
        self._batch.restype = ctypes.c_double
        self._batch.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p,
                                ctypes.c_void_p, ctypes.c_int]
        self._decode_only = getattr(self.lib, pre + "_last_decode_only", None)      # (reference harness only)
        if self._decode_only is not None:
            self._decode_only.restype = ctypes.c_double
            self._decode_only.argtypes = []

    def last_decode_only_seconds(self):
        """Seconds the slowest thread of the last decode_batch spent inside X_decompress (no X_init / X_free); None for the port."""
        return float(self._decode_only()) if self._decode_only is not None else None

    def decode_batch(self, units: np.ndarray, in_bytes: np.ndarray, out_size: int | None = None,
                     threads: int = 1, out_init: np.ndarray | None = None):
        """Decode every unit.  Returns (out uint8[out_size], status int32[n], seconds).
        out_init: initial contents of the output buffer (LZX DELTA reference data in front of the units)."""
        units = np.ascontiguousarray(units, dtype=UNIT_DTYPE)
        in_bytes = np.ascontiguousarray(in_bytes, dtype=np.uint8)
        if out_size is None:
            out_size = int((units["out_off"] + units["out_len"]).max()) if len(units) else 0
        out = np.zeros(max(out_size, 1), dtype=np.uint8)
        if out_init is not None:
            out[:len(out_init)] = out_init
        status = np.full(len(units), -1, dtype=np.int32)
        secs = self._batch(units.ctypes.data, len(units), in_bytes.ctypes.data, out.ctypes.data,
                           status.ctypes.data, int(threads))
        return out[:out_size], status, float(secs)

    def decode_one(self, codec: int, data: bytes, out_len: int, window_bits: int = 0,
                   reset_interval: int = 0, flags: int = 0):
        """Decode a single stream.  Returns (bytes, err)."""
        u = np.zeros(1, dtype=UNIT_DTYPE)
        u["codec"], u["window_bits"], u["reset_interval"], u["flags"] = codec, window_bits, reset_interval, flags
        u["in_len"], u["out_len"] = len(data), out_len
        buf = np.frombuffer(bytes(data) + b"\0", dtype=np.uint8)
        out, st, _ = self.decode_batch(u, buf, out_len)
        return out.tobytes(), int(st[0])


def load(prefer: str = "reference") -> Oracle:
    """Load the preferred oracle, falling back to the other kind if it is not built."""
    order = [prefer] + [k for k in ("reference", "port") if k != prefer]
    last = None
    for k in order:
        try:
            return Oracle(k)
        except (FileNotFoundError, OSError) as e:  # pragma: no cover - depends on build state
            last = e
    raise last
