"""Host-side Python binding of the C-ABI (include/msgpu.h) - plumbing for tests and bench.

PyTorch is used only for device memory and streams; every decode call goes through
libmsgpu.so.  There is no CPU fallback: if the CUDA library is missing or no GPU is present,
constructing a BatchDecoder raises.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from .units import UNIT_DTYPE

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MSGPU_LIB") or os.path.join(PKG, "libmsgpu.so")      # MSGPU_LIB: an experimental build (tools/variant_bench.py)

# every symbol include/msgpu.h declares
ABI_SYMBOLS = [
    "msgpu_create", "msgpu_destroy", "msgpu_last_error", "msgpu_decode_batch_device",
    "msgpu_decode_batch_device_units", "msgpu_decode_batch_host", "msgpu_launch_count",
    "msgpu_scratch_bytes", "msgpu_last_kernel_ms", "msgpu_set_stage_timing", "msgpu_stage_ms", "msgpu_version", "msgpu_last_produced",
    "msgpu_shard_range", "msgpu_decode_batch_host_multi", "msgpu_digest_device", "msgpu_decode_batch_host_digest", "msgpu_plan_batch",
]

_lib = None


def load_library() -> ctypes.CDLL:
    """dlopen libmsgpu.so (built in-tree by libmspack_b200/build.py) and declare the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -m libmspack_b200.build` "
                           "(the CUDA extension is required, there is no CPU path)")
    lib = ctypes.CDLL(LIB_PATH)
    vp, sz, i32p = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p
    lib.msgpu_create.restype = vp
    lib.msgpu_create.argtypes = [ctypes.c_int]
    lib.msgpu_destroy.restype = None
    lib.msgpu_destroy.argtypes = [vp]
    lib.msgpu_last_error.restype = ctypes.c_char_p
    lib.msgpu_last_error.argtypes = [vp]
    lib.msgpu_decode_batch_device.restype = ctypes.c_int
    lib.msgpu_decode_batch_device.argtypes = [vp, vp, sz, vp, sz, vp, sz, i32p, vp]
    lib.msgpu_decode_batch_device_units.restype = ctypes.c_int
    lib.msgpu_decode_batch_device_units.argtypes = [vp, vp, sz, vp, sz, vp, sz, i32p, vp]
    lib.msgpu_decode_batch_host.restype = ctypes.c_int
    lib.msgpu_decode_batch_host.argtypes = [vp, vp, sz, vp, sz, vp, sz, i32p]
    lib.msgpu_last_produced.restype = ctypes.c_int
    lib.msgpu_last_produced.argtypes = [vp, vp, sz]
    lib.msgpu_digest_device.restype = ctypes.c_int
    lib.msgpu_digest_device.argtypes = [vp, vp, sz, vp, sz, vp, ctypes.c_int, vp, vp]
    lib.msgpu_decode_batch_host_digest.restype = ctypes.c_int
    lib.msgpu_decode_batch_host_digest.argtypes = [vp, vp, sz, vp, sz, sz, ctypes.c_int, vp, i32p]
    lib.msgpu_shard_range.restype = ctypes.c_int
    lib.msgpu_shard_range.argtypes = [vp, sz, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t)]
    lib.msgpu_decode_batch_host_multi.restype = ctypes.c_int
    lib.msgpu_decode_batch_host_multi.argtypes = [vp, ctypes.c_int, vp, sz, vp, sz, vp, sz, i32p]
    lib.msgpu_launch_count.restype = ctypes.c_uint64
    lib.msgpu_launch_count.argtypes = [vp]
    lib.msgpu_scratch_bytes.restype = ctypes.c_size_t
    lib.msgpu_scratch_bytes.argtypes = [vp]
    lib.msgpu_last_kernel_ms.restype = ctypes.c_float
    lib.msgpu_last_kernel_ms.argtypes = [vp]
    lib.msgpu_set_stage_timing.restype = ctypes.c_int
    lib.msgpu_set_stage_timing.argtypes = [vp, ctypes.c_int]
    lib.msgpu_stage_ms.restype = ctypes.c_float
    lib.msgpu_stage_ms.argtypes = [vp, ctypes.c_int]
    lib.msgpu_version.restype = ctypes.c_char_p
    lib.msgpu_version.argtypes = []
    lib.msgpu_plan_batch.restype = ctypes.c_int
    lib.msgpu_plan_batch.argtypes = [vp, sz, sz, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint32),
                                     ctypes.POINTER(ctypes.c_uint32)]
    _lib = lib
    return lib


def plan_batch(units: np.ndarray, scratch_budget_bytes: int) -> dict:
    """msgpu_plan_batch (host only, no GPU needed): how a batch would be cut under a scratch budget."""
    lib = load_library()
    units = np.ascontiguousarray(units)
    fmax, slots, waves, rounds = ctypes.c_uint32(0), ctypes.c_uint64(0), ctypes.c_uint32(0), ctypes.c_uint32(0)
    rc = lib.msgpu_plan_batch(units.ctypes.data, len(units), int(scratch_budget_bytes), ctypes.byref(fmax), ctypes.byref(slots), ctypes.byref(waves),
                              ctypes.byref(rounds))
    if rc:
        raise RuntimeError(f"msgpu_plan_batch failed: {rc}")
    return dict(fmax=fmax.value, frame_slots=slots.value, waves=waves.value, rounds=rounds.value)


class BatchDecoder:
    """One msgpu context on one CUDA device."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        self.device = int(device)
        self.ctx = self.lib.msgpu_create(self.device)
        if not self.ctx:
            raise RuntimeError("msgpu_create failed: no usable CUDA device (libmspack_b200 has no CPU fallback)")

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.msgpu_destroy(self.ctx)
            self.ctx = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise RuntimeError(f"msgpu error {rc}: {self.lib.msgpu_last_error(self.ctx).decode()}")

    @property
    def launches(self) -> int:
        return int(self.lib.msgpu_launch_count(self.ctx))

    @property
    def scratch_bytes(self) -> int:
        return int(self.lib.msgpu_scratch_bytes(self.ctx))

    def last_kernel_ms(self) -> float:
        return float(self.lib.msgpu_last_kernel_ms(self.ctx))

    def set_stage_timing(self, on: bool) -> None:
        self.lib.msgpu_set_stage_timing(self.ctx, 1 if on else 0)

    def stage_ms(self, stage: int) -> float:
        return float(self.lib.msgpu_stage_ms(self.ctx, stage))

    def decode_device(self, units: np.ndarray, d_in, d_out, d_status=None, stream=None) -> None:
        """Inputs already resident: d_in / d_out / d_status are torch CUDA tensors (uint8 / uint8 / int32)."""
        units = np.ascontiguousarray(units, dtype=UNIT_DTYPE)
        sp = ctypes.c_void_p(stream.cuda_stream) if stream is not None else None
        rc = self.lib.msgpu_decode_batch_device(
            self.ctx, units.ctypes.data, len(units), ctypes.c_void_p(d_in.data_ptr()), d_in.numel(),
            ctypes.c_void_p(d_out.data_ptr()), d_out.numel(),
            ctypes.c_void_p(d_status.data_ptr()) if d_status is not None else None, sp)
        self._check(rc)

    def decode_host(self, units: np.ndarray, comp: np.ndarray, out_bytes: int, out_init: np.ndarray | None = None):
        """Host buffers in, host buffers out (H2D + decode + D2H inside the call).
        out_init: initial contents of the output buffer - LZX DELTA units expect their reference data in front of their
        output (include/msgpu.h MSGPU_FLAG_REF_SHIFT); the call uploads those bytes."""
        units = np.ascontiguousarray(units, dtype=UNIT_DTYPE)
        comp = np.ascontiguousarray(comp, dtype=np.uint8)
        out = np.empty(max(out_bytes, 1), dtype=np.uint8)
        if out_init is not None:
            out[:len(out_init)] = out_init
        status = np.full(len(units), -1, dtype=np.int32)
        rc = self.lib.msgpu_decode_batch_host(self.ctx, units.ctypes.data, len(units), comp.ctypes.data, comp.size,
                                              out.ctypes.data, out_bytes, status.ctypes.data)
        self._check(rc)
        return out[:out_bytes], status

    def decode_host_digest(self, units: np.ndarray, comp: np.ndarray, out_bytes: int, kind: int = 1):
        """Host buffers in, per-unit digests out (kind 1 = MD5: uint8[n, 16]; 2 = CRC-32: uint32[n]); the output stays on the device."""
        units = np.ascontiguousarray(units, dtype=UNIT_DTYPE)
        comp = np.ascontiguousarray(comp, dtype=np.uint8)
        dig = np.zeros((len(units), 16), dtype=np.uint8) if kind == 1 else np.zeros(len(units), dtype=np.uint32)
        status = np.full(len(units), -1, dtype=np.int32)
        self._check(self.lib.msgpu_decode_batch_host_digest(self.ctx, units.ctypes.data, len(units), comp.ctypes.data, comp.size, out_bytes, kind,
                                                            dig.ctypes.data, status.ctypes.data))
        return dig, status

    def decode_host_digest_into(self, units: np.ndarray, comp_ptr: int, comp_bytes: int, out_bytes: int, kind: int, dig: np.ndarray, status: np.ndarray):
        self._check(self.lib.msgpu_decode_batch_host_digest(self.ctx, units.ctypes.data, len(units), ctypes.c_void_p(comp_ptr), comp_bytes, out_bytes, kind,
                                                            dig.ctypes.data, status.ctypes.data))

    def digest_device(self, units: np.ndarray, d_out, kind: int, d_digest, d_status=None, stream=None) -> None:
        """Digests of units already decoded into the CUDA tensor d_out; d_digest: uint8 CUDA tensor of n * 16 (MD5) / n * 4 (CRC-32) bytes."""
        units = np.ascontiguousarray(units, dtype=UNIT_DTYPE)
        sp = ctypes.c_void_p(stream.cuda_stream) if stream is not None else None
        self._check(self.lib.msgpu_digest_device(self.ctx, units.ctypes.data, len(units), ctypes.c_void_p(d_out.data_ptr()), d_out.numel(),
                                                 ctypes.c_void_p(d_status.data_ptr()) if d_status is not None else None, kind,
                                                 ctypes.c_void_p(d_digest.data_ptr()), sp))

    def last_produced(self, n: int) -> np.ndarray:
        """Bytes each unit of the most recent (single-wave) batch produced - the way to learn a KWAJ unit's size."""
        out = np.zeros(n, dtype=np.uint32)
        self._check(self.lib.msgpu_last_produced(self.ctx, out.ctypes.data, n))
        return out

    def decode_host_into(self, units: np.ndarray, comp_ptr: int, comp_bytes: int, out_ptr: int, out_bytes: int, status: np.ndarray):
        """Same, with caller-owned (e.g. pinned) host buffers given as raw addresses."""
        rc = self.lib.msgpu_decode_batch_host(self.ctx, units.ctypes.data, len(units), ctypes.c_void_p(comp_ptr), comp_bytes,
                                              ctypes.c_void_p(out_ptr), out_bytes, status.ctypes.data)
        self._check(rc)
