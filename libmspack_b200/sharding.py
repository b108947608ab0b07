"""Multi-GPU layer.  Units are independent (a CAB folder / a CHM reset interval never reads another unit's bytes),
so R devices share a batch by unit index - shard r owns units [floor(r n / R), floor((r + 1) n / R)), moved forward where that would
cut an MSZIP block chain - and there is no data-path collective.  Two ways to use it:

* one process, several devices: `MultiDecoder` -> `msgpu_decode_batch_host_multi` (include/msgpu.h): one thread and one context
  per device inside the C library, host buffers in, host buffers out;
* one process per device (torchrun): every rank calls `shard_range` for its slice and decodes it with its own `BatchDecoder`;
  torch.distributed then carries only the control plane (barriers, max-over-ranks timing, bench.py's optional output gather).
"""
from __future__ import annotations

import ctypes

import numpy as np

from .units import UNIT_DTYPE


def shard_range(n: int, rank: int, world: int, units: np.ndarray | None = None):
    """Half-open unit-index range owned by `rank` (msgpu_shard_range when a unit table is given: chains stay whole)."""
    if units is None:
        return (n * rank) // world, (n * (rank + 1)) // world
    from .codec import load_library
    units = np.ascontiguousarray(units, dtype=UNIT_DTYPE)
    lo, hi = ctypes.c_size_t(0), ctypes.c_size_t(0)
    rc = load_library().msgpu_shard_range(units.ctypes.data, len(units), rank, world, ctypes.byref(lo), ctypes.byref(hi))
    if rc:
        raise ValueError(f"msgpu_shard_range: error {rc}")
    return int(lo.value), int(hi.value)


class MultiDecoder:
    """One msgpu context per CUDA device; one call decodes a host-resident batch on all of them."""

    def __init__(self, devices):
        from .codec import BatchDecoder
        self.decoders = [BatchDecoder(d) for d in devices]
        self.lib = self.decoders[0].lib
        self._ctxs = (ctypes.c_void_p * len(self.decoders))(*[d.ctx for d in self.decoders])

    def close(self):
        for d in self.decoders:
            d.close()
        self.decoders = []

    @property
    def launches(self) -> int:
        return sum(d.launches for d in self.decoders)

    def decode_host_into(self, units: np.ndarray, comp_ptr: int, comp_bytes: int, out_ptr: int, out_bytes: int, status: np.ndarray) -> None:
        units = np.ascontiguousarray(units, dtype=UNIT_DTYPE)
        rc = self.lib.msgpu_decode_batch_host_multi(self._ctxs, len(self.decoders), units.ctypes.data, len(units), ctypes.c_void_p(comp_ptr), comp_bytes,
                                                    ctypes.c_void_p(out_ptr), out_bytes, status.ctypes.data)
        if rc:
            raise RuntimeError(f"msgpu_decode_batch_host_multi: error {rc}: " + "; ".join(d.lib.msgpu_last_error(d.ctx).decode() for d in self.decoders))

    def decode_host(self, units: np.ndarray, comp: np.ndarray, out_bytes: int, out_init: np.ndarray | None = None):
        comp = np.ascontiguousarray(comp, dtype=np.uint8)
        out = np.empty(max(out_bytes, 1), dtype=np.uint8)
        if out_init is not None:
            out[:len(out_init)] = out_init
        status = np.full(len(units), -1, dtype=np.int32)
        self.decode_host_into(units, comp.ctypes.data, comp.size, out.ctypes.data, out_bytes, status)
        return out[:out_bytes], status
