"""CPU-only: the SAME per-lane P1 state machines and per-lane P2 functions the CUDA kernels run
(libmspack_b200/csrc/*.cuh compiled as plain C++, tests/emul/emul.cpp) against the oracle.  This checks the
decoder logic without a GPU; the GPU parity tests (-m gpu) check the real kernels through the C-ABI."""
import ctypes
import hashlib
import os

import numpy as np
import pytest

from libmspack_b200 import gen
from libmspack_b200.units import CODEC_LZX, CODEC_MSZIP, CODEC_QUANTUM
from util import assert_same, golden_manifest, golden_unit

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emul():
    from libmspack_b200 import build
    so = build.build_emul()
    lib = ctypes.CDLL(so)
    lib.emul_decode_batch.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]

    def run(units, comp, out_bytes, frames_per_round=1, out_shift=0):
        units = np.ascontiguousarray(units)
        comp = np.concatenate([np.ascontiguousarray(comp, dtype=np.uint8), np.zeros(64, np.uint8)])
        out = np.zeros(out_bytes + 128, np.uint8)
        base = (-out.ctypes.data) % 16 + out_shift               # 16-byte aligned + the requested misalignment
        st = np.full(len(units), -1, np.int32)
        lib.emul_decode_batch(units.ctypes.data, len(units), comp.ctypes.data, out.ctypes.data + base, st.ctypes.data, frames_per_round)
        return out[base:base + out_bytes], st
    return run


@pytest.mark.parametrize("entry", golden_manifest(), ids=lambda e: e["name"])
@pytest.mark.parametrize("frames_per_round", [1, 2])
def test_device_logic_on_golden_vectors(emul, entry, frames_per_round):
    u, comp = golden_unit(entry)
    out, st = emul(u, comp, entry["out_len"], frames_per_round)
    assert int(st[0]) == entry["err"]
    if entry["err"] == 0:
        assert hashlib.md5(out.tobytes()).hexdigest() == entry["md5"]


CASES = [(CODEC_MSZIP, dict()), (CODEC_MSZIP, dict(data="random", unit_bytes=40000)), (CODEC_MSZIP, dict(unit_bytes=65536, level=1)),
         (CODEC_QUANTUM, dict()), (CODEC_QUANTUM, dict(window_bits=12, unit_bytes=65536, data="binary")),
         (CODEC_LZX, dict()), (CODEC_LZX, dict(block_mode=4, split=3)), (CODEC_LZX, dict(unit_bytes=65536, reset_interval=2, block_mode=4)),
         (CODEC_LZX, dict(intel=1, data="binary", unit_bytes=70000)), (CODEC_LZX, dict(window_bits=15, unit_bytes=100000, block_mode=4))]


@pytest.mark.parametrize("codec,kw", CASES, ids=lambda x: str(x))
def test_device_logic_matches_oracle(emul, oracle_ref, codec, kw):
    b = gen.make_batch(codec, 20, **kw)
    o1, s1, _ = oracle_ref.decode_batch(b.units, b.comp, b.out_bytes, threads=4)
    for fpr in (1, 2):
        o2, s2 = emul(b.units, b.comp, b.out_bytes, fpr)
        assert_same(b.units, o1, s1, o2, s2, f"emulation {codec} {kw} F={fpr}")


def test_device_logic_on_corrupt_streams(emul, oracle_ref):
    rng = np.random.default_rng(5)
    for codec in (CODEC_MSZIP, CODEC_LZX, CODEC_QUANTUM):
        b = gen.make_batch(codec, 48)
        comp = b.comp.copy()
        for i, u in enumerate(b.units):
            lo, n = int(u["in_off"]), int(u["in_len"])
            if i % 2 == 0:
                comp[lo + int(rng.integers(0, n))] ^= 1 << int(rng.integers(0, 8))
            else:
                b.units["in_len"][i] = max(1, n - int(rng.integers(1, 40)))
        o1, s1, _ = oracle_ref.decode_batch(b.units, comp, b.out_bytes)
        o2, s2 = emul(b.units, comp, b.out_bytes, 1)
        assert_same(b.units, o1, s1, o2, s2, f"corrupt {codec}")


def _shifted(b, shift):
    """The same batch with every unit's input moved `shift` bytes (unit inputs no longer 4-byte aligned)."""
    comp = np.concatenate([np.zeros(shift, np.uint8), b.comp])
    units = b.units.copy()
    units["in_off"] += np.uint64(shift)
    return units, comp


@pytest.mark.parametrize("shift", [1, 2, 3])
def test_device_logic_unaligned_input(emul, oracle_ref, shift):
    """Unit inputs at any byte offset: the word loads of the bit readers fall back to byte loads."""
    for codec, kw in ((CODEC_LZX, dict(block_mode=4, split=2)), (CODEC_MSZIP, dict(data="random", unit_bytes=40000)), (CODEC_QUANTUM, dict())):
        b = gen.make_batch(codec, 12, **kw)
        units, comp = _shifted(b, shift)
        o1, s1, _ = oracle_ref.decode_batch(units, comp, b.out_bytes, threads=4)
        o2, s2 = emul(units, comp, b.out_bytes, 1)
        assert_same(units, o1, s1, o2, s2, f"unaligned {codec} shift {shift}")


@pytest.mark.parametrize("shift", [1, 2])
def test_device_logic_unaligned_output(emul, oracle_ref, shift):
    """An output base that is not 4/16-byte aligned: literal words and resolved 16-byte groups fall back to byte stores."""
    for codec, kw in ((CODEC_LZX, dict(block_mode=4, split=2, unit_bytes=40000)), (CODEC_MSZIP, dict(data="random", unit_bytes=40000)),
                      (CODEC_MSZIP, dict()), (CODEC_QUANTUM, dict())):
        b = gen.make_batch(codec, 8, **kw)
        o1, s1, _ = oracle_ref.decode_batch(b.units, b.comp, b.out_bytes, threads=4)
        o2, s2 = emul(b.units, b.comp, b.out_bytes, 1, out_shift=shift)
        assert_same(b.units, o1, s1, o2, s2, f"unaligned output {codec} shift {shift}")
