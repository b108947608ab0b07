"""KWAJ/MSZIP streams (random block lengths, damage, truncation, missing terminator, tight or too small output areas): the device
logic (host emulation) against the reference's mszipd_decompress_kwaj - status, produced size, bytes.  usage: fuzz_kwaj.py [seed] [cases]
Development aid (CPU only): TEST INFRASTRUCTURE, like everything that loads oracle/.  Run from the repository root."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np, ctypes
from libmspack_b200 import build
from libmspack_b200.units import UNIT_DTYPE
from oracle import oracle as orc
from util import kwaj_mszip_stream
ref = orc.load("reference")
lib = ctypes.CDLL(build.build_emul())
lib.emul_decode_batch.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
lib.emul_last_produced.restype = ctypes.c_uint32
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
bad = 0; total = 0
for trial in range(int(sys.argv[2]) if len(sys.argv) > 2 else 200):
    lens = [int(x) for x in rng.integers(1, 32769, int(rng.integers(1, 6)))]
    if rng.random() < 0.3: lens = [32768] * len(lens)
    comp, out = kwaj_mszip_stream(lens, seed=int(rng.integers(0, 1000)), terminator=bool(rng.random() < 0.8))
    comp = bytearray(comp)
    k = int(rng.integers(0, 5))
    if k == 0: comp[int(rng.integers(0, len(comp)))] ^= 1 << int(rng.integers(0, 8))
    elif k == 1: comp = comp[:max(1, len(comp) - int(rng.integers(1, 200)))]
    elif k == 2:
        p = int(rng.integers(0, len(comp))); comp[p:p + 8] = bytes(rng.integers(0, 256, min(8, len(comp) - p), dtype=np.uint8))
    cap = int(sum(lens)) + int(rng.choice([0, 16, 40000])) if rng.random() < 0.8 else max(1, int(sum(lens)) // 2)
    u = np.zeros(1, dtype=UNIT_DTYPE); u[0] = (1, 0, 0, 0x10, 0, len(comp), cap, 0)
    buf = np.frombuffer(bytes(comp) + b"\0" * 80, dtype=np.uint8).copy()
    BIG = 400000
    ub = u.copy(); ub["out_len"] = BIG
    o1 = np.zeros(BIG + 64, np.uint8); prod = ctypes.c_uint32(0)
    s1 = ref._decode(ub.ctypes.data, buf.ctypes.data, o1.ctypes.data, ctypes.byref(prod))
    for mode in (1, 2):
        o2 = np.zeros(cap + 192, np.uint8); base = (-o2.ctypes.data) % 16; st = np.full(1, -1, np.int32)
        lib.emul_decode_batch(u.ctypes.data, 1, buf.ctypes.data, o2.ctypes.data + base, st.ctypes.data, mode)
        p2 = lib.emul_last_produced(); total += 1
        # the reference's writer silently drops what does not fit; the device says CAPACITY instead
        truncated = prod.value > cap
        if truncated:
            ok = st[0] == 101 or (s1 != 0 and st[0] == s1 and p2 <= cap)      # (an error in front of the block that would not fit is the error)
        else:
            ok = st[0] == s1 and p2 == prod.value and np.array_equal(o2[base:base + prod.value], o1[:prod.value])
        if not ok:
            bad += 1
            if bad < 6: print("MISMATCH trial", trial, "mode", mode, "ref", s1, prod.value, "emul", st[0], p2, "cap", cap, "kind", k)
print("kwaj fuzz:", total, "decodes,", bad, "mismatches")
