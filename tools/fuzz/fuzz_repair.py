"""MSZIP repair mode (mszipd_init(repair_mode=1)): damaged multi-block folders, input buffer sizes from 2 bytes to 4 KiB: the device
logic (host emulation) against the reference - status and bytes.  usage: fuzz_repair.py [seed] [rounds]
Development aid (CPU only): TEST INFRASTRUCTURE, like everything that loads oracle/.  Run from the repository root."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np, ctypes
from libmspack_b200 import gen, build
from oracle import oracle as orc
ref = orc.load("reference")
lib = ctypes.CDLL(build.build_emul())
lib.emul_decode_batch.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
def emul(units, comp, out_bytes, mode=1):
    units = np.ascontiguousarray(units); comp = np.concatenate([np.ascontiguousarray(comp, dtype=np.uint8), np.zeros(64, np.uint8)])
    out = np.zeros(out_bytes + 128, np.uint8); base = (-out.ctypes.data) % 16; st = np.full(len(units), -1, np.int32)
    lib.emul_decode_batch(units.ctypes.data, len(units), comp.ctypes.data, out.ctypes.data + base, st.ctypes.data, mode)
    return out[base:base+out_bytes], st
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
bad = 0; total = 0; okc = 0
for trial in range(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    nblk = int(rng.integers(2, 7)); ub = 32768 * nblk
    b = gen.make_batch(1, 16, unit_bytes=ub, first_unit=int(rng.integers(0, 1 << 20)), data=("text", "binary")[trial % 2], level=(6, 1, 0)[trial % 3])
    comp = b.comp.copy(); units = b.units.copy()
    bufsize = int(rng.choice([4096, 4096, 2048, 512, 64, 16, 6, 2]))
    units["flags"] = 0x1 | (bufsize << 6)
    for i, u in enumerate(units):
        lo, n = int(u["in_off"]), int(u["in_len"])
        for _ in range(int(rng.integers(1, 4))):
            k = int(rng.integers(0, 4)); pos = lo + int(rng.integers(2, n))
            if k == 0: comp[pos] ^= 1 << int(rng.integers(0, 8))
            elif k == 1: comp[pos:pos + 16] = rng.integers(0, 256, min(16, lo + n - pos), dtype=np.uint8)
            elif k == 2: comp[pos:min(pos + 300, lo + n)] = 0
            else: units["in_len"][i] = max(4, n - int(rng.integers(1, 3000)))
    o1, s1, _ = ref.decode_batch(units, comp, b.out_bytes, threads=8)
    for mode in (1, 2):
        o2, s2 = emul(units, comp, b.out_bytes, mode)
        for i, u in enumerate(units):
            lo, n = int(u["out_off"]), int(u["out_len"]); total += 1; okc += int(s1[i] == 0)
            if s1[i] != s2[i] or (s1[i] == 0 and not np.array_equal(o1[lo:lo+n], o2[lo:lo+n])):
                bad += 1
                if bad <= 6:
                    d = np.nonzero(o1[lo:lo+n] != o2[lo:lo+n])[0]
                    print("MISMATCH trial", trial, "unit", i, "mode", mode, "buf", bufsize, "ref", s1[i], "emul", s2[i], "first diff", d[:1], "ndiff", len(d))
print("repair emul-vs-ref:", total, "decodes,", bad, "mismatches;", okc, "with status 0")
