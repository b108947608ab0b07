"""Random mutations of generated units (bit flips, truncation, header bytes, shorter requests): the device logic (host emulation of
the per-lane kernels code, tests/emul) against the reference's decoders - status and bytes.  usage: fuzz_units.py [seed] [rounds]
Development aid (CPU only): TEST INFRASTRUCTURE, like everything that loads oracle/.  Run from the repository root."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np, ctypes, zlib, time
from libmspack_b200 import gen, build
from libmspack_b200.units import UNIT_DTYPE
from oracle import oracle as orc
ref = orc.load("reference")
lib = ctypes.CDLL(build.build_emul())
lib.emul_decode_batch.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
def emul(units, comp, out_bytes, mode=1, out_init=None):
    units = np.ascontiguousarray(units); comp = np.concatenate([np.ascontiguousarray(comp, dtype=np.uint8), np.zeros(64, np.uint8)])
    out = np.zeros(out_bytes + 128, np.uint8); base = (-out.ctypes.data) % 16
    if out_init is not None: out[base:base+len(out_init)] = out_init
    st = np.full(len(units), -1, np.int32)
    lib.emul_decode_batch(units.ctypes.data, len(units), comp.ctypes.data, out.ctypes.data + base, st.ctypes.data, mode)
    return out[base:base+out_bytes], st
def compare(units, comp, out_bytes, what, modes=(1,), out_init=None):
    o1, s1, _ = ref.decode_batch(units, comp, out_bytes, threads=8, out_init=out_init)
    bad = 0
    for m in modes:
        o2, s2 = emul(units, comp, out_bytes, m, out_init)
        for i,u in enumerate(units):
            lo,n = int(u["out_off"]), int(u["out_len"])
            if s1[i] != s2[i] or (s1[i]==0 and not np.array_equal(o1[lo:lo+n], o2[lo:lo+n])):
                bad += 1
                if bad <= 3: print("MISMATCH", what, "mode", hex(m), "unit", i, "ref", s1[i], "emul", s2[i])
    return bad
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 0)
t0=time.time(); total=0; bad=0
cases = [(3, dict(block_mode=4, split=3, unit_bytes=20000), (1,0x401,0x4401,0x6401)), (3, dict(block_mode=4, unit_bytes=70000, window_bits=15), (2,0x402,0x4402,0x6402)),
         (3, dict(block_mode=3, unit_bytes=9999), (1,0x401,0x4401,0x6401)), (3, dict(unit_bytes=65536, reset_interval=1, block_mode=4, slack=4), (1,0x401,0x4401,0x6401)),
         (1, dict(unit_bytes=40000), (1,2,0x4001)), (1, dict(unit_bytes=70000, data="random"), (1,0x4001)), (1, dict(unit_bytes=5000, level=1), (1,0x4001)),
         (2, dict(unit_bytes=40000, window_bits=10), (1,2,0x801,0x802,0x4001,0x4002,0x6001)), (2, dict(unit_bytes=20000, window_bits=15, data="binary"), (1,0x801,0x4001,0x2001,0x6001)),
         (3, dict(delta=1, window_bits=17, ref_bytes=5000, unit_bytes=40000, block_mode=4), (1,))]
for rnd in range(int(sys.argv[2]) if len(sys.argv)>2 else 6):
    for codec, kw, modes in cases:
        b = gen.make_batch(codec, 24, first_unit=int(rng.integers(0, 1<<20)), **kw)
        comp = b.comp.copy(); units = b.units.copy()
        for i,u in enumerate(units):
            lo,n = int(u["in_off"]), int(u["in_len"])
            k = int(rng.integers(0, 6))
            if k == 0: comp[lo + int(rng.integers(0, n))] ^= 1 << int(rng.integers(0, 8))
            elif k == 1: units["in_len"][i] = max(1, n - int(rng.integers(1, 300)))
            elif k == 2:
                p = lo + int(rng.integers(0, min(n, 64))); comp[p] = rng.integers(0, 256)          # header area
            elif k == 3:
                for _ in range(int(rng.integers(2, 6))): comp[lo + int(rng.integers(0, n))] = rng.integers(0, 256)
            elif k == 4: units["out_len"][i] = max(1, int(u["out_len"]) - int(rng.integers(1, 5000)))   # ask for less than encoded
            # k == 5: intact
        bad += compare(units, comp, b.out_bytes, f"{codec} {kw}", modes, b.out_init); total += b.n*len(modes)
print("fuzz done", total, "unit-decodes, mismatches", bad, "in %.1fs"%(time.time()-t0))
