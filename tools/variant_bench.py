"""A/B of P1 kernel variants on the headline batch (development aid): one generated batch, one context per MSGPU_LZX_VARIANT,
stage timing P1 / P2, round trip verified.  usage: variant_bench.py [units] [variant ids...]
VB_CODEC=1 / 2 / 3 picks MSZIP (MSGPU_ZIP_VARIANT) / Quantum (MSGPU_QTM_VARIANT: 0, 1) / LZX (default)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from libmspack_b200 import gen
from libmspack_b200 import codec as _codec
from libmspack_b200.codec import BatchDecoder

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
variants = [int(v) for v in sys.argv[2:]] or [11, 20, 21, 22]
data = os.environ.get("VB_DATA", "text")
codec = int(os.environ.get("VB_CODEC", "3"))
b = gen.make_batch(codec, n, keep_raw=True, data=data)
d_in = torch.from_numpy(b.comp).cuda(); d_out = torch.zeros(b.out_bytes, dtype=torch.uint8, device="cuda"); d_st = torch.zeros(n, dtype=torch.int32, device="cuda")
stream = torch.cuda.Stream(); torch.cuda.synchronize()
libs = os.environ.get("VB_LIBS", "").split(",") if os.environ.get("VB_LIBS") else [None]
p2s = [int(x) for x in os.environ.get("VB_P2", "0").split(",")]          # MSGPU_P2_VARIANT values (1 = the byte-parallel pass A)
for lib, v, p2v in [(l, v, q) for l in libs for q in p2s for v in variants]:
    os.environ["MSGPU_LZX_VARIANT" if codec == 3 else ("MSGPU_QTM_VARIANT" if codec == 2 else "MSGPU_ZIP_VARIANT")] = str(v)
    os.environ["MSGPU_P2_VARIANT"] = str(p2v)
    if lib:
        _codec._lib = None; _codec.LIB_PATH = os.path.abspath(lib)       # another build of the library (dlopen keeps both)
    dec = BatchDecoder(0)
    d_out.zero_()
    best = 1e9
    for it in range(6):
        dec.decode_device(b.units, d_in, d_out, d_st, stream); torch.cuda.synchronize()
        best = min(best, dec.last_kernel_ms())
    ok = bool((d_st == 0).all().item()) and np.array_equal(d_out.cpu().numpy(), b.raw)
    dec.set_stage_timing(True)
    p1 = p2 = 1e9
    for it in range(3):
        dec.decode_device(b.units, d_in, d_out, d_st, stream); torch.cuda.synchronize()
        p1 = min(p1, dec.stage_ms(0)); p2 = min(p2, dec.stage_ms(1))
    print(json.dumps({"lib": lib, "variant": v, "p2_variant": p2v, "units": n, "data": data, "codec": codec, "best_ms": round(best, 3), "GB_per_s": round(b.out_bytes / best / 1e6, 1),
                      "p1_ms": round(p1, 3), "p2_ms": round(p2, 3), "verified": ok}), flush=True)
    dec.close()
