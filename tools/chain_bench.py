"""Cabinet with long MSZIP folders through the cabinet front end: block chains against one stream per folder
(MSGPU_CAB_NOCHAIN=1).  Development record; prints one JSON line."""
import json, os, sys, time, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from libmspack_b200 import cab, gen
from libmspack_b200.codec import BatchDecoder
from cabfile import build_cab

nblocks = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nfolders = int(sys.argv[2]) if len(sys.argv) > 2 else 4
folders, raws = [], []
for k in range(nfolders):
    n = 32768 * nblocks - 77 * k
    raw = gen.raw_units(1, n, first_unit=1000 * k).tobytes()
    blocks = []
    for off in range(0, n, 32768):
        kw = {"zdict": raw[off - 32768:off]} if off else {}
        c = zlib.compressobj(6, zlib.DEFLATED, -15, **kw)
        blocks.append((b"CK" + c.compress(raw[off:off + 32768]) + c.flush(), len(raw[off:off + 32768])))
    folders.append(dict(comp_type=1, blocks=blocks, files=[(f"f{k}.bin", 0, n)]))
    raws.append(raw)
img = build_cab(folders)
plan = cab.scan(img)
dec = BatchDecoder(0)
res = {}
for mode in ("chain", "one_stream"):
    if mode == "one_stream":
        os.environ["MSGPU_CAB_NOCHAIN"] = "1"
    best = 1e9
    for it in range(3):
        t0 = time.perf_counter(); out, st = plan.decode(dec); dt = time.perf_counter() - t0
        best = min(best, dt)
    ok = bool((st == 0).all()) and all(out[int(plan.folders["out_off"][k]):int(plan.folders["out_off"][k]) + len(r)].tobytes() == r for k, r in enumerate(raws))
    res[mode] = {"seconds": round(best, 4), "MB_per_s": round(sum(len(r) for r in raws) / best / 1e6, 1), "verified": ok}
print(json.dumps({"workload": f"{nfolders} MSZIP folders x {nblocks} CFDATA blocks, msgpu_cab_decode_host (host image in, host bytes out)", **res}))
