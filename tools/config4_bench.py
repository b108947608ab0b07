"""One GPU's share of BASELINE configs[3]: 131 072 LZX wb21 reset intervals of 64 KiB (8 GiB of output per GPU)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from libmspack_b200 import gen
from libmspack_b200.codec import BatchDecoder
from libmspack_b200.units import CODEC_LZX
n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
b = gen.make_batch(CODEC_LZX, n, unit_bytes=65536, reset_interval=2, slack=4, keep_raw=True, threads=os.cpu_count())
dec = BatchDecoder(0); stream = torch.cuda.Stream()
d_in = torch.from_numpy(b.comp).cuda(); d_out = torch.zeros(b.out_bytes, dtype=torch.uint8, device="cuda"); d_st = torch.full((n,), -1, dtype=torch.int32, device="cuda")
torch.cuda.synchronize(); best = 1e9
for _ in range(3):
    dec.decode_device(b.units, d_in, d_out, d_st, stream); torch.cuda.synchronize(); best = min(best, dec.last_kernel_ms())
ok = bool((d_st == 0).all().item()) and bool(torch.equal(d_out.cpu(), torch.from_numpy(b.raw)))
print(json.dumps({"config": "configs[3] per-GPU share: LZX wb21, reset_interval 2 frames, 64 KiB units", "units": n, "out_bytes": n * 65536, "kernels_ms": round(best, 3),
                  "GB_per_s": round(n * 65536 / best / 1e6, 2), "verified": ok, "scratch_GiB": round(dec.scratch_bytes / 2**30, 1)}))
