"""A/B of the Quantum P1 kernel forms on one generated batch (development aid; bench.py is the contract).
usage: qtm_ab.py [units]   - MSGPU_QTM_CONV=0 / 1 (msgpu_create reads it), stage timing, round trip against the generator's data"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from libmspack_b200 import gen
from libmspack_b200.codec import BatchDecoder

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
t0 = time.time(); b = gen.make_batch(2, n, keep_raw=True); print(f"generated {n} Quantum units in {time.time() - t0:.1f} s, ratio {b.in_bytes / b.out_bytes:.3f}")
d_in = torch.from_numpy(b.comp).cuda(); d_out = torch.zeros(b.out_bytes, dtype=torch.uint8, device="cuda"); d_st = torch.zeros(n, dtype=torch.int32, device="cuda")
stream = torch.cuda.Stream()
for conv in (0, 1, 0, 1):
    os.environ["MSGPU_QTM_CONV"] = str(conv)
    dec = BatchDecoder(0)
    dec.set_stage_timing(True)
    best = (1e9, 0, 0)
    for it in range(3):
        d_out.zero_(); torch.cuda.synchronize()
        dec.decode_device(b.units, d_in, d_out, d_st, stream); torch.cuda.synchronize()
        best = min(best, (dec.last_kernel_ms(), dec.stage_ms(0), dec.stage_ms(1)))
    ok = bool((d_st == 0).all().item()) and np.array_equal(d_out.cpu().numpy(), b.raw)
    print(f"MSGPU_QTM_CONV={conv}: total {best[0]:.3f} ms  P1 {best[1]:.3f} ms  P2 {best[2]:.3f} ms  -> {b.out_bytes / best[0] / 1e6:.1f} GB/s  roundtrip_ok={ok}")
    dec.close()
