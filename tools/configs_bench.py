"""Device-resident timings of the other BASELINE.json configs at reduced unit counts (development record;
bench.py is the contract and covers configs[2]).  Prints one JSON line per config."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from libmspack_b200 import gen
from libmspack_b200.codec import BatchDecoder
from libmspack_b200.units import CODEC_LZX, CODEC_MSZIP, CODEC_QUANTUM

dec = BatchDecoder(0)
stream = torch.cuda.Stream(); torch.cuda.synchronize()

def run(name, b, iters=3):
    d_in = torch.from_numpy(b.comp).cuda(); d_out = torch.zeros(b.out_bytes, dtype=torch.uint8, device="cuda")
    d_st = torch.full((b.n,), -1, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        dec.decode_device(b.units, d_in, d_out, d_st, stream); torch.cuda.synchronize()
        best = min(best, dec.last_kernel_ms())
    ok = bool((d_st == 0).all().item())
    if b.raw is not None:
        ub = int(b.units["out_len"][0]); stride = (ub + 15) & ~15
        ok = ok and np.array_equal(d_out.cpu().numpy().reshape(b.n, stride)[:, :ub].reshape(-1), b.raw)
    U = int(b.units["out_len"].astype(np.int64).sum())
    print(json.dumps({"config": name, "units": b.n, "out_bytes": U, "kernels_ms": round(best, 3), "GB_per_s": round(U / best / 1e6, 2), "verified": ok}), flush=True)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
run("configs[1]: MSZIP 32 KiB blocks (zlib level 6, Zipf text)", gen.make_batch(CODEC_MSZIP, n, keep_raw=True))
run("configs[2]: LZX wb21, one 32 KiB frame per unit", gen.make_batch(CODEC_LZX, n, keep_raw=True))
run("configs[3]-like: LZX wb21 reset intervals of 64 KiB (reset_interval 2, 4 slack bytes)", gen.make_batch(CODEC_LZX, n // 2, unit_bytes=65536, reset_interval=2, slack=4, keep_raw=True))
parts = [gen.make_batch(c, n // 3, first_unit=k * n) for k, c in enumerate((CODEC_MSZIP, CODEC_LZX, CODEC_QUANTUM))]
mixed = gen.concat_batches(parts)
perm = np.random.default_rng(0x51544D31).permutation(mixed.n)
mixed.units = mixed.units[perm].copy()
run("configs[4]-like: mixed MSZIP / LZX wb21 / Quantum wb21, per-unit dispatch", mixed)
run("Quantum wb21 only", gen.make_batch(CODEC_QUANTUM, n // 2, keep_raw=True))
