"""A/B of kernel builds / knobs on one generated batch (development aid): one context per job, stage timing P1 / P2, round trip verified.
usage: variant_bench.py [units] [job ...]      a job is a comma-separated list of ENV=value settings ("-" = none), e.g. MSGPU_EXP=1
VB_CODEC=1 / 2 / 3 picks MSZIP / Quantum / LZX (default); VB_LIBS=a.so,b.so compares builds; VB_DATA picks the corpus; VB_KW='dict(...)' extra generator arguments.
Every shape runs in a process of its own (VB_ISOLATE=0 turns that off): a shape that faults or hangs costs its own line and a
timeout, not the rest of the run - the batch is generated once and handed over through a file in /dev/shm."""
import json, os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def run_one(units, comp, raw, out_bytes, codec, data, lib, v, p2v):
    import torch
    from libmspack_b200 import codec as _codec
    from libmspack_b200.codec import BatchDecoder
    n = len(units)
    for kv in (v.split(",") if v != "-" else []):
        k, _, val = kv.partition("=")
        os.environ[k] = val
    if lib:
        _codec._lib = None; _codec.LIB_PATH = os.path.abspath(lib)       # another build of the library (dlopen keeps both)
    d_in = torch.from_numpy(comp).cuda(); d_out = torch.zeros(out_bytes, dtype=torch.uint8, device="cuda"); d_st = torch.zeros(n, dtype=torch.int32, device="cuda")
    stream = torch.cuda.Stream(); torch.cuda.synchronize()
    dec = BatchDecoder(0)
    best = 1e9
    for it in range(6):
        dec.decode_device(units, d_in, d_out, d_st, stream); torch.cuda.synchronize()
        best = min(best, dec.last_kernel_ms())
    ok = bool((d_st == 0).all().item()) and np.array_equal(d_out.cpu().numpy(), raw)
    dec.set_stage_timing(True)
    p1 = p2 = 1e9
    for it in range(3):
        dec.decode_device(units, d_in, d_out, d_st, stream); torch.cuda.synchronize()
        p1 = min(p1, dec.stage_ms(0)); p2 = min(p2, dec.stage_ms(1))
    print(json.dumps({"lib": lib, "variant": v, "p2_variant": p2v, "units": n, "data": data, "codec": codec, "best_ms": round(best, 3), "GB_per_s": round(out_bytes / best / 1e6, 1),
                      "p1_ms": round(p1, 3), "p2_ms": round(p2, 3), "verified": ok}), flush=True)
    dec.close()


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        path, codec, data, lib, v, p2v = sys.argv[2], int(sys.argv[3]), sys.argv[4], (sys.argv[5] if sys.argv[5] != "-" else None), sys.argv[6], int(sys.argv[7])
        z = np.load(path)
        run_one(z["units"], z["comp"], z["raw"], int(z["out_bytes"]), codec, data, lib, v, p2v)
        return
    from libmspack_b200 import gen
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    variants = sys.argv[2:] or ["-"]
    data = os.environ.get("VB_DATA", "text")
    codec = int(os.environ.get("VB_CODEC", "3"))
    b = gen.make_batch(codec, n, keep_raw=True, data=data, **eval(os.environ.get("VB_KW", "dict()")))
    libs = os.environ.get("VB_LIBS", "").split(",") if os.environ.get("VB_LIBS") else [None]
    p2s = [0]
    jobs = [(l, v, q) for l in libs for q in p2s for v in variants]
    if os.environ.get("VB_ISOLATE", "1") == "0":
        for lib, v, p2v in jobs:
            run_one(b.units, b.comp, b.raw, b.out_bytes, codec, data, lib, v, p2v)
        return
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    path = os.path.join(shm, f"vb_batch_{os.getpid()}.npz")
    np.savez(path, units=b.units, comp=b.comp, raw=b.raw, out_bytes=np.int64(b.out_bytes))
    try:
        for lib, v, p2v in jobs:
            t0 = time.time()
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", path, str(codec), data, lib or "-", str(v), str(p2v)],
                                   capture_output=True, timeout=int(os.environ.get("VB_TIMEOUT", "90")))
                line = [l for l in r.stdout.decode(errors="replace").splitlines() if l.startswith("{")]
                if r.returncode == 0 and line:
                    print(line[-1], flush=True)
                else:
                    print(json.dumps({"lib": lib, "variant": v, "p2_variant": p2v, "codec": codec, "failed": r.returncode, "stderr": r.stderr.decode(errors="replace")[-400:]}), flush=True)
            except subprocess.TimeoutExpired:
                print(json.dumps({"lib": lib, "variant": v, "p2_variant": p2v, "codec": codec, "failed": "timeout", "seconds": round(time.time() - t0, 1)}), flush=True)
    finally:
        try:
            os.remove(path)
        except OSError:
            pass


if __name__ == "__main__":
    main()
