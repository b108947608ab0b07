"""Quick device-resident timing of one codec batch (development aid; bench.py is the contract)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from libmspack_b200 import gen
from libmspack_b200.codec import BatchDecoder

codec = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
kw = {}
if len(sys.argv) > 4:
    kw = eval(sys.argv[4])
t0 = time.time(); b = gen.make_batch(codec, n, keep_raw=(codec != 3 or True), **kw); tg = time.time() - t0
dec = BatchDecoder(0)
d_in = torch.from_numpy(b.comp).cuda(); d_out = torch.zeros(b.out_bytes, dtype=torch.uint8, device="cuda"); d_st = torch.zeros(n, dtype=torch.int32, device="cuda")
stream = torch.cuda.Stream(); torch.cuda.synchronize()
best = 1e9
for it in range(iters):
    dec.decode_device(b.units, d_in, d_out, d_st, stream)
    torch.cuda.synchronize()
    ms = dec.last_kernel_ms(); best = min(best, ms)
    print(f"iter {it}: kernels {ms:.3f} ms  -> {b.out_bytes/ms/1e6:.1f} GB/s out")
ok = bool((d_st == 0).all().item()) and np.array_equal(d_out.cpu().numpy().reshape(n, -1)[:, :b.units['out_len'][0]].reshape(-1), b.raw)
print(f"codec {codec} n {n} gen {tg:.1f}s ratio {b.in_bytes/(n*int(b.units['out_len'][0])):.3f} best {best:.3f} ms = {b.out_bytes/best/1e6:.1f} GB/s  roundtrip_ok={ok} launches={dec.launches} scratch={dec.scratch_bytes/2**30:.2f} GiB")
if os.environ.get("QB_STAGE"):
    for mode in (False, True, False, True):
        dec.set_stage_timing(mode)
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record(stream)
        dec.decode_device(b.units, d_in, d_out, d_st, stream)
        t1.record(stream)
        torch.cuda.synchronize()
        print(f"stage_timing={mode}: outer {t0.elapsed_time(t1):.3f} ms, ctx total {dec.last_kernel_ms():.3f} ms, P1 {dec.stage_ms(0):.3f} P2 {dec.stage_ms(1):.3f} E8 {dec.stage_ms(2):.3f}")
if os.environ.get("QB_HOST"):
    # host-buffer path (msgpu_decode_batch_host): pinned buffers, H2D + kernels + D2H; plus the raw copy times for scale
    h_in = torch.from_numpy(b.comp).pin_memory(); h_out = torch.empty(b.out_bytes, dtype=torch.uint8).pin_memory()
    h_st = np.zeros(n, np.int32)
    for it in range(4):
        t0 = time.perf_counter()
        dec.decode_host_into(b.units, h_in.data_ptr(), h_in.numel(), h_out.data_ptr(), h_out.numel(), h_st)
        dt = time.perf_counter() - t0
        print(f"host path iter {it}: {dt*1e3:.2f} ms = {b.out_bytes/dt/1e9:.1f} GB/s  ok={bool((h_st == 0).all())}")
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
    e0.record(); d_in.copy_(h_in, non_blocking=True); e1.record(); h_out.copy_(d_out, non_blocking=True); e2.record(); torch.cuda.synchronize()
    print(f"raw copies: H2D {h_in.numel()/1e6:.0f} MB in {e0.elapsed_time(e1):.2f} ms ({h_in.numel()/e0.elapsed_time(e1)/1e6:.1f} GB/s), "
          f"D2H {h_out.numel()/1e6:.0f} MB in {e1.elapsed_time(e2):.2f} ms ({h_out.numel()/e1.elapsed_time(e2)/1e6:.1f} GB/s)")
