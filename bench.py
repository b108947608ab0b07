#!/usr/bin/env python
"""bench.py - the headline benchmark: decompressed GB/s of a batch of LZX (21-bit window) units.

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): per GPU, 65 536
independent LZX units, window_bits 21, one 32 KiB frame each, synthetic Zipf text,
compressed by this repository's own LZX encoder (libmspack_b200/gen - the reference has no encoder).
A "step" is one pass of the hot path over that batch.  Weak scaling: every rank decodes its own 65 536
units (different corpus blocks); there is no collective in the data path (units are independent).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--units U] [--impl reference]
    torchrun --nproc-per-node N ... bench.py --gpus N ...

`value`  = whole-job decompressed GB/s with the compressed units already resident in HBM (CUDA events
           around the K timed steps, max over ranks).
`e2e`    = the same metric through the reference-facing C-ABI call with HOST buffers
           (msgpu_decode_batch_host: H2D of units + compressed bytes, decode, D2H of output + status inside
           the timed region).
`roofline` is for the dominant kernel (the P1 entropy kernel): algorithmic bytes (U + C per unit,
           SURVEY.md 8d) / its launch time measured live with CUDA events, against the measured HBM copy
           bandwidth of MEASURED_PEAKS.json.
`cpu_baseline` / `--impl reference` time the reference's own decoders (oracle/_ref/libmspack_ref.so,
           built from /root/reference/libmspack/mspack/{lzxd,qtmd,mszipd}.c) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "decompressed GB/s (batch LZX 21-bit window)"
UNIT_BYTES = 32768
WINDOW_BITS = 21


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--units", type=int, default=65536, help="units per GPU (BASELINE: 65536)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--gather", action="store_true", help="also time the optional NCCL all-gather of the outputs")
    ap.add_argument("--cpu-sample", type=int, default=32768, help="units in the cpu_baseline sample")
    ap.add_argument("--e2e-inflight", type=int, default=1, choices=[1, 2],
                    help="2: also measure e2e with two batches in flight (two host threads, one context each; every step still one "
                         "msgpu_decode_batch_host call with its own H2D and D2H) and report it as e2e.inflight2 - not validated on a GPU yet")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic():
    """DRAM bytes per P1 launch from the committed ncu capture summary, if there is one for this unit count."""
    p = os.path.join(ROOT, "profiles", "p1_lzx_dram.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            pass
    return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 2 + k and r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def run_reference(args):
    """--impl reference: the reference's own CPU decoders on the same workload, all host threads."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    from libmspack_b200 import gen
    from libmspack_b200.units import CODEC_LZX
    from oracle import oracle as orc
    ora = orc.load("reference")
    cores = os.cpu_count() or 1
    sample = min(args.units, args.cpu_sample)
    b = gen.make_batch(CODEC_LZX, sample, unit_bytes=UNIT_BYTES, window_bits=WINDOW_BITS, threads=cores)
    for _ in range(max(args.warmup, 1)):
        ora.decode_batch(b.units, b.comp, b.out_bytes, threads=cores)
    secs = []
    for _ in range(args.steps):
        _, st, s = ora.decode_batch(b.units, b.comp, b.out_bytes, threads=cores)
        assert (st == 0).all()
        secs.append(s)
    t = float(np.mean(secs))
    gbs = sample * UNIT_BYTES / t / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": round(gbs, 4), "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(t * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"{sample} LZX units (window_bits {WINDOW_BITS}, one 32 KiB frame each) per step, Zipf text - a bounded sample of the "
                                   f"{args.units}-unit batch of the b200 arm", "units_per_step": sample, "threads": cores},
            "cpu_baseline": {"value": round(gbs, 4), "unit": "GB/s", "cores": cores, "kind": ora.kind,
                             "sample": f"{sample} units x 32 KiB per step, lzxd_init + lzxd_decompress + lzxd_free per unit, one pthread per core"},
            "e2e": {"value": round(gbs, 4), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist
    from libmspack_b200 import gen
    from libmspack_b200.codec import BatchDecoder
    from libmspack_b200.units import CODEC_LZX

    rank, local_rank, world = dist_env()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libmspack_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cores = os.cpu_count() or 1
    n = args.units

    # ---- workload: this rank's units (weak scaling: per-GPU work is fixed) ----
    t0 = time.time()
    b = gen.make_batch(CODEC_LZX, n, unit_bytes=UNIT_BYTES, window_bits=WINDOW_BITS, first_unit=rank * n,
                       threads=max(1, cores // max(world, 1)), keep_raw=True)
    gen_s = time.time() - t0
    U, C = n * UNIT_BYTES, b.in_bytes

    dec = BatchDecoder(local_rank)
    d_in = torch.from_numpy(b.comp).to(dev)
    d_out = torch.zeros(b.out_bytes, dtype=torch.uint8, device=dev)
    d_st = torch.full((n,), -1, dtype=torch.int32, device=dev)
    # an explicit (non-default) stream: its handle is what msgpu_decode_batch_device launches on, so the CUDA events
    # below really bracket the kernels (a NULL stream would mean "the context's own stream")
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- correctness of what is being timed: decode(encode(x)) == x over the whole batch ----
    dec.decode_device(b.units, d_in, d_out, d_st, stream)
    torch.cuda.synchronize()
    ok = bool((d_st == 0).all().item()) and bool(torch.equal(d_out.cpu(), torch.from_numpy(b.raw)))
    if not ok:
        raise SystemExit("bench.py: GPU output differs from the generator's raw data")

    # ---- device-resident timing: K steps, CUDA events, max over ranks ----
    clocks = ClockSampler(local_rank)      # sampling spans the warm-up and the timed steps (nvidia-smi needs ~0.3 s to start)
    clocks.start()
    t_w = time.time()
    for _ in range(max(args.warmup, 3)):
        dec.decode_device(b.units, d_in, d_out, d_st, stream)
    while time.time() - t_w < 1.0:         # keep the GPU under the same load until the sampler is certainly running
        dec.decode_device(b.units, d_in, d_out, d_st, stream)
    barrier()
    launches0 = dec.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    ctx_ms = 0.0
    for _ in range(args.steps):
        dec.decode_device(b.units, d_in, d_out, d_st, stream)
    e1.record(stream)
    barrier()
    clk = clocks.stop()
    ms_total = e0.elapsed_time(e1)
    ctx_ms = dec.last_kernel_ms()          # the library's own events around the last step's kernels (cross-check)
    launches = dec.launches - launches0
    t_ms = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_step = float(t_ms.item()) / args.steps
    value = world * U / (ms_step * 1e-3) / 1e9

    # ---- dominant-kernel roofline: stage timing (stages serialised, each launch between CUDA events) ----
    dec.set_stage_timing(True)
    p1_ms, p2_ms = [], []
    for _ in range(3):
        dec.decode_device(b.units, d_in, d_out, d_st, stream)
        torch.cuda.synchronize()
        p1_ms.append(dec.stage_ms(0)); p2_ms.append(dec.stage_ms(1))
    dec.set_stage_timing(False)
    p1, p2 = float(np.median(p1_ms)), float(np.median(p2_ms))
    peak, peak_src = load_peaks()
    achieved = (U + C) / (p1 * 1e-3) / 1e9
    traffic = load_traffic()
    roofline = {"bound": "hbm", "kernel": "k_p1_lzx (entropy stage, one thread per unit)", "achieved": round(achieved, 2), "peak": peak,
                "unit": "GB/s", "frac": round(achieved / peak, 5), "peak_source": peak_src,
                "traffic": (traffic or {}).get("dram_bytes_per_launch_set"),
                "algorithmic_bytes_per_step": U + C, "kernel_ms_per_step": round(p1, 3),
                "p2_resolve_ms_per_step": round(p2, 3), "p2_achieved_gbs": round((2 * U) / (p2 * 1e-3) / 1e9, 2),
                "hbm_write_fraction": round(value / world / peak, 5), "last_step_kernels_ms": round(ctx_ms, 3),
                "note": "latency/issue bound integer path: the practical limiter is serial symbol decode x resident warps, not DRAM (SURVEY.md 8d)"}

    # ---- end to end through the C-ABI with host buffers (pinned), H2D + D2H inside the timed region ----
    h_in = torch.from_numpy(b.comp).pin_memory()
    h_out = torch.empty(b.out_bytes, dtype=torch.uint8).pin_memory()
    h_st = np.full(n, -1, dtype=np.int32)
    for _ in range(2):
        dec.decode_host_into(b.units, h_in.data_ptr(), h_in.numel(), h_out.data_ptr(), h_out.numel(), h_st)
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        dec.decode_host_into(b.units, h_in.data_ptr(), h_in.numel(), h_out.data_ptr(), h_out.numel(), h_st)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_ok = bool((h_st == 0).all()) and bool(torch.equal(h_out, torch.from_numpy(b.raw)))
    e2e = {"value": round(world * U / float(t_e.item()) / 1e9, 3), "unit": "GB/s", "h2d_bytes_per_step": int(b.comp.size + b.units.nbytes),
           "d2h_bytes_per_step": int(b.out_bytes + h_st.nbytes), "steps": e2e_steps, "verified": e2e_ok,
           "api": "msgpu_decode_batch_host (include/msgpu.h), pinned host buffers"}

    # ---- optional: the same with two batches in flight (the host call is synchronous; a caller with a stream of batches overlaps
    # one batch's D2H with the next one's H2D + kernels by calling from two threads, one context each) ----
    if args.e2e_inflight == 2:
        try:
            import threading
            dec2 = BatchDecoder(local_rank)
            h_out2 = torch.empty(b.out_bytes, dtype=torch.uint8).pin_memory()
            h_st2 = np.full(n, -1, dtype=np.int32)
            lanes = [(dec, h_out, h_st, (e2e_steps + 1) // 2), (dec2, h_out2, h_st2, e2e_steps // 2)]
            dec2.decode_host_into(b.units, h_in.data_ptr(), h_in.numel(), h_out2.data_ptr(), h_out2.numel(), h_st2)
            errs = []

            def work(d, ho, hs, k):
                try:
                    for _ in range(k):
                        d.decode_host_into(b.units, h_in.data_ptr(), h_in.numel(), ho.data_ptr(), ho.numel(), hs)
                except Exception as ex:      # noqa: BLE001 - reported below
                    errs.append(repr(ex))
            barrier()
            h_out.zero_(); h_out2.zero_()
            ths = [threading.Thread(target=work, args=l) for l in lanes]
            t0 = time.perf_counter()
            for t in ths:
                t.start()
            for t in ths:
                t.join()
            torch.cuda.synchronize()
            s2 = (time.perf_counter() - t0) / e2e_steps
            t_2 = torch.tensor([s2], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t_2, op=dist.ReduceOp.MAX)
            raw_t = torch.from_numpy(b.raw)
            ok2 = not errs and bool((h_st == 0).all()) and bool((h_st2 == 0).all()) and bool(torch.equal(h_out, raw_t)) and bool(torch.equal(h_out2, raw_t))
            e2e["inflight2"] = {"value": round(world * U / float(t_2.item()) / 1e9, 3), "unit": "GB/s", "steps": e2e_steps, "verified": ok2,
                                "how": "two host threads, one msgpu context each, alternate steps; every step one msgpu_decode_batch_host call", "errors": errs[:2]}
            dec2.close()
        except Exception as ex:      # noqa: BLE001 - the sequential e2e above stands
            e2e["inflight2"] = {"error": repr(ex)[:300]}

    # ---- optional output gather over NCCL (off the data path; reported separately) ----
    gather = None
    if args.gather and world > 1:
        outs = torch.empty(world * b.out_bytes, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(outs, d_out)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        dist.all_gather_into_tensor(outs, d_out)
        g1.record()
        torch.cuda.synchronize()
        gather = {"ms": round(g0.elapsed_time(g1), 3), "bytes_per_rank_in": (world - 1) * b.out_bytes}

    # ---- CPU baseline: the reference's own decoders on this host's cores (rank 0, N == 1 only) ----
    cpu = None
    if rank == 0 and world == 1:
        try:
            from oracle import oracle as orc
            ora = orc.load("reference")
            ns = min(n, args.cpu_sample)
            sub = b.units[:ns].copy()
            ora.decode_batch(sub[:2048], b.comp, int(sub["out_off"][-1]) + UNIT_BYTES, threads=cores)
            out_c, st_c, secs = ora.decode_batch(sub, b.comp, ns * UNIT_BYTES, threads=cores)
            same = bool((st_c == 0).all()) and np.array_equal(out_c, b.raw[:ns * UNIT_BYTES])
            cpu = {"value": round(ns * UNIT_BYTES / secs / 1e9, 4), "unit": "GB/s", "cores": cores, "kind": ora.kind,
                   "sample": f"first {ns} units of the same batch, lzxd_init + lzxd_decompress + lzxd_free per unit, one pthread per core",
                   "gpu_output_identical_on_sample": same}
        except Exception as e:  # the oracle is a reported baseline, never a dependency of the product path
            cpu = {"value": None, "unit": "GB/s", "cores": cores, "kind": "unavailable", "sample": str(e)}

    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 3), "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": f"{n} LZX units per GPU (window_bits {WINDOW_BITS}, one 32 KiB frame each, BASELINE configs[2]), "
                                       f"Zipf text corpus seed 0x4D534346, own LZX encoder", "units_per_gpu": n, "unit_bytes": UNIT_BYTES,
                           "compressed_bytes_per_gpu": int(C), "ratio": round(C / U, 4), "parallelism": f"units sharded by index over {world} GPU(s), no data-path collective",
                           "l2": "inputs (compressed + output + intermediate records) are far larger than the 126 MB L2", "generate_s": round(gen_s, 1),
                           "output_gather": gather},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
