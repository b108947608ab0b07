"""What the host side of the end-to-end path can carry: pinned H2D and D2H copies, no kernels, on 1 / 2 / 4 / 8
GPUs at once - alone, then both directions together - with the byte counts of bench.py's headline step (1.0 GB in, 2.1 GB out per GPU).
One JSON line per GPU count; `e2e_ceiling_gbs` is what bench.py's e2e could reach at that GPU count if the kernels cost nothing.
usage: python tools/pcie_probe.py [max_gpus]"""
import json, os, sys, threading, time
import torch

IN_B, OUT_B = 1_000_000_000, 2_147_483_648


def worker(dev, res, bar, mode):
    torch.cuda.set_device(dev)
    h_in = torch.empty(IN_B, dtype=torch.uint8).pin_memory(); h_out = torch.empty(OUT_B, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(IN_B, dtype=torch.uint8, device=f"cuda:{dev}"); d_out = torch.empty(OUT_B, dtype=torch.uint8, device=f"cuda:{dev}")
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    best = 1e9
    for it in range(4):
        bar.wait()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        if it:
            best = min(best, dt)
    res[dev] = best


def main():
    maxg = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
    for n in (1, 2, 4, 8):
        if n > maxg or n > torch.cuda.device_count():
            break
        line = {"gpus": n}
        for mode in ("h2d", "d2h", "both"):
            res, bar = {}, threading.Barrier(n)
            ths = [threading.Thread(target=worker, args=(d, res, bar, mode)) for d in range(n)]
            [t.start() for t in ths]; [t.join() for t in ths]
            t = max(res.values())
            if mode == "h2d":
                line["h2d_gbs_per_gpu"] = round(IN_B / t / 1e9, 2)
            elif mode == "d2h":
                line["d2h_gbs_per_gpu"] = round(OUT_B / t / 1e9, 2)
            else:
                line["both_s"] = round(t, 4); line["e2e_ceiling_gbs"] = round(n * OUT_B / t / 1e9, 2); line["e2e_ceiling_gbs_per_gpu"] = round(OUT_B / t / 1e9, 2)
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
