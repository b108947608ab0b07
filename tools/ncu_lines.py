"""Aggregate an ncu `--page source --print-source cuda,sass --csv` dump per CUDA source line.
usage: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv | python tools/ncu_lines.py [topN]"""
import csv, sys
top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rows = list(csv.reader(sys.stdin))
cur_file = None; hdr = None; agg = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or r[0] in ("Function Name",): continue
    if r[0] != "":   # a source line row (aggregated over its SASS)
        try:
            key = (cur_file, int(r[0]), r[1].strip()[:110])
        except ValueError:
            continue
        d = dict(zip(hdr[4:], r[4:]))
        def f(k):
            try: return float(d.get(k, "0") or 0)
            except ValueError: return 0.0
        agg[key] = (f("# Samples"), f("Instructions Executed"), f("Thread Instructions Executed"))
tot_s = sum(v[0] for v in agg.values()) or 1; tot_i = sum(v[1] for v in agg.values()) or 1
print(f"total samples {tot_s:.0f}  total warp-instructions {tot_i:.0f}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*v[0]/tot_s:5.1f}% smp {100*v[1]/tot_i:5.1f}% ins  thr/ins {v[2]/max(v[1],1):4.1f}  {k[0]}:{k[1]}  {k[2]}")
