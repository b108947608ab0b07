"""Rank the lines tools/variant_bench.py printed (gpurun_out/r2_variants*.log): per codec, shapes sorted by whole-batch time, with the
stage times and the change against the default shape of that codec.  usage: rank_variants.py gpurun_out/r2_variants*.log"""
import json, sys

DEFAULT = {3: 30, 1: 14, 2: 0}
rows = []
for path in sys.argv[1:]:
    for line in open(path, errors="replace"):
        line = line.strip()
        if line.startswith("{"):
            try:
                rows.append(json.loads(line))
            except ValueError:
                pass
for codec in sorted({r.get("codec") for r in rows}):
    rs = [r for r in rows if r.get("codec") == codec]
    ok = [r for r in rs if "best_ms" in r]
    base = next((r for r in ok if r["variant"] == DEFAULT.get(codec) and r.get("p2_variant", 0) == 0), None)
    print(f"codec {codec}: {len(ok)} measured, {len(rs) - len(ok)} failed" + (f", default {base['best_ms']} ms (P1 {base['p1_ms']} + P2 {base['p2_ms']})" if base else ""))
    for r in sorted(ok, key=lambda r: r["best_ms"]):
        d = f"{(r['best_ms'] / base['best_ms'] - 1) * 100:+6.1f} %" if base else ""
        print(f"  variant {r['variant']:3d} p2 {r.get('p2_variant', 0)}  {r['best_ms']:8.3f} ms {d}  P1 {r['p1_ms']:7.3f}  P2 {r['p2_ms']:7.3f}  {r['GB_per_s']:7.1f} GB/s  {'verified' if r.get('verified') else 'NOT VERIFIED'}")
    for r in rs:
        if "best_ms" not in r:
            print(f"  variant {r.get('variant')} p2 {r.get('p2_variant')}  FAILED: {r.get('failed')}  {str(r.get('stderr', ''))[-160:]!r}")
