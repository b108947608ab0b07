"""Static size of the lockstep loops of the entropy kernels in the built library: for every k_p1_* instantiation, the loops that contain
the warp vote of p1_run (backward branch over a VOTE), smallest first - the hot decode loop(s) and the outer service loop.
usage: sass_loops.py [kernel-name-substring]      (cuobjdump -sass of libmspack_b200/libmsgpu.so; executed counts need the GPU)"""
import os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "libmspack_b200", "libmsgpu.so")], capture_output=True, text=True, check=True).stdout
funcs, cur = {}, None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []; continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?)\s*;", line)
    if m:
        funcs[cur].append((int(m.group(1), 16), m.group(2)))
names = subprocess.run(["c++filt"], input="\n".join(funcs), capture_output=True, text=True).stdout.splitlines()
want = sys.argv[1] if len(sys.argv) > 1 else "k_p1_"
for k, dem in sorted(zip(funcs, names), key=lambda x: x[1]):
    if want not in dem:
        continue
    ins = funcs[k]
    votes = [a for a, t in ins if "VOTE" in t]
    loops = []
    for a, t in ins:
        m = re.search(r"BRA\S*\s+(?:\S+\s+)?(0x[0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and any(tgt <= v <= a for v in votes):
                loops.append(sum(1 for x, _ in ins if tgt <= x <= a))
    print(f"{dem.split('(')[0].replace('void ', ''):48s} {len(ins):5d} instructions; vote loops {sorted(loops)}")
