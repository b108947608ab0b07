"""Does the built libmsgpu.so still hold, instruction for instruction, the kernels that were verified on a B200?

    python tools/sass_check.py            compare libmspack_b200/libmsgpu.so with profiles/sass_verified.json
    python tools/sass_check.py --record [file.sass]   (re)write the baseline from the current build (or a cuobjdump -sass dump)

New code paths go into template instantiations of their own, so a kernel that has been measured and parity-tested
on the GPU must come out of the compiler unchanged when a feature is added next to it.  The baseline stores, per verified kernel,
a hash of its SASS (opcodes and operands, no addresses); kernels are matched by hash, so a renamed instantiation (an extra
template argument) still matches."""
import hashlib, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = os.path.join(ROOT, "profiles", "sass_verified.json")
# the kernels of the default path (what bench.py and the gpu test tier launch), by the start of their demangled name
VERIFIED = ["k_p1_lzx<448, 256, false, 104", "k_p1_lzx<448, 72, true, 0", "k_p1_mszip<448, 124", "k_p1_qtm<", "k_p2_resolve<false", "k_p2_resolve<true",
            "k_p2_ring", "k_p2_chain", "k_e8", "k_status", "k_set_status"]


def kernels(sass_text):
    out, cur = {}, None
    for line in sass_text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1); out[cur] = []; continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?)\s*;", line)
        if m:
            out[cur].append(m.group(2))
    return {k: (hashlib.sha256("\n".join(v).encode()).hexdigest()[:24], len(v)) for k, v in out.items()}


def demangle(names):
    r = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, r))


def dump(path):
    if path.endswith(".sass"):
        return open(path).read()
    return subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout


def main():
    if "--record" in sys.argv:
        src = sys.argv[sys.argv.index("--record") + 1] if len(sys.argv) > sys.argv.index("--record") + 1 else os.path.join(ROOT, "libmspack_b200", "libmsgpu.so")
        ks = kernels(dump(src)); dm = demangle(list(ks))
        base = {}
        for k, (h, n) in ks.items():
            d = dm[k].replace("void ", "")
            if any(d.startswith(v) for v in VERIFIED):
                base[d.split("(")[0]] = {"sha": h, "instructions": n}
        json.dump(base, open(BASE, "w"), indent=1, sort_keys=True)
        print(f"recorded {len(base)} kernels from {src}")
        return 0
    base = json.load(open(BASE))
    ks = kernels(dump(os.path.join(ROOT, "libmspack_b200", "libmsgpu.so")))
    have = {h for h, _ in ks.values()}
    bad = [k for k, v in base.items() if v["sha"] not in have]
    print(f"{len(base) - len(bad)} of {len(base)} verified kernels unchanged" + ("" if not bad else "; CHANGED: " + ", ".join(bad)))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
