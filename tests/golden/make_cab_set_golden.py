#!/usr/bin/env python
"""Generate tests/golden/cab/manifest_sets.json (+ the synthetic set's cabinets): cabinet SETS and salvage mode.

Run HERE (the container that has /root/reference), after make_cab_golden.py.  Every scenario - the reference's own split-1..5.cab
fixture set (whole, and with a cabinet missing at either end), a synthetic three-cabinet set whose LZX and Quantum folders are cut in
the MIDDLE of a CFDATA block, and the damaged synthetic cabinets in salvage mode - is extracted file by file with the UNMODIFIED
reference (oracle/_ref/ref_cabx: open + append + extract); per file the folder, offset, length, extract() error and MD5 are recorded.
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)

from cabfile import build_cab  # noqa: E402
from make_cab_golden import split_blocks, unit_stream  # noqa: E402

OUT = os.path.join(HERE, "cab")
CABX = os.path.join(ROOT, "oracle", "_ref", "ref_cabx")


def synthetic_set():
    """three cabinets: [MSZIP, LZX part 1] [LZX part 2, Quantum part 1] [Quantum part 2, stored] - both cuts inside a block"""
    import numpy as np
    zs, _ = unit_stream(1, 40000)
    ck = [i for i in range(1, len(zs) - 1) if zs[i:i + 2] == b"CK"]
    ls, _ = unit_stream(3, 120000, window_bits=17, block_mode=4, split=2)
    qs, _ = unit_stream(2, 30000, window_bits=18)       # one frame: cabd appends the 0xFF trailer after the JOINED block
    assert qs[-1] == 0xFF
    stored = bytes(np.random.default_rng(9).integers(0, 256, 9000, dtype=np.uint8))
    lb = split_blocks(ls, 120000, [len(ls) // 4, len(ls) // 2, 3 * len(ls) // 4])
    qblock = (qs[:-1], 30000)

    def cut(block, at):
        payload, usize = block
        return (payload[:at], 0), (payload[at:], usize)
    l_a, l_b = cut(lb[2], len(lb[2][0]) // 3)
    q_a, q_b = cut(qblock, len(qblock[0]) // 2)
    cab1 = build_cab([
        dict(comp_type=1, blocks=split_blocks(zs, 40000, ck), files=[("zip_a.bin", 0, 15000), ("zip_b.bin", 15000, 25000)]),
        dict(comp_type=3 | (17 << 8), blocks=[lb[0], lb[1], l_a], files=[("lzx_a.bin", 0, 50000), ("lzx_b.bin", 50000, 70000, 0xFFFE)]),
    ], next=("set2.cab", "disk2"), set_index=0)
    cab2 = build_cab([
        dict(comp_type=3 | (17 << 8), blocks=[l_b, lb[3]], files=[("lzx_b.bin", 50000, 70000, 0xFFFD)]),
        dict(comp_type=2 | (18 << 8), blocks=[q_a], files=[("qtm_a.bin", 0, 10000, 0xFFFE), ("qtm_b.bin", 10000, 20000, 0xFFFE)]),
    ], prev=("set1.cab", "disk1"), next=("set3.cab", "disk3"), set_index=1)
    cab3 = build_cab([
        dict(comp_type=2 | (18 << 8), blocks=[q_b], files=[("qtm_a.bin", 0, 10000, 0xFFFD), ("qtm_b.bin", 10000, 20000, 0xFFFD)]),
        dict(comp_type=0, blocks=[(stored, 9000)], files=[("raw.bin", 0, 9000)]),
    ], prev=("set2.cab", "disk2"), set_index=2)
    return {"synth_set1.cab": cab1, "synth_set2.cab": cab2, "synth_set3.cab": cab3}


def run(cabs, salvage):
    tmp = tempfile.mkdtemp()
    try:
        paths = [os.path.join(OUT, c) for c in cabs]
        cmd = [CABX] + (["--salvage"] if salvage else []) + [paths[0], tmp] + paths[1:]
        # A damaged stream can copy from window positions nothing was decoded to yet; the reference's window is malloc'd and never
        # cleared (qtmd.c:396-409 reads it as it is), so what such a match yields is whatever the heap held.  Every allocation of a
        # page or more straight from mmap (zero pages) makes the reference's answer the deterministic one the GPU path defines.
        env = dict(os.environ, MALLOC_MMAP_THRESHOLD_="4096")
        lines = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=300, env=env).stdout.decode("ascii", "replace").split("\n")
        entry = {"cabs": list(cabs), "salvage": int(salvage), "open": lines[0].strip(), "files": []}
        for ln in lines[1:]:
            if not ln.strip():
                continue
            idx, fol, off, length, err = (int(x) for x in ln.split())
            rec = {"index": idx, "folder": fol, "offset": off, "length": length, "err": err}
            fp = os.path.join(tmp, str(idx))
            if os.path.exists(fp):
                data = open(fp, "rb").read()
                rec["md5"] = hashlib.md5(data).hexdigest()
                rec["written"] = len(data)
            entry["files"].append(rec)
        return entry
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "cabx"], stdout=subprocess.DEVNULL)
    for name, img in synthetic_set().items():
        open(os.path.join(OUT, name), "wb").write(img)
    split = [f"split-{k}.cab" for k in range(1, 6)]
    synth = ["synth_set1.cab", "synth_set2.cab", "synth_set3.cab"]
    scenarios = [(split, 0), (split[:2], 0), (split[1:], 0), (split[2:4], 0), (synth, 0), (synth[:2], 0), (synth[1:], 0), (synth, 1)]
    for name in ("synth_badsum_lzx.cab", "synth_badsum_qtm.cab", "synth_badsum_zip_raw.cab", "synth_bigblock.cab", "synth_nosum_corrupt.cab", "synth_truncated.cab", "synth_multi.cab"):
        scenarios.append(([name], 1))
    manifest = [run(c, s) for c, s in scenarios]
    json.dump(manifest, open(os.path.join(OUT, "manifest_sets.json"), "w"), indent=1)
    for e in manifest:
        print(e["cabs"], "salvage" if e["salvage"] else "", e["open"], [(f["folder"], f["err"]) for f in e["files"]])


if __name__ == "__main__":
    main()
