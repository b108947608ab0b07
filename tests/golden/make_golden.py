#!/usr/bin/env python
"""Generate tests/golden/{manifest.json,units/*} from the reference's own fixture cabinets.

Run HERE (the container that has /root/reference); the outputs are committed so the GPU box,
which has no /root/reference, can run the parity tests.

For every folder of every listed cabinet the script cuts the bytes the codec would be fed
(tests/cabfile.py), decodes them with the UNMODIFIED reference decoders (oracle/_ref/
libmspack_ref.so, built by oracle/Makefile from /root/reference/libmspack/mspack/*.c) and records
the error code, length and MD5 of the output.  Where the reference's own tests assert an MD5 for
that folder the script checks the oracle against it (``asserted``) - that is what pins the oracle.
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from cabfile import parse_cab  # noqa: E402
from oracle import oracle as orc  # noqa: E402

REF = "/root/reference"
T = REF + "/libmspack/test/test_files/cabd/"
B = REF + "/cabextract/test/bugs/"
C = REF + "/cabextract/test/cabs/"

# (fixture path, {folder index: md5 asserted by the reference's tests}, where asserted)
FIXTURES = [
    (T + "mszip_lzx_qtm.cab", {0: "940cba86658fbceb582faecd2b5975d1", 1: "703474293b614e7110b3eb8ac2762b53",
                               2: "98fcfa4962a0f169a3c7fdbcb445cf17"}, "libmspack/test/cabd_test.c:472-478"),
    (T + "normal_2files_2folders.cab", {}, "libmspack/test/cabd_test.c:486-520 (order independence only)"),
    (T + "normal_2files_1folder.cab", {}, ""),
    (C + "large-files-cab.cab", {0: "ac923e14971324651015ba44ceb59b36"},
     "cabextract/test/large-files.test:14-25 (MD5 of the intermediate large-files.cab, SURVEY.md 8c)"),
    (C + "simple.cab", {}, ""),
    (C + "dir.cab", {}, ""),
    (B + "cve-2010-2801-qtm-flush.cab", {}, "cabextract/test/bugs.test (must not crash)"),
    (T + "cve-2010-2800-mszip-infinite-loop.cab", {}, "libmspack/test/cabd_test.c:410-421 (must fail)"),
    (T + "cve-2014-9556-qtm-infinite-loop.cab", {}, "libmspack/test/cabd_test.c:410-421 (must fail)"),
    (T + "cve-2015-4470-mszip-over-read.cab", {}, "libmspack/test/cabd_test.c:410-421 (must fail)"),
    (T + "cve-2015-4471-lzx-under-read.cab", {}, "libmspack/test/cabd_test.c:410-421 (must fail)"),
    (T + "cve-2018-18584-qtm-max-size-block.cab", {}, "libmspack/test/cabd_test.c:410-421 (must fail)"),
    (T + "lzx-main-tree-no-lengths.cab", {}, "libmspack/test/cabd_test.c:410-421 (must fail)"),
    (T + "lzx-premature-matches.cab", {}, "libmspack/test/cabd_test.c:410-421 (must fail)"),
]
KEEP_OUT_BELOW = 1 << 16   # commit expected output bytes only when small; MD5 otherwise


def main():
    o = orc.Oracle("reference")
    units_dir = os.path.join(HERE, "units")
    os.makedirs(units_dir, exist_ok=True)
    manifest = []
    for path, asserted, where in FIXTURES:
        data = open(path, "rb").read()
        try:
            folders = parse_cab(data)
        except Exception as e:  # malformed container: nothing for the codec path
            print("skip", path, e)
            continue
        base = os.path.basename(path)[:-4]
        for i, f in enumerate(folders):
            if f.method not in (1, 2, 3) or not f.blocks:
                continue
            if f.method == 2 and not (10 <= f.window_bits <= 21):
                continue
            if f.method == 3 and not (15 <= f.window_bits <= 21):
                continue
            unit = f.unit_bytes()
            out_len = f.out_len
            if out_len == 0 or out_len > (64 << 20):
                continue
            out, err = o.decode_one(f.method, unit, out_len, f.window_bits)
            md5 = hashlib.md5(out).hexdigest()
            name = f"{base}.f{i}"
            ent = {"name": name, "codec": f.method, "window_bits": f.window_bits, "reset_interval": 0,
                   "in_len": len(unit), "out_len": out_len, "err": err, "md5": md5 if err == 0 else None,
                   "nblocks": len(f.blocks), "source": path.replace(REF + "/", ""), "asserted_by": where}
            if i in asserted:
                assert err == 0 and md5 == asserted[i], (name, err, md5, asserted[i])
                ent["asserted_md5"] = asserted[i]
            open(os.path.join(units_dir, name + ".in"), "wb").write(unit)
            if err == 0 and out_len <= KEEP_OUT_BELOW:
                open(os.path.join(units_dir, name + ".out"), "wb").write(out)
            manifest.append(ent)
            print(f"{name:48s} codec={f.method} wb={f.window_bits:2d} blocks={len(f.blocks):4d} "
                  f"in={len(unit):6d} out={out_len:9d} err={err:2d} md5={md5 if err == 0 else '-'}"
                  f"{'  ASSERTED-OK' if i in asserted else ''}")
    json.dump(manifest, open(os.path.join(HERE, "manifest.json"), "w"), indent=1)
    print(len(manifest), "units")


if __name__ == "__main__":
    main()
