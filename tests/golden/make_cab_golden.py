#!/usr/bin/env python
"""Generate tests/golden/cab/{*.cab,manifest.json}: cabinet-level expectations for the f1 front end (include/msgpu_cab.h).

Run HERE (the container that has /root/reference).  Every cabinet - the reference's own fixtures plus synthetic ones written
by tests/cabfile.build_cab from this repository's encoders - is opened and extracted file by file with the UNMODIFIED
reference (oracle/_ref/ref_cabx = cabd.c + system.c + the three codecs, `make -C oracle cabx`); the open error, and per
member file (folder, offset, length, extract() error, MD5 of the extracted bytes) are recorded.  The GPU tests check
msgpu_cab_scan / msgpu_cab_decode_host against these records.
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

import numpy as np  # noqa: E402
from cabfile import build_cab  # noqa: E402
from libmspack_b200 import gen  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(HERE, "cab")
CABX = os.path.join(ROOT, "oracle", "_ref", "ref_cabx")


def unit_stream(codec, nbytes, **kw):
    b = gen.make_batch(codec, 1, unit_bytes=nbytes, keep_raw=True, **kw)
    return bytes(b.comp[:int(b.units["in_len"][0])]), bytes(b.raw)


def split_blocks(stream, total, cuts):
    """cut the codec stream at the given byte positions; every block claims 32 KiB of output, the last one the rest"""
    pos = [0] + list(cuts) + [len(stream)]
    blocks, left = [], total
    for a, b in zip(pos[:-1], pos[1:]):
        u = min(32768, left) if b != len(stream) else left
        blocks.append((stream[a:b], u))
        left -= u
    return blocks


def two_files(name, total):
    h = total // 3
    return [(name + "_a.bin", 0, h), (name + "_b.bin", h, total - h)]


def synthetic():
    zs, zraw = unit_stream(1, 70000)
    ck = [i for i in range(1, len(zs) - 1) if zs[i:i + 2] == b"CK"]
    ls, lraw = unit_stream(3, 100000, window_bits=16, block_mode=4, split=2)
    qs, qraw = unit_stream(2, 20000, window_bits=17)
    assert qs[-1] == 0xFF
    stored = bytes(np.random.default_rng(5).integers(0, 256, 50000, dtype=np.uint8))
    folders = [
        dict(comp_type=1, blocks=split_blocks(zs, 70000, ck), files=two_files("zip", 70000)),
        dict(comp_type=3 | (16 << 8), blocks=split_blocks(ls, 100000, [len(ls) // 3, 2 * len(ls) // 3, len(ls) - 7]), files=two_files("lzx", 100000)),
        dict(comp_type=2 | (17 << 8), blocks=[(qs[:-1], 20000)], files=two_files("qtm", 20000)),
        dict(comp_type=0, blocks=[(stored[:32768], 32768), (stored[32768:], 50000 - 32768)], files=two_files("raw", 50000)),
    ]
    out = {}
    good = build_cab(folders)
    out["synth_multi.cab"] = good
    out["synth_nosum.cab"] = build_cab(folders, with_checksums=False)

    def flip(img, folder, block, byte=100):
        # locate the block's payload by re-walking the writer's layout
        import struct
        img = bytearray(img)
        off = struct.unpack_from("<I", img, 0x24 + 8 * folder)[0]
        for _ in range(block):
            off += 8 + struct.unpack_from("<H", img, off + 4)[0]
        img[off + 8 + byte] ^= 0x40
        return bytes(img)
    out["synth_badsum_lzx.cab"] = flip(good, 1, 1)
    out["synth_badsum_zip_raw.cab"] = flip(flip(good, 0, 2), 3, 0)
    out["synth_badsum_qtm.cab"] = flip(good, 2, 0)
    out["synth_nosum_corrupt.cab"] = flip(flip(out["synth_nosum.cab"], 0, 1, 300), 1, 2, 50)
    out["synth_truncated.cab"] = good[:len(good) - 20000]
    big = [dict(f) for f in folders]
    big[1] = dict(big[1], blocks=[big[1]["blocks"][0], (big[1]["blocks"][1][0] + b"\0" * 40000, 32768)] + big[1]["blocks"][2:])
    out["synth_bigblock.cab"] = build_cab(big)
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "cabx"], stdout=subprocess.DEVNULL)
    cabs = {}
    for d in (REF + "/libmspack/test/test_files/cabd", REF + "/cabextract/test/cabs", REF + "/cabextract/test/bugs"):
        for fn in sorted(os.listdir(d)):
            p = os.path.join(d, fn)
            if fn.endswith(".cab") and os.path.getsize(p) < (1 << 20):
                cabs.setdefault(fn, open(p, "rb").read())
    cabs.update(synthetic())
    manifest = []
    for name, img in sorted(cabs.items()):
        path = os.path.join(OUT, name)
        with open(path, "wb") as f:
            f.write(img)
        tmp = tempfile.mkdtemp()
        try:
            lines = subprocess.run([CABX, path, tmp], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=300).stdout.decode("ascii", "replace").split("\n")
            entry = {"name": name, "open": int(lines[0].split()[1]), "files": []}
            for ln in lines[1:]:
                if not ln.strip():
                    continue
                idx, fol, off, length, err = (int(x) for x in ln.split())
                rec = {"index": idx, "folder": fol, "offset": off, "length": length, "err": err}
                fp = os.path.join(tmp, str(idx))
                if err == 0 and os.path.exists(fp):
                    rec["md5"] = hashlib.md5(open(fp, "rb").read()).hexdigest()
                entry["files"].append(rec)
            manifest.append(entry)
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    json.dump(manifest, open(os.path.join(OUT, "manifest.json"), "w"), indent=1)
    nerr = sum(1 for e in manifest if e["open"])
    print(f"{len(manifest)} cabinets ({nerr} the reference refuses to open), "
          f"{sum(len(e['files']) for e in manifest)} files, {sum(1 for e in manifest for f in e['files'] if f['err'])} failing extracts")


if __name__ == "__main__":
    main()
