"""Cabinet front end (include/msgpu_cab.h) against the reference's own cabd.c.

tests/golden/cab/manifest.json holds, for the reference's fixture cabinets and for synthetic ones (multi-block folders of all
four methods, wrong checksums, missing checksums + corrupt data, truncated image, oversized block), what the UNMODIFIED
reference did: the error of mscab_decompressor::open() and, per member file, extract()'s error and the MD5 of the bytes
(tests/golden/make_cab_golden.py).  CPU tests cover the header scan; GPU tests decode every cabinet as one batch."""
import hashlib
import json
import os

import numpy as np
import pytest

from libmspack_b200 import cab
from cabfile import cab_checksum, parse_cab

HERE = os.path.dirname(os.path.abspath(__file__))
CABDIR = os.path.join(HERE, "golden", "cab")
MANIFEST = json.load(open(os.path.join(CABDIR, "manifest.json")))


def _image(entry):
    return open(os.path.join(CABDIR, entry["name"]), "rb").read()


@pytest.mark.parametrize("entry", MANIFEST, ids=lambda e: e["name"])
def test_scan_matches_reference_open(entry):
    """msgpu_cab_scan accepts exactly the cabinets the reference opens and fails with the reference's error otherwise;
    folder / file tables agree with the reference's (and with the independent Python parser of tests/cabfile.py)."""
    img = _image(entry)
    if entry["open"]:
        with pytest.raises(cab.CabError) as e:
            cab.scan(img)
        assert e.value.code == entry["open"]
        return
    plan = cab.scan(img)
    assert len(plan.files) == len(entry["files"])
    for rec, f in zip(entry["files"], plan.files):
        assert (int(f["offset"]), int(f["length"])) == (rec["offset"], rec["length"])
        if rec["folder"] >= 0 and int(f["folder"]) != 0xFFFFFFFF:
            assert int(f["folder"]) == rec["folder"]
    pf = parse_cab(img)
    assert len(pf) == len(plan.folders)
    for a, b in zip(pf, plan.folders):
        assert a.comp_type == int(b["comp_type"]) and a.method == int(b["codec"]) and a.window_bits == int(b["window_bits"])


def test_checksum_definition():
    """cabd_checksum restated in tests/cabfile.py == the values stored in a reference fixture (pins the test-side writer)."""
    import struct
    img = open(os.path.join(CABDIR, "normal_2files_1folder.cab"), "rb").read()
    off = struct.unpack_from("<I", img, 0x24)[0]
    csum, cs, _ = struct.unpack_from("<IHH", img, off)
    assert csum and cab_checksum(img[off + 4:off + 8], cab_checksum(img[off + 8:off + 8 + cs])) == csum


def _expected_folder_status(entry, plan):
    """folder -> (expected status, comparable).  The reference extracts FILES; a folder is comparable when its files tile it
    up to the end of what its blocks hold (then the first failing file's error is the folder's error)."""
    res = {}
    for fi in range(len(plan.folders)):
        recs = [r for r in entry["files"] if r["folder"] == fi]
        out_len = int(plan.folders["out_len"][fi])
        if not recs:
            res[fi] = (None, False)
            continue
        errs = [r["err"] for r in recs if r["err"]]
        end = max(r["offset"] + r["length"] for r in recs)
        # cabd.c:1071-1079 refuses a file that cannot fit the folder's blocks before decoding anything (a file-level check; when
        # the scan stopped early, num_blocks is what it found, not the header's count, and the check does not apply)
        over = int(plan.folders["scan_status"][fi]) == 0 and any(r["offset"] + r["length"] > int(plan.folders["num_blocks"][fi]) * 32768 for r in recs)
        res[fi] = (errs[0] if errs else 0, (end == out_len or bool(errs)) and not over)
    return res


@pytest.mark.gpu
@pytest.mark.parametrize("entry", [e for e in MANIFEST if not e["open"]], ids=lambda e: e["name"])
def test_cabinet_decode_matches_reference(decoder, entry):
    img = _image(entry)
    plan = cab.scan(img)
    out, st = plan.decode(decoder)
    exp = _expected_folder_status(entry, plan)
    checked = 0
    for fi, (want, comparable) in exp.items():
        if not comparable:
            continue
        got = int(st[fi])
        if int(plan.folders["scan_status"][fi]) and want:
            assert got != 0, (entry["name"], fi)          # split / unsupported folders: any error will do
        else:
            assert got == want, (entry["name"], fi, got, want)
        checked += 1
        if got == 0:
            base = int(plan.folders["out_off"][fi])
            for r in entry["files"]:
                if r["folder"] == fi and r["err"] == 0:
                    data = out[base + r["offset"]: base + r["offset"] + r["length"]].tobytes()
                    assert hashlib.md5(data).hexdigest() == r["md5"], (entry["name"], fi, r["index"])
    if entry["name"].startswith("synth"):
        assert checked == len(plan.folders)


# ---- cabinet sets and salvage mode (tests/golden/cab/manifest_sets.json: the reference's open + append + extract) ----
SETS = json.load(open(os.path.join(CABDIR, "manifest_sets.json")))


def _set_id(e):
    return "+".join(c.replace(".cab", "") for c in e["cabs"]) + ("/salvage" if e["salvage"] else "")


def _set_plan(e):
    return cab.scan([open(os.path.join(CABDIR, c), "rb").read() for c in e["cabs"]], cab.SALVAGE if e["salvage"] else 0)


@pytest.mark.parametrize("entry", SETS, ids=_set_id)
def test_set_scan_merges_folders_like_the_reference(entry):
    """msgpu_cab_scan_set: as many folders as the reference has after append() (cabd.c:878-1000), every member file inside its folder,
    and a folder whose other half is not among the images refused (the reference fails its files with DATAFORMAT / DECRUNCH)."""
    assert entry["open"] == "open 0"
    plan = _set_plan(entry)
    assert len(plan.folders) == max(f["folder"] for f in entry["files"]) + 1
    for fi in range(len(plan.folders)):
        recs = [r for r in entry["files"] if r["folder"] == fi]
        if all(r["err"] == 0 for r in recs) and not entry["salvage"]:
            assert int(plan.folders["scan_status"][fi]) == 0, (fi, plan.folders[fi])
            assert max(r["offset"] + r["length"] for r in recs) <= int(plan.folders["out_len"][fi])


@pytest.mark.gpu
@pytest.mark.parametrize("entry", SETS, ids=_set_id)
def test_set_decode_matches_reference(decoder, entry):
    """Every file the reference extracts from the set (or from the damaged cabinet in salvage mode) has the same bytes in the
    merged folder's output; a folder none of whose files the reference can extract reports an error."""
    plan = _set_plan(entry)
    out, st = plan.decode(decoder)
    for fi in range(len(plan.folders)):
        recs = [r for r in entry["files"] if r["folder"] == fi]
        base, olen = int(plan.folders["out_off"][fi]), int(plan.folders["out_len"][fi])
        good = [r for r in recs if r["err"] == 0]
        if recs and not good:
            assert int(st[fi]) != 0, (fi, int(st[fi]))
        if recs and len(good) == len(recs):
            assert int(st[fi]) == 0, (fi, int(st[fi]), plan.folders[fi])
        for r in good:
            if r["offset"] + r["length"] <= olen and r.get("written") == r["length"]:
                data = out[base + r["offset"]: base + r["offset"] + r["length"]].tobytes()
                assert hashlib.md5(data).hexdigest() == r["md5"], (_set_id(entry), fi, r["index"])
