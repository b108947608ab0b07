"""gpu: the reference's OWN regression programs (libmspack/test/cabd_test.c, chmd_test.c) compiled from the reference
sources where they lie together with the reference's container parsers (cabd.c / chmd.c / system.c), but with
lzxd_* / qtmd_* / mszipd_* coming from libmspack_dropin.so, i.e. from the GPU kernels (oracle/Makefile targets
cabd_gpu / chmd_gpu; the binaries are built in this container and travel to the GPU box).  This is the drop-in
boundary of SURVEY.md 8(b) exercised by the reference's own tests."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("binary,expect", [("cabd_test_gpu", "ALL 433 TESTS PASSED"), ("chmd_test_gpu", "ALL 195 TESTS PASSED")])
def test_reference_test_program_passes_on_the_dropin(binary, expect):
    exe = os.path.join(ROOT, "oracle", "_ref", binary)
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (needs /root/reference at build time)")
    r = subprocess.run([exe], capture_output=True, timeout=900, cwd=ROOT)   # fixture paths are relative to the repo root
    out = (r.stdout + r.stderr).decode("utf-8", "replace")                   # (the programs print raw member names)
    assert r.returncode == 0, out[-2000:]
    assert expect in out, out[-2000:]


def _cab_manifest():
    import json
    return json.load(open(os.path.join(ROOT, "tests", "golden", "cab", "manifest.json")))


@pytest.mark.parametrize("entry", [e for e in _cab_manifest() if not e["open"] and e["files"]], ids=lambda e: e["name"])
def test_reference_cabd_extracts_through_the_dropin(entry, tmp_path):
    """The reference's cabd.c + system.c extracting every member file of the golden cabinets with the codecs coming from the GPU
    drop-in (oracle/_ref/cabx_gpu = oracle/ref_cabx.c): per file the MSPACK_ERR_* and the bytes the unmodified reference produced
    (tests/golden/cab/manifest.json).  Covers what cabd does to a codec that a batch never sees: two X_decompress calls per file
    (skip, extract), many files per folder served from ONE decode, the late lzxd_set_output_length, and - the synth_badsum
    cabinets - a read error from a damaged late block that must not fail the files in front of it (cabd.c:1322-1324)."""
    import hashlib
    exe = os.path.join(ROOT, "oracle", "_ref", "cabx_gpu")
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (needs /root/reference at build time)")
    r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "cab", entry["name"]), str(tmp_path)], capture_output=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.decode().split("\n")
    assert lines[0] == "open 0"
    got = {}
    for ln in lines[1:]:
        f = ln.split()
        if len(f) == 5:
            got[int(f[0])] = int(f[4])
    for rec in entry["files"]:
        assert got.get(rec["index"]) == rec["err"], (entry["name"], rec["index"], got.get(rec["index"]), rec["err"])
        if rec["err"] == 0:
            data = open(os.path.join(str(tmp_path), str(rec["index"])), "rb").read()
            assert hashlib.md5(data).hexdigest() == rec["md5"], (entry["name"], rec["index"])
