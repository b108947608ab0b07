"""cpu: the drop-in's HOST logic (libmspack_b200/csrc/mspack_dropin.c: input slurp in input_buffer_size reads, one decode-ahead per
stream, lazy errors, exact decode behind a failed decode-ahead, replay through system->write) driven by the reference's own cabd.c +
system.c, with the device code answered by the host emulation (oracle/_ref/cabx_emul = oracle/ref_cabx.c + tests/emul/emul_abi.cpp).
Same expectations as the gpu test tests/test_y_reference_suites_gpu.py::test_reference_cabd_extracts_through_the_dropin: per member
file the MSPACK_ERR_* and the bytes the UNMODIFIED reference produced (tests/golden/cab/manifest.json)."""
import hashlib
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "cabx_emul")


def _cab_manifest():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "cab", "manifest.json")))


@pytest.mark.parametrize("entry", [e for e in _cab_manifest() if not e["open"] and e["files"]], ids=lambda e: e["name"])
def test_reference_cabd_extracts_through_the_dropin_on_the_emulation(entry, tmp_path):
    if not os.path.exists(EXE):
        pytest.skip(f"{EXE} not built (needs /root/reference at build time)")
    r = subprocess.run([EXE, os.path.join(ROOT, "tests", "golden", "cab", entry["name"]), str(tmp_path)], capture_output=True, timeout=600, cwd=ROOT)
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
