"""KWAJ files with MSZIP data (mszipd_decompress_kwaj, mszipd.c:462-495, called by kwajd.c:320-322) through
the reference's own kwajd.c: with the reference's mszipd.c (CPU: pins the KWAJ writer below to the reference) and with
mszipd_* coming from the GPU drop-in (gpu).  The reference ships no MSZIP KWAJ fixture, so the files are generated."""
import os
import struct
import subprocess

import pytest

from util import kwaj_mszip_stream

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KWAJX_REF = os.path.join(ROOT, "oracle", "_ref", "kwajx_ref")
KWAJX_GPU = os.path.join(ROOT, "oracle", "_ref", "kwajx_gpu")


def kwaj_file(payload: bytes) -> bytes:
    """kwaj.h:16-21: "KWAJ" 88 F0 27 D1, method (4 = MSZIP), data offset, header flags (none)."""
    return b"KWAJ\x88\xf0\x27\xd1" + struct.pack("<HHH", 4, 14, 0) + payload


def _run(tool, path_in, path_out):
    r = subprocess.run([tool, path_in, path_out], capture_output=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    return r.stdout.decode().strip(), (open(path_out, "rb").read() if os.path.exists(path_out) else b"")


def _cases():
    good, out = kwaj_mszip_stream([32768, 32768, 100, 32768, 9000], seed=3)
    big, bigout = kwaj_mszip_stream([32768] * 40 + [123], seed=4)              # output 12x the input: the drop-in grows its area
    zeros, zout = kwaj_mszip_stream([32768] * 4, seed=5)
    bad = bytearray(good); bad[len(bad) // 2:len(bad) // 2 + 64] = bytes(64)        # (deflate has no checksum: what comes of it is whatever the reference says)
    sig = bytearray(good); sig[3] = ord("X")
    return [("good", good, out), ("big", big, bigout), ("more", zeros, zout), ("noterm", good[:-2], out), ("damaged", bytes(bad), None),
            ("badsig", bytes(sig), b""), ("empty", b"\0\0", b"")]


def _check(tool, tmp_path):
    res = {}
    for name, payload, expect in _cases():
        pin, pout = str(tmp_path / (name + ".kwj")), str(tmp_path / (name + ".out"))
        open(pin, "wb").write(kwaj_file(payload))
        err, data = _run(tool, pin, pout)
        res[name] = (err, data)
        if expect is not None and name != "badsig":
            assert err == "err 0", (name, err)
            assert data == expect, name
    assert res["badsig"][0] == "err 8"
    return res


@pytest.mark.skipif(not os.path.exists(KWAJX_REF), reason="oracle/_ref/kwajx_ref not built (needs /root/reference)")
def test_kwaj_files_decode_with_the_reference(tmp_path):
    _check(KWAJX_REF, tmp_path)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(KWAJX_GPU), reason="oracle/_ref/kwajx_gpu not built")
def test_kwaj_files_decode_through_the_gpu_dropin(tmp_path):
    got = _check(KWAJX_GPU, tmp_path)
    if os.path.exists(KWAJX_REF):
        (tmp_path / "r").mkdir()
        want = _check(KWAJX_REF, tmp_path / "r")
        for name in want:
            assert got[name][0] == want[name][0], name                    # same error code ...
            assert got[name][1] == want[name][1], name                    # ... and the same bytes written before it


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(4))
def test_mszip_repair_mode_on_the_gpu(decoder, oracle_ref, seed):
    """mszipd_init(repair_mode = 1) through the batch ABI (MSGPU_FLAG_MSZIP_REPAIR + the input buffer size): damaged MSZIP folders
    against the reference in repair mode - zero-filled blocks, stale-state restarts, overflowing blocks."""
    import numpy as np
    from util import assert_same, damaged_mszip_batch
    units, comp, out_bytes = damaged_mszip_batch(300 + seed, n=64, level=(6, 1, 0)[seed % 3], data=("text", "binary")[seed % 2])
    out_g, st_g = decoder.decode_host(units, comp, out_bytes)
    out_o, st_o, _ = oracle_ref.decode_batch(units, comp, out_bytes, threads=8)
    assert_same(units, out_o, st_o, out_g, st_g, f"repair seed {seed}")
    assert (st_o == 0).any()
