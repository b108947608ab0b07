"""The reference's Offline Address Book path (oabd.c) - the caller of lzxd_init(is_delta=1) and lzxd_set_reference_data
 - over generated OAB files: with the reference's own lzxd.c (CPU: pins the OAB writer and the
LZX DELTA encoder to the reference) and with lzxd_* coming from the GPU drop-in (gpu: the kernels behind the same caller)."""
import os
import subprocess

import numpy as np
import pytest

from libmspack_b200 import gen
import oabfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OABX_REF = os.path.join(ROOT, "oracle", "_ref", "oabx_ref")
OABX_GPU = os.path.join(ROOT, "oracle", "_ref", "oabx_gpu")


def _data():
    target = gen.raw_units(1, 700000, data="text").tobytes()
    rng = np.random.default_rng(21)
    base = bytearray(target[3000:] + target[:1000])                 # the "older version": moved, with scattered edits
    for p in rng.integers(0, len(base), 300):
        base[int(p)] ^= 0x55
    return bytes(base), target


def _run(tool, args, tmp_path):
    r = subprocess.run([tool] + args, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    return r.stdout.strip()


def _check(tool, tmp_path):
    base, target = _data()
    p = lambda n: str(tmp_path / n)
    open(p("base"), "wb").write(base)
    # full file: LZX DELTA blocks (window_bits 18 for 200 000-byte blocks), every third block stored
    open(p("full.oab"), "wb").write(oabfile.full_oab(target, stored_every=3))
    assert _run(tool, ["full", p("full.oab"), p("full.out")], tmp_path) == "err 0"
    assert open(p("full.out"), "rb").read() == target
    # incremental patch: every block decodes against its slice of the base file (reference data)
    open(p("patch.oab"), "wb").write(oabfile.patch_oab(base, target))
    assert _run(tool, ["patch", p("patch.oab"), p("base"), p("patch.out")], tmp_path) == "err 0"
    assert open(p("patch.out"), "rb").read() == target
    # a damaged LZX block: DECRUNCH (11) or the block CRC (CHECKSUM 9) - whatever the reference says, the drop-in must say too
    bad = bytearray(oabfile.patch_oab(base, target, block_mode=1))
    bad[len(bad) // 2] ^= 0x10
    open(p("bad.oab"), "wb").write(bytes(bad))
    return _run(tool, ["patch", p("bad.oab"), p("base"), p("bad.out")], tmp_path)


@pytest.mark.skipif(not os.path.exists(OABX_REF), reason="oracle/_ref/oabx_ref not built (needs /root/reference)")
def test_oab_files_decode_with_the_reference(tmp_path):
    assert _check(OABX_REF, tmp_path) in ("err 11", "err 9", "err 3")


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(OABX_GPU), reason="oracle/_ref/oabx_gpu not built")
def test_oab_files_decode_through_the_gpu_dropin(tmp_path):
    got = _check(OABX_GPU, tmp_path)
    if os.path.exists(OABX_REF):
        (tmp_path / "r").mkdir()
        assert got == _check(OABX_REF, tmp_path / "r")
    else:
        assert got in ("err 11", "err 9", "err 3")
