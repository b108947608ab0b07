"""SURVEY.md 8 f3 for LZX / Quantum folders: long units (many 32 KiB frames decoded in order by one lane).

A batch of few long units gets up to 64 frame slots per unit and launch round (msgpu.cu pick_fmax / frame_slots), so the
reference's 65 535-block folders (cabextract/test/cabs/large-files-cab.cab -> large-files.cab: mszip-2gb.txt, lzx15-2gb.txt,
lzx21-2gb.txt, MD5 d64bf04a56027b97ac17d751aba2d291 each, cabextract/test/large-files.test:20-22) decode in 1 024 launch rounds
instead of 32 768.  Runs after the core parity tests."""
import hashlib
import os
import time

import numpy as np
import pytest

from libmspack_b200 import cab, gen
from libmspack_b200.units import CODEC_LZX, CODEC_MSZIP, CODEC_QUANTUM
from util import golden_manifest, golden_unit
from test_gpu_parity import _check_batch

pytestmark = pytest.mark.gpu

LARGE_MD5 = "d64bf04a56027b97ac17d751aba2d291"      # cabextract/test/large-files.test:20-22, all three member files


def _long_mixed_batch():
    """long LZX / Quantum / MSZIP units (13 - 40 frames, ragged ends, reset intervals, uncompressed blocks, E8) next to short ones"""
    parts = [gen.make_batch(CODEC_LZX, 3, unit_bytes=40 * 32768 + 777, block_mode=4, split=2, intel=1, data="binary"),
             gen.make_batch(CODEC_LZX, 2, unit_bytes=33 * 32768, reset_interval=4, window_bits=16, first_unit=10),
             gen.make_batch(CODEC_QUANTUM, 3, unit_bytes=20 * 32768 + 5, window_bits=17, first_unit=20),
             gen.make_batch(CODEC_MSZIP, 3, unit_bytes=30 * 32768 + 100, first_unit=30),
             gen.make_batch(CODEC_LZX, 40, first_unit=100), gen.make_batch(CODEC_QUANTUM, 40, unit_bytes=4097, first_unit=200),
             gen.make_batch(CODEC_MSZIP, 40, unit_bytes=65536, first_unit=300)]
    m = gen.concat_batches(parts)
    perm = np.random.default_rng(11).permutation(m.n)
    m.units = m.units[perm].copy()
    return m


@pytest.mark.parametrize("fmax", [None, "2", "5", "64", "4096"])
def test_long_units_frames_per_round(decoder, oracle_ref, fmax, monkeypatch):
    """the same bytes and status as the reference whatever the number of frame slots per unit (default rule, 2 = the round-1
    behaviour, an odd count, more slots than any unit has frames)"""
    if fmax is None:
        monkeypatch.delenv("MSGPU_FMAX", raising=False)
    else:
        monkeypatch.setenv("MSGPU_FMAX", fmax)
    m = _long_mixed_batch()
    out, st = _check_batch(decoder, oracle_ref, m, f"long units MSGPU_FMAX={fmax}")
    assert (st == 0).all()


def test_long_units_corrupt(decoder, oracle_ref, monkeypatch):
    """damaged long units fail in the frame where the reference fails, with its error, after the same bytes"""
    monkeypatch.delenv("MSGPU_FMAX", raising=False)
    m = _long_mixed_batch()
    rng = np.random.default_rng(3)
    comp, units = m.comp.copy(), m.units.copy()
    for i, u in enumerate(units):
        lo, n = int(u["in_off"]), int(u["in_len"])
        if n < 64 or int(u["out_len"]) < 5 * 32768:
            continue
        if i % 2 == 0:
            comp[lo + n // 2 + int(rng.integers(0, n // 4))] ^= 1 << int(rng.integers(0, 8))
        else:
            units["in_len"][i] = n - n // 3
    c = gen.Batch(units, comp, None, m.out_bytes, m.out_init)
    _check_batch(decoder, oracle_ref, c, "corrupt long units")


@pytest.mark.skipif(os.environ.get("MSGPU_SKIP_LARGE") == "1", reason="MSGPU_SKIP_LARGE=1")
def test_large_files_cab_65535_block_folders(decoder):
    """cabextract/test/large-files.test: the outer cabinet's 449-frame LZX folder is large-files.cab (MD5 asserted by the golden
    manifest); its three folders of 65 535 CFDATA blocks each (MSZIP as a block chain, LZX window 2^15 and 2^21 as one lane each)
    are decoded as ONE batch through the cabinet front end and must give the MD5 the reference's test asserts.  Measured: 27.6 s on
    a B200 (profiles/r2_pytest_long_units_p.txt); the bound below is only a guard against falling back to two frames per round."""
    entry = [e for e in golden_manifest() if e["name"] == "large-files-cab.f0"][0]
    u, comp = golden_unit(entry)
    inner, st = decoder.decode_host(u, comp, entry["out_len"])
    assert int(st[0]) == 0 and hashlib.md5(inner.tobytes()).hexdigest() == entry["md5"]
    plan = cab.scan(inner.tobytes())
    assert [int(x) for x in plan.folders["num_blocks"]] == [65535, 65535, 65535]
    assert [plan.file_name(i) for i in range(3)] == [b"mszip-2gb.txt", b"lzx15-2gb.txt", b"lzx21-2gb.txt"]
    t0 = time.perf_counter()
    out, fst = plan.decode(decoder)
    dt = time.perf_counter() - t0
    print(f"large-files.cab: {plan.out_bytes / 1e9:.2f} GB in {dt:.1f} s through msgpu_cab_decode_host, {decoder.launches} launches so far")
    assert [int(x) for x in fst] == [0, 0, 0]
    for f in plan.files:
        fo = int(plan.folders["out_off"][int(f["folder"])]) + int(f["offset"])
        assert hashlib.md5(out[fo:fo + int(f["length"])]).hexdigest() == LARGE_MD5
    assert dt < 240.0, f"65 535-block folders took {dt:.1f} s"
