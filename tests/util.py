"""Shared helpers for the tests."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def golden_manifest():
    return json.load(open(os.path.join(GOLDEN, "manifest.json")))


def golden_unit(entry):
    """(units ndarray[1], comp uint8[]) for one golden manifest entry."""
    from libmspack_b200.units import UNIT_DTYPE
    data = open(os.path.join(GOLDEN, "units", entry["name"] + ".in"), "rb").read()
    u = np.zeros(1, dtype=UNIT_DTYPE)
    u["codec"], u["window_bits"], u["reset_interval"] = entry["codec"], entry["window_bits"], entry["reset_interval"]
    u["in_len"], u["out_len"] = len(data), entry["out_len"]
    comp = np.frombuffer(data + b"\0" * 16, dtype=np.uint8).copy()
    return u, comp


def golden_expected(entry):
    p = os.path.join(GOLDEN, "units", entry["name"] + ".out")
    return open(p, "rb").read() if os.path.exists(p) else None


def assert_same(units, out_a, st_a, out_b, st_b, what=""):
    """Bit-exact on every unit both decode without error; identical status everywhere."""
    assert np.array_equal(st_a, st_b), (f"{what}: status differs at {np.nonzero(st_a != st_b)[0][:8]}: "
                                        f"{st_a[st_a != st_b][:8]} vs {st_b[st_a != st_b][:8]}")
    if np.array_equal(out_a, out_b):
        return
    for i, u in enumerate(units):
        if st_a[i] != 0:
            continue
        lo, n = int(u["out_off"]), int(u["out_len"])
        if not np.array_equal(out_a[lo:lo + n], out_b[lo:lo + n]):
            d = np.nonzero(out_a[lo:lo + n] != out_b[lo:lo + n])[0]
            raise AssertionError(f"{what}: unit {i} differs at byte {d[0]} ({len(d)} bytes differ)")
