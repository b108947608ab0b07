"""ms_div_rcp (msgpu_p1_qtm.cuh, QtmLane OPT bit 2): the division through a float reciprocal is exact whenever it does not fall back to
the plain division - also with a reciprocal that is a few ulp off (the GPU's approximate MUFU.RCP; the host emulation divides exactly).
Restated here in numpy float32 arithmetic, checked on random and on boundary operands."""
import numpy as np


def div_rcp(n, d, ulp_off):
    n = n.astype(np.uint32); d = d.astype(np.uint32)
    rd = (np.float32(1.0) / d.astype(np.float32)).astype(np.float32)
    rd = np.nextafter(rd, np.float32(np.inf if ulp_off > 0 else -np.inf), dtype=np.float32) if ulp_off else rd
    for _ in range(max(0, abs(ulp_off) - 1)):
        rd = np.nextafter(rd, np.float32(np.inf if ulp_off > 0 else -np.inf), dtype=np.float32)
    qf = (n.astype(np.float32) * rd).astype(np.float32)
    fast = qf < np.float32(1048576.0)
    q = np.where(fast, qf, 0).astype(np.uint32)
    r = (n.astype(np.int64) - q.astype(np.int64) * d.astype(np.int64))
    r = ((r + 2**31) % 2**32 - 2**31)                      # the int32 view of the uint32 difference
    q = np.where(r < 0, q - 1, np.where(r >= d.astype(np.int64), q + 1, q)).astype(np.uint32)
    return np.where(fast, q, n // d), fast


def test_div_rcp_exact():
    rng = np.random.default_rng(1)
    N = 400000
    cases = []
    # GET_SYMBOL's operands: (C - L + 1) * total - 1 over range (quotient < total), and cum * range over total (quotient <= range)
    rngs = rng.integers(1, 65537, N, dtype=np.int64); tot = rng.integers(7, 3909, N, dtype=np.int64)
    cl = (rng.random(N) * rngs).astype(np.int64) + 1
    cases.append((cl * tot - 1, rngs))
    cum = (rng.random(N) * (tot + 1)).astype(np.int64)
    cases.append((cum * rngs, tot))
    # quotients right at multiples of the divisor, and arbitrary operands (damaged streams)
    d = rng.integers(1, 65537, N, dtype=np.int64); k = rng.integers(0, 1 << 20, N, dtype=np.int64)
    for delta in (-1, 0, 1):
        n = np.clip(k * d + delta, 0, 2**32 - 1)
        cases.append((n, d))
    cases.append((rng.integers(0, 2**32, N, dtype=np.int64), rng.integers(1, 65537, N, dtype=np.int64)))
    cases.append((rng.integers(0, 2**32, N, dtype=np.int64), rng.integers(1, 4000, N, dtype=np.int64)))
    for n, d in cases:
        for off in (-3, -1, 0, 1, 3):
            q, fast = div_rcp(n, d, off)
            assert np.array_equal(q.astype(np.int64), n // d), (off, int(np.argmax(q.astype(np.int64) != n // d)))
    assert fast.any() and not fast.all()
