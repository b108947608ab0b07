"""CPU-only, world_size 2 over gloo: the multi-GPU layer is "shard units by index, no data-path collective"
(SURVEY.md 8e).  Two processes each take their slice of one batch, decode it (with the oracle standing in for
the device on this GPU-less box), and the concatenation must equal the single-process result."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from libmspack_b200 import gen
    from libmspack_b200.sharding import shard_range
    from libmspack_b200.units import CODEC_LZX
    from oracle import oracle as orc
    n = 101
    lo, hi = shard_range(n, rank, world)
    b = gen.make_batch(CODEC_LZX, hi - lo, first_unit=lo, threads=2)        # every rank generates ONLY its shard
    out, st, _ = orc.load("reference").decode_batch(b.units, b.comp, b.out_bytes)
    t = torch.tensor([float((st == 0).sum())])
    dist.all_reduce(t)                                                       # control plane only (counts / timings)
    q.put((rank, lo, hi, out.tobytes(), float(t.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_equals_single_process():
    from libmspack_b200 import gen
    from libmspack_b200.sharding import shard_range
    from libmspack_b200.units import CODEC_LZX
    from oracle import oracle as orc
    world, n = 2, 101
    assert [shard_range(n, r, world) for r in range(world)] == [(0, 50), (50, 101)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, 29611, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(60)
    full = gen.make_batch(CODEC_LZX, n, threads=2)
    ref, st, _ = orc.load("reference").decode_batch(full.units, full.comp, full.out_bytes)
    assert (st == 0).all()
    assert b"".join(r[3] for r in res) == ref.tobytes()
    assert all(r[4] == n for r in res)


def test_shard_ranges_cover_the_batch_and_keep_chains_whole():
    """msgpu_shard_range (the split msgpu_decode_batch_host_multi and bench.py's ranks use): the shards tile [0, n) in order for any
    shard count, equal floor(r n / R) without chains, and never start on a CHAIN_NEXT unit."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from libmspack_b200 import gen
    from libmspack_b200.sharding import shard_range
    from libmspack_b200.units import CODEC_MSZIP
    from util import chain_batch
    b = gen.make_batch(CODEC_MSZIP, 101, unit_bytes=2000)
    for world in (1, 2, 3, 8, 64):
        rs = [shard_range(b.n, r, world, b.units) for r in range(world)]
        assert rs == [shard_range(b.n, r, world) for r in range(world)]
        assert rs[0][0] == 0 and rs[-1][1] == b.n and all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
    chain, _, _ = chain_batch([32768 * 5 + 100, 32768 * 3, 32768 * 7 + 1])
    for world in (2, 3, 5):
        rs = [shard_range(chain.n, r, world, chain.units) for r in range(world)]
        assert rs[0][0] == 0 and rs[-1][1] == chain.n and all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
        for lo, hi in rs:
            assert lo == chain.n or not (int(chain.units["flags"][lo]) & 0x8), (world, lo)
