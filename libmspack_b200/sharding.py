"""Multi-GPU layer: units are independent (a CAB folder / a CHM reset interval never reads another unit's
bytes), so rank r of R simply owns units [floor(r*n/R), floor((r+1)*n/R)) - no data-path
collective.  torch.distributed is used for the control plane only (barriers, max-over-ranks timing, the
optional output gather in bench.py)."""


def shard_range(n: int, rank: int, world: int):
    """Half-open unit-index range owned by `rank`."""
    return (n * rank) // world, (n * (rank + 1)) // world
