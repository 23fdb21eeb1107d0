"""Batch sharding for independent transforms: contiguous split of the batch over ranks,
one process per GPU, no data-path collective.  Pure host logic (no GPU needed)."""
from __future__ import annotations


def shard_range(batch: int, rank: int, world: int) -> tuple[int, int]:
    """[start, stop) of rank's contiguous shard; the first batch % world ranks get one extra transform."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(batch, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def all_shards(batch: int, world: int) -> list[tuple[int, int]]:
    return [shard_range(batch, r, world) for r in range(world)]


def job_throughput(units_per_rank: list[int], ms_per_rank: list[float]) -> float:
    """units of all ranks divided by the slowest rank's time (units per millisecond)."""
    return sum(units_per_rank) / max(ms_per_rank)
