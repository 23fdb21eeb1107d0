"""Batch sharding for independent transforms: contiguous split of the batch over devices / ranks, no
data-path collective.  The arithmetic lives in the library (`CkFftB200ShardRange`, csrc/multi.cu) -- it is what the
multi-device scheduler behind `CkFft*BatchMulti` uses to cut a batch -- and this module is its Python mirror, so that
a torchrun job (one process per GPU, bench.py) and the one-process scheduler shard a batch identically.
Pure host logic (no GPU needed)."""
from __future__ import annotations

import ctypes as C

from . import _lib


def shard_range(batch: int, rank: int, world: int) -> tuple[int, int]:
    """[start, stop) of rank's contiguous shard; the first batch % world ranks get one extra transform."""
    first, count = C.c_size_t(0), C.c_size_t(0)
    if batch < 0 or not _lib.load().CkFftB200ShardRange(int(batch), int(rank), int(world), C.byref(first), C.byref(count)):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return int(first.value), int(first.value + count.value)


def all_shards(batch: int, world: int) -> list[tuple[int, int]]:
    return [shard_range(batch, r, world) for r in range(world)]


def job_throughput(units_per_rank: list[int], ms_per_rank: list[float]) -> float:
    """units of all ranks divided by the slowest rank's time (units per millisecond)."""
    return sum(units_per_rank) / max(ms_per_rank)
