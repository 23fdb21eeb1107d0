"""CPU tests of the multi-GPU host logic: shard arithmetic, and a world-size-2 gloo run in which each
rank transforms its shard (the oracle stands in for the GPU here: tests may use it) and rank 0 checks
that the gathered shards equal the single-process result, with the max-over-ranks timing reduction
bench.py uses."""
import os
import socket
import sys

import numpy as np
import pytest

from ckfft_b200.sharding import all_shards, job_throughput, shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("batch", [0, 1, 7, 8, 1000, 1 << 20])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_shards_partition_the_batch(batch, world):
    shards = all_shards(batch, world)
    assert shards[0][0] == 0 and shards[-1][1] == batch
    for (a0, a1), (b0, b1) in zip(shards, shards[1:]):
        assert a1 == b0
    sizes = [b - a for a, b in shards]
    assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(batch, world, world)


def test_job_throughput_uses_slowest_rank():
    assert job_throughput([10, 10], [1.0, 2.0]) == 10.0


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, batch = 256, 37
    rng = np.random.default_rng(42)          # same data on every rank
    x = (rng.uniform(-1, 1, (batch, n)) + 1j * rng.uniform(-1, 1, (batch, n))).astype(np.complex64)
    lo, hi = shard_range(batch, rank, world)
    orc = oracle.Restatement(n, 3)
    mine = orc.complex(x[lo:hi])
    # no collective on the data path; gather only to verify
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, mine))
    t = torch.tensor([1.0 + rank], dtype=torch.float64)   # pretend device time: max over ranks
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        full = np.concatenate([g[2] for g in sorted(gathered, key=lambda g: g[0])])
        want = orc.complex(x)
        ok = np.array_equal(full.view(np.uint32), want.view(np.uint32)) and float(t.item()) == float(world)
        open(os.path.join(tmpdir, "ok"), "w").write("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_batch_sharding(tmp_path):
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").read_text() == "1"
