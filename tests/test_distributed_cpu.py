"""CPU (gloo, world size 2 and 4) test of the distributed six-step host logic: slab arithmetic, exchange order
and the index algebra of ckfft_b200/distributed.py, with the oracle standing in for the GPU kernels."""
import os
import socket
import sys

import numpy as np
import pytest

from ckfft_b200.distributed import NumpyBackend, six_step, split_n

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_split():
    assert split_n(1 << 30, 8) == (1 << 15, 1 << 15)
    assert split_n(1 << 21, 2) == (1 << 10, 1 << 11)
    with pytest.raises(ValueError):
        split_n(1000, 2)
    with pytest.raises(ValueError):
        split_n(16, 8)


def test_single_rank_matches_fft():
    rng = np.random.default_rng(0)
    n = 1 << 12
    x = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)
    be = NumpyBackend(lambda a, inv: (np.fft.ifft(a, axis=1) * a.shape[1] if inv else np.fft.fft(a, axis=1)).astype(np.complex64))
    y = six_step(x.copy(), n, 0, 1, be)
    assert np.linalg.norm(y - np.fft.fft(x.astype(np.complex128))) / np.linalg.norm(np.fft.fft(x.astype(np.complex128))) < 1e-6


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    import oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 1 << 14
    rng = np.random.default_rng(7)
    x = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)      # same on every rank
    orc = oracle.Restatement(n, 3)
    be = NumpyBackend(lambda a, inv: orc.complex(np.ascontiguousarray(a), inv), dist)
    per = n // world
    ok = True
    for inverse in (False, True):
        mine = six_step(x[rank * per:(rank + 1) * per].copy(), n, rank, world, be, inverse)
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        if rank == 0:
            full = np.concatenate(parts)
            want = orc.complex(x, inverse)
            err = np.linalg.norm(full.astype(np.complex128) - want) / np.linalg.norm(want)
            ok = ok and err < 1e-6 * 14
    if rank == 0:
        open(os.path.join(tmpdir, f"ok{world}"), "w").write("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_gloo_six_step(tmp_path, world):
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert (tmp_path / f"ok{world}").read_text() == "1"
