"""CPU (gloo, world size 2 and 4) test of the distributed six-step host logic: slab arithmetic, exchange order
and the index algebra of ckfft_b200/distributed.py, with the oracle standing in for the GPU kernels."""
import os
import socket
import sys

import numpy as np
import pytest

from ckfft_b200.distributed import six_step, split_n
from dist_replay import NumpyBackend

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
    sys.path.insert(0, os.path.join(ROOT, "tests"))
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


# ---- fused variant (peer stores instead of collectives): layout + routing algebra of csrc/dist_fused.cu -----------
def _synthetic_layout(la, lb, lc, ld, world):
    """A layout with small factors (the kernels need L >= 128, the host arithmetic does not)."""
    from ckfft_b200 import _lib

    lay = _lib.DistLayout()
    n1, n2 = la * lb, lc * ld
    lay.log2n1, lay.log2n2 = n1.bit_length() - 1, n2.bit_length() - 1
    lay.log2n, lay.world = lay.log2n1 + lay.log2n2, world
    lay.la, lay.lb, lay.lc, lay.ld = la, lb, lc, ld
    lay.passes = 2 + (la > 1) + (lc > 1)
    return lay


def test_fused_layout_rules():
    from ckfft_b200.distributed import fused_layout

    for lg in range(14, 31):
        for world in (1, 2, 4, 8):
            lay = fused_layout(1 << lg, world)
            assert lay.la * lay.lb == 1 << lay.log2n1 and lay.lc * lay.ld == 1 << lay.log2n2
            assert lay.log2n1 + lay.log2n2 == lg
            assert lay.passes == 2 + (lay.la > 1) + (lay.lc > 1)
            for L in (lay.la, lay.lb, lay.lc, lay.ld):
                assert L == 1 or 128 <= L <= 1024          # a tile-kernel length
            assert (1 << lay.log2n1) // world >= 16 and (1 << lay.log2n2) // world >= 16
    assert fused_layout(1 << 30, 8).passes == 4 and fused_layout(1 << 30, 8, 3).passes == 3
    assert fused_layout(1 << 20, 2).passes == 2 and fused_layout(1 << 24, 2).passes == 3
    for bad in (1 << 13, 1 << 31, 3 << 14):
        with pytest.raises(ValueError):
            fused_layout(bad, 2)
    with pytest.raises(ValueError):
        fused_layout(1 << 20, 3)


@pytest.mark.parametrize("shape", [(1, 128, 1, 128, 1), (1, 128, 1, 128, 8), (1, 128, 128, 128, 2), (16, 16, 16, 16, 2),
                                   (8, 32, 16, 8, 4), (1, 64, 8, 16, 2), (16, 8, 1, 128, 8)])
@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("pull", [False, True])
def test_fused_replay_matches_fft(shape, inverse, pull):
    """All ranks simulated in one process: exchange (push layouts) or peer reads of the first pass (pull layouts) +
    the pass descriptors reproduce the N-point transform."""
    from ckfft_b200.distributed import fused_layout
    from dist_replay import replay_fused

    la, lb, lc, ld, world = shape
    lay = fused_layout(la * lb * lc * ld, world) if min(lb, ld) >= 128 and la in (1,) and lc in (1, 128) else \
        _synthetic_layout(la, lb, lc, ld, world)
    lay.pull = int(pull)
    assert (lay.la, lay.lb, lay.lc, lay.ld) == (la, lb, lc, ld)
    n = 1 << lay.log2n
    rng = np.random.default_rng(3)
    x = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)
    per = n // world
    y = np.concatenate(replay_fused([x[r * per:(r + 1) * per] for r in range(world)], lay, inverse))
    want = np.fft.ifft(x.astype(np.complex128)) * n if inverse else np.fft.fft(x.astype(np.complex128))
    assert np.linalg.norm(y - want) / np.linalg.norm(want) < 1e-6


def _fused_worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    from ckfft_b200.distributed import fused_layout
    from dist_replay import replay_fused

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def deliver(buf_id, outgoing, my_bufs):
        # every rank sends the same amount to every peer: one all-to-all of the indices, one of the values
        idx = torch.from_numpy(np.concatenate([outgoing[q][0].astype(np.int64) for q in range(world)]))
        val = torch.from_numpy(np.concatenate([outgoing[q][1].astype(np.complex64) for q in range(world)]).view(np.float32).copy())
        ridx, rval = torch.empty_like(idx), torch.empty_like(val)
        dist.all_to_all_single(ridx, idx)
        dist.all_to_all_single(rval, val)
        my_bufs[buf_id][ridx.numpy()] = rval.numpy().view(np.complex64)

    ok = True
    for n, pull in ((1 << 14, False), (1 << 21, False), (1 << 21, True)):
        lay = fused_layout(n, world, pull=pull)
        rng = np.random.default_rng(11)
        x = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)      # same on every rank
        per = n // world
        # push layouts touch only this rank's slice; a pull layout reads the peers' input arrays (here: their copies)
        slices = {q: x[q * per:(q + 1) * per] for q in (range(world) if pull else [rank])}
        mine = replay_fused(slices, lay, False, deliver=deliver, ranks=[rank])[0]
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        if rank == 0:
            want = np.fft.fft(x.astype(np.complex128))
            ok = ok and np.linalg.norm(np.concatenate(parts) - want) / np.linalg.norm(want) < 1e-6
    if rank == 0:
        open(os.path.join(tmpdir, f"fused{world}"), "w").write("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_fused_routing_world2(tmp_path):
    """World-size-2 run of the fused transform's host logic: every routed store crosses the process boundary."""
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_fused_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "fused2").read_text() == "1"
