"""GPU tests of the multi-device batched scheduler behind the C ABI (CkFftB200MultiInit / CkFft*BatchMulti,
csrc/multi.cu; north_star item 5): one host call, contiguous shards, one context replica and one host
thread per device, no collective.  Independent transforms share no state (inc/ckfft/ckfft.h:39-41), so the result of
a sharded call must be BIT-identical to the single-device call on the same arrays, whatever the device list is.

On a one-GPU box the device list [0, 0, 0] exercises the same code (three replicas, three threads, three shards on
the one device); with two or more GPUs the same tests also run across distinct devices."""
import ctypes as C
import os

import numpy as np
import pytest

import ckfft_b200 as ck
import oracle
from ckfft_b200 import _lib
from ckfft_b200.sharding import all_shards
from conftest import rel_rms, tolerance, uniform_complex

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def device_lists():
    nd = torch.cuda.device_count()
    lists = [[0], [0, 0, 0]]
    if nd >= 2:
        lists += [list(range(nd)), [nd - 1, 0]]
    return lists


@pytest.mark.parametrize("devices", device_lists(), ids=lambda d: "dev" + "_".join(map(str, d)))
def test_multi_matches_single_device_bitwise(devices):
    rng = np.random.default_rng(len(devices))
    n, batch = 1024, 1001                              # ragged: shards of different length
    x = uniform_complex(rng, (batch, n))
    xr = rng.uniform(-1, 1, (batch, n)).astype(np.float32)
    with ck.Context(n, ck.BOTH) as ctx, ck.MultiContext(n, ck.BOTH, devices) as mc:
        assert mc.devices == devices
        for name, arg in (("complex_forward", x), ("complex_inverse", x), ("real_forward", xr)):
            want = getattr(ctx, name)(arg)
            got = getattr(mc, name)(arg)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), name
        spec = ctx.real_forward(xr)
        assert np.array_equal(mc.real_inverse(spec, n).view(np.uint32), ctx.real_inverse(spec, n).view(np.uint32))
    orc = oracle.Restatement(n, 3)
    with ck.MultiContext(n, ck.BOTH, devices) as mc:
        assert rel_rms(mc.complex_forward(x), orc.complex(x, False)) <= tolerance(n)
    orc.close()


def test_multi_batch_smaller_than_device_list_and_empty():
    rng = np.random.default_rng(5)
    n = 256
    with ck.Context(n, ck.BOTH) as ctx, ck.MultiContext(n, ck.BOTH, [0, 0, 0, 0]) as mc:
        for batch in (1, 2, 3, 5):
            x = uniform_complex(rng, (batch, n))
            assert np.array_equal(mc.complex_forward(x).view(np.uint32), ctx.complex_forward(x).view(np.uint32))
        lib = _lib.load()
        x = uniform_complex(rng, (1, n))
        y = np.empty_like(x)
        assert lib.CkFftComplexForwardBatchMulti(mc._m, n, x.ctypes.data, y.ctypes.data, 0) == 1     # empty batch: nothing to do


def test_multi_shard_plan_is_the_library_shard_range():
    """every replica transforms exactly the rows CkFftB200ShardRange assigns to it: transform rows that differ per
    shard through per-device contexts and compare with the one call"""
    n, batch, devices = 64, 37, [0, 0, 0]
    rng = np.random.default_rng(9)
    x = uniform_complex(rng, (batch, n))
    with ck.MultiContext(n, ck.FORWARD, devices) as mc:
        got = mc.complex_forward(x)
        lib = _lib.load()
        for i, (lo, hi) in enumerate(all_shards(batch, len(devices))):
            part = np.empty((hi - lo, n), np.complex64)
            ctx_i = lib.CkFftB200MultiContext(mc._m, i)
            assert ctx_i and lib.CkFftB200ContextDevice(ctx_i) == devices[i]
            assert lib.CkFftComplexForwardBatch(ctx_i, n, x[lo:hi].ctypes.data, part.ctypes.data, hi - lo) == 1
            assert np.array_equal(part.view(np.uint32), got[lo:hi].view(np.uint32))


def test_multi_error_returns():
    lib = _lib.load()
    nd = torch.cuda.device_count()
    assert not lib.CkFftB200MultiInit(1000, ck.BOTH, None, 0)                      # not a power of two
    assert not lib.CkFftB200MultiInit(1024, 0, None, 0)                            # bad direction
    bad = (C.c_int * 1)(nd)
    assert not lib.CkFftB200MultiInit(1024, ck.BOTH, bad, 1)                       # no such device
    assert "does not exist" in ck.last_error()
    n = 1024
    x = np.zeros((4, n), np.complex64)
    y = np.zeros_like(x)
    with ck.MultiContext(n, ck.FORWARD, [0]) as mc:
        f = lib.CkFftComplexForwardBatchMulti
        assert f(mc._m, n, x.ctypes.data, y.ctypes.data, 4) == 1
        assert f(mc._m, 2 * n, x.ctypes.data, y.ctypes.data, 1) == 0               # n > nMax
        assert f(mc._m, 1000, x.ctypes.data, y.ctypes.data, 1) == 0                # not a power of two
        assert f(mc._m, n, x.ctypes.data, x.ctypes.data, 4) == 0                   # in == out (src/ckfft/ckfft.cpp:88)
        assert f(mc._m, n, None, y.ctypes.data, 4) == 0
        assert f(None, n, x.ctypes.data, y.ctypes.data, 4) == 0
        assert lib.CkFftComplexInverseBatchMulti(mc._m, n, x.ctypes.data, y.ctypes.data, 4) == 0    # forward-only handle
        assert "direction" in ck.last_error()
        d = torch.zeros((4, n), dtype=torch.complex64, device="cuda")
        assert f(mc._m, n, d.data_ptr(), y.ctypes.data, 4) == 0                    # device array: belongs to one GPU
    assert lib.CkFftB200MultiDeviceCount(None) == 0 and lib.CkFftB200MultiDevice(None, 0) == -1


@pytest.mark.parametrize("pin", ["1", "0"])
def test_multi_large_pageable_call(pin, monkeypatch):
    """>= 64 MiB of pageable memory: with CKFFT_B200_PIN=1 the call page-locks the arrays for its duration, by default it
    does not; either way the result is the single-device result and the arrays are ordinary memory again afterwards"""
    monkeypatch.setenv("CKFFT_B200_PIN", pin)
    n, batch = 4096, 3000                                # 98 MB in + 98 MB out
    rng = np.random.default_rng(11)
    x = uniform_complex(rng, (batch, n))
    devices = list(range(torch.cuda.device_count())) if torch.cuda.device_count() >= 2 else [0, 0]
    with ck.Context(n, ck.BOTH) as ctx, ck.MultiContext(n, ck.BOTH, devices) as mc:
        want = ctx.complex_forward(x)
        for _ in range(2):                               # twice: registration / unregistration must be repeatable
            got = mc.complex_forward(x)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    x[0, 0] = 7.0                                        # still writable, still ours
    assert x[0, 0] == 7.0


@pytest.mark.parametrize("static", ["0", "1"])
def test_multi_dynamic_schedule_matches_static(static, monkeypatch):
    """Batches of >= 256 MiB are scheduled dynamically: every device draws 32 MiB chunks of the one batch from a shared
    counter (a device behind a slower host link takes fewer).  Which device transforms which row cannot matter:
    bit-identical to the single-device call, as with static shards (CKFFT_B200_MULTI_STATIC=1)."""
    monkeypatch.setenv("CKFFT_B200_MULTI_STATIC", static)
    n, batch = 4096, 5003                                # 164 MB in + 164 MB out, ragged last chunk
    rng = np.random.default_rng(17)
    x = rng.uniform(-1, 1, (batch, n)).astype(np.float32)
    devices = list(range(torch.cuda.device_count())) if torch.cuda.device_count() >= 2 else [0, 0, 0]
    with ck.Context(n, ck.BOTH) as ctx, ck.MultiContext(n, ck.BOTH, devices) as mc:
        want = ctx.real_forward(x)
        got = mc.real_forward(x)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        back = mc.real_inverse(got, n)
        assert np.array_equal(back.view(np.uint32), ctx.real_inverse(want, n).view(np.uint32))
        xc = (x[:, ::2] + 1j * x[:, 1::2]).astype(np.complex64)     # 2048-point complex rows, 82 MB: static path inside the same handle
        assert np.array_equal(mc.complex_forward(xc).view(np.uint32), ctx.complex_forward(xc).view(np.uint32))


def test_multi_two_callers_share_one_handle():
    import threading

    n, batch = 512, 600
    rng = np.random.default_rng(13)
    xs = [uniform_complex(rng, (batch, n)) for _ in range(3)]
    with ck.Context(n, ck.BOTH) as ctx, ck.MultiContext(n, ck.BOTH, [0, 0]) as mc:
        want = [ctx.complex_forward(x) for x in xs]
        got = [None] * 3

        def work(i):
            got[i] = mc.complex_forward(xs[i])

        th = [threading.Thread(target=work, args=(i,)) for i in range(3)]
        [t.start() for t in th]
        [t.join() for t in th]
        for g, w in zip(got, want):
            assert np.array_equal(g.view(np.uint32), w.view(np.uint32))
