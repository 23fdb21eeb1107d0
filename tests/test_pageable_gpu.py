"""Host-pointer calls on PAGEABLE (malloc'ed) arrays -- what a drop-in caller of the reference hands over
(src/test/test.cpp:254-301 passes plain heap arrays).  Calls that move >= 64 MiB are staged by the library through pinned
slots with two teams of host threads (api.cu, run_host_pageable); results must be bit-identical to the device-resident path
and to the driver-staged path (CKFFT_B200_PAGEABLE_PIPE=0), whatever the chunking."""
import numpy as np
import pytest

import ckfft_b200 as ck
from conftest import uniform_complex

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("threads", ["1", "3"])
def test_pageable_complex_batch_is_staged_and_bit_identical(threads, monkeypatch):
    monkeypatch.setenv("CKFFT_B200_HOST_THREADS", threads)      # (read once per process: the first value wins; both are legal)
    n, batch = 1024, 9001                                       # 73.7 MB in + 73.7 MB out, ragged last 16 MiB chunk
    rng = np.random.default_rng(5)
    x = uniform_complex(rng, (batch, n))
    with ck.Context(n, ck.BOTH) as ctx:
        want = ctx.complex_forward(torch.from_numpy(x).cuda()).cpu().numpy()
        got = ctx.complex_forward(x)                            # numpy arrays: pageable host memory
        assert np.array_equal(bits(got), bits(want))
        monkeypatch.setenv("CKFFT_B200_PAGEABLE_PIPE", "0")
        ref = ctx.complex_forward(x)
        assert np.array_equal(bits(ref), bits(want))
        monkeypatch.delenv("CKFFT_B200_PAGEABLE_PIPE")
        back = ctx.complex_inverse(got)
        assert np.array_equal(bits(back), bits(ctx.complex_inverse(torch.from_numpy(want).cuda()).cpu().numpy()))


def test_pageable_real_frames_and_small_chunks(monkeypatch):
    monkeypatch.setenv("CKFFT_B200_PAGEABLE_CHUNK_MB", "16")
    n, batch = 4096, 4500                                       # 73.7 MB in + 73.8 MB out; output rows of 2049 complex values
    rng = np.random.default_rng(6)
    x = rng.uniform(-1, 1, (batch, n)).astype(np.float32)
    with ck.Context(n, ck.BOTH) as ctx:
        want = ctx.real_forward(torch.from_numpy(x).cuda()).cpu().numpy()
        got = ctx.real_forward(x)
        assert np.array_equal(bits(got), bits(want))
        back = ctx.real_inverse(got, n)
        want_back = ctx.real_inverse(torch.from_numpy(want).cuda(), n).cpu().numpy()
        assert np.array_equal(bits(back), bits(want_back))


def test_pageable_long_transforms():
    """one transform per chunk (n = 2^22: 32 MiB rows), multi-pass kernels behind the staged pipeline"""
    n, batch = 1 << 22, 3
    rng = np.random.default_rng(7)
    x = uniform_complex(rng, (batch, n))
    with ck.Context(n, ck.BOTH) as ctx:
        want = ctx.complex_forward(torch.from_numpy(x).cuda()).cpu().numpy()
        got = ctx.complex_forward(x)
        assert np.array_equal(bits(got), bits(want))


@pytest.mark.parametrize("static", ["0", "1"])
def test_pageable_arrays_through_the_multi_device_scheduler(static, monkeypatch):
    """>= 256 MiB on pageable arrays: the workers of one BatchMulti call draw chunks of the shared batch (dynamic schedule) or take
    contiguous shards (static) and stage them through their own pinned slots, dividing the host's copy threads among them"""
    monkeypatch.setenv("CKFFT_B200_MULTI_STATIC", static)
    n, batch = 4096, 9001                                       # 147 MB in + 148 MB out
    rng = np.random.default_rng(8)
    x = rng.uniform(-1, 1, (batch, n)).astype(np.float32)
    devices = list(range(torch.cuda.device_count())) if torch.cuda.device_count() >= 2 else [0, 0, 0]
    with ck.Context(n, ck.BOTH) as ctx, ck.MultiContext(n, ck.BOTH, devices) as mc:
        want = ctx.real_forward(torch.from_numpy(x).cuda()).cpu().numpy()
        got = mc.real_forward(x)
        assert np.array_equal(bits(got), bits(want))
        back = mc.real_inverse(got, n)
        assert np.array_equal(bits(back), bits(ctx.real_inverse(torch.from_numpy(want).cuda(), n).cpu().numpy()))
