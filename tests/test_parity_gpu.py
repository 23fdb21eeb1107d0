"""GPU parity tests: the CUDA path, called through the C ABI (host pointers and device pointers),
against the oracle restatement, the committed golden vectors of the reference, fp64 truth and
size-independent properties at BASELINE.json's full sizes.

Tolerance (BASELINE.json north_star): relative RMS error <= 1e-6 * log2(N), written in conftest.tolerance.
"""
import ctypes as C
import threading

import numpy as np
import pytest

import ckfft_b200 as ck
import oracle
from ckfft_b200 import _lib
from conftest import rel_rms, tolerance, uniform_complex

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

ALL_SIZES = [1 << k for k in range(0, 15)]            # complex 1 .. 16384 (single pass)
REAL_SIZES = [1 << k for k in range(0, 16)]           # real 1 .. 32768


def harness_rms(a, b):
    d = (np.asarray(a, np.complex64) - np.asarray(b, np.complex64)).view(np.float32)
    return float(np.sqrt(np.sum(d.astype(np.float64) ** 2) / d.size))


@pytest.fixture(scope="module")
def ctx_big():
    c = ck.Context(32768, ck.BOTH)
    yield c
    c.close()


@pytest.fixture(scope="module")
def orc_big():
    o = oracle.Restatement(32768, 3)
    yield o
    o.close()


# ---------------------------------------------------------------------------------------------
# oracle parity, every size, host path and device path, awkward batch counts
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", ALL_SIZES)
@pytest.mark.parametrize("inverse", [False, True])
def test_complex_vs_oracle(ctx_big, orc_big, n, inverse):
    rng = np.random.default_rng(n + inverse)
    for batch in (1, 3, 37):
        if n * batch > (1 << 19):
            continue
        x = uniform_complex(rng, (batch, n))
        want = orc_big.complex(x, inverse)
        f = ctx_big.complex_inverse if inverse else ctx_big.complex_forward
        got_host = f(x)
        got_dev = f(torch.from_numpy(x).cuda()).cpu().numpy()
        assert np.array_equal(got_host.view(np.uint32), got_dev.view(np.uint32)), "host and device paths differ"
        assert rel_rms(got_host, want) <= tolerance(n), (n, batch)
        assert rel_rms(got_host, oracle.fp64_c2c(x, inverse)) <= tolerance(n)


@pytest.mark.parametrize("n", REAL_SIZES)
def test_real_vs_oracle(ctx_big, orc_big, n):
    rng = np.random.default_rng(1000 + n)
    for batch in (1, 2, 3, 33):
        if n * batch > (1 << 19):
            continue
        x = rng.uniform(-1, 1, (batch, n)).astype(np.float32)
        want = orc_big.real_forward(x)
        got = ctx_big.real_forward(x)
        got_dev = ctx_big.real_forward(torch.from_numpy(x).cuda()).cpu().numpy()
        assert np.array_equal(got.view(np.uint32), got_dev.view(np.uint32))
        assert rel_rms(got, want) <= tolerance(n), (n, batch)
        assert rel_rms(got, oracle.fp64_real_forward(x)) <= tolerance(n)
        # inverse on the oracle's spectrum, and on a generic (non-Hermitian-edge) spectrum: the imaginary
        # parts of bins 0 and n/2 are NOT ignored by the reference (fft_real_default.cpp:79-107)
        for spec in (want, uniform_complex(rng, want.shape)):
            want_inv = orc_big.real_inverse(spec, n)
            got_inv = ctx_big.real_inverse(spec, n)
            got_inv_dev = ctx_big.real_inverse(torch.from_numpy(spec).cuda(), n).cpu().numpy()
            assert np.array_equal(got_inv.view(np.uint32), got_inv_dev.view(np.uint32))
            assert rel_rms(got_inv, want_inv) <= tolerance(n), (n, batch)


# ---------------------------------------------------------------------------------------------
# multi-pass lengths (four-step: two passes up to 2^20, three passes above)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("log2n", [15, 16, 17, 18, 19, 20, 21, 22, 24])
def test_large_complex_vs_oracle(log2n):
    n = 1 << log2n
    batch = 3 if log2n <= 18 else 1
    rng = np.random.default_rng(log2n)
    x = uniform_complex(rng, (batch, n))
    orc = oracle.Restatement(n, 3)
    with ck.Context(n, ck.BOTH) as ctx:
        for inverse in (False, True):
            want = orc.complex(x, inverse)
            f = ctx.complex_inverse if inverse else ctx.complex_forward
            got = f(torch.from_numpy(x).cuda()).cpu().numpy()
            assert rel_rms(got, want) <= tolerance(n), (n, inverse)
            assert rel_rms(got, oracle.fp64_c2c(x, inverse)) <= tolerance(n)
        if log2n <= 18:
            assert np.array_equal(ctx.complex_forward(x).view(np.uint32),
                                  ctx.complex_forward(torch.from_numpy(x).cuda()).cpu().numpy().view(np.uint32))
    orc.close()


@pytest.mark.parametrize("log2n,batch", [(15, 300), (16, 150), (17, 100), (18, 80), (19, 40), (20, 20)])
def test_pipelined_two_pass_matches_two_kernel_path(log2n, batch):
    """The L2-resident dataflow kernel (pipe_kernel.cuh) with enough problems to wrap its ring several times:
    bit-identical to the two-kernel four-step path, and within tolerance of fp64 on a few transforms."""
    import os

    n = 1 << log2n
    g = torch.Generator(device="cuda").manual_seed(log2n)
    x = torch.view_as_complex(torch.empty((batch, n, 2), dtype=torch.float32, device="cuda").uniform_(-1, 1, generator=g))
    with ck.Context(n, ck.BOTH) as ctx:
        for inverse in (False, True):
            f = ctx.complex_inverse if inverse else ctx.complex_forward
            os.environ["CKFFT_B200_PIPE"] = "1"
            try:
                for rep in range(3):                     # repeated launches: counters and ring are per launch
                    y = f(x)
                torch.cuda.synchronize()
                os.environ["CKFFT_B200_PIPE"] = "0"
                y0 = f(x)
                torch.cuda.synchronize()
            finally:
                os.environ.pop("CKFFT_B200_PIPE", None)
            assert torch.equal(torch.view_as_real(y), torch.view_as_real(y0)), (n, inverse)
            for b in (0, batch // 2, batch - 1):
                assert rel_rms(y[b].cpu().numpy(), oracle.fp64_c2c(x[b].cpu().numpy(), inverse)) <= tolerance(n)


@pytest.mark.parametrize("log2n,batch", [(16, 130), (17, 70), (18, 40), (19, 24), (20, 12), (21, 6)])
def test_pipelined_real_forward_fused_split(log2n, batch):
    """Real forward transforms above the single-pass limit: the split runs inside pass 2 of the dataflow kernel
    (mirror-paired tile columns).  Checked against the oracle, fp64, and the separate split pass."""
    import os

    n = 1 << log2n
    rng = np.random.default_rng(log2n)
    x = rng.uniform(-1, 1, (batch, n)).astype(np.float32)
    xd = torch.from_numpy(x).cuda()
    with ck.Context(n, ck.BOTH) as ctx:
        os.environ["CKFFT_B200_PIPE_REAL"] = "1"
        try:
            for rep in range(2):
                y = ctx.real_forward(xd)
            torch.cuda.synchronize()
            os.environ["CKFFT_B200_PIPE_REAL"] = "0"
            y0 = ctx.real_forward(xd)
            torch.cuda.synchronize()
        finally:
            os.environ.pop("CKFFT_B200_PIPE_REAL", None)
        assert y.shape == (batch, n // 2 + 1)
        assert rel_rms(y.cpu().numpy(), y0.cpu().numpy()) <= 3e-7        # same operations; the split factors are one rounding apart
        orc = oracle.Restatement(n, 3)
        for b in (0, batch // 2, batch - 1):
            got = y[b].cpu().numpy()
            assert rel_rms(got, orc.real_forward(x[b:b + 1])[0]) <= tolerance(n)
            assert rel_rms(got, oracle.fp64_real_forward(x[b:b + 1])[0]) <= tolerance(n)
            assert got[0].imag == 0.0 and got[-1].imag == 0.0          # DC and Nyquist bins are real
        orc.close()


@pytest.mark.parametrize("log2n,batch", [(16, 130), (17, 70), (18, 40), (19, 24), (20, 12), (21, 6)])
def test_pipelined_real_inverse_fused_twist(log2n, batch):
    """Real inverse transforms above the single-pass limit: the twist runs inside pass 1 of the dataflow kernel (the
    consumers read Y[k] and Y[M-k] with plain loads -- rows of n/2+1 values are only 8-byte aligned -- and twist in
    registers; the tile is fetched by TMA although rows of n/2+1 values are only 8-byte aligned).  Agrees with the separate
    twist pass + complex transform to rounding (its twist factors are one table value times a compile-time constant);
    checked against the oracle."""
    import os

    n = 1 << log2n
    rng = np.random.default_rng(300 + log2n)
    spec = uniform_complex(rng, (batch, n // 2 + 1))             # a generic spectrum: bins 0 and n/2 need not be real
    sd = torch.from_numpy(spec).cuda()
    with ck.Context(n, ck.BOTH) as ctx:
        os.environ["CKFFT_B200_PIPE_REAL"] = "1"
        try:
            for rep in range(2):                                 # repeated launches: counters and ring are per launch
                x = ctx.real_inverse(sd, n)
            torch.cuda.synchronize()
            launches = ck.kernel_launches()
            ctx.real_inverse(sd, n)
            assert ck.kernel_launches() - launches == 1          # one kernel: no twist pass
            os.environ["CKFFT_B200_PIPE_REAL"] = "0"
            x0 = ctx.real_inverse(sd, n)
            torch.cuda.synchronize()
        finally:
            os.environ.pop("CKFFT_B200_PIPE_REAL", None)
        assert x.shape == (batch, n)
        assert rel_rms(x.cpu().numpy(), x0.cpu().numpy()) <= 3e-7, n
        orc = oracle.Restatement(n, 3)
        for b in (0, batch // 2, batch - 1):
            got = x[b].cpu().numpy()
            assert rel_rms(got, orc.real_inverse(spec[b:b + 1], n)[0]) <= tolerance(n)
        orc.close()
        # odd offsets: a batch that starts at an odd row of a larger array has the frame alignments swapped
        y = ctx.real_inverse(sd[1:], n)
        assert torch.equal(y, x[1:])


@pytest.mark.parametrize("log2n", [16, 17, 20, 22])
def test_large_real_vs_oracle(log2n):
    n = 1 << log2n
    batch = 2 if log2n <= 17 else 1
    rng = np.random.default_rng(50 + log2n)
    x = rng.uniform(-1, 1, (batch, n)).astype(np.float32)
    orc = oracle.Restatement(n, 3)
    with ck.Context(n, ck.BOTH) as ctx:
        want = orc.real_forward(x)
        got = ctx.real_forward(torch.from_numpy(x).cuda()).cpu().numpy()
        assert rel_rms(got, want) <= tolerance(n)
        assert rel_rms(got, oracle.fp64_real_forward(x)) <= tolerance(n)
        back = ctx.real_inverse(torch.from_numpy(want).cuda(), n).cpu().numpy()
        assert rel_rms(back, orc.real_inverse(want, n)) <= tolerance(n)
        assert rel_rms(back, 2.0 * n * x.astype(np.float64)) <= tolerance(n)
    orc.close()


def test_very_large_round_trip_2_26():
    """N = 2^26 (three passes, 512 MiB per array): inverse(forward(x)) = N x, Parseval, and two analytic bins."""
    n = 1 << 26
    g = torch.Generator(device="cuda").manual_seed(99)
    x = torch.view_as_complex(torch.empty((1, n, 2), dtype=torch.float32, device="cuda").uniform_(-1, 1, generator=g))
    with ck.Context(n, ck.BOTH) as ctx:
        y = ctx.complex_forward(x)
        e_in = float((x.real.double() ** 2 + x.imag.double() ** 2).sum())
        e_out = float((y.real.double() ** 2 + y.imag.double() ** 2).sum()) / n
        assert abs(e_out - e_in) <= 1e-5 * e_in
        dc = complex(x.real.double().sum().item(), x.imag.double().sum().item())
        assert abs(complex(y[0, 0].item()) - dc) <= 1e-3 * np.sqrt(n)
        sign = torch.ones(n, dtype=torch.float64, device="cuda"); sign[1::2] = -1
        nyq = complex((x.real.double()[0] * sign).sum().item(), (x.imag.double()[0] * sign).sum().item())
        assert abs(complex(y[0, n // 2].item()) - nyq) <= 1e-3 * np.sqrt(n)
        z = ctx.complex_inverse(y)
        err = float(torch.linalg.vector_norm((z / n - x).abs().double()) / torch.linalg.vector_norm(x.abs().double()))
        assert err <= tolerance(n)


# ---------------------------------------------------------------------------------------------
# the reference's own fixture and acceptance criteria (golden vectors generated from the reference)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1 << k for k in range(0, 13)])
@pytest.mark.parametrize("nmax", ["n", 8192])
def test_golden_fixture(golden, n, nmax):
    """src/test/test.cpp:866-921: counts 4096 .. 1 of input.txt, contexts with maxCount == count and
    maxCount == 2*4096.  Checked three ways: north_star tolerance vs the reference's outputs, the
    harness's absolute RMS <= 0.001 vs KISS FFT, and the harness's real-vs-complex self-consistency."""
    nmax = n if nmax == "n" else nmax
    x = golden["input"][:n]
    with ck.Context(max(nmax, 1), ck.BOTH) as ctx:
        f, i = ctx.complex_forward(x), ctx.complex_inverse(x)
        assert rel_rms(f, golden[f"cfwd_{n}"]) <= tolerance(n)
        assert rel_rms(i, golden[f"cinv_{n}"]) <= tolerance(n)
        assert harness_rms(f, golden[f"kfwd_{n}"]) <= 1e-3
        assert harness_rms(i, golden[f"kinv_{n}"]) <= 1e-3
        xr = np.ascontiguousarray(x.real)
        rf = ctx.real_forward(xr)
        assert rel_rms(rf, golden[f"rfwd_{n}"]) <= tolerance(n)
        k = n // 2 + 1
        cref = ctx.complex_forward(xr.astype(np.complex64))
        assert harness_rms(rf * np.float32(0.5), cref[:k]) <= 1e-3
        ri = ctx.real_inverse(np.ascontiguousarray(cref[:k]), n)
        assert rel_rms(ri, golden[f"rinv_{n}"]) <= tolerance(n)
        d = ri[:k] - ctx.complex_inverse(cref)[:k].real
        assert float(np.sqrt(np.mean(d.astype(np.float64) ** 2))) <= 1e-3


def test_table_stride_contexts_are_bit_identical(golden):
    """A larger nMax only changes the table stride (fft.cpp:35); results must not change at all."""
    x = golden["input"][:1024]
    outs = []
    for nmax in (1024, 2048, 8192, 1 << 16):
        with ck.Context(nmax, ck.BOTH) as ctx:
            outs.append((ctx.complex_forward(x), ctx.real_forward(np.ascontiguousarray(x.real))))
    for a, b in outs[1:]:
        assert np.array_equal(a.view(np.uint32), outs[0][0].view(np.uint32))
        assert np.array_equal(b.view(np.uint32), outs[0][1].view(np.uint32))


def test_example_round_trip(golden):
    """src/example/main.cpp:44-84"""
    x = golden["example_in"]
    with ck.Context(1024, ck.BOTH) as ctx:
        f = ctx.real_forward(x)
        assert rel_rms(f, golden["example_fwd"]) <= tolerance(1024)
        rt = ctx.real_inverse(f, 1024)
        assert float(np.sum((rt / 2048.0 - x) ** 2)) < 1e-6


# ---------------------------------------------------------------------------------------------
# properties at BASELINE.json's full sizes (device resident)
# ---------------------------------------------------------------------------------------------
def test_full_size_c2c_1024_x_1M_round_trip_and_spot_parity(orc_big):
    """config 2: N=1024 x 2^20 transforms.  inverse(forward(x)) = N x on every transform; Parseval on
    every transform; 4096 leading + 4096 random transforms against the oracle."""
    n, batch = 1024, 1 << 20
    g = torch.Generator(device="cuda").manual_seed(1234)
    x = torch.empty((batch, n, 2), dtype=torch.float32, device="cuda")
    x.uniform_(-1, 1, generator=g)
    x = torch.view_as_complex(x)
    with ck.Context(n, ck.BOTH) as ctx:
        y = ctx.complex_forward(x)
        energy_in = (x.real.double() ** 2 + x.imag.double() ** 2).sum(dim=1)
        energy_out = (y.real.double() ** 2 + y.imag.double() ** 2).sum(dim=1)
        assert float(((energy_out / n - energy_in).abs() / energy_in).max()) < 1e-5
        idx = torch.cat([torch.arange(4096, device="cuda"),
                         torch.randint(0, batch, (4096,), device="cuda", generator=g)])
        xs, ys = x[idx].cpu().numpy(), y[idx].cpu().numpy()
        assert rel_rms(ys, orc_big.complex(xs, False)) <= tolerance(n)
        assert rel_rms(ys, oracle.fp64_c2c(xs, False)) <= tolerance(n)
        z = ctx.complex_inverse(y)
        del y
        err = torch.linalg.vector_norm((z / n - x).view(batch, -1).abs().double(), dim=1)
        ref = torch.linalg.vector_norm(x.view(batch, -1).abs().double(), dim=1)
        assert float((err / ref).max()) <= tolerance(n)
    torch.cuda.synchronize()


def test_full_size_r2c_4096_x_256K_round_trip_and_spot_parity(orc_big):
    """config 3: N=4096 real x 2^18 frames, R2C then C2R = 2N x; forward scale 2; spot parity."""
    n, batch = 4096, 1 << 18
    g = torch.Generator(device="cuda").manual_seed(1235)
    x = torch.empty((batch, n), dtype=torch.float32, device="cuda")
    x.uniform_(-1, 1, generator=g)
    with ck.Context(n, ck.BOTH) as ctx:
        y = ctx.real_forward(x)
        assert y.shape == (batch, n // 2 + 1)
        # DC bin = 2 * sum(x), imaginary parts of DC and Nyquist are exactly zero-sum rounding
        dc = 2.0 * x.double().sum(dim=1)
        assert float((y[:, 0].real.double() - dc).abs().max()) < 1e-2
        idx = torch.randint(0, batch, (2048,), device="cuda", generator=g)
        xs, ys = x[idx].cpu().numpy(), y[idx].cpu().numpy()
        assert rel_rms(ys, orc_big.real_forward(xs)) <= tolerance(n)
        assert rel_rms(ys, oracle.fp64_real_forward(xs)) <= tolerance(n)
        z = ctx.real_inverse(y, n)
        err = torch.linalg.vector_norm((z / (2.0 * n) - x).double(), dim=1)
        ref = torch.linalg.vector_norm(x.double(), dim=1)
        assert float((err / ref).max()) <= tolerance(n)
    torch.cuda.synchronize()


@pytest.mark.parametrize("n", [16, 256, 4096, 16384])
def test_linearity_and_impulse(ctx_big, n):
    rng = np.random.default_rng(n)
    a, b = uniform_complex(rng, (4, n)), uniform_complex(rng, (4, n))
    fa, fb = ctx_big.complex_forward(a), ctx_big.complex_forward(b)
    fab = ctx_big.complex_forward((a + np.complex64(2) * b).astype(np.complex64))
    assert rel_rms(fab, fa + 2 * fb) <= 2 * tolerance(n)
    imp = np.zeros((1, n), np.complex64); imp[0, 1] = 1
    k = np.arange(n)
    assert np.allclose(ctx_big.complex_forward(imp)[0], np.exp(-2j * np.pi * k / n), atol=1e-6)
    assert np.allclose(ctx_big.complex_inverse(imp)[0], np.exp(+2j * np.pi * k / n), atol=1e-6)


# ---------------------------------------------------------------------------------------------
# error behaviour of the boundary (reference: src/ckfft/ckfft.cpp:36-114)
# ---------------------------------------------------------------------------------------------
def test_error_returns():
    lib = _lib.load()
    fwd = ck.Context(64, ck.FORWARD)
    inv = ck.Context(64, ck.INVERSE)
    a = np.zeros(64, np.complex64); b = np.zeros(64, np.complex64)
    pa, pb = a.ctypes.data, b.ctypes.data
    assert lib.CkFftComplexForward(fwd.handle, 64, pa, pb) == 1
    assert lib.CkFftComplexInverse(fwd.handle, 64, pa, pb) == 0        # wrong direction
    assert lib.CkFftComplexForward(inv.handle, 64, pa, pb) == 0
    assert lib.CkFftRealForward(inv.handle, 64, pa, pb) == 0
    assert lib.CkFftComplexInverse(inv.handle, 64, pa, pb) == 1
    assert lib.CkFftComplexForward(fwd.handle, 128, pa, pb) == 0        # n > nMax
    assert lib.CkFftComplexForward(fwd.handle, 48, pa, pb) == 0         # not a power of two
    assert lib.CkFftComplexForward(fwd.handle, 0, pa, pb) == 0
    assert lib.CkFftComplexForward(fwd.handle, -64, pa, pb) == 0
    assert lib.CkFftComplexForward(fwd.handle, -(2 ** 31), pa, pb) == 0  # the reference's INT_MIN hole
    assert lib.CkFftComplexForward(fwd.handle, 64, pa, pa) == 0         # in == out
    assert lib.CkFftComplexForward(fwd.handle, 64, None, pb) == 0
    assert lib.CkFftComplexForward(fwd.handle, 64, pa, None) == 0
    assert lib.CkFftRealInverse(inv.handle, 64, pa, pb, None) == 0      # tmpBuf NULL (ckfft.cpp:57-60)
    assert lib.CkFftRealInverse(inv.handle, 64, pa, pb, pa) == 1
    assert lib.CkFftRealInverseBatch(inv.handle, 64, pa, pb, None, 1) == 1   # batched variant ignores tmpBuf
    assert lib.CkFftComplexForwardBatch(fwd.handle, 16, pa, pb, 0) == 1      # empty batch is a no-op
    # device pointer + host pointer mixed, and a misaligned device pointer
    d = torch.zeros(130, dtype=torch.complex64, device="cuda")
    assert lib.CkFftComplexForward(fwd.handle, 64, d.data_ptr(), pb) == 0
    assert lib.CkFftComplexForwardBatchAsync(fwd.handle, 64, d.data_ptr() + 4, d.data_ptr() + 8 * 64, 1, 0, 0, None) == 0
    assert lib.CkFftComplexForwardBatchAsync(fwd.handle, 64, d.data_ptr(), d.data_ptr() + 8 * 64, 1, 32, 0, None) == 0  # stride < n
    assert ck.last_error() != ""
    fwd.close(); inv.close()


def test_user_buffer_context(golden):
    """CkFftInit with caller storage (inc/ckfft/ckfft.h:51-55): the context lives in the caller's buffer and
    CkFftShutdown does not free it (context.cpp:111,116-122)."""
    lib = _lib.load()
    sz = C.c_size_t(0)
    assert not lib.CkFftInit(256, ck.BOTH, None, C.byref(sz))
    buf = C.create_string_buffer(sz.value)
    h = lib.CkFftInit(256, ck.BOTH, buf, C.byref(sz))
    assert h == C.addressof(buf)
    x = np.ascontiguousarray(golden["input"][:256]); out = np.empty_like(x)
    assert lib.CkFftComplexForward(h, 256, x.ctypes.data, out.ctypes.data) == 1
    assert rel_rms(out, golden["cfwd_256"]) <= tolerance(256)
    lib.CkFftShutdown(h)
    buf.raw   # still ours


def test_strided_async(ctx_big, orc_big):
    """stream-ordered variant with padded strides (aligned spectrum stride for R2C)."""
    lib = _lib.load()
    n, batch = 1024, 9
    rng = np.random.default_rng(5)
    x = rng.uniform(-1, 1, (batch, n + 8)).astype(np.float32)
    xd = torch.from_numpy(x).cuda()
    yd = torch.zeros((batch, n // 2 + 2), dtype=torch.complex64, device="cuda")   # padded: 514 bins per row
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        ok = lib.CkFftRealForwardBatchAsync(ctx_big.handle, n, xd.data_ptr(), yd.data_ptr(), batch, n + 8, n // 2 + 2, s.cuda_stream)
    assert ok == 1
    s.synchronize()
    want = orc_big.real_forward(np.ascontiguousarray(x[:, :n]))
    got = yd.cpu().numpy()
    assert rel_rms(got[:, : n // 2 + 1], want) <= tolerance(n)
    assert np.all(got[:, n // 2 + 1] == 0)     # padding untouched
    xo = torch.zeros((batch, n + 2), dtype=torch.float32, device="cuda")
    ok = lib.CkFftRealInverseBatchAsync(ctx_big.handle, n, yd.data_ptr(), xo.data_ptr(), batch, n // 2 + 2, n + 2, None)
    assert ok == 1
    torch.cuda.synchronize()
    assert rel_rms(xo.cpu().numpy()[:, :n], 2.0 * n * x[:, :n]) <= tolerance(n)


def test_context_shared_between_threads(orc_big):
    """'contexts can be used simultaneously on different threads' (inc/ckfft/ckfft.h:39-41)"""
    ctx = ck.Context(2048, ck.BOTH)
    rng = np.random.default_rng(11)
    xs = [uniform_complex(rng, (17, 2048)) for _ in range(4)]
    want = [orc_big.complex(x) for x in xs]
    errs = [None] * 4

    def work(i):
        for _ in range(5):
            errs[i] = rel_rms(ctx.complex_forward(xs[i]), want[i])

    ts = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert all(e is not None and e <= tolerance(2048) for e in errs), errs
    ctx.close()


def test_large_host_batch_is_chunked(orc_big):
    """host path with more data than one staging chunk (3 chunks in flight, api.cu run_host)"""
    n, batch = 4096, 3000          # 98 MB in, > 32 MiB chunk
    rng = np.random.default_rng(2)
    x = uniform_complex(rng, (batch, n))
    with ck.Context(n, ck.FORWARD) as ctx:
        before = ck.kernel_launches()
        y = ctx.complex_forward(x)
        assert ck.kernel_launches() - before >= 3
    pick = [0, 1, 1023, 1024, 1025, 2047, 2048, 2999]
    assert rel_rms(y[pick], orc_big.complex(x[pick])) <= tolerance(n)
    assert rel_rms(y[::97], oracle.fp64_c2c(x[::97])) <= tolerance(n)


# ---------------------------------------------------------------------------------------------
# distributed six-step: the local kernels (pack / tiled transpose / twiddle) on one GPU, world = 1
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("log2n", [16, 21, 24])
def test_six_step_single_rank(log2n):
    from ckfft_b200.distributed import DistributedFFT

    n = 1 << log2n
    rng = np.random.default_rng(log2n)
    x = uniform_complex(rng, (n,))
    want = oracle.fp64_c2c(x)
    d = DistributedFFT(n)
    xd = torch.from_numpy(x).cuda()
    y = d.forward(xd)
    assert rel_rms(y.cpu().numpy(), want) <= tolerance(n)
    z = d.inverse(y)
    assert rel_rms(z.cpu().numpy() / n, x) <= tolerance(n)
    d.close()


def test_six_step_glue_kernels():
    lib = _lib.load()
    rows, parts, w = 24, 4, 40          # ragged on purpose (not multiples of the 32x32 tile)
    rng = np.random.default_rng(1)
    a = uniform_complex(rng, (rows, parts, w))
    ad = torch.from_numpy(a).cuda()
    packed = torch.empty_like(ad)
    assert lib.CkFftB200PackColumnsAsync(ad.data_ptr(), packed.data_ptr(), rows, parts, w, None) == 1
    assert np.array_equal(packed.cpu().numpy().reshape(parts, rows, w), a.transpose(1, 0, 2))
    out = torch.empty_like(ad)
    assert lib.CkFftB200UnpackTransposeAsync(packed.data_ptr(), out.data_ptr(), parts, rows, w, None) == 1
    assert np.array_equal(out.cpu().numpy().reshape(w, parts * rows), a.transpose(1, 0, 2).transpose(2, 0, 1).reshape(w, parts * rows))
    n = 1 << 20
    with ck.Context(n, ck.BOTH) as ctx:
        t = torch.ones((64, 1024), dtype=torch.complex64, device="cuda")
        assert lib.CkFftB200TwiddleRowsAsync(ctx.handle, n, t.data_ptr(), 64, 1024, 960, 0, None) == 1
        i = (960 + np.arange(64))[:, None].astype(np.float64); k = np.arange(1024)[None, :].astype(np.float64)
        assert np.allclose(t.cpu().numpy(), np.exp(-2j * np.pi * i * k / n), atol=3e-7)
        assert lib.CkFftB200TwiddleRowsAsync(ctx.handle, n, t.data_ptr(), 64, 1024, 1 << 19, 0, None) == 0   # exponent overflow


# ---------------------------------------------------------------------------------------------
# fused distributed transform on one GPU (world = 1: the routed passes store into the GPU's own arrays, so the
# exchange kernel, the flag barrier and every routed tile kernel run exactly as they do across NVLink)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("log2n", [14, 15, 17, 20, 21, 23, 24])
@pytest.mark.parametrize("pull", [False, True])
def test_fused_distributed_single_rank(log2n, pull):
    from ckfft_b200.distributed import FusedDistributedFFT

    n = 1 << log2n
    rng = np.random.default_rng(log2n)
    x = uniform_complex(rng, (n,))
    want = oracle.fp64_c2c(x)
    d = FusedDistributedFFT(n, pull=pull)
    xd = torch.from_numpy(x).cuda()
    if pull and log2n % 2:
        d.input.copy_(xd)            # the zero-copy way: the caller fills the plan's own input array
        xd = d.input
    y = d.forward(xd).clone()
    assert rel_rms(y.cpu().numpy(), want) <= tolerance(n)
    z = d.inverse(y)
    assert rel_rms(z.cpu().numpy() / n, x) <= tolerance(n)
    d.check()
    d.close()


@pytest.mark.parametrize("prefer,pull", [(3, False), (4, False), (4, True)])
def test_fused_distributed_2_28_analytic(prefer, pull):
    """Four-pass and three-pass layouts at 2^28 points: closed-form spectrum of exponentials + an impulse."""
    from ckfft_b200.distributed import FusedDistributedFFT

    n = 1 << 28
    d = FusedDistributedFFT(n, prefer_passes=prefer, pull=pull)
    assert d.layout.passes == prefer
    idx = torch.arange(n, device="cuda", dtype=torch.float64)
    freqs, amps, n0 = [3, n // 3 + 1, n - 7], [1.0, 0.5, 0.25], 5
    x = torch.zeros(n, dtype=torch.complex64, device="cuda")
    for f, a in zip(freqs, amps):
        ph = 2.0 * np.pi * ((idx * f) % n) / n
        x += (a * torch.complex(torch.cos(ph), torch.sin(ph))).to(torch.complex64)
    x[n0] += 1.0
    y = d.forward(x)
    ph = -2.0 * np.pi * ((idx * n0) % n) / n
    want = torch.complex(torch.cos(ph), torch.sin(ph))
    del ph, idx
    for f, a in zip(freqs, amps):
        want[f] += a * n
    err = float(torch.linalg.vector_norm(y.to(torch.complex128) - want) / torch.linalg.vector_norm(want))
    del want
    assert err <= tolerance(n)
    z = d.inverse(y.clone())
    assert float(torch.linalg.vector_norm(z / n - x) / torch.linalg.vector_norm(x)) <= tolerance(n)
    d.check()
    d.close()
    torch.cuda.empty_cache()


def test_fused_distributed_2_30_analytic():
    """BASELINE config 5 at its named size on ONE GPU (world = 1: the routed kernels store into the GPU's own arrays, the
    same kernels, barriers and layouts as across NVLink).  The reference cannot be the oracle here -- it cannot even
    create a context for nMax >= 2^28 (src/ckfft/context.cpp:37-45) -- so the transform is checked against the
    closed-form spectrum of exponentials + an impulse on every one of the 2^30 bins, by Parseval and by the round trip.
    Needs ~50 GB of device memory (4 x 8 GiB plan arrays + the check's temporaries)."""
    import os
    import sys

    from ckfft_b200.distributed import FusedDistributedFFT

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from analytic import analytic_error, analytic_signal

    free, _ = torch.cuda.mem_get_info()
    if free < (56 << 30):
        pytest.skip("needs 56 GiB of free device memory")
    n = 1 << 30
    dev = torch.device("cuda", torch.cuda.current_device())
    d = FusedDistributedFFT(n)                   # default layout for 2^30: four passes, pull
    assert d.layout.passes == 4 and d.input is not None
    freqs, amps, n0 = [3, n // 3 + 1, n - 7], [1.0, 0.5, 0.25], 5
    d.input.copy_(analytic_signal(n, 0, n, dev, freqs, amps, n0))
    torch.cuda.empty_cache()
    x = d.input
    ex = float(torch.linalg.vector_norm(x).double() ** 2)
    y = d.forward(x)
    d.check()
    num, den = analytic_error(y, n, 0, dev, freqs, amps, n0)
    assert float(torch.sqrt(num / den)) <= tolerance(n)
    ey = float(torch.linalg.vector_norm(y).double() ** 2) / n
    assert abs(ey - ex) <= 1e-5 * ex                                  # Parseval
    spot = y[[0, 3, n // 3 + 1, n - 7, n // 2]].cpu().numpy()        # the three spectral lines stand n * a_i above the impulse's unit circle
    for got, want in zip(np.abs(spot[1:4]).astype(np.float64), (n, n / 2, n / 4)):
        assert abs(got - want) <= 1e-6 * want
    # round trip: inverse(forward(x)) = n * x.  The inverse reads `out` through a copy into the pull array.
    keep = x[: 1 << 20].clone()
    z = d.inverse(y.clone())
    d.check()
    err = float(torch.linalg.vector_norm(z[: 1 << 20] / n - keep) / torch.linalg.vector_norm(keep))
    assert err <= tolerance(n)
    d.close()
    torch.cuda.empty_cache()


def test_fused_distributed_flag_block_serves_a_second_plan():
    """A C caller may destroy a plan and create another one on the same work / mid / out / flag buffers:
    the new plan must continue from the epoch the flag block holds, otherwise its barriers pass at once on stale values."""
    from ckfft_b200.distributed import FusedDistributedFFT

    lib = _lib.load()
    n = 1 << 22
    rng = np.random.default_rng(22)
    x = uniform_complex(rng, (n,))
    want = oracle.fp64_c2c(x)
    xd = torch.from_numpy(x).cuda()
    d = FusedDistributedFFT(n, pull=False)
    for _ in range(3):
        d.forward(xd)
    d.check()
    arrs = [(C.c_void_p * 1)(d._own[b]) for b in range(4)]
    for prefer in (3, 0):
        plan2 = lib.CkFftB200DistPlanCreate(d.ctx.handle, n, 0, 1, prefer, arrs[0], arrs[1], arrs[2], arrs[3], None)
        assert plan2, ck.last_error()
        assert lib.CkFftB200DistExecAsync(plan2, xd.data_ptr(), 0, torch.cuda.current_stream().cuda_stream) == 1
        assert lib.CkFftB200DistPlanStatus(plan2) == 1
        assert rel_rms(d.out.cpu().numpy(), want) <= tolerance(n)
        lib.CkFftB200DistPlanDestroy(plan2)
    d.close()


def test_single_chunk_host_call_allocates_one_staging_pair():
    """a host call that fits one chunk needs one pair of staging buffers, not three, and staging larger
    than 256 MiB per slot is handed back when the call ends."""
    import threading

    n = 1 << 24                                   # one transform = 128 MiB in + 128 MiB out = one chunk
    rng = np.random.default_rng(3)
    x = uniform_complex(rng, (n,))
    result = {}

    def work():                                   # a fresh thread: fresh per-thread staging
        with ck.Context(n, ck.FORWARD) as ctx:
            torch.cuda.synchronize()
            free0, _ = torch.cuda.mem_get_info()
            y = ctx.complex_forward(x)
            free1, _ = torch.cuda.mem_get_info()
            result["grew"] = free0 - free1
            result["y"] = y

    t = threading.Thread(target=work)
    t.start()
    t.join()
    assert rel_rms(result["y"], oracle.fp64_c2c(x)) <= tolerance(n)
    assert result["grew"] <= (320 << 20), result["grew"]        # one pair (256 MiB + slack), not three (768 MiB)


def test_out_arguments_are_validated(ctx_big):
    """a caller-supplied `out` travels to the C ABI as a bare pointer -- wrong dtype / shape / layout /
    side must raise instead of writing out of bounds."""
    x = uniform_complex(np.random.default_rng(0), (4, 256))
    xd = torch.from_numpy(x).cuda()
    good = np.empty_like(x)
    assert ctx_big.complex_forward(x, good) is good
    for bad in (np.empty((4, 255), np.complex64), np.empty((4, 256), np.complex128), np.empty((256, 4), np.complex64).T,
                np.empty((3, 256), np.complex64), torch.empty((4, 256), dtype=torch.complex64, device="cuda")):
        with pytest.raises(ck.CkFftError):
            ctx_big.complex_forward(x, bad)
    for bad in (torch.empty((4, 256), dtype=torch.complex64), torch.empty((4, 512), dtype=torch.complex64, device="cuda")[:, ::2],
                torch.empty((4, 256), dtype=torch.float32, device="cuda"), good):
        with pytest.raises(ck.CkFftError):
            ctx_big.complex_forward(xd, bad)
    xr = np.zeros((2, 64), np.float32)
    with pytest.raises(ck.CkFftError):
        ctx_big.real_forward(xr, np.empty((2, 32), np.complex64))               # n/2 + 1 = 33 bins
    with pytest.raises(ck.CkFftError):
        ctx_big.real_inverse(np.zeros((2, 33), np.complex64), 64, np.empty((2, 63), np.float32))
    with pytest.raises(ck.CkFftError):
        ctx_big.real_forward_power(torch.zeros((2, 64), device="cuda"), None, torch.empty((2, 33), dtype=torch.float64, device="cuda"))
    re = torch.zeros((2, 64), device="cuda")
    with pytest.raises(ck.CkFftError):
        ctx_big.complex_planar(re, re.clone(), False, (torch.empty((2, 64), device="cuda"), torch.empty((2, 32), device="cuda")))


def test_fused_distributed_rejects_bad_arguments():
    from ckfft_b200.distributed import FusedDistributedFFT

    lib = _lib.load()
    d = FusedDistributedFFT(1 << 14)
    x = torch.zeros(1 << 14, dtype=torch.complex64, device="cuda")
    assert lib.CkFftB200DistExecAsync(d._plan, None, 0, None) == 0
    assert lib.CkFftB200DistExecAsync(d._plan, d.out.data_ptr(), 0, None) == 0      # input aliases a plan buffer
    assert lib.CkFftB200DistExecAsync(None, x.data_ptr(), 0, None) == 0
    with ck.Context(1 << 12, ck.BOTH) as small:                                      # context too small for n
        arr = (C.c_void_p * 1)(d._own[0])
        assert not lib.CkFftB200DistPlanCreate(small.handle, 1 << 14, 0, 1, 0, arr, arr, arr, arr, None)
    d.close()


# ---------------------------------------------------------------------------------------------
# audio front end: window + real forward + power spectrum fused in one kernel
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [32, 64, 256, 1024, 2048, 4096, 8192, 32768])
def test_power_spectrum_vs_oracle(ctx_big, orc_big, n):
    rng = np.random.default_rng(n + 7)
    for batch in (1, 5, 34):
        if n * batch > (1 << 19):
            continue
        x = rng.uniform(-1, 1, (batch, n)).astype(np.float32)
        w = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n) / n)).astype(np.float32)     # Hann
        for window in (None, w):
            xw = x if window is None else (x * window).astype(np.float32)
            y = orc_big.real_forward(xw)
            want = (y.real.astype(np.float64) ** 2 + y.imag.astype(np.float64) ** 2)
            got = ctx_big.real_forward_power(torch.from_numpy(x).cuda(),
                                             None if window is None else torch.from_numpy(window).cuda()).cpu().numpy()
            assert got.shape == (batch, n // 2 + 1)
            # power = |Y|^2: relative error doubles; measured against the spectrum's total power
            assert np.linalg.norm(got - want) / np.linalg.norm(want) <= 4 * tolerance(n), (n, batch)


# ---------------------------------------------------------------------------------------------
# fused distributed transform, world = 2: two PROCESSES (CUDA IPC handles, flag barriers over mapped peer memory,
# every routed store crossing a process boundary).  On a one-GPU box both ranks share the device -- the driver
# time-slices their kernels, so a barrier costs a few time slices instead of microseconds, which is fine for a test.
# ---------------------------------------------------------------------------------------------
def _fused_two_rank_worker(rank, world, port, tmpdir, ndev):
    import os
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist

    from ckfft_b200.distributed import FusedDistributedFFT

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank % ndev)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = True
    for log2n, pull in ((14, False), (21, False), (14, True), (21, True)):
        n = 1 << log2n
        rng = np.random.default_rng(100 + log2n)
        x = uniform_complex(rng, (n,))                                   # same on every rank
        per = n // world
        d = FusedDistributedFFT(n, pull=pull)
        xd = torch.from_numpy(x[rank * per:(rank + 1) * per]).cuda()
        y = d.forward(xd).clone()
        z = d.inverse(y).clone()
        d.check()
        parts, back = [None] * world, [None] * world
        dist.all_gather_object(parts, y.cpu().numpy())
        dist.all_gather_object(back, z.cpu().numpy())
        if rank == 0:
            want = np.fft.fft(x.astype(np.complex128))
            ok = ok and rel_rms(np.concatenate(parts), want) <= tolerance(n)
            ok = ok and rel_rms(np.concatenate(back) / n, x) <= tolerance(n)
        dist.barrier()
        d.close()
    if rank == 0:
        open(os.path.join(tmpdir, "ok"), "w").write("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


def test_fused_distributed_two_processes(tmp_path):
    import socket

    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_fused_two_rank_worker, args=(2, port, str(tmp_path), torch.cuda.device_count()), nprocs=2, join=True)
    assert (tmp_path / "ok").read_text() == "1"
