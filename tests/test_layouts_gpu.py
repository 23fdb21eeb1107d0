"""GPU tests of the data layouts either side of the hot path: split-complex ("planar") arrays and
in-place stream-ordered calls, through the C ABI, against the oracle restatement and -- bit for bit -- against the
interleaved / out-of-place calls that run the same arithmetic.

Tolerance (BASELINE.json north_star): relative RMS error <= 1e-6 * log2(N) (conftest.tolerance).
"""
import numpy as np
import pytest

import ckfft_b200 as ck
import oracle
from ckfft_b200 import _lib
from conftest import rel_rms, tolerance, uniform_complex

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


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


def bits(t):
    return t.cpu().numpy().view(np.uint32)


# ---------------------------------------------------------------------------------------------
# split-complex arrays
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1 << k for k in range(0, 15)])
@pytest.mark.parametrize("inverse", [False, True])
def test_planar_vs_oracle_and_interleaved(ctx_big, orc_big, n, inverse):
    rng = np.random.default_rng(100 + n + inverse)
    for batch in (1, 5, 37):
        if n * batch > (1 << 19):
            continue
        x = uniform_complex(rng, (batch, n))
        re = torch.from_numpy(np.ascontiguousarray(x.real)).cuda()
        im = torch.from_numpy(np.ascontiguousarray(x.imag)).cuda()
        ore, oim = ctx_big.complex_planar(re, im, inverse)
        torch.cuda.synchronize()
        got = (ore.cpu().numpy() + 1j * oim.cpu().numpy()).astype(np.complex64)
        assert rel_rms(got, orc_big.complex(x, inverse)) <= tolerance(n), (n, batch)
        # Below 2048 points the planar kernels are the plain-load instantiations of the plan table; an interleaved input
        # that is only 8-byte aligned takes the same instantiation (bulk prefetch needs 16-byte rows).  From 2048 points
        # on, 16-byte aligned planes are bulk-prefetched like 16-byte aligned interleaved rows (same plan, same
        # arithmetic).  Either way the two layouts must agree bit for bit.
        f = ctx_big.complex_inverse if inverse else ctx_big.complex_forward
        off = 1 if n < 2048 else 2
        pad = torch.empty(batch * n + 2, dtype=torch.complex64, device="cuda")
        xi = pad[off:off + batch * n].view(batch, n)
        xi.copy_(torch.from_numpy(x))
        assert xi.data_ptr() % 16 == (8 if n < 2048 else 0) and re.data_ptr() % 16 == 0 and im.data_ptr() % 16 == 0
        inter = f(xi).cpu().numpy()
        assert np.array_equal(got.view(np.uint32), inter.view(np.uint32)), "planar and interleaved kernels differ"
        assert np.array_equal(re.cpu().numpy(), x.real) and np.array_equal(im.cpu().numpy(), x.imag), "input modified"


@pytest.mark.parametrize("n", [2048, 4096, 8192, 16384])
def test_planar_unaligned_rows_take_the_plain_kernels(ctx_big, orc_big, n):
    # planes whose rows are not 16-byte aligned (odd pitch / offset pointer) cannot be bulk-copied: plain-load kernels
    lib = _lib.load()
    rng = np.random.default_rng(n)
    batch, pitch = 3, n + 1
    x = uniform_complex(rng, (batch, n))
    re = torch.zeros(batch * pitch + 1, dtype=torch.float32, device="cuda")[1:].view(batch, pitch)
    im = torch.zeros(batch * pitch + 1, dtype=torch.float32, device="cuda")[1:].view(batch, pitch)
    re[:, :n] = torch.from_numpy(np.ascontiguousarray(x.real)).cuda()
    im[:, :n] = torch.from_numpy(np.ascontiguousarray(x.imag)).cuda()
    ore = torch.empty((batch, n), dtype=torch.float32, device="cuda")
    oim = torch.empty((batch, n), dtype=torch.float32, device="cuda")
    assert lib.CkFftB200ComplexForwardPlanarBatchAsync(ctx_big.handle, n, re.data_ptr(), im.data_ptr(), ore.data_ptr(), oim.data_ptr(),
                                                       batch, pitch, n, None) == 1, ck.last_error()
    torch.cuda.synchronize()
    got = (ore.cpu().numpy() + 1j * oim.cpu().numpy()).astype(np.complex64)
    assert rel_rms(got, orc_big.complex(x, False)) <= tolerance(n)


def test_planar_in_place_strided_and_errors(ctx_big, orc_big):
    lib = _lib.load()
    n, batch, pitch = 512, 11, 520
    rng = np.random.default_rng(8)
    x = uniform_complex(rng, (batch, n))
    re = torch.zeros((batch, pitch), dtype=torch.float32, device="cuda"); re[:, :n] = torch.from_numpy(np.ascontiguousarray(x.real)).cuda()
    im = torch.zeros((batch, pitch), dtype=torch.float32, device="cuda"); im[:, :n] = torch.from_numpy(np.ascontiguousarray(x.imag)).cuda()
    h = ctx_big.handle
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        ok = lib.CkFftB200ComplexForwardPlanarBatchAsync(h, n, re.data_ptr(), im.data_ptr(), re.data_ptr(), im.data_ptr(), batch,
                                                         pitch, pitch, s.cuda_stream)
    assert ok == 1, ck.last_error()
    s.synchronize()
    got = (re.cpu().numpy()[:, :n] + 1j * im.cpu().numpy()[:, :n]).astype(np.complex64)
    assert rel_rms(got, orc_big.complex(x, False)) <= tolerance(n)
    assert np.all(re.cpu().numpy()[:, n:] == 0) and np.all(im.cpu().numpy()[:, n:] == 0)     # padding untouched
    # inverse in place brings n * x back
    assert lib.CkFftB200ComplexInversePlanarBatchAsync(h, n, re.data_ptr(), im.data_ptr(), re.data_ptr(), im.data_ptr(), batch,
                                                       pitch, pitch, None) == 1
    torch.cuda.synchronize()
    back = (re.cpu().numpy()[:, :n] + 1j * im.cpu().numpy()[:, :n]).astype(np.complex64)
    assert rel_rms(back / n, x) <= tolerance(n)
    # error returns
    o1 = torch.empty_like(re); o2 = torch.empty_like(im)
    p = [t.data_ptr() for t in (re, im, o1, o2)]
    assert lib.CkFftB200ComplexForwardPlanarBatchAsync(h, n, p[0], p[1], p[0], p[3], batch, pitch, pitch, None) == 0   # one plane aliased
    assert lib.CkFftB200ComplexForwardPlanarBatchAsync(h, n, p[0], p[1], p[1], p[0], batch, pitch, pitch, None) == 0   # planes swapped
    assert lib.CkFftB200ComplexForwardPlanarBatchAsync(h, n, p[0], p[0], p[2], p[3], batch, pitch, pitch, None) == 0   # re is im
    assert lib.CkFftB200ComplexForwardPlanarBatchAsync(h, n, p[0], p[1], p[0], p[1], batch, pitch, pitch + 2, None) == 0   # in place, strides differ
    assert lib.CkFftB200ComplexForwardPlanarBatchAsync(h, n, p[0], None, p[2], p[3], batch, pitch, pitch, None) == 0
    assert lib.CkFftB200ComplexForwardPlanarBatchAsync(h, n, p[0], p[1], p[2], p[3], batch, n - 1, pitch, None) == 0   # stride < n
    assert lib.CkFftB200ComplexForwardPlanarBatchAsync(h, 48, p[0], p[1], p[2], p[3], batch, pitch, pitch, None) == 0  # not a power of two
    assert lib.CkFftB200ComplexForwardPlanarBatchAsync(h, n, p[0] + 2, p[1], p[2], p[3], 1, pitch, pitch, None) == 0   # misaligned
    assert lib.CkFftB200ComplexForwardPlanarBatchAsync(h, 32768, p[0], p[1], p[2], p[3], 1, 0, 0, None) == 0          # single-pass lengths only
    assert lib.CkFftB200ComplexForwardPlanarBatchAsync(h, n, p[0], p[1], p[2], p[3], 0, 0, 0, None) == 1               # empty batch
    fwd_only = ck.Context(64, ck.FORWARD)
    assert lib.CkFftB200ComplexInversePlanarBatchAsync(fwd_only.handle, 64, p[0], p[1], p[2], p[3], 1, 0, 0, None) == 0  # wrong direction
    fwd_only.close()


# ---------------------------------------------------------------------------------------------
# in-place stream-ordered calls
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 8, 16, 64, 256, 1024, 2048, 4096, 8192, 16384])
@pytest.mark.parametrize("inverse", [False, True])
def test_complex_in_place_matches_out_of_place(ctx_big, n, inverse):
    """Same kernels, so the in-place result must equal the out-of-place one bit for bit -- on batches large enough
    for every persistent CTA to walk several rows with its prefetch running ahead."""
    lib = _lib.load()
    batch = max(3, min(1 << 13, (1 << 24) // n))
    rng = np.random.default_rng(n)
    x = torch.from_numpy(uniform_complex(rng, (batch, n))).cuda()
    f = ctx_big.complex_inverse if inverse else ctx_big.complex_forward
    want = f(x)
    fn = lib.CkFftComplexInverseBatchAsync if inverse else lib.CkFftComplexForwardBatchAsync
    assert fn(ctx_big.handle, n, x.data_ptr(), x.data_ptr(), batch, 0, 0, torch.cuda.current_stream().cuda_stream) == 1, ck.last_error()
    torch.cuda.synchronize()
    assert np.array_equal(bits(x), bits(want))


@pytest.mark.parametrize("log2n", [15, 16, 18, 20, 21, 22])
def test_large_complex_in_place(log2n):
    """multi-pass lengths: the dataflow kernel (two passes) and the three-pass path write `output` only after the pass
    that reads `input` has consumed the transform"""
    lib = _lib.load()
    n = 1 << log2n
    batch = max(2, (1 << 24) // n)
    rng = np.random.default_rng(log2n)
    with ck.Context(n, ck.BOTH) as ctx:
        x = torch.from_numpy(uniform_complex(rng, (batch, n))).cuda()
        want = ctx.complex_forward(x)
        for rep in range(2):
            y = x.clone()
            assert lib.CkFftComplexForwardBatchAsync(ctx.handle, n, y.data_ptr(), y.data_ptr(), batch, 0, 0,
                                                     torch.cuda.current_stream().cuda_stream) == 1, ck.last_error()
            torch.cuda.synchronize()
            assert np.array_equal(bits(y), bits(want)), rep
        assert lib.CkFftComplexInverseBatchAsync(ctx.handle, n, y.data_ptr(), y.data_ptr(), batch, 0, 0, None) == 1
        torch.cuda.synchronize()
        assert rel_rms(y.cpu().numpy() / n, x.cpu().numpy()) <= tolerance(n)


@pytest.mark.parametrize("n", [1, 2, 4, 8, 16, 32, 64, 128, 1024, 4096, 8192, 16384, 32768])
def test_real_in_place_padded_rows(ctx_big, orc_big, n):
    """rows of n + 2 floats = n/2 + 1 complex: the forward transform overwrites the samples with the half spectrum, the
    inverse overwrites the spectrum with 2n * the samples (the reference's scaling, inc/ckfft/ckfft.h:123-125)"""
    lib = _lib.load()
    batch = 7 if n >= 4096 else 300
    bins = n // 2 + 1
    rng = np.random.default_rng(n + 1)
    x = rng.uniform(-1, 1, (batch, n)).astype(np.float32)
    buf = torch.zeros((batch, 2 * bins), dtype=torch.float32, device="cuda")
    buf[:, :n] = torch.from_numpy(x).cuda()
    stream = torch.cuda.current_stream().cuda_stream
    assert lib.CkFftRealForwardBatchAsync(ctx_big.handle, n, buf.data_ptr(), buf.data_ptr(), batch, 2 * bins, bins, stream) == 1, ck.last_error()
    torch.cuda.synchronize()
    got = buf.cpu().numpy().view(np.complex64)
    assert rel_rms(got, orc_big.real_forward(x)) <= tolerance(n)
    assert lib.CkFftRealInverseBatchAsync(ctx_big.handle, n, buf.data_ptr(), buf.data_ptr(), batch, bins, 2 * bins, stream) == 1, ck.last_error()
    torch.cuda.synchronize()
    back = buf.cpu().numpy()[:, :n]
    assert rel_rms(back / (2.0 * n), x) <= tolerance(n)


def test_in_place_argument_rules(ctx_big):
    lib = _lib.load()
    h = ctx_big.handle
    d = torch.zeros(4 * 1030, dtype=torch.complex64, device="cuda")
    p = d.data_ptr()
    assert lib.CkFftComplexForwardBatchAsync(h, 1024, p, p, 4, 1024, 1030, None) == 0       # rows would not coincide
    assert lib.CkFftComplexForwardBatchAsync(h, 1024, p, p, 4, 1030, 1030, None) == 1
    assert lib.CkFftRealForwardBatchAsync(h, 1024, p, p, 4, 1024, 513, None) == 0           # dense real rows != spectrum rows
    assert lib.CkFftRealForwardBatchAsync(h, 1024, p, p, 4, 1026, 513, None) == 1
    assert lib.CkFftRealInverseBatchAsync(h, 1024, p, p, 4, 513, 1024, None) == 0
    assert lib.CkFftRealInverseBatchAsync(h, 1024, p, p, 4, 513, 1026, None) == 1
    torch.cuda.synchronize()
    # the classic and the synchronous batched calls keep the reference's rule: in == out -> 0 (ckfft.cpp:46,69,88,107)
    assert lib.CkFftComplexForward(h, 1024, p, p) == 0
    assert lib.CkFftComplexForwardBatch(h, 1024, p, p, 4) == 0
    assert lib.CkFftRealForward(h, 1024, p, p) == 0


# ---------------------------------------------------------------------------------------------
# short rows (tile kernels, small_kernel.cuh): padded strides, ragged last tile, exactly one tile
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("batch", [1, 63, 64, 65, 127, 128, 129, 300])
@pytest.mark.parametrize("n", [8, 16, 32, 64])
def test_short_complex_rows_strided_and_ragged(ctx_big, orc_big, n, batch):
    lib = _lib.load()
    rng = np.random.default_rng(1000 * n + batch)
    pin, pout = n + 3, n + 5                      # rows start on 8-byte boundaries only
    x = uniform_complex(rng, (batch, n))
    xin = torch.zeros((batch, pin), dtype=torch.complex64, device="cuda")
    xin[:, :n] = torch.from_numpy(x).cuda()
    out = torch.full((batch, pout), 7.0 + 0j, dtype=torch.complex64, device="cuda")
    for inverse in (False, True):
        fn = lib.CkFftComplexInverseBatchAsync if inverse else lib.CkFftComplexForwardBatchAsync
        assert fn(ctx_big.handle, n, xin.data_ptr(), out.data_ptr(), batch, pin, pout, None) == 1, ck.last_error()
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        assert rel_rms(got[:, :n], orc_big.complex(x, inverse)) <= tolerance(n), (n, batch, inverse)
        assert np.all(got[:, n:] == 7.0), "padding between rows was written"


@pytest.mark.parametrize("batch", [1, 127, 128, 129, 300])
@pytest.mark.parametrize("n", [16, 32, 64])
def test_short_real_rows_strided_and_ragged(ctx_big, orc_big, n, batch):
    lib = _lib.load()
    rng = np.random.default_rng(2000 * n + batch)
    bins = n // 2 + 1
    pin, pspec, pout = n + 6, bins + 3, n + 2     # even strides of the real arrays (8-byte rows), odd spectrum pitch
    x = rng.uniform(-1, 1, (batch, n)).astype(np.float32)
    xin = torch.zeros((batch, pin), dtype=torch.float32, device="cuda")
    xin[:, :n] = torch.from_numpy(x).cuda()
    spec = torch.full((batch, pspec), 7.0 + 0j, dtype=torch.complex64, device="cuda")
    assert lib.CkFftRealForwardBatchAsync(ctx_big.handle, n, xin.data_ptr(), spec.data_ptr(), batch, pin, pspec, None) == 1, ck.last_error()
    torch.cuda.synchronize()
    got = spec.cpu().numpy()
    want = orc_big.real_forward(x)
    assert rel_rms(got[:, :bins], want) <= tolerance(n), (n, batch)
    assert np.all(got[:, bins:] == 7.0), "padding between spectrum rows was written"
    back = torch.full((batch, pout), 7.0, dtype=torch.float32, device="cuda")
    assert lib.CkFftRealInverseBatchAsync(ctx_big.handle, n, spec.data_ptr(), back.data_ptr(), batch, pspec, pout, None) == 1, ck.last_error()
    torch.cuda.synchronize()
    b = back.cpu().numpy()
    assert rel_rms(b[:, :n], orc_big.real_inverse(want, n)) <= tolerance(n), (n, batch)
    assert np.all(b[:, n:] == 7.0), "padding between sample rows was written"


def test_real_16_points_odd_strides_take_the_thread_per_transform_kernels(ctx_big, orc_big):
    """rows of an odd number of floats are only 4-byte aligned: the 8-byte tile copies do not apply (tiny_kernel.cuh)"""
    lib = _lib.load()
    n, batch, pin = 16, 200, 19
    rng = np.random.default_rng(77)
    x = rng.uniform(-1, 1, (batch, n)).astype(np.float32)
    xin = torch.zeros((batch, pin), dtype=torch.float32, device="cuda")
    xin[:, :n] = torch.from_numpy(x).cuda()
    spec = torch.zeros((batch, n // 2 + 1), dtype=torch.complex64, device="cuda")
    assert lib.CkFftRealForwardBatchAsync(ctx_big.handle, n, xin.data_ptr(), spec.data_ptr(), batch, pin, 0, None) == 1, ck.last_error()
    torch.cuda.synchronize()
    assert rel_rms(spec.cpu().numpy(), orc_big.real_forward(x)) <= tolerance(n)


# ---------------------------------------------------------------------------------------------
# multi-pass lengths (> 16384 complex points) with padded rows: one transform per launch sequence (api.cu)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("log2n", [15, 16, 18, 21])
def test_large_complex_rows_strided(log2n):
    """padded input and output rows above the single-pass limit, bit-identical to the dense call; the padding stays untouched"""
    lib = _lib.load()
    n = 1 << log2n
    batch = 3
    pin, pout = n + 6, n + 3                      # 16-byte aligned input rows (dataflow kernel) / 8-byte aligned output rows
    rng = np.random.default_rng(log2n)
    with ck.Context(n, ck.BOTH) as ctx:
        x = torch.from_numpy(uniform_complex(rng, (batch, n))).cuda()
        xin = torch.zeros((batch, pin), dtype=torch.complex64, device="cuda")
        xin[:, :n] = x
        for inverse in (False, True):
            want = ctx.complex_inverse(x) if inverse else ctx.complex_forward(x)
            out = torch.full((batch, pout), 7.0 + 0j, dtype=torch.complex64, device="cuda")
            fn = lib.CkFftComplexInverseBatchAsync if inverse else lib.CkFftComplexForwardBatchAsync
            assert fn(ctx.handle, n, xin.data_ptr(), out.data_ptr(), batch, pin, pout, None) == 1, ck.last_error()
            torch.cuda.synchronize()
            assert np.array_equal(bits(out[:, :n].contiguous()), bits(want)), (log2n, inverse)
            assert bool((out[:, n:] == 7.0).all()), "padding between rows was written"
        # rows that start 8 bytes off a 16-byte boundary (odd pitch): the plain-load tile kernels
        xodd = torch.zeros((batch, n + 1), dtype=torch.complex64, device="cuda")
        xodd[:, :n] = x
        out = torch.empty((batch, n + 1), dtype=torch.complex64, device="cuda")
        assert lib.CkFftComplexForwardBatchAsync(ctx.handle, n, xodd.data_ptr(), out.data_ptr(), batch, n + 1, n + 1, None) == 1, ck.last_error()
        torch.cuda.synchronize()
        assert rel_rms(out[:, :n].cpu().numpy(), ctx.complex_forward(x).cpu().numpy()) <= tolerance(n)


@pytest.mark.parametrize("n", [1 << 16, 1 << 17, 1 << 19, 1 << 22])
def test_large_real_rows_strided_and_in_place(n):
    lib = _lib.load()
    batch = 3
    bins = n // 2 + 1
    pin, pspec, pout = n + 4, bins + 2, n + 6
    rng = np.random.default_rng(n)
    with ck.Context(n, ck.BOTH) as ctx:
        x = torch.from_numpy(rng.uniform(-1, 1, (batch, n)).astype(np.float32)).cuda()
        want = ctx.real_forward(x)
        xin = torch.zeros((batch, pin), dtype=torch.float32, device="cuda")
        xin[:, :n] = x
        spec = torch.full((batch, pspec), 7.0 + 0j, dtype=torch.complex64, device="cuda")
        assert lib.CkFftRealForwardBatchAsync(ctx.handle, n, xin.data_ptr(), spec.data_ptr(), batch, pin, pspec, None) == 1, ck.last_error()
        torch.cuda.synchronize()
        assert np.array_equal(bits(spec[:, :bins].contiguous()), bits(want)), n
        assert bool((spec[:, bins:] == 7.0).all()), "padding between spectrum rows was written"
        back_want = ctx.real_inverse(want, n)
        back = torch.full((batch, pout), 7.0, dtype=torch.float32, device="cuda")
        assert lib.CkFftRealInverseBatchAsync(ctx.handle, n, spec.data_ptr(), back.data_ptr(), batch, pspec, pout, None) == 1, ck.last_error()
        torch.cuda.synchronize()
        assert rel_rms(back[:, :n].cpu().numpy(), back_want.cpu().numpy()) <= tolerance(n)
        assert bool((back[:, n:] == 7.0).all()), "padding between sample rows was written"
        # in place: rows of n + 2 floats = n/2 + 1 complex values, forward then inverse over the same bytes
        buf = torch.zeros((batch, n + 2), dtype=torch.float32, device="cuda")
        buf[:, :n] = x
        assert lib.CkFftRealForwardBatchAsync(ctx.handle, n, buf.data_ptr(), buf.data_ptr(), batch, n + 2, bins, None) == 1, ck.last_error()
        torch.cuda.synchronize()
        got = buf.view(torch.complex64)
        assert rel_rms(got.cpu().numpy(), want.cpu().numpy()) <= tolerance(n)
        assert lib.CkFftRealInverseBatchAsync(ctx.handle, n, buf.data_ptr(), buf.data_ptr(), batch, bins, n + 2, None) == 1, ck.last_error()
        torch.cuda.synchronize()
        assert rel_rms(buf[:, :n].cpu().numpy() / (2 * n), x.cpu().numpy()) <= tolerance(n)      # inverse(forward(x)) = 2 n x
