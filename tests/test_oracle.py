"""CPU tests: the oracle (oracle/ckfft_oracle.c) is pinned against the reference's own fixture,
its golden outputs, the harness's acceptance criteria and fp64 truth.  No GPU, no product code."""
import numpy as np
import pytest

import oracle
from conftest import rel_rms, tolerance, uniform_complex

SIZES = [1 << k for k in range(0, 13)]   # 1 .. 4096, the range of the reference's regression test


@pytest.fixture(scope="module")
def orc():
    o = oracle.Restatement(8192, 3)
    yield o
    o.close()


def harness_rms(a, b):
    """compare() of the reference harness (src/test/test.cpp:140-159): RMS over all components."""
    d = (np.asarray(a, np.complex64) - np.asarray(b, np.complex64)).view(np.float32)
    return float(np.sqrt(np.sum(d.astype(np.float32) ** 2) / d.size))


@pytest.mark.parametrize("n", SIZES)
def test_restatement_bit_exact_vs_golden(orc, golden, n):
    """The restatement reproduces the compiled reference bit for bit on the reference's fixture."""
    x = golden["input"][:n]
    for inv, key in ((False, "cfwd"), (True, "cinv")):
        got = orc.complex(x, inv)
        assert np.array_equal(got.view(np.uint32), golden[f"{key}_{n}"].view(np.uint32)), (n, key)
    xr = np.ascontiguousarray(x.real)
    rf = orc.real_forward(xr)
    assert np.array_equal(rf.view(np.uint32), golden[f"rfwd_{n}"].view(np.uint32))
    spec = orc.complex(xr.astype(np.complex64), False)[: n // 2 + 1]
    ri = orc.real_inverse(spec, n)
    assert np.array_equal(ri.view(np.uint32), golden[f"rinv_{n}"].view(np.uint32))


@pytest.mark.parametrize("n", SIZES)
def test_harness_complex_criterion(orc, golden, n):
    """regressionTestComplex (src/test/test.cpp:754-786): RMS difference to KISS FFT <= 0.001."""
    x = golden["input"][:n]
    assert harness_rms(orc.complex(x, False), golden[f"kfwd_{n}"]) <= 1e-3
    assert harness_rms(orc.complex(x, True), golden[f"kinv_{n}"]) <= 1e-3


@pytest.mark.parametrize("n", SIZES)
def test_harness_real_criterion(orc, golden, n):
    """regressionTestReal (src/test/test.cpp:788-864): real forward * 0.5 vs complex forward of the real
    data (first n/2+1 bins); real inverse vs the real part of the complex inverse (first n/2+1 samples)."""
    xr = np.ascontiguousarray(golden["input"][:n].real)
    ref = orc.complex(xr.astype(np.complex64), False)
    out = orc.real_forward(xr) * np.float32(0.5)
    k = n // 2 + 1
    assert harness_rms(out[:k], ref[:k]) <= 1e-3
    inv_ref = orc.complex(ref, True)
    fl = orc.real_inverse(ref[:k], n)
    d = fl[:k] - inv_ref[:k].real
    assert float(np.sqrt(np.sum(d * d) / k)) <= 1e-3


def test_example_known_answer(orc, golden):
    """src/example/main.cpp:44-84: forward then inverse of 1024 real samples returns 2048 * input."""
    x = golden["example_in"]
    f = orc.real_forward(x)
    assert np.array_equal(f.view(np.uint32), golden["example_fwd"].view(np.uint32))
    rt = orc.real_inverse(f, 1024)
    assert np.array_equal(rt.view(np.uint32), golden["example_rt"].view(np.uint32))
    err = float(np.sum((rt / 2048.0 - x) ** 2))
    assert err < 1e-6   # the example prints "error: 0.000000"


@pytest.mark.parametrize("n", [1, 2, 4, 8, 64, 1024, 4096, 16384, 1 << 17])
def test_against_fp64_truth(n):
    rng = np.random.default_rng(n)
    o = oracle.Restatement(n, 3)
    x = uniform_complex(rng, (4, n))
    assert rel_rms(o.complex(x, False), oracle.fp64_c2c(x, False)) <= tolerance(n)
    assert rel_rms(o.complex(x, True), oracle.fp64_c2c(x, True)) <= tolerance(n)
    xr = np.ascontiguousarray(x.real)
    y = o.real_forward(xr)
    assert rel_rms(y, oracle.fp64_real_forward(xr)) <= tolerance(n)
    assert rel_rms(o.real_inverse(y, n), oracle.fp64_real_inverse(y, n)) <= tolerance(n)
    assert rel_rms(o.real_inverse(y, n), 2.0 * n * xr.astype(np.float64)) <= tolerance(n)
    o.close()


def test_fftw_matches_numpy_fp64():
    rng = np.random.default_rng(3)
    x = uniform_complex(rng, (3, 2048))
    a = oracle.fp64_c2c(x)
    b = np.fft.fft(x.astype(np.complex128))
    assert rel_rms(a, b) < 1e-14


@pytest.mark.skipif(not oracle.reference_available(), reason="oracle/_ref/libckfft_ref.so not built")
@pytest.mark.parametrize("n", [1, 2, 4, 8, 16, 32, 128, 1024, 2048, 8192, 1 << 16])
def test_restatement_bit_exact_vs_compiled_reference(n):
    """Random data, both table sizes (maxCount == n and maxCount == 2n, the harness's two passes)."""
    rng = np.random.default_rng(100 + n)
    x = uniform_complex(rng, (5, n))
    xr = np.ascontiguousarray(x.real)
    for nmax in (n, 2 * n):
        R, O = oracle.Reference(nmax, 3), oracle.Restatement(nmax, 3)
        for inv in (False, True):
            assert np.array_equal(R.complex(x, inv).view(np.uint32), O.complex(x, inv).view(np.uint32))
        y = R.real_forward(xr)
        assert np.array_equal(y.view(np.uint32), O.real_forward(xr).view(np.uint32))
        assert np.array_equal(R.real_inverse(y, n).view(np.uint32), O.real_inverse(y, n).view(np.uint32))
        R.close(); O.close()


def test_twiddle_table_is_independent_of_nmax(orc):
    """table_n[k] == table_nmax[k * nmax / n] bit for bit (power-of-two scaling is exact)."""
    big = orc.twiddles(8192)
    for n in (8, 64, 1024, 4096):
        small = orc.twiddles(n)
        assert np.array_equal(small.view(np.uint32), np.ascontiguousarray(big[:: 8192 // n]).view(np.uint32))
    assert np.array_equal(orc.twiddles(64, True), np.conj(orc.twiddles(64)))


def test_closed_forms_and_conventions(orc):
    """Values established by running the reference (SURVEY 8a 'Conventions')."""
    imp = np.zeros(8, np.complex64); imp[1] = 1
    k = np.arange(8)
    want = np.cos(np.pi * k / 4) - 1j * np.sin(np.pi * k / 4)
    assert np.allclose(orc.complex(imp), want, atol=1e-6)
    r = orc.real_forward(np.arange(8, dtype=np.float32))
    assert np.allclose(r, [56, -8 + 19.3137j, -8 + 8j, -8 + 3.3137j, -8], atol=1e-3)
    assert np.allclose(orc.real_forward(np.array([3.0], np.float32)), [6.0])
    assert np.allclose(orc.real_forward(np.array([1.0, 2.0], np.float32)), [6.0, -2.0])
    x4 = np.array([1, 2, 3, 4], np.float32)
    assert np.allclose(orc.real_forward(x4), 2 * np.fft.rfft(x4))
    assert np.allclose(orc.real_inverse(np.fft.rfft(x4).astype(np.complex64), 4), 4 * x4)


def test_argument_checks():
    """src/ckfft/ckfft.cpp:14-114: NULL / 0 on bad arguments."""
    with pytest.raises(ValueError):
        oracle.Restatement(1000, 3)
    with pytest.raises(ValueError):
        oracle.Restatement(1024, 4)
    fwd_only = oracle.Restatement(16, 1)
    x = np.ones(16, np.complex64)
    fwd_only.complex(x, False)
    with pytest.raises(ValueError):
        fwd_only.complex(x, True)           # no inverse table
    with pytest.raises(ValueError):
        fwd_only.complex(np.ones(32, np.complex64))   # n > nmax
    with pytest.raises(ValueError):
        fwd_only.complex(np.ones(12, np.complex64))   # not a power of two
    fwd_only.close()
