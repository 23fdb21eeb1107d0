"""ctypes front end for the parity oracle.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package.  The product package (ckfft_b200) never does.

Three checkers live here:
  Restatement  -- oracle/ckfft_oracle.c, our iterative C restatement of the reference's scalar
                  path (bit-identical to the reference; see the header of that file).
  Reference    -- oracle/_ref/libckfft_ref.so, the UNMODIFIED reference compiled from
                  /root/reference by oracle/build.py (travels to the GPU box as a binary).
  fftw_c2c     -- oracle/_ref/libfftw3_ref.so, vendored FFTW 3.3.2 in double precision, the
                  fp64 truth named by BASELINE.json; numpy's fp64 pocketfft when absent.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_FP = C.POINTER(C.c_float)


def _fp(a):
    return a.ctypes.data_as(_FP)


def _ensure_built():
    path = os.path.join(HERE, "libckfft_oracle.so")
    src = os.path.join(HERE, "ckfft_oracle.c")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        from . import build  # noqa: PLC0415

        build.build_restatement()
    return path


class Restatement:
    """oracle/ckfft_oracle.c.  Mirrors CkFftInit/.../CkFftShutdown (inc/ckfft/ckfft.h:59-158)."""

    def __init__(self, nmax: int, direction: int = 3):
        lib = C.CDLL(_ensure_built())
        lib.ckfft_oracle_init.restype = C.c_void_p
        lib.ckfft_oracle_init.argtypes = [C.c_int, C.c_int]
        lib.ckfft_oracle_shutdown.argtypes = [C.c_void_p]
        lib.ckfft_oracle_complex_batch.argtypes = [C.c_void_p, C.c_int, _FP, _FP, C.c_long, C.c_int]
        lib.ckfft_oracle_real_forward_batch.argtypes = [C.c_void_p, C.c_int, _FP, _FP, C.c_long]
        lib.ckfft_oracle_real_inverse_batch.argtypes = [C.c_void_p, C.c_int, _FP, _FP, C.c_long]
        lib.ckfft_oracle_twiddles.argtypes = [C.c_int, C.c_int, _FP]
        self.lib = lib
        self.nmax = nmax
        self.ctx = lib.ckfft_oracle_init(nmax, direction)
        if not self.ctx:
            raise ValueError(f"oracle init rejected nmax={nmax} direction={direction}")

    def close(self):
        if self.ctx:
            self.lib.ckfft_oracle_shutdown(self.ctx)
            self.ctx = None

    __del__ = close

    def twiddles(self, nmax: int, inverse: bool = False) -> np.ndarray:
        t = np.empty(2 * nmax, np.float32)
        self.lib.ckfft_oracle_twiddles(nmax, int(inverse), _fp(t))
        return t.view(np.complex64)

    def complex(self, x: np.ndarray, inverse: bool = False) -> np.ndarray:
        """x: complex64 [..., n] -> same shape."""
        x = np.ascontiguousarray(x, np.complex64)
        n = x.shape[-1]
        out = np.empty_like(x)
        ok = self.lib.ckfft_oracle_complex_batch(self.ctx, n, _fp(x.view(np.float32)), _fp(out.view(np.float32)),
                                                 x.size // n, int(inverse))
        if not ok:
            raise ValueError("oracle rejected arguments")
        return out

    def real_forward(self, x: np.ndarray) -> np.ndarray:
        """x: float32 [..., n] -> complex64 [..., n/2+1] (= 2 * rfft)."""
        x = np.ascontiguousarray(x, np.float32)
        n = x.shape[-1]
        out = np.empty(x.shape[:-1] + (n // 2 + 1,), np.complex64)
        ok = self.lib.ckfft_oracle_real_forward_batch(self.ctx, n, _fp(x), _fp(out.view(np.float32)), x.size // n)
        if not ok:
            raise ValueError("oracle rejected arguments")
        return out

    def real_inverse(self, y: np.ndarray, n: int) -> np.ndarray:
        """y: complex64 [..., n/2+1] -> float32 [..., n]."""
        y = np.ascontiguousarray(y, np.complex64)
        assert y.shape[-1] == n // 2 + 1
        out = np.empty(y.shape[:-1] + (n,), np.float32)
        ok = self.lib.ckfft_oracle_real_inverse_batch(self.ctx, n, _fp(y.view(np.float32)), _fp(out),
                                                      y.size // (n // 2 + 1))
        if not ok:
            raise ValueError("oracle rejected arguments")
        return out


def reference_available() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libckfft_ref.so"))


class Reference:
    """The unmodified reference library (oracle/_ref/libckfft_ref.so) through ref_driver.cpp."""

    def __init__(self, nmax: int, direction: int = 3):
        lib = C.CDLL(os.path.join(HERE, "_ref", "libckfft_ref.so"))
        lib.ckref_init.restype = C.c_void_p
        lib.ckref_init.argtypes = [C.c_int, C.c_int]
        lib.ckref_shutdown.argtypes = [C.c_void_p]
        lib.ckref_complex_batch.argtypes = [C.c_void_p, C.c_int, _FP, _FP, C.c_long, C.c_int, C.c_int]
        lib.ckref_real_forward_batch.argtypes = [C.c_void_p, C.c_int, _FP, _FP, C.c_long, C.c_int]
        lib.ckref_real_inverse_batch.argtypes = [C.c_void_p, C.c_int, _FP, _FP, C.c_long, C.c_int]
        lib.ckref_kiss_complex.argtypes = [C.c_int, _FP, _FP, C.c_int]
        self.lib = lib
        self.nmax = nmax
        self.ctx = lib.ckref_init(nmax, direction)
        if not self.ctx:
            raise ValueError(f"reference CkFftInit returned NULL for nmax={nmax} direction={direction}")

    def close(self):
        if self.ctx:
            self.lib.ckref_shutdown(self.ctx)
            self.ctx = None

    __del__ = close

    def max_threads(self) -> int:
        return int(self.lib.ckref_max_threads())

    def complex(self, x, inverse=False, threads=0, out=None):
        x = np.ascontiguousarray(x, np.complex64)
        n = x.shape[-1]
        if out is None:
            out = np.empty_like(x)
        ok = self.lib.ckref_complex_batch(self.ctx, n, _fp(x.view(np.float32)), _fp(out.view(np.float32)),
                                          x.size // n, int(inverse), threads)
        if not ok:
            raise ValueError("reference returned 0")
        return out

    def real_forward(self, x, threads=0, out=None):
        x = np.ascontiguousarray(x, np.float32)
        n = x.shape[-1]
        if out is None:
            out = np.empty(x.shape[:-1] + (n // 2 + 1,), np.complex64)
        ok = self.lib.ckref_real_forward_batch(self.ctx, n, _fp(x), _fp(out.view(np.float32)), x.size // n, threads)
        if not ok:
            raise ValueError("reference returned 0")
        return out

    def real_inverse(self, y, n, threads=0, out=None):
        y = np.ascontiguousarray(y, np.complex64)
        assert y.shape[-1] == n // 2 + 1
        if out is None:
            out = np.empty(y.shape[:-1] + (n,), np.float32)
        ok = self.lib.ckref_real_inverse_batch(self.ctx, n, _fp(y.view(np.float32)), _fp(out),
                                               y.size // (n // 2 + 1), threads)
        if not ok:
            raise ValueError("reference returned 0")
        return out

    def kiss(self, x, inverse=False):
        x = np.ascontiguousarray(x, np.complex64)
        assert x.ndim == 1
        out = np.empty_like(x)
        self.lib.ckref_kiss_complex(x.shape[0], _fp(x.view(np.float32)), _fp(out.view(np.float32)), int(inverse))
        return out


# ---------------------------------------------------------------------------------------------
# fp64 truth
# ---------------------------------------------------------------------------------------------
_fftw = None


def fftw_available() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libfftw3_ref.so"))


def _load_fftw():
    global _fftw
    if _fftw is None:
        lib = C.CDLL(os.path.join(HERE, "_ref", "libfftw3_ref.so"))
        lib.fftw_plan_many_dft.restype = C.c_void_p
        lib.fftw_plan_many_dft.argtypes = [C.c_int, C.POINTER(C.c_int), C.c_int,
                                           C.c_void_p, C.POINTER(C.c_int), C.c_int, C.c_int,
                                           C.c_void_p, C.POINTER(C.c_int), C.c_int, C.c_int,
                                           C.c_int, C.c_uint]
        lib.fftw_execute.argtypes = [C.c_void_p]
        lib.fftw_destroy_plan.argtypes = [C.c_void_p]
        _fftw = lib
    return _fftw


def fp64_c2c(x: np.ndarray, inverse: bool = False) -> np.ndarray:
    """Unnormalised fp64 DFT over the last axis; sign convention of ckfft (forward = exp(-2*pi*i*jk/N)).

    Uses the vendored FFTW 3.3.2 (ext/fftw-3.3.2, FFTW_FORWARD = -1, api/fftw3.h) when
    oracle/_ref/libfftw3_ref.so exists, numpy's fp64 pocketfft otherwise.
    """
    x = np.ascontiguousarray(x, np.complex128)
    n = x.shape[-1]
    if not fftw_available():
        return np.fft.ifft(x, axis=-1) * n if inverse else np.fft.fft(x, axis=-1)
    lib = _load_fftw()
    out = np.empty_like(x)
    howmany = x.size // n
    nn = (C.c_int * 1)(n)
    FFTW_ESTIMATE = 1 << 6
    plan = lib.fftw_plan_many_dft(1, nn, howmany, x.ctypes.data, None, 1, n, out.ctypes.data, None, 1, n,
                                  +1 if inverse else -1, FFTW_ESTIMATE)
    if not plan:
        raise RuntimeError("fftw_plan_many_dft failed")
    lib.fftw_execute(plan)
    lib.fftw_destroy_plan(plan)
    return out


def fp64_real_forward(x: np.ndarray) -> np.ndarray:
    """fp64 truth for CkFftRealForward: 2 * rfft(x) (src/ckfft/fft_real_default.cpp:13-63)."""
    x = np.asarray(x, np.float64)
    n = x.shape[-1]
    return 2.0 * fp64_c2c(x.astype(np.complex128))[..., : n // 2 + 1]


def fp64_real_inverse(y: np.ndarray, n: int) -> np.ndarray:
    """fp64 truth for CkFftRealInverse on a Hermitian half-spectrum y[..., n/2+1]: n * irfft(y).

    Imaginary parts of y[0] and y[n/2] are used as given by the reference
    (fft_real_default.cpp:79-107); for spectra produced by a real forward transform they are 0.
    """
    y = np.asarray(y, np.complex128)
    h = n // 2
    if n == 1:
        return y[..., :1].real.copy()
    full = np.concatenate([y, np.conj(y[..., h - 1:0:-1])], axis=-1)
    return fp64_c2c(full, inverse=True).real
