"""CPU checks of the compile-time constants the CUDA kernels embed (ckfft_b200/csrc/fft_regs.cuh, small_kernel.cuh):
the butterfly factors cos(2 pi k / 32), the split / twist factors cos(2 pi k / 64), and the symmetry helpers that expand
them to the full circle (restated here in Python exactly as the constexpr functions are written)."""
import math
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "ckfft_b200", "csrc", "fft_regs.cuh")).read()


def table(name, count):
    m = re.search(r"constexpr float " + name + r"\(int k\)\s*\{\s*constexpr float t\[(\d+)\] = \{(.*?)\};", SRC, re.S)
    assert m and int(m.group(1)) == count
    vals = [float(v.strip().rstrip("f")) for v in m.group(2).replace("\n", " ").split(",") if v.strip()]
    assert len(vals) == count
    return vals


def test_cos32_table_is_correctly_rounded():
    t = table("cos32", 9)
    for k, v in enumerate(t):
        assert np.float32(v) == np.float32(math.cos(2 * math.pi * k / 32)) or (k == 8 and v == 0.0), (k, v)


def test_cos64_table_is_correctly_rounded():
    t = table("cos64", 17)
    for k, v in enumerate(t):
        assert np.float32(v) == np.float32(math.cos(2 * math.pi * k / 64)) or (k == 16 and v == 0.0), (k, v)


def _expand(t, n):
    """the constexpr cosN(k): first octant/quadrant table + symmetry, as written in fft_regs.cuh"""
    def cos(k):
        k &= n - 1
        if k > n // 2:
            k = n - k
        return -t[n // 2 - k] if k > n // 4 else t[k]
    return cos


def test_symmetry_expansion_covers_the_circle():
    for name, n in (("cos32", 32), ("cos64", 64)):
        t = table(name, n // 4 + 1)
        cos = _expand(t, n)
        sin = lambda k: cos(k - n // 4)
        for k in range(-2 * n, 2 * n):
            assert abs(cos(k) - math.cos(2 * math.pi * k / n)) < 6e-8, (name, k)
            assert abs(sin(k) - math.sin(2 * math.pi * k / n)) < 6e-8, (name, k)


def test_two_threads_per_row_decomposition():
    """the 64-point rows of small_kernel.cuh: X[k] = A0[k] + W_64^k A1[k], X[k+32] = A0[k] - W_64^k A1[k] with A_t the
    32-point transforms of the even / odd samples; thread t keeps the bins of its own parity"""
    rng = np.random.default_rng(3)
    x = rng.uniform(-1, 1, 64) + 1j * rng.uniform(-1, 1, 64)
    for sign in (-1, 1):
        a0 = np.fft.fft(x[0::2]) if sign < 0 else np.fft.ifft(x[0::2]) * 32
        a1 = np.fft.fft(x[1::2]) if sign < 0 else np.fft.ifft(x[1::2]) * 32
        w = np.exp(sign * 2j * np.pi * np.arange(32) / 64)
        got = np.empty(64, complex)
        for t in (0, 1):
            for i in range(16):
                k = 2 * i + t
                got[k] = a0[k] + w[k] * a1[k]
                got[k + 32] = a0[k] - w[k] * a1[k]
        want = np.fft.fft(x) if sign < 0 else np.fft.ifft(x) * 64
        assert np.allclose(got, want, rtol=0, atol=1e-12)


def test_split_factor_factorisation():
    """RTWC (fft_kernel.cuh): W_2M^(j + s*T + u*STR) = W_2M^j * W_64^(s*32/E + u*32/R) for every plan with a paired epilogue"""
    for M, E, R in ((2048, 32, 2), (4096, 32, 4), (8192, 32, 8), (16384, 32, 16), (128, 16, 8), (32, 8, 4)):
        T, STR, B = M // E, M // R, E // R
        for j in (0, 1, T - 1):
            for s in range(B // 2):
                for u in range(R):
                    k = j + s * T + u * STR
                    lhs = np.exp(-2j * np.pi * k / (2 * M))
                    rhs = np.exp(-2j * np.pi * j / (2 * M)) * np.exp(-2j * np.pi * (s * 32 // E + u * 32 // R) / 64)
                    assert (s * 32) % E == 0 and (u * 32) % R == 0
                    assert abs(lhs - rhs) < 1e-12, (M, j, s, u)


def test_split_factor_of_the_self_mirrored_butterflies():
    """thread 0's first pair slot (fft_kernel.cuh, r2c_paired_epilogue): bins STR/2 + w*STR carry W_64^(16/R + w*32/R)"""
    for M, R in ((2048, 2), (4096, 4), (8192, 8), (16384, 16), (128, 8), (32, 4)):
        STR = M // R
        for w in range(R // 2):
            k = STR // 2 + w * STR
            assert 16 % R == 0 and 32 % R == 0
            assert abs(np.exp(-2j * np.pi * k / (2 * M)) - np.exp(-2j * np.pi * (16 // R + w * (32 // R)) / 64)) < 1e-12
