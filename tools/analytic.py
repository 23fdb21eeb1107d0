"""Closed-form test signal for transforms too long for any CPU oracle (BASELINE config 5, N = 2^30: the reference cannot
even create a context there, src/ckfft/context.cpp:37-45 overflows its int byte count at nMax >= 2^28).

    x[j] = sum_i a_i exp(+2 pi i f_i j / n) + delta[j - n0]
    X[k] = exp(-2 pi i k n0 / n) + n * sum_i a_i delta[k - f_i]          (forward transform, ckfft sign convention)

Both are generated slice by slice on the GPU in float64 (chunked, so a 2^30-point signal never needs a 2^30-point
float64 temporary); a rank of a distributed run builds only its own slice.  Used by bench.py (`secondary.dist30`) and
tests/test_parity_gpu.py."""
from __future__ import annotations

import numpy as np


def analytic_signal(n, lo, cnt, dev, freqs, amps, n0, chunk=1 << 26):
    """x[j] = sum_i a_i exp(+2 pi i f_i j / n) + delta[j - n0] for j in [lo, lo + cnt), complex64, built in chunks"""
    import torch

    x = torch.empty(cnt, dtype=torch.complex64, device=dev)
    for c0 in range(0, cnt, chunk):
        c1 = min(cnt, c0 + chunk)
        idx = torch.arange(lo + c0, lo + c1, device=dev, dtype=torch.int64)
        acc = torch.zeros(c1 - c0, dtype=torch.complex128, device=dev)
        for f, a in zip(freqs, amps):
            ph = (2.0 * np.pi / n) * ((idx * f) % n).to(torch.float64)
            acc += a * torch.complex(torch.cos(ph), torch.sin(ph))
        x[c0:c1] = acc.to(torch.complex64)
    if lo <= n0 < lo + cnt:
        x[n0 - lo] += 1.0
    return x


def analytic_error(y, n, lo, dev, freqs, amps, n0, chunk=1 << 26):
    """(sum |y - X|^2, sum |X|^2) over this slice of the closed-form spectrum X[k] = exp(-2 pi i k n0 / n) + n a_i delta[k - f_i]"""
    import torch

    num = torch.zeros((), dtype=torch.float64, device=dev)
    den = torch.zeros((), dtype=torch.float64, device=dev)
    cnt = y.numel()
    for c0 in range(0, cnt, chunk):
        c1 = min(cnt, c0 + chunk)
        k = torch.arange(lo + c0, lo + c1, device=dev, dtype=torch.int64)
        ph = (-2.0 * np.pi / n) * ((k * n0) % n).to(torch.float64)
        want = torch.complex(torch.cos(ph), torch.sin(ph))
        for f, a in zip(freqs, amps):
            if lo + c0 <= f < lo + c1:
                want[f - lo - c0] += a * n
        d = y[c0:c1].to(torch.complex128) - want
        num += (d.real ** 2 + d.imag ** 2).sum()
        den += (want.real ** 2 + want.imag ** 2).sum()
    return num, den
