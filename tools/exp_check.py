#!/usr/bin/env python3
"""Numerics of an A/B library against torch.fft (fp64) for a few lengths: python tools/exp_check.py <kind> <n>..."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ckfft_b200 as ck  # noqa: E402
kind = sys.argv[1]
for n in [int(a) for a in sys.argv[2:]]:
    batch = 1237
    ctx = ck.Context(n, ck.BOTH)
    g = torch.Generator(device="cuda").manual_seed(n)
    if kind == "c2c":
        x = torch.view_as_complex(torch.empty((batch, n, 2), dtype=torch.float32, device="cuda").uniform_(-1, 1, generator=g))
        for inv in (False, True):
            got = (ctx.complex_inverse(x) if inv else ctx.complex_forward(x)).to(torch.complex128)
            want = torch.fft.ifft(x.to(torch.complex128), norm="forward") if inv else torch.fft.fft(x.to(torch.complex128))
            err = float(torch.linalg.norm(got - want) / torch.linalg.norm(want))
            print(f"{kind} n={n} inv={inv} rel={err:.2e} {'ok' if err < 1e-6 * np.log2(n) else 'FAIL'}")
    else:
        x = torch.empty((batch, n), dtype=torch.float32, device="cuda").uniform_(-1, 1, generator=g)
        y = ctx.real_forward(x)
        want = 2 * torch.fft.rfft(x.to(torch.float64))
        e1 = float(torch.linalg.norm(y.to(torch.complex128) - want) / torch.linalg.norm(want))
        back = ctx.real_inverse(y, n)
        e2 = float(torch.linalg.norm(back.to(torch.float64) / (2 * n) - x.to(torch.float64)) / torch.linalg.norm(x.to(torch.float64)))
        ok = e1 < 1e-6 * np.log2(n) and e2 < 2e-6 * np.log2(n)
        print(f"real n={n} fwd rel={e1:.2e} round trip={e2:.2e} {'ok' if ok else 'FAIL'}")
    ctx.close()
