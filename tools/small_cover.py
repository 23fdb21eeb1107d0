#!/usr/bin/env python3
"""Touch every kernel variant once with small batches (for compute-sanitizer runs)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ckfft_b200 as ck
rng = np.random.default_rng(0)
worst = 0.0
for lg in list(range(0, 16)) + [16, 21]:
    n = 1 << lg
    batch = 5 if n <= 4096 else (3 if n <= 32768 else 1)
    ctx = ck.Context(max(n, 2), ck.BOTH)
    x = (rng.uniform(-1, 1, (batch, n)) + 1j * rng.uniform(-1, 1, (batch, n))).astype(np.complex64)
    xd = torch.from_numpy(x).cuda()
    y = ctx.complex_forward(xd)
    z = ctx.complex_inverse(y)
    torch.cuda.synchronize()
    e = float((z / n - xd).abs().max())
    xr = torch.from_numpy(np.ascontiguousarray(x.real)).cuda()
    yr = ctx.real_forward(xr)
    zr = ctx.real_inverse(yr, n)
    torch.cuda.synchronize()
    e2 = float((zr / (2.0 * n) - xr).abs().max())
    worst = max(worst, e, e2)
    ctx.close()
print("max round-trip abs error", worst)
assert worst < 1e-4
