#!/usr/bin/env python3
"""Touch every kernel variant once with small, ragged batches (for compute-sanitizer runs):
   compute-sanitizer --tool memcheck python tools/small_cover.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ckfft_b200 as ck
rng = np.random.default_rng(0)
worst = 0.0
for lg in list(range(0, 16)) + [16, 21]:
    n = 1 << lg
    # short rows: a partial tile, exactly one tile, a full tile plus a ragged one (the tile kernels stage 64 / 128 rows)
    batches = (5, 128, 197) if n <= 64 else ((5,) if n <= 4096 else ((3,) if n <= 32768 else (1,)))
    ctx = ck.Context(max(n, 2), ck.BOTH)
    for batch in batches:
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
        e3 = 0.0
        if n <= 16384:
            re, im = xd.real.contiguous(), xd.imag.contiguous()
            ore, oim = ctx.complex_planar(re, im, False)
            torch.cuda.synchronize()
            e3 = float((torch.complex(ore, oim) - y).abs().max()) / max(1.0, float(y.abs().max()))
        worst = max(worst, e, e2, e3)
    ctx.close()
print("max round-trip / planar-vs-interleaved abs error", worst)
assert worst < 1e-4
