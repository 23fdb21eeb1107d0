#!/usr/bin/env python3
"""Pipeline-depth grid of the dataflow kernel (development): time per launch for NBUF x LAG x RING_MB at one length.
    python tools/pipe_grid.py <log2n> [kind]        kind: c2c (default) | r2c | c2r"""
import itertools
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ckfft_b200 as ck  # noqa: E402

lg = int(sys.argv[1])
kind = sys.argv[2] if len(sys.argv) > 2 else "c2c"
n = 1 << lg
total = 1 << 27
batch = total // n
ctx = ck.Context(n, ck.BOTH)
x = torch.view_as_complex(torch.empty((batch, n, 2), dtype=torch.float32, device="cuda").uniform_(-1, 1))
y = torch.empty_like(x)
if kind == "c2c":
    f = lambda: ctx.complex_forward(x, y)
    nbytes = 16 * total
elif kind == "r2c":
    xr = x.view(torch.float32).view(-1)[: batch * n].view(batch, n)
    yo = y.view(-1)[: batch * (n // 2 + 1)].view(batch, n // 2 + 1)
    f = lambda: ctx.real_forward(xr, yo)
    nbytes = (4 * n + 8 * (n // 2 + 1)) * batch
else:
    yi = x.view(-1)[: batch * (n // 2 + 1)].view(batch, n // 2 + 1)
    xo = y.view(torch.float32).view(-1)[: batch * n].view(batch, n)
    f = lambda: ctx.real_inverse(yi, n, xo)
    nbytes = (4 * n + 8 * (n // 2 + 1)) * batch


def run():
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(8)]
    for a, b in evs:
        a.record(); f(); b.record()
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in evs]))


print(f"n=2^{lg} {kind} batch={batch}")
base = run()
print(f"default: {base:.4f} ms  frac {nbytes / base / 1e6 / 6545:.3f}")
lags = [int(a) for a in os.environ.get("GRID_LAGS", "0").split(",")]
for nbuf, ring_mb, lag in itertools.product((1, 2), (64, 96), lags):
    os.environ["CKFFT_B200_PIPE_NBUF"] = str(nbuf)
    os.environ["CKFFT_B200_PIPE_RING_MB"] = str(ring_mb)
    if lag:
        os.environ["CKFFT_B200_PIPE_LAG"] = str(lag)
    else:
        os.environ.pop("CKFFT_B200_PIPE_LAG", None)
    ms = run()
    print(f"nbuf={nbuf} ring_mb={ring_mb:3d} lag={lag:4d}: {ms:.4f} ms  frac {nbytes / ms / 1e6 / 6545:.3f}", flush=True)
