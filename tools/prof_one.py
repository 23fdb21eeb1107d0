#!/usr/bin/env python3
"""Run one transform shape a few times (for ncu):  python tools/prof_one.py <kind> <n> [log2_total_elems]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ckfft_b200 as ck  # noqa: E402

kind, n = sys.argv[1], int(sys.argv[2])
total = 1 << (int(sys.argv[3]) if len(sys.argv) > 3 else 27)
batch = max(1, total // n)
ctx = ck.Context(n, ck.BOTH)
x = torch.view_as_complex(torch.empty((batch, n, 2), dtype=torch.float32, device="cuda").uniform_(-1, 1))
out = torch.empty_like(x)
for _ in range(4):
    if kind == "c2c":
        ctx.complex_forward(x, out)
    elif kind == "c2cp":        # split-complex arrays
        f32 = x.view(torch.float32).view(-1)
        o32 = out.view(torch.float32).view(-1)
        ctx.complex_planar(f32[: batch * n].view(batch, n), f32[batch * n:].view(batch, n), False,
                           (o32[: batch * n].view(batch, n), o32[batch * n:].view(batch, n)))
    elif kind == "c2ci":
        ctx.complex_inverse(x, out)
    elif kind == "r2c":
        xr = x.view(torch.float32).view(-1)[: batch * n].view(batch, n)
        yo = out.view(-1)[: batch * (n // 2 + 1)].view(batch, n // 2 + 1)
        ctx.real_forward(xr, yo)
    else:
        yi = x.view(-1)[: batch * (n // 2 + 1)].view(batch, n // 2 + 1)
        xo = out.view(torch.float32).view(-1)[: batch * n].view(batch, n)
        ctx.real_inverse(yi, n, xo)
torch.cuda.synchronize()
