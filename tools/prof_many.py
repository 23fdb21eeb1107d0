#!/usr/bin/env python3
"""Run several transform shapes in one process (for one ncu session):
    python tools/prof_many.py r2c:8192 c2r:4096 c2c:16384 [--log2-total 25] [--reps 2]
kinds: c2c, c2ci (inverse), r2c, c2r, c2cp (split-complex arrays)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ckfft_b200 as ck  # noqa: E402

args = sys.argv[1:]
log2_total, reps, shapes = 25, 2, []
while args:
    a = args.pop(0)
    if a == "--log2-total":
        log2_total = int(args.pop(0))
    elif a == "--reps":
        reps = int(args.pop(0))
    else:
        kind, n = a.split(":")
        shapes.append((kind, int(n)))

for kind, n in shapes:
    batch = max(1, (1 << log2_total) // n)
    ctx = ck.Context(n, ck.BOTH)
    x = torch.view_as_complex(torch.empty((batch, n, 2), dtype=torch.float32, device="cuda").uniform_(-1, 1))
    out = torch.empty_like(x)
    for _ in range(reps):
        if kind == "c2c":
            ctx.complex_forward(x, out)
        elif kind == "c2ci":
            ctx.complex_inverse(x, out)
        elif kind == "c2cp":
            f32 = x.view(torch.float32).view(-1)
            o32 = out.view(torch.float32).view(-1)
            ctx.complex_planar(f32[: batch * n].view(batch, n), f32[batch * n:].view(batch, n), False,
                               (o32[: batch * n].view(batch, n), o32[batch * n:].view(batch, n)))
        elif kind == "r2c":
            xr = x.view(torch.float32).view(-1)[: batch * n].view(batch, n)
            yo = out.view(-1)[: batch * (n // 2 + 1)].view(batch, n // 2 + 1)
            ctx.real_forward(xr, yo)
        else:
            yi = x.view(-1)[: batch * (n // 2 + 1)].view(batch, n // 2 + 1)
            xo = out.view(torch.float32).view(-1)[: batch * n].view(batch, n)
            ctx.real_inverse(yi, n, xo)
    torch.cuda.synchronize()
    print(kind, n, batch, flush=True)
    ctx.close()
    del x, out
