#!/usr/bin/env python3
"""e2e (pinned host buffers through the C ABI) for a few chunk sizes: CKFFT_B200_CHUNK_MB=<mb> python tools/e2e_probe.py"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ckfft_b200 as ck
n, batch = 1024, 1 << 19
hx = torch.empty((batch, n), dtype=torch.complex64, pin_memory=True); torch.view_as_real(hx).uniform_(-1, 1)
hy = torch.empty_like(hx).pin_memory()
ctx = ck.Context(n, ck.BOTH)
nx, ny = hx.numpy(), hy.numpy()
ctx.complex_forward(nx, ny)
t = time.perf_counter()
for _ in range(3): ctx.complex_forward(nx, ny)
dt = (time.perf_counter() - t) / 3
print(os.environ.get("CKFFT_B200_CHUNK_MB", "32"), "MiB chunks:", round(16 * n * batch / dt / 1e9, 1), "GB/s")
