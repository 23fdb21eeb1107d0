#!/usr/bin/env python3
"""Real transforms with dense (n/2+1) versus padded spectrum rows: python tools/stride_probe.py [n ...]   (GPU)"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ckfft_b200 as ck
from ckfft_b200 import _lib

lib = _lib.load()
peak = 6546.9
sizes = [int(a) for a in sys.argv[1:]] or [1024, 4096, 8192, 16384, 32768]
for n in sizes:
    batch = (1 << 27) // n
    ctx = ck.Context(n, ck.BOTH)
    x = torch.empty((batch, n), dtype=torch.float32, device="cuda").uniform_(-1, 1)
    for pad in (1, 2, 4, 16):
        stride = n // 2 + pad
        y = torch.zeros((batch, stride), dtype=torch.complex64, device="cuda")
        xo = torch.empty_like(x)
        s = torch.cuda.current_stream().cuda_stream
        fwd = lambda: lib.CkFftRealForwardBatchAsync(ctx.handle, n, x.data_ptr(), y.data_ptr(), batch, n, stride, s)
        inv = lambda: lib.CkFftRealInverseBatchAsync(ctx.handle, n, y.data_ptr(), xo.data_ptr(), batch, stride, n, s)
        res = []
        for f in (fwd, inv):
            for _ in range(3):
                assert f() == 1, ck.last_error()
            torch.cuda.synchronize()
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
            for a, b in evs:
                a.record(); f(); b.record()
            torch.cuda.synchronize()
            ms = float(np.median([a.elapsed_time(b) for a, b in evs]))
            res.append((4 * n + 8 * (n // 2 + 1)) * batch / ms / 1e6 / peak)
        err = float((xo / (2 * n) - x).norm() / x.norm())
        print(f"n={n:6d} spectrum stride n/2+{pad:<2d}  r2c {res[0]:.3f}  c2r {res[1]:.3f}  roundtrip err {err:.1e}", flush=True)
    ctx.close()
