#!/usr/bin/env python3
"""Latency of ONE classic call on host buffers (BASELINE config 1: what the reference's harness times, src/test/test.cpp:611-662),
through the C ABI with ctypes overhead excluded as far as possible (pre-bound function, raw pointers).
    python tools/latency_probe.py [n ...]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ckfft_b200 as ck  # noqa: E402
from ckfft_b200 import _lib  # noqa: E402

lib = _lib.load()
sizes = [int(a) for a in sys.argv[1:]] or [64, 1024, 4096, 16384]
for n in sizes:
    ctx = ck.Context(n, ck.BOTH)
    rng = np.random.default_rng(n)
    x = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64)
    y = np.empty_like(x)
    f, h, px, py = lib.CkFftComplexForward, ctx.handle, x.ctypes.data, y.ctypes.data
    for _ in range(200):
        f(h, n, px, py)
    ts = []
    for _ in range(3000):
        t0 = time.perf_counter_ns()
        f(h, n, px, py)
        ts.append(time.perf_counter_ns() - t0)
    ts = np.array(ts) / 1e3
    xr = np.ascontiguousarray(x.real)
    yr = np.empty(n // 2 + 1, np.complex64)
    fr, pxr, pyr = lib.CkFftRealForward, xr.ctypes.data, yr.ctypes.data
    for _ in range(200):
        fr(h, n, pxr, pyr)
    tr = []
    for _ in range(3000):
        t0 = time.perf_counter_ns()
        fr(h, n, pxr, pyr)
        tr.append(time.perf_counter_ns() - t0)
    tr = np.array(tr) / 1e3
    print(f"n={n:6d} spin_sync={os.environ.get('CKFFT_B200_SPIN_SYNC', '1')}: CkFftComplexForward median {np.median(ts):6.2f} us  p10 {np.percentile(ts, 10):6.2f}  p90 {np.percentile(ts, 90):6.2f} | "
          f"CkFftRealForward median {np.median(tr):6.2f} us", flush=True)
    ctx.close()
