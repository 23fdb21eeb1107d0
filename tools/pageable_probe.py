#!/usr/bin/env python3
"""Pageable (malloc'ed) host arrays through CkFftComplexForwardBatch: the staging copy of host_copy.cpp, A/B over
CKFFT_B200_NT_COPY (0 memcpy, 1 widest non-temporal stores, 2 SSE2, 3 AVX2, 4 AVX-512) in one process (the switch is read
per call; CKFFT_B200_HOST_THREADS, read once per process, sets the copy threads per team).  2^18 transforms of 1024 points = 2 GiB in + 2 GiB out per call; every result is compared bit for bit with the
device-resident path.   python tools/pageable_probe.py [levels...]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ckfft_b200 as ck
n, batch = 1024, 1 << int(os.environ.get("PROBE_LOG2_BATCH", "18"))
levels = [int(a) for a in sys.argv[1:]] or [0, 1, 2, 3, 4, 1, 0]
rng = np.random.default_rng(1)
x = np.empty((batch, n), np.complex64)
x.view(np.float32)[...] = rng.random((batch, 2 * n), dtype=np.float32) * 2 - 1
y = np.zeros((batch, n), np.complex64)
ctx = ck.Context(n, ck.BOTH)
want = ctx.complex_forward(torch.from_numpy(x).cuda()).cpu().numpy()
flags = open("/proc/cpuinfo").read().split("flags")[1].split("\n")[0] if os.path.exists("/proc/cpuinfo") else ""
print("host threads per team:", os.environ.get("CKFFT_B200_HOST_THREADS", "default"), "| cpu:", os.cpu_count(), "threads; avx2", " avx2 " in flags + " ", "avx512f", " avx512f " in flags + " ")
for lv in levels:
    os.environ["CKFFT_B200_NT_COPY"] = str(lv)
    y[...] = 0
    ctx.complex_forward(x, out=y)
    same = bool(np.array_equal(y.view(np.uint32), want.view(np.uint32)))
    best = 1e9
    for _ in range(3):
        t = time.perf_counter(); ctx.complex_forward(x, out=y); best = min(best, time.perf_counter() - t)
    print(f"CKFFT_B200_NT_COPY={lv}: {16 * n * batch / best / 1e9:6.1f} GB/s  bit-identical {same}", flush=True)
