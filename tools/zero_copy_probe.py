#!/usr/bin/env python3
"""Does running the kernel directly on pinned (mapped) host memory beat the staged copy pipeline over PCIe?"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ckfft_b200 as ck
from ckfft_b200 import _lib
lib = _lib.load()
n, batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1024, 1 << 19
hx = torch.empty((batch, n), dtype=torch.complex64, pin_memory=True); torch.view_as_real(hx).uniform_(-1, 1)
hy = torch.empty((batch, n), dtype=torch.complex64, pin_memory=True)
ctx = ck.Context(n, ck.BOTH)
nx, ny = hx.numpy(), hy.numpy()
ctx.complex_forward(nx, ny)
ref = ny.copy()
t = time.perf_counter()
for _ in range(3): ctx.complex_forward(nx, ny)
dt = (time.perf_counter() - t) / 3
print("staged pipeline :", round(16 * n * batch / dt / 1e9, 1), "GB/s")
hy.zero_()
s = torch.cuda.current_stream().cuda_stream
ok = lib.CkFftComplexForwardBatchAsync(ctx.handle, n, hx.data_ptr(), hy.data_ptr(), batch, 0, 0, s); torch.cuda.synchronize()
print("zero-copy ok", ok, "equal", bool(np.array_equal(ref.view(np.uint32), ny.view(np.uint32))))
t = time.perf_counter()
for _ in range(3):
    lib.CkFftComplexForwardBatchAsync(ctx.handle, n, hx.data_ptr(), hy.data_ptr(), batch, 0, 0, s)
torch.cuda.synchronize()
dt = (time.perf_counter() - t) / 3
print("zero-copy kernel:", round(16 * n * batch / dt / 1e9, 1), "GB/s")
