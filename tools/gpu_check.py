#!/usr/bin/env python3
"""Quick on-GPU parity + throughput sweep (development tool; the judged numbers come from bench.py)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ckfft_b200 as ck  # noqa: E402
import oracle  # noqa: E402


def rel(a, b):
    a = np.asarray(a).astype(np.complex128); b = np.asarray(b).astype(np.complex128)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-30))


def parity():
    rng = np.random.default_rng(7)
    worst = 0.0
    for lg in (range(0, 16) if len(sys.argv) <= 1 else []):
        n = 1 << lg
        nb = 37 if n <= 4096 else 5
        ctx = ck.Context(max(n, 2), ck.BOTH)
        orc = oracle.Restatement(max(n, 2), 3)
        x = (rng.uniform(-1, 1, (nb, n)) + 1j * rng.uniform(-1, 1, (nb, n))).astype(np.complex64)
        line = [f"n={n:6d}"]
        if n <= 16384:
            for inv in (False, True):
                want = orc.complex(x, inv)
                got_h = ctx.complex_inverse(x) if inv else ctx.complex_forward(x)
                xd = torch.from_numpy(x).cuda()
                got_d = (ctx.complex_inverse(xd) if inv else ctx.complex_forward(xd)).cpu().numpy()
                e1, e2 = rel(got_h, want), rel(got_d, want)
                line.append(f"c2c{'i' if inv else 'f'} {e1:.1e}/{e2:.1e}")
                worst = max(worst, e1 / max(1, lg), e2 / max(1, lg))
        xr = np.ascontiguousarray(x.real)
        want = orc.real_forward(xr)
        got = ctx.real_forward(xr)
        got_d = ctx.real_forward(torch.from_numpy(xr).cuda()).cpu().numpy()
        e1, e2 = rel(got, want), rel(got_d, want)
        line.append(f"r2c {e1:.1e}/{e2:.1e}")
        worst = max(worst, e1 / max(1, lg), e2 / max(1, lg))
        want2 = orc.real_inverse(want, n)
        got2 = ctx.real_inverse(want, n)
        got2_d = ctx.real_inverse(torch.from_numpy(want).cuda(), n).cpu().numpy()
        e1, e2 = rel(got2, want2), rel(got2_d, want2)
        line.append(f"c2r {e1:.1e}/{e2:.1e}")
        worst = max(worst, e1 / max(1, lg), e2 / max(1, lg))
        print(" ".join(line), flush=True)
        ctx.close(); orc.close()
    print(f"worst err/log2n = {worst:.2e} (budget 1e-6)", flush=True)
    return worst


def sweep(sizes, total_log2=27, iters=10):
    peak = 6546.9
    print(f"{'n':>7} {'kind':>4} {'batch':>9} {'ms':>8} {'GB/s':>8} {'frac':>6} {'GFLOP/s':>9}")
    for n in sizes:
        batch = max(1, (1 << total_log2) // n)
        ctx = ck.Context(n, ck.BOTH)
        x = torch.empty((batch, n), dtype=torch.complex64, device="cuda")
        x.real.uniform_(-1, 1); x.imag.uniform_(-1, 1)
        out = torch.empty_like(x)
        kinds = ("c2c", "r2c", "c2r") + (("c2cp", "c2ci") if os.environ.get("CKFFT_SWEEP_LAYOUTS") and n <= 16384 else ())
        for kind in kinds:
            if kind == "c2cp":          # split-complex arrays
                xre = x.view(torch.float32).view(-1)[: batch * n].view(batch, n)
                xim = x.view(torch.float32).view(-1)[batch * n:].view(batch, n)
                ore = out.view(torch.float32).view(-1)[: batch * n].view(batch, n)
                oim = out.view(torch.float32).view(-1)[batch * n:].view(batch, n)
                f = lambda: ctx.complex_planar(xre, xim, False, (ore, oim))
                nbytes = 16 * n * batch
                flops = 5 * n * np.log2(n) * batch
            elif kind == "c2ci":        # in place
                lib = ck._lib.load()
                f = lambda: lib.CkFftComplexForwardBatchAsync(ctx.handle, n, out.data_ptr(), out.data_ptr(), batch, 0, 0,
                                                              torch.cuda.current_stream().cuda_stream)
                nbytes = 16 * n * batch
                flops = 5 * n * np.log2(n) * batch
            elif kind == "c2c":
                f = lambda: ctx.complex_forward(x, out)
                nbytes = 16 * n * batch
                flops = 5 * n * np.log2(n) * batch
            elif kind == "r2c":
                xr = x.view(torch.float32).view(-1)[: batch * n].view(batch, n)
                yo = out.view(-1)[: batch * (n // 2 + 1)].view(batch, n // 2 + 1)
                f = lambda: ctx.real_forward(xr, yo)
                nbytes = (4 * n + 8 * (n // 2 + 1)) * batch
                flops = 2.5 * n * np.log2(n) * batch
            else:
                yi = x.view(-1)[: batch * (n // 2 + 1)].view(batch, n // 2 + 1)
                xo = out.view(torch.float32).view(-1)[: batch * n].view(batch, n)
                f = lambda: ctx.real_inverse(yi, n, xo)
                nbytes = (4 * n + 8 * (n // 2 + 1)) * batch
                flops = 2.5 * n * np.log2(n) * batch
            for _ in range(3):
                f()
            torch.cuda.synchronize()
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
            for a, b in evs:
                a.record(); f(); b.record()
            torch.cuda.synchronize()
            ms = float(np.median([a.elapsed_time(b) for a, b in evs]))
            gbs = nbytes / ms / 1e6
            print(f"{n:7d} {kind:>4} {batch:9d} {ms:8.3f} {gbs:8.1f} {gbs / peak:6.3f} {flops / ms / 1e6:9.1f}", flush=True)
        ctx.close()
        del x, out


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    t = time.time()
    w = parity()
    print(f"parity took {time.time() - t:.1f}s")
    sizes = [1 << k for k in range(4, 23)]
    if len(sys.argv) > 1:
        sizes = [int(a) for a in sys.argv[1:]]
    sweep(sizes)
    sys.exit(0 if w < 1e-6 else 1)
