#!/usr/bin/env python3
"""Multi-GPU check + timing of the distributed six-step transform (run under torchrun):
    python -m torch.distributed.run --nproc-per-node P --master-addr 127.0.0.1 tools/dist_check.py [log2n ...]
Parity: analytic inputs (a few complex exponentials + an impulse, whose spectrum is known in closed form), the
round trip inverse(forward(x)) = N x, Parseval; for N <= 2^24 also the full fp64 spectrum on rank 0."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ckfft_b200.distributed import DistributedFFT  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    sizes = [int(a) for a in sys.argv[1:]] or [20, 24, 26]
    for lg in sizes:
        n = 1 << lg
        per = n // world
        d = DistributedFFT(n)
        idx = torch.arange(rank * per, (rank + 1) * per, device=dev, dtype=torch.float64)
        # x = sum_f a_f exp(2 pi i f n / N) + impulse at n0  ->  X[k] = N a_f at k = f, plus exp(-2 pi i k n0 / N)
        freqs = [3, n // 3 + 1, n - 7]
        amps = [1.0, 0.5, 0.25]
        x = torch.zeros(per, dtype=torch.complex128, device=dev)
        for f, a in zip(freqs, amps):
            ph = 2.0 * np.pi * ((idx * f) % n) / n
            x += a * torch.complex(torch.cos(ph), torch.sin(ph))
        n0 = 5
        if rank * per <= n0 < (rank + 1) * per:
            x[n0 - rank * per] += 1.0
        g = torch.Generator(device=dev).manual_seed(1 + rank)
        noise = torch.view_as_complex(torch.empty((per, 2), dtype=torch.float32, device=dev).uniform_(-1, 1, generator=g))
        xs = x.to(torch.complex64)
        y = d.forward(xs)
        # closed form
        kk = torch.arange(rank * per, (rank + 1) * per, device=dev, dtype=torch.float64)
        ph = -2.0 * np.pi * ((kk * n0) % n) / n
        want = torch.complex(torch.cos(ph), torch.sin(ph))
        for f, a in zip(freqs, amps):
            if rank * per <= f < (rank + 1) * per:
                want[f - rank * per] += a * n
        num = torch.linalg.vector_norm((y.to(torch.complex128) - want)) ** 2
        den = torch.linalg.vector_norm(want) ** 2
        t = torch.stack([num, den]); dist.all_reduce(t)
        err_analytic = float(torch.sqrt(t[0] / t[1]))
        # round trip + Parseval on noise
        yn = d.forward(noise)
        zn = d.inverse(yn)
        num = torch.linalg.vector_norm((zn / n - noise).to(torch.complex128)) ** 2
        den = torch.linalg.vector_norm(noise.to(torch.complex128)) ** 2
        e_out = torch.linalg.vector_norm(yn.to(torch.complex128)) ** 2 / n
        t = torch.stack([num, den, e_out]); dist.all_reduce(t)
        err_rt = float(torch.sqrt(t[0] / t[1])); pars = float(abs(t[2] - t[1]) / t[1])
        err_full = None
        if lg <= 24:
            nr, yr = torch.view_as_real(noise).contiguous(), torch.view_as_real(yn).contiguous()
            parts = [torch.empty_like(nr) for _ in range(world)] if rank == 0 else None
            dist.gather(nr, parts, dst=0)
            yparts = [torch.empty_like(yr) for _ in range(world)] if rank == 0 else None
            dist.gather(yr, yparts, dst=0)
            if rank == 0:
                full = torch.view_as_complex(torch.cat(parts)).cpu().numpy().astype(np.complex128)
                ref = np.fft.fft(full)
                got = torch.view_as_complex(torch.cat(yparts)).cpu().numpy()
                err_full = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
        # timing
        for _ in range(2):
            d.forward(noise)
        dist.barrier(device_ids=[local]); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 5
        e0.record()
        for _ in range(iters):
            d.forward(noise)
        e1.record()
        dist.barrier(device_ids=[local]); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev, dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            tol = 1e-6 * lg
            ms = float(ms.item())
            print(f"N=2^{lg} P={world}: analytic {err_analytic:.2e} roundtrip {err_rt:.2e} parseval {pars:.1e} "
                  f"full-fp64 {err_full if err_full is None else format(err_full, '.2e')} (tol {tol:.1e}) | {ms:.3f} ms "
                  f"= {16.0 * n / world / ms / 1e6:.1f} GB/s per GPU algorithmic, {5.0 * n * lg / ms / 1e6:.0f} GFLOP/s, "
                  f"exchange {d.bytes_per_exchange() / 1e6:.1f} MB/GPU x3", flush=True)
            assert err_analytic <= tol and err_rt <= tol and pars < 1e-5 and (err_full is None or err_full <= tol)
        d.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
