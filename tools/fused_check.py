#!/usr/bin/env python3
"""Check + timing of the FUSED distributed transform (peer stores over NVLink, no collective on the data path).
    python tools/fused_check.py [log2n ...]                                   one GPU (world 1: every "peer" is the GPU itself)
    python -m torch.distributed.run --nproc-per-node P --master-addr 127.0.0.1 tools/fused_check.py [log2n ...]
Options: --passes 3|4 (force a pass count), --no-nccl (skip the comparison with the NCCL six-step), --push / --pull (exchange kernel first,
or a first pass that pulls its tiles from the peers; default: pull for four-pass layouts).
Parity: analytic input (complex exponentials + an impulse, closed-form spectrum), round trip, Parseval, the full fp64
spectrum for N <= 2^24, and element-wise agreement with the NCCL six-step of ckfft_b200.distributed on the same input."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ckfft_b200.distributed import DistributedFFT, FusedDistributedFFT  # noqa: E402


def main():
    args = sys.argv[1:]
    prefer = 0
    if "--passes" in args:
        i = args.index("--passes"); prefer = int(args[i + 1]); del args[i:i + 2]
    with_nccl = "--no-nccl" not in args
    pull = False if "--push" in args else (True if "--pull" in args else None)
    args = [a for a in args if not a.startswith("--")]
    multi = "RANK" in os.environ
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if multi:
        dist.init_process_group("nccl", device_id=dev)

    def allsum(t):
        if multi:
            dist.all_reduce(t)
        return t

    def sync():
        if multi:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    sizes = [int(a) for a in args] or [14, 20, 24]
    ok_all = True
    for lg in sizes:
        n = 1 << lg
        per = n // world
        d = FusedDistributedFFT(n, prefer_passes=prefer, pull=pull)
        lay = d.layout
        pull_now = bool(lay.pull)
        idx = torch.arange(rank * per, (rank + 1) * per, device=dev, dtype=torch.float64)
        freqs, amps, n0 = [3, n // 3 + 1, n - 7], [1.0, 0.5, 0.25], 5
        x = torch.zeros(per, dtype=torch.complex128, device=dev)
        for f, a in zip(freqs, amps):
            ph = 2.0 * np.pi * ((idx * f) % n) / n
            x += a * torch.complex(torch.cos(ph), torch.sin(ph))
        if rank * per <= n0 < (rank + 1) * per:
            x[n0 - rank * per] += 1.0
        xs = x.to(torch.complex64)
        del x
        y = d.forward(xs)
        ph = -2.0 * np.pi * ((idx * n0) % n) / n
        want = torch.complex(torch.cos(ph), torch.sin(ph))
        for f, a in zip(freqs, amps):
            if rank * per <= f < (rank + 1) * per:
                want[f - rank * per] += a * n
        t = allsum(torch.stack([torch.linalg.vector_norm(y.to(torch.complex128) - want) ** 2, torch.linalg.vector_norm(want) ** 2]))
        err_analytic = float(torch.sqrt(t[0] / t[1]))
        del want, ph, idx, xs
        g = torch.Generator(device=dev).manual_seed(1 + rank)
        noise = torch.view_as_complex(torch.empty((per, 2), dtype=torch.float32, device=dev).uniform_(-1, 1, generator=g))
        yn = d.forward(noise).clone()
        zn = d.inverse(yn)
        num = torch.linalg.vector_norm(zn / n - noise) ** 2
        den = torch.linalg.vector_norm(noise) ** 2
        e_out = torch.linalg.vector_norm(yn) ** 2 / n
        t = allsum(torch.stack([num, den, e_out]).double())
        err_rt = float(torch.sqrt(t[0] / t[1])); pars = float(abs(t[2] - t[1]) / t[1])
        d.check()
        err_full = None
        if lg <= 24:
            nr, yr = torch.view_as_real(noise).contiguous(), torch.view_as_real(yn).contiguous()
            if multi:
                parts = [torch.empty_like(nr) for _ in range(world)] if rank == 0 else None
                dist.gather(nr, parts, dst=0)
                yparts = [torch.empty_like(yr) for _ in range(world)] if rank == 0 else None
                dist.gather(yr, yparts, dst=0)
            else:
                parts, yparts = [nr], [yr]
            if rank == 0:
                full = torch.view_as_complex(torch.cat(parts)).cpu().numpy().astype(np.complex128)
                ref = np.fft.fft(full)
                got = torch.view_as_complex(torch.cat(yparts)).cpu().numpy()
                err_full = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
        err_nccl = None
        ms_nccl = None
        if with_nccl and (multi or True):
            dn = DistributedFFT(n) if multi else None
            if dn is not None:
                yref = dn.forward(noise)
                t = allsum(torch.stack([torch.linalg.vector_norm(yn - yref) ** 2, torch.linalg.vector_norm(yref) ** 2]).double())
                err_nccl = float(torch.sqrt(t[0] / t[1]))
                del yref
                for _ in range(2):
                    dn.forward(noise)
                sync()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    dn.forward(noise)
                e1.record()
                sync()
                ms = torch.tensor([e0.elapsed_time(e1) / 5], device=dev, dtype=torch.float64)
                if multi:
                    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
                ms_nccl = float(ms.item())
                dn.close()
        if pull_now:
            d.input.copy_(noise)                 # timed the zero-copy way: the input lives in the plan's peer-visible array
            noise = d.input
        for _ in range(3):
            d.forward(noise)
        sync()
        iters = 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            d.forward(noise)
        e1.record()
        sync()
        d.check()
        ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev, dtype=torch.float64)
        if multi:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        phases = d.profile(noise)
        sync()
        if rank == 0:
            print("   phases (rank 0, ms): " + "  ".join(f"{k} {v:.3f}" for k, v in phases), flush=True)
        if rank == 0:
            tol = 1e-6 * lg
            ms = float(ms.item())
            fmt = lambda v: "None" if v is None else format(v, ".2e")
            print(f"FUSED{'(pull)' if pull_now else '(push)'} N=2^{lg} P={world} passes={lay.passes} ({lay.la}x{lay.lb} . {lay.lc}x{lay.ld}): analytic {err_analytic:.2e} "
                  f"roundtrip {err_rt:.2e} parseval {pars:.1e} full-fp64 {fmt(err_full)} vs-nccl-six-step {fmt(err_nccl)} (tol {tol:.1e}) | "
                  f"{ms:.3f} ms = {16.0 * n / world / ms / 1e6:.1f} GB/s per GPU algorithmic, {5.0 * n * lg / ms / 1e6:.0f} GFLOP/s, "
                  f"exchange {d.bytes_per_exchange() / 1e6:.1f} MB/GPU x3"
                  + (f" | NCCL six-step {ms_nccl:.3f} ms ({ms_nccl / ms:.2f}x)" if ms_nccl else ""), flush=True)
            ok = err_analytic <= tol and err_rt <= tol and pars < 1e-5 and (err_full is None or err_full <= tol) and \
                (err_nccl is None or err_nccl <= tol)
            if not ok:
                print("FAILED", flush=True)
            ok_all = ok_all and ok
        d.close()
        del noise, yn, zn
        torch.cuda.empty_cache()
    if multi:
        dist.destroy_process_group()
    if rank == 0 and not ok_all:
        sys.exit(1)


if __name__ == "__main__":
    main()
