#!/usr/bin/env python3
"""bench.py -- the headline benchmark of BASELINE.json:
"Batched fp32 C2C FFT HBM GB/s + 5N*log2N GFLOP/s at N=1024, 1/2/4/8 B200".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

One "step" = one pass of the hot path over one batch: CkFftComplexForward on 2^20 transforms of
N=1024 (config[1] of BASELINE.json; 8.59 GB in + 8.59 GB out per GPU, far larger than the 126 MB L2,
so no L2 flush is needed between iterations).  With N>1 GPUs (launched under torchrun, one rank per
GPU) every rank runs the same per-GPU batch (weak scaling, independent transforms, no collective on
the data path); `value` is the work of all ranks divided by the slowest rank's device time.

One JSON line on stdout (rank 0):
  value / ms_per_step   device-resident throughput, CUDA events on the launching stream, max over ranks
  e2e                   same metric through the C ABI with pinned HOST buffers (CkFftComplexForwardBatch:
                        chunked H2D -> kernel -> D2H inside the timed region), with its own roofline: the
                        concurrent host<->device copy rate of this box measured in the same run (tools/pcie_peak.py)
  e2e_pageable          the same call on malloc'ed (pageable) host arrays -- what a drop-in caller hands over
  e2e_multi             the same work through ONE call of the multi-device scheduler (CkFftComplexForwardBatchMulti)
                        from one process (rank 0) over all N GPUs
  roofline              algorithmic bytes / kernel time vs the measured HBM copy peak
  cpu_baseline          the reference ckfft (oracle/_ref, compiled unmodified) on this box's host cores
  clocks                nvidia-smi samples taken while the steps ran
  secondary             the other BASELINE configs under the same clock: config 3 (R2C / C2R N=4096 x 2^18 frames),
                        config 4 (size sweep N=2^4..2^20, batch = 2^28/N per GPU), large real transforms, and -- with
                        N > 1 GPUs -- config 5 (one 2^30-point transform over the N GPUs, fused distributed path) with
                        its phases, per-phase NVLink rate and an analytic + Parseval check
`--impl reference` times the reference CPU implementation instead (rank 0 only), same metric/config.
Other workloads (`--workload r2c4096|c2r4096|stft4096|c2c<N>|sweep|dist<log2N>`) exist for the parity configs, the size
sweep and the distributed single transform (config 5);
the default is the judged one.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Batched fp32 C2C FFT HBM GB/s + 5N*log2N GFLOP/s at N=1024"
FALLBACK_HBM_GBS = 6650.0     # /opt/skills/guides/B200_PROFILING.md, used only if MEASURED_PEAKS.json is absent


# ---------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------
def workload_spec(name: str):
    """-> dict(kind, n, batch, bytes_per_transform, flops_per_transform, description)"""
    if name == "c2c1024":
        n, batch, kind = 1024, 1 << 20, "c2c"
    elif name == "r2c4096":
        n, batch, kind = 4096, 1 << 18, "r2c"
    elif name == "c2r4096":
        n, batch, kind = 4096, 1 << 18, "c2r"
    elif name == "stft4096":
        n, batch, kind = 4096, 1 << 18, "pow"       # audio front end: window + R2C + power spectrum
    elif name.startswith("c2c"):
        n, kind = int(name[3:]), "c2c"
        batch = max(1, (1 << 28) // n)          # config 4: constant 2 GiB in + 2 GiB out
    else:
        raise SystemExit(f"unknown workload {name}")
    if kind == "c2c":
        nbytes, flops = 16 * n, 5.0 * n * np.log2(n)
    elif kind == "pow":
        nbytes, flops = 4 * n + 4 * (n // 2 + 1), 2.5 * n * np.log2(n)
    else:
        nbytes, flops = 4 * n + 8 * (n // 2 + 1), 2.5 * n * np.log2(n)
    desc = {"c2c": f"batched complex C2C forward N={n} x {batch} transforms fp32 per GPU",
            "r2c": f"batched real R2C N={n} x {batch} frames fp32 per GPU",
            "c2r": f"batched real C2R N={n} x {batch} frames fp32 per GPU",
            "pow": f"Hann window + R2C + power spectrum N={n} x {batch} frames fp32 per GPU (fused)"}[kind]
    return dict(name=name, kind=kind, n=n, batch=batch, bytes=nbytes, flops=flops, desc=desc)


def make_config(spec, world):
    """the `config` object both arms print (the driver compares them)"""
    return {"workload": spec["desc"], "n": spec["n"], "kind": spec["kind"], "batch_per_gpu": spec["batch"],
            "bytes_per_transform": spec["bytes"], "l2": "inputs larger than L2 (no flush needed)",
            "parallelism": f"batch-sharded x{world}, no collective"}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, device copy)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic(name):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f).get(name)
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index: int):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(gpu_index)], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        rows = []
        with open(self.tmp.name) as f:
            for line in f:
                parts = [p.strip() for p in line.split(",")]
                if len(parts) >= 7:
                    try:
                        rows.append((float(parts[0]), float(parts[1]), float(parts[2]), parts[3:7]))
                    except ValueError:
                        pass
        os.unlink(self.tmp.name)
        if not rows:
            return out
        # "under load": samples drawing more than half of the largest power seen
        pmax = max(r[2] for r in rows)
        loaded = [r for r in rows if r[2] >= 0.5 * pmax] or rows
        out["sm_mhz"] = statistics.median(r[0] for r in loaded)
        out["sm_max_mhz"] = max(r[1] for r in rows)
        out["power_w_max"] = pmax
        out["samples"] = len(loaded)
        out["reasons"] = [n for i, n in enumerate(self.NAMES) if any(r[3][i].lower() == "active" for r in loaded)]
        return out


# ---------------------------------------------------------------------------------------------
# CPU reference (oracle/_ref): the cpu_baseline leg and the --impl reference arm
# ---------------------------------------------------------------------------------------------
def host_threads() -> int:
    """threads the CPU arm may use: the cores this process is allowed on (torchrun exports
    OMP_NUM_THREADS=1, so the count is passed to OpenMP explicitly)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_run(spec, n_transforms, threads, repeats=1, seed=1234, min_seconds=0.0):
    """Time the reference ckfft (unmodified, compiled by oracle/build.py) on `n_transforms` transforms.
    threads <= 0: all host threads.  Returns (best seconds, kind)."""
    import oracle

    if threads <= 0:
        threads = host_threads()

    n, kind = spec["n"], spec["kind"]
    rng = np.random.default_rng(seed)

    def rand_c64(rows, cols):
        return (rng.random((rows, cols, 2), dtype=np.float32) * np.float32(2) - np.float32(1)).view(np.complex64)[..., 0]
    if oracle.reference_available():
        impl, tag = oracle.Reference(n, 3), "reference"
    else:   # the compiled reference did not travel: time our bit-identical C port instead
        impl, tag = oracle.Restatement(n, 3), "port"
    if kind == "c2c":
        x = rand_c64(n_transforms, n)
        out = np.empty_like(x)
        run = (lambda: impl.complex(x, False, threads, out)) if tag == "reference" else (lambda: impl.complex(x, False))
    elif kind == "r2c":
        x = rng.random((n_transforms, n), dtype=np.float32) * np.float32(2) - np.float32(1)
        out = np.empty((n_transforms, n // 2 + 1), np.complex64)
        run = (lambda: impl.real_forward(x, threads, out)) if tag == "reference" else (lambda: impl.real_forward(x))
    else:
        x = rand_c64(n_transforms, n // 2 + 1)
        out = np.empty((n_transforms, n), np.float32)
        run = (lambda: impl.real_inverse(x, n, threads, out)) if tag == "reference" else (lambda: impl.real_inverse(x, n))
    run()   # touch pages, warm caches
    best = float("inf")
    spent, done = 0.0, 0
    while done < repeats or (spent < min_seconds and done < 200):
        t = time.perf_counter()
        run()
        dt = time.perf_counter() - t
        best = min(best, dt)
        spent += dt
        done += 1
    impl.close()
    return best, tag


def cpu_baseline(spec):
    """Bounded sample on the host cores: all threads and one thread, best of 3."""
    cores = host_threads()
    sample_all = min(spec["batch"], 1 << 18)
    sample_one = min(spec["batch"], 1 << 14)
    t_all, tag = cpu_reference_run(spec, sample_all, 0, repeats=3, min_seconds=6.0)
    t_one, _ = cpu_reference_run(spec, sample_one, 1, repeats=3, min_seconds=3.0)
    gbs_all = spec["bytes"] * sample_all / t_all / 1e9
    gbs_one = spec["bytes"] * sample_one / t_one / 1e9
    return {"value": round(gbs_all, 3), "unit": "GB/s", "cores": cores, "kind": tag,
            "sample": f"{sample_all} of {spec['batch']} transforms (same seeded uniform(-1,1) data), OpenMP over "
                      f"the batch on {cores} host threads, one shared context, best pass of >= 6 s of repeats",
            "gflops": round(spec["flops"] * sample_all / t_all / 1e9, 2),
            "us_per_transform": round(t_all / sample_all * 1e6, 4),
            "single_core": {"value": round(gbs_one, 3), "unit": "GB/s", "us_per_transform": round(t_one / sample_one * 1e6, 3),
                            "sample": f"{sample_one} transforms, 1 thread"}}


def run_reference_arm(args, spec, rank):
    """--impl reference: the reference's own CPU implementation, all host threads, rank 0 only."""
    if rank != 0:
        return
    cores = host_threads()
    # size one step so that (steps + warmup) steps take about a minute in total
    probe = 1 << 12
    t_probe, tag = cpu_reference_run(spec, probe, 0, repeats=2)
    per_transform = t_probe / probe
    budget = 60.0 / max(1, args.steps + args.warmup)
    sample = int(min(spec["batch"], max(1 << 10, budget / per_transform)))
    sample = 1 << (sample.bit_length() - 1)
    import oracle

    n = spec["n"]
    impl = oracle.Reference(n, 3) if oracle.reference_available() else oracle.Restatement(n, 3)
    rng = np.random.default_rng(1234)
    x = (rng.random((sample, n, 2), dtype=np.float32) * np.float32(2) - np.float32(1)).view(np.complex64)[..., 0]
    if spec["kind"] == "r2c":
        xin, out = np.ascontiguousarray(x.real), np.empty((sample, n // 2 + 1), np.complex64)
        step = (lambda: impl.real_forward(xin, cores, out)) if tag == "reference" else (lambda: impl.real_forward(xin))
    elif spec["kind"] == "c2r":
        xin, out = np.ascontiguousarray(x[:, : n // 2 + 1]), np.empty((sample, n), np.float32)
        step = (lambda: impl.real_inverse(xin, n, cores, out)) if tag == "reference" else (lambda: impl.real_inverse(xin, n))
    else:
        out = np.empty_like(x)
        step = (lambda: impl.complex(x, False, cores, out)) if tag == "reference" else (lambda: impl.complex(x, False))
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    impl.close()
    gbs = spec["bytes"] * sample / dt / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": round(gbs, 3), "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "gflops": round(spec["flops"] * sample / dt / 1e9, 2),
        "config": make_config(spec, args.gpus),
        "cpu_baseline": {"value": round(gbs, 3), "unit": "GB/s", "cores": cores, "kind": tag,
                         "sample": f"bounded sample of the workload: {sample} of {spec['batch']} transforms per step, OpenMP over "
                                   f"the batch on {cores} host threads"},
        "e2e": {"value": round(gbs, 3), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# device-timed building blocks (shared by the main line, `secondary` and the stand-alone workloads)
# ---------------------------------------------------------------------------------------------
def reduce_max(value, world, dev):
    if world <= 1:
        return float(value)
    import torch
    import torch.distributed as dist

    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(value, world, dev):
    if world <= 1:
        return float(value)
    import torch
    import torch.distributed as dist

    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def time_steps(step, steps, warmup, barrier, world, dev):
    """mean ms per step: CUDA events on torch's current stream (the stream the library launches on), warm-up first,
    barrier + synchronize on both sides, max over ranks"""
    import torch

    for _ in range(warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    return reduce_max(e0.elapsed_time(e1) / steps, world, dev)


def measure_sweep(world, dev, barrier, steps=5, lgs=range(4, 21)):
    """BASELINE config 4: C2C forward, N = 2^4 .. 2^20, batch = 2^28 / N per GPU (2 GiB in + 2 GiB out per launch)."""
    import torch

    import ckfft_b200 as ck

    peak, _ = measured_peak()
    total = 1 << 28
    x = torch.view_as_complex(torch.empty((total, 2), dtype=torch.float32, device=dev).uniform_(-1, 1))
    y = torch.empty_like(x)
    rows = []
    for lg in lgs:
        n = 1 << lg
        ctx = ck.Context(n, ck.FORWARD)
        xv, yv = x.view(total // n, n), y.view(total // n, n)
        ms = time_steps(lambda: ctx.complex_forward(xv, yv), steps, 3, barrier, world, dev)
        gbs = 16.0 * total * world / ms / 1e6
        rows.append({"n": n, "ms": round(ms, 4), "gbs": round(gbs, 1), "frac": round(gbs / world / peak, 4),
                     "gflops": round(5.0 * total * lg * world / ms / 1e6, 1)})
        ctx.close()
    del x, y
    torch.cuda.empty_cache()
    return rows


def measure_real_large(world, dev, barrier, steps=5, lgs=range(17, 22)):
    """real transforms beyond the single-pass limit (n = 2^17 .. 2^21, 2^27 samples per launch): R2C and C2R"""
    import torch

    import ckfft_b200 as ck

    peak, _ = measured_peak()
    total = 1 << 27
    rows = []
    for lg in lgs:
        n = 1 << lg
        batch = total // n
        ctx = ck.Context(n, ck.BOTH)
        xr = torch.empty((batch, n), dtype=torch.float32, device=dev).uniform_(-1, 1)
        yc = torch.empty((batch, n // 2 + 1), dtype=torch.complex64, device=dev)
        nbytes = (4 * n + 8 * (n // 2 + 1)) * batch
        ms_f = time_steps(lambda: ctx.real_forward(xr, yc), steps, 3, barrier, world, dev)
        ms_i = time_steps(lambda: ctx.real_inverse(yc, n, xr), steps, 3, barrier, world, dev)
        rows.append({"n": n, "r2c_ms": round(ms_f, 4), "r2c_frac": round(nbytes / ms_f / 1e6 / peak, 4),
                     "c2r_ms": round(ms_i, 4), "c2r_frac": round(nbytes / ms_i / 1e6 / peak, 4)})
        ctx.close()
        del xr, yc
    torch.cuda.empty_cache()
    return rows


def measure_config3(name, world, dev, barrier, steps=20, warmup=5):
    """BASELINE config 3 (and the fused audio front end): N = 4096 real, 2^18 frames per GPU"""
    import torch

    import ckfft_b200 as ck

    spec = workload_spec(name)
    n, batch, kind = spec["n"], spec["batch"], spec["kind"]
    peak, _ = measured_peak()
    ctx = ck.Context(n, ck.BOTH)
    g = torch.Generator(device=dev).manual_seed(1235)
    if kind == "c2r":
        x = torch.view_as_complex(torch.empty((batch, n // 2 + 1, 2), dtype=torch.float32, device=dev).uniform_(-1, 1, generator=g))
        y = torch.empty((batch, n), dtype=torch.float32, device=dev)
        step = lambda: ctx.real_inverse(x, n, y)   # noqa: E731
    else:
        x = torch.empty((batch, n), dtype=torch.float32, device=dev).uniform_(-1, 1, generator=g)
        if kind == "pow":
            wnd = torch.hann_window(n, periodic=True, dtype=torch.float32, device=dev)
            y = torch.empty((batch, n // 2 + 1), dtype=torch.float32, device=dev)
            step = lambda: ctx.real_forward_power(x, wnd, y)   # noqa: E731
        else:
            y = torch.empty((batch, n // 2 + 1), dtype=torch.complex64, device=dev)
            step = lambda: ctx.real_forward(x, y)   # noqa: E731
    ms = time_steps(step, steps, warmup, barrier, world, dev)
    gbs = spec["bytes"] * batch * world / ms / 1e6
    ctx.close()
    del x, y
    torch.cuda.empty_cache()
    return {"workload": spec["desc"], "value": round(gbs, 1), "unit": "GB/s", "ms_per_step": round(ms, 4), "steps": steps,
            "frac": round(gbs / world / peak, 4), "gflops": round(spec["flops"] * batch * world / ms / 1e6, 1),
            "bytes_per_transform": spec["bytes"]}


def measure_dist(lg, rank, world, dev, barrier, steps=20, warmup=3):
    """BASELINE config 5: ONE complex forward transform of N = 2^lg points spread over the GPUs of the box (strong
    scaling) through the fused distributed path (peer stores / TMA reads over NVLink, no collective on the data path).
    Timed on random data; checked on a closed-form signal (exponentials + an impulse) and by Parseval."""
    import torch
    import torch.distributed as dist

    from ckfft_b200.distributed import FusedDistributedFFT

    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from analytic import analytic_error, analytic_signal

    n = 1 << lg
    per = n // world
    d = FusedDistributedFFT(n)
    # ---- correctness first: closed-form spectrum, every bin of every rank ----
    freqs, amps, n0 = [3, n // 3 + 1, n - 7], [1.0, 0.5, 0.25], 5
    xa = analytic_signal(n, rank * per, per, dev, freqs, amps, n0)
    if d.input is not None:
        d.input.copy_(xa)
        xa = d.input
    ya = d.forward(xa)
    d.check()
    num, den = analytic_error(ya, n, rank * per, dev, freqs, amps, n0)
    nd = torch.stack([num, den])
    if world > 1:
        dist.all_reduce(nd)
    analytic_err = float(torch.sqrt(nd[0] / nd[1]).item())
    del xa, ya
    torch.cuda.empty_cache()
    # ---- timing on random data ----
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = torch.view_as_complex(torch.empty((per, 2), dtype=torch.float32, device=dev).uniform_(-1, 1, generator=g))
    if d.input is not None:          # pull layouts: the input lives in the plan's peer-visible array (no local copy)
        d.input.copy_(x)
        x = d.input
    out = {}
    ms = time_steps(lambda: out.__setitem__("y", d.forward(x)), steps, warmup, barrier, world, dev)
    d.check()
    y = out["y"]
    e = torch.stack([torch.linalg.vector_norm(x).double() ** 2, torch.linalg.vector_norm(y).double() ** 2 / n])
    if world > 1:
        dist.all_reduce(e)
    parseval = abs(float(e[1] - e[0])) / float(e[0])
    phases = d.profile(x)
    barrier()
    lay = d.layout
    exch = d.bytes_per_exchange()
    peak, peak_src = measured_peak()
    per_gpu = 16.0 * n / world / ms / 1e6
    tol = 1e-6 * lg
    res = {
        "workload": f"one complex forward FFT of 2^{lg} points over {world} GPU(s), fused distributed path",
        "ms": round(ms, 4), "steps": steps, "value": round(16.0 * n / ms / 1e6, 1), "unit": "GB/s (algorithmic 16*N bytes / time, all GPUs)",
        "gflops": round(5.0 * n * lg / ms / 1e6, 1),
        "layout": f"n1 = {lay.la}x{lay.lb}, n2 = {lay.lc}x{lay.ld}, {lay.passes} passes, {'pull' if lay.pull else 'push'}",
        "frac_hbm_per_gpu": round(per_gpu / peak, 4),
        "exchange_bytes_per_gpu": exch,
        "phases_ms": {f"{i}:{k}": round(v, 4) for i, (k, v) in enumerate(phases)},
        # every phase whose kernel crosses NVLink moves (P-1)/P * 8N/P bytes out of (and into) each GPU
        "nvlink_gbs": {f"{i}:{k}": round(exch / v / 1e6, 1) for i, (k, v) in enumerate(phases)
                       if world > 1 and v > 0 and ("push" in k or "pull" in k or k == "exchange")},
        "nvlink_floor_ms": round(3 * exch / 900e9 * 1e3, 4) if world > 1 else None,             # three exchanges at the nominal 900 GB/s per direction
        "nvlink_floor_measured_ms": round(3 * exch / 770e9 * 1e3, 4) if world > 1 else None,    # ... at the 770 GB/s a peer copy reaches (B200_PROFILING.md)
        "frac_of_measured_floor": round(3 * exch / 770e9 * 1e3 / ms, 4) if world > 1 else None,
        "check": {"analytic_rel_rms": analytic_err, "parseval_rel": parseval, "tolerance": tol,
                  "ok": bool(analytic_err <= tol and parseval <= 1e-5)},
        "gpu_launches_per_step": len(phases),
    }
    d.close()
    torch.cuda.empty_cache()
    return res


def run_sweep(args, rank, world, local_rank, dev, barrier):
    rows = measure_sweep(world, dev, barrier, steps=max(3, min(args.steps, 20)))
    if rank == 0:
        peak, peak_src = measured_peak()
        mean = sum(r["gbs"] for r in rows) / len(rows)
        print(json.dumps({
            "metric": "Batched fp32 C2C FFT HBM GB/s, size sweep N=2^4..2^20 (mean over sizes)", "value": round(mean, 1), "unit": "GB/s",
            "n_gpus": world, "steps": max(3, min(args.steps, 20)), "warmup": 3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "size sweep, batch = 2^28/N complex transforms per GPU (2 GiB in + 2 GiB out), forward",
                       "parallelism": f"batch-sharded x{world}, no collective", "peak": peak, "peak_source": peak_src},
            "sweep": rows}), flush=True)


def run_dist(args, lg, rank, world, local_rank, dev, barrier):
    res = measure_dist(lg, rank, world, dev, barrier, steps=min(args.steps, 50), warmup=max(3, min(args.warmup, 10)))
    if rank == 0:
        if not res["check"]["ok"]:
            raise SystemExit(f"sanity check failed: {res['check']}")
        peak, peak_src = measured_peak()
        print(json.dumps({
            "metric": f"single 1-D complex FFT N=2^{lg}, natural order in and out, algorithmic 16*N bytes / time", "value": res["value"],
            "unit": "GB/s", "n_gpus": world, "steps": res["steps"], "warmup": max(3, min(args.warmup, 10)), "ms_per_step": res["ms"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "gflops": res["gflops"],
            "config": {"workload": res["workload"], "layout": res["layout"], "l2": "inputs larger than L2",
                       "parallelism": f"slices of N/{world}, peer stores over NVLink, 3 flag barriers, no collective"},
            "roofline": {"bound": "hbm", "achieved": round(res["frac_hbm_per_gpu"] * peak, 1), "peak": peak, "unit": "GB/s",
                         "frac": res["frac_hbm_per_gpu"], "traffic": None, "peak_source": peak_src,
                         "note": "per GPU; the exchange phases are NVLink-bound (see phases_ms / nvlink_gbs)"},
            "exchange_bytes_per_gpu": res["exchange_bytes_per_gpu"], "phases_ms": res["phases_ms"], "nvlink_gbs": res["nvlink_gbs"],
            "check": res["check"], "gpu_launches": res["gpu_launches_per_step"] * res["steps"]}), flush=True)


# ---------------------------------------------------------------------------------------------
# end to end: the reference-facing call on HOST buffers
# ---------------------------------------------------------------------------------------------
def measure_e2e(args, spec, ctx, rank, world, local_rank, dev, barrier, host_barrier):
    """-> (e2e, e2e_pageable, e2e_multi).  Wall clock around the synchronous C-ABI call, max over ranks."""
    import torch

    import ckfft_b200 as ck

    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import pcie_peak

    n, batch, kind = spec["n"], spec["batch"], spec["kind"]
    in_shape = (batch, n) if kind != "c2r" else (batch, n // 2 + 1)
    out_shape = (batch, n) if kind != "r2c" else (batch, n // 2 + 1)
    in_dtype = torch.float32 if kind == "r2c" else torch.complex64
    out_dtype = torch.float32 if kind == "c2r" else torch.complex64
    hx = torch.empty(in_shape, dtype=in_dtype, pin_memory=True)
    hy = torch.empty(out_shape, dtype=out_dtype, pin_memory=True)
    torch.view_as_real(hx).uniform_(-1, 1) if hx.is_complex() else hx.uniform_(-1, 1)
    nx, ny = hx.numpy(), hy.numpy()

    def call(c, a, b):
        if kind == "c2c":
            return c.complex_forward(a, b)
        if kind == "r2c":
            return c.real_forward(a, b)
        return c.real_inverse(a, n, b)

    def timed(fn, steps):
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        return reduce_max((time.perf_counter() - t0) / steps, world, dev)

    in_bytes, out_bytes = int(hx.numel() * hx.element_size()), int(hy.numel() * hy.element_size())
    # ---- pinned host arrays, one process per GPU ----
    dt = timed(lambda: call(ctx, nx, ny), args.e2e_steps)
    value = spec["bytes"] * batch * world / dt / 1e9
    e2e = {"value": round(value, 2), "unit": "GB/s", "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes,
           "ms_per_step": round(dt * 1e3, 3), "steps": args.e2e_steps,
           "gflops": round(spec["flops"] * batch * world / dt / 1e9, 1),
           "path": "CkFft*Batch on pinned host arrays: 32 MiB chunks, up to 4 in flight, H2D / kernel / D2H on three role streams; "
                   "wall clock around the synchronous call, max over ranks"}
    # ---- pageable host arrays (what a drop-in caller's malloc gives): bounded sample of the same workload ----
    sub = min(batch, 1 << 18)
    px = np.empty((sub,) + tuple(in_shape[1:]), dtype=nx.dtype)
    py = np.empty((sub,) + tuple(out_shape[1:]), dtype=ny.dtype)
    px[...] = nx[:sub]
    py[...] = 0
    dtp = timed(lambda: call(ctx, px, py), 2)
    same = bool(np.array_equal(py.view(np.uint32), ny[:sub].view(np.uint32)))
    os.environ["CKFFT_B200_PIN"] = "1"
    try:
        dtu = timed(lambda: call(ctx, px, py), 2)
    finally:
        os.environ.pop("CKFFT_B200_PIN", None)
    os.environ["CKFFT_B200_PAGEABLE_PIPE"] = "0"
    try:
        dtd = timed(lambda: call(ctx, px, py), 2)
    finally:
        os.environ.pop("CKFFT_B200_PAGEABLE_PIPE", None)
    pageable = {"value": round(spec["bytes"] * sub * world / dtp / 1e9, 2), "unit": "GB/s", "ms_per_step": round(dtp * 1e3, 3),
                "sample": f"{sub} of {batch} transforms per GPU in numpy (malloc) arrays",
                "path": "library-side staging: two teams of host threads copy chunks (non-temporal stores, host_copy.cpp) between the caller's "
                        "pageable arrays and pinned slots while the copy engines and the GPU work on the neighbouring chunks (api.cu, run_host_pageable)",
                "driver_staged": {"value": round(spec["bytes"] * sub * world / dtd / 1e9, 2), "ms_per_step": round(dtd * 1e3, 3),
                                  "note": "CKFFT_B200_PAGEABLE_PIPE=0: cudaMemcpyAsync straight on the pageable arrays (the driver "
                                          "stages them synchronously, H2D / D2H do not overlap)"},
                "with_registration": {"value": round(spec["bytes"] * sub * world / dtu / 1e9, 2), "ms_per_step": round(dtu * 1e3, 3),
                                      "note": "CKFFT_B200_PIN=1: the call page-locks the arrays for its duration (cudaHostRegister + "
                                              "unregister inside the timed region)"},
                "h2d_bytes_per_step": int(px.nbytes), "d2h_bytes_per_step": int(py.nbytes)}
    pageable["bit_identical_to_pinned_path"] = same
    del px, py
    # ---- one process, one call, all GPUs: the multi-device scheduler behind the C ABI (rank 0 drives, the others wait) ----
    multi = None
    host_barrier()
    if rank == 0:
        ndev = min(world, torch.cuda.device_count())
        mc = ck.MultiContext(n, ck.BOTH, devices=list(range(ndev)))
        fn = {"c2c": lambda: mc.complex_forward(nx, ny), "r2c": lambda: mc.real_forward(nx, ny),
              "c2r": lambda: mc.real_inverse(nx, n, ny)}[kind]
        fn()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            fn()
        dtm = (time.perf_counter() - t0) / args.e2e_steps
        multi = {"value": round(spec["bytes"] * batch / dtm / 1e9, 2), "unit": "GB/s", "devices": mc.devices,
                 "ms_per_step": round(dtm * 1e3, 3), "steps": args.e2e_steps, "transforms": batch,
                 "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes,
                 "path": f"CkFft*BatchMulti from ONE process: {batch} transforms in pinned host arrays sharded over {ndev} GPU(s) "
                         "(strong scaling of one batch), one host thread + H2D / kernel / D2H pipeline per device"}
        mc.close()
    host_barrier()
    barrier()
    # ---- the platform's copy ceiling, measured last (it overwrites the arrays): same pinned arrays, same chunking,
    # every rank copying at once ----
    link = pcie_peak.measure(local_rank, min(in_bytes, out_bytes, 4 << 30), 32 << 20, True, 2, hx, hy)
    link_sum = reduce_sum(link["both"], world, dev)
    e2e["roofline"] = {"bound": "host link (PCIe)", "achieved": e2e["value"], "peak": round(link_sum, 2), "unit": "GB/s",
                       "frac": round(e2e["value"] / link_sum, 4),
                       "peak_source": "plain cudaMemcpyAsync H2D + D2H at once on the same pinned arrays, 32 MiB chunks, "
                                      f"all {world} GPU(s) copying concurrently, bytes of both directions / time (tools/pcie_peak.py)",
                       "per_direction_gbs_rank0": {"h2d_alone": link["h2d"], "d2h_alone": link["d2h"], "both": link["both"]}}
    barrier()
    del hx, hy
    return e2e, pageable, multi


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2c1024")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    spec = workload_spec("c2c1024" if args.workload == "sweep" or args.workload.startswith("dist") else args.workload)

    if args.impl == "reference":
        run_reference_arm(args, spec, rank)
        return
    if world > 1 and "CKFFT_B200_HOST_THREADS" not in os.environ:
        # the library's pageable-array staging uses two teams of host threads per calling process (default: half the hardware
        # threads each, at most 8); N ranks share one host, so every rank takes its share of the cores
        os.environ["CKFFT_B200_HOST_THREADS"] = str(max(1, min(8, (os.cpu_count() or 16) // (2 * world))))

    import torch
    import torch.distributed as dist

    import ckfft_b200 as ck

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # NUMA placement of the host arrays of the end-to-end leg: torchrun lets the ranks float over all cores, so first-touch
    # puts a rank's pinned arrays on whatever node it happens to run on and the DMA of most GPUs crosses the socket link.
    # Each rank therefore binds itself to the CPUs next to its GPU (sysfs local_cpulist) while it allocates them
    # (CKFFT_BENCH_NUMA=0 disables it); the CPU baseline runs with the original affinity restored.
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import pcie_peak

    numa = pcie_peak.bind_to_gpu_node(local_rank) if os.environ.get("CKFFT_BENCH_NUMA", "1") != "0" else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    # Host-side barrier (gloo: the waiting ranks block in a socket poll).  Used where rank 0 works alone for a while --
    # the one-process multi-device call, the CPU baseline -- so that the other ranks neither keep a spinning NCCL
    # kernel on their GPU nor burn a host core each while they wait.
    host_group = dist.new_group(backend="gloo") if world > 1 else None

    def host_barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=host_group)

    if args.workload == "sweep":
        run_sweep(args, rank, world, local_rank, dev, barrier)
        if world > 1:
            dist.destroy_process_group()
        return
    if args.workload.startswith("dist"):
        run_dist(args, int(args.workload[4:] or 30), rank, world, local_rank, dev, barrier)
        if world > 1:
            dist.destroy_process_group()
        return
    n, batch, kind = spec["n"], spec["batch"], spec["kind"]
    sampler = ClockSampler(local_rank) if rank == 0 else None     # started early: nvidia-smi needs ~1 s to come up
    ctx = ck.Context(n, ck.BOTH)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    if kind == "c2c":
        x = torch.view_as_complex(torch.empty((batch, n, 2), dtype=torch.float32, device=dev).uniform_(-1, 1, generator=g))
        y = torch.empty_like(x)
        step = lambda: ctx.complex_forward(x, y)   # noqa: E731
    elif kind == "r2c":
        x = torch.empty((batch, n), dtype=torch.float32, device=dev).uniform_(-1, 1, generator=g)
        y = torch.empty((batch, n // 2 + 1), dtype=torch.complex64, device=dev)
        step = lambda: ctx.real_forward(x, y)      # noqa: E731
    elif kind == "pow":
        x = torch.empty((batch, n), dtype=torch.float32, device=dev).uniform_(-1, 1, generator=g)
        wnd = torch.hann_window(n, periodic=True, dtype=torch.float32, device=dev)
        y = torch.empty((batch, n // 2 + 1), dtype=torch.float32, device=dev)
        step = lambda: ctx.real_forward_power(x, wnd, y)   # noqa: E731
        args.no_e2e = True     # device-resident API only
    else:
        x = torch.view_as_complex(torch.empty((batch, n // 2 + 1, 2), dtype=torch.float32, device=dev).uniform_(-1, 1, generator=g))
        y = torch.empty((batch, n), dtype=torch.float32, device=dev)
        step = lambda: ctx.real_inverse(x, n, y)   # noqa: E731

    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = ck.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    launches = ck.kernel_launches() - launches0
    ms_total = reduce_max(ev0.elapsed_time(ev1), world, dev)
    ms_step = ms_total / args.steps
    if ms_total < 1000.0:
        # the timed region is shorter than a few sampling periods: keep the same step running (untimed) so that
        # the clock record describes this workload under load rather than the idle gaps around it
        for _ in range(int(min(2000, 1000.0 / max(ms_step, 1e-3)))):
            step()
        torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    if clocks is not None:
        clocks["window"] = "warm-up + timed steps" + (" + untimed continuation of the same step to 1 s" if ms_total < 1000.0 else "")
    gbs = spec["bytes"] * batch * world / ms_step / 1e6
    gflops = spec["flops"] * batch * world / ms_step / 1e6

    # cheap in-run sanity check of the timed buffers (full parity lives in tests/): Parseval on a slice
    if kind == "c2c":
        xin, yout = x[:64], y[:64]
        ein = float((xin.real.double() ** 2 + xin.imag.double() ** 2).sum())
        eout = float((yout.real.double() ** 2 + yout.imag.double() ** 2).sum()) / n
        if not abs(eout - ein) <= 1e-5 * ein:
            raise SystemExit(f"sanity check failed: Parseval {ein} vs {eout}")
    del x, y
    torch.cuda.empty_cache()

    # ---- end to end through the C ABI with host buffers ----
    e2e = pageable = multi = None
    if not args.no_e2e:
        e2e, pageable, multi = measure_e2e(args, spec, ctx, rank, world, local_rank, dev, barrier, host_barrier)

    # ---- the other BASELINE configs under the same clock ----
    secondary = None
    if not args.no_secondary and spec["name"] == "c2c1024":
        secondary = {"note": "device-resident, CUDA events, max over ranks, same conventions as the main line; "
                             "frac = per-GPU GB/s / measured HBM copy peak"}
        secondary["r2c4096"] = measure_config3("r2c4096", world, dev, barrier)
        secondary["c2r4096"] = measure_config3("c2r4096", world, dev, barrier)
        secondary["stft4096"] = measure_config3("stft4096", world, dev, barrier)
        sweep = measure_sweep(world, dev, barrier)
        secondary["sweep"] = {"workload": "C2C forward, N = 2^4 .. 2^20, batch = 2^28/N per GPU, 5 steps each",
                              "n": [r["n"] for r in sweep], "frac": [r["frac"] for r in sweep], "gbs": [r["gbs"] for r in sweep]}
        real = measure_real_large(world, dev, barrier)
        secondary["real_large"] = {"workload": "real n = 2^17 .. 2^21 (multi-pass), 2^27 samples per GPU per launch, 5 steps each",
                                   "n": [r["n"] for r in real], "r2c_frac": [r["r2c_frac"] for r in real],
                                   "c2r_frac": [r["c2r_frac"] for r in real]}
        if world > 1:
            try:
                secondary["dist30"] = measure_dist(30, rank, world, dev, barrier, steps=20, warmup=3)
            except Exception as exc:            # keep the main line: a failure here is reported, not fatal
                secondary["dist30"] = {"error": f"{type(exc).__name__}: {exc}"}

    host_barrier()
    if numa is not None:
        os.sched_setaffinity(0, numa[0])          # the CPU baseline gets every host core back
    if rank == 0:
        peak, peak_src = measured_peak()
        per_gpu = gbs / world
        line = {
            "metric": METRIC if spec["name"] == "c2c1024" else f"Batched fp32 {kind} FFT HBM GB/s + GFLOP/s at N={n}",
            "value": round(gbs, 1), "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "gflops": round(gflops, 1), "gflops_convention": "5*N*log2(N) per complex transform (2.5*N*log2(N) real)",
            "config": make_config(spec, world),
            "roofline": {"bound": "hbm", "achieved": round(per_gpu, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(per_gpu / peak, 4), "traffic": ncu_traffic(spec["name"]),
                         "peak_source": peak_src,
                         "note": "achieved = 16*N bytes x transforms per launch / mean kernel time (one launch per step, "
                                 "CUDA events on the launching stream), per GPU"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if e2e is not None:
            e2e["numa"] = (f"rank bound to the CPUs of its GPU's NUMA node ({numa[1]}) while allocating the pinned arrays"
                           if numa is not None else "no NUMA binding (single node, or CKFFT_BENCH_NUMA=0)")
            line["e2e"] = e2e
            line["e2e_pageable"] = pageable
            line["e2e_multi"] = multi
        if secondary is not None:
            line["secondary"] = secondary
        if not args.no_cpu_baseline and kind != "pow":
            line["cpu_baseline"] = cpu_baseline(spec)
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        host_barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
