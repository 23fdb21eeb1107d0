#!/usr/bin/env python3
"""Host <-> device copy ceiling of this box: the denominator of bench.py's `e2e.roofline`.

    python tools/pcie_peak.py [--gpus N] [--mb 2048] [--chunk-mb 32] [--pageable]

Plain `cudaMemcpyAsync` traffic with the chunking of the library's host-buffer pipeline (32 MiB chunks, a few in
flight): host->device alone, device->host alone, and both directions at once on two streams -- the pattern the
CkFft*Batch host path generates, minus the transform.  With --gpus N the N devices copy concurrently (one thread per
device), which is what tells a per-link ceiling (PCIe Gen5 x16: ~55 GB/s per direction in practice) from a platform
ceiling (host DRAM / root complex / IOMMU) when the per-GPU figure drops as N grows.

`measure()` is imported by bench.py; it runs in the calling process on the current device (under torchrun every rank
calls it at the same time after a barrier, so the per-rank numbers are the concurrent ones)."""
from __future__ import annotations

import argparse
import json
import threading
import time


def gpu_local_cpus(device: int):
    """CPUs of the NUMA node the GPU hangs off (sysfs), or None if the platform does not say."""
    try:
        import subprocess

        # CUDA_VISIBLE_DEVICES is not set by torchrun, so the CUDA ordinal is nvidia-smi's index
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(device)],
                             capture_output=True, text=True, timeout=20).stdout.strip().splitlines()[0]
        bus = bus.lower()
        if bus.count(":") == 2 and len(bus.split(":")[0]) == 8:
            bus = bus[4:]                                   # 00000000:1B:00.0 -> 0000:1b:00.0
        with open(f"/sys/bus/pci/devices/{bus}/local_cpulist") as f:
            text = f.read().strip()
        cpus = set()
        for part in text.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        return (sorted(cpus), node) if cpus else None
    except Exception:
        return None


def bind_to_gpu_node(device: int):
    """Pin the calling thread to the CPUs next to `device`, so that the host memory it allocates next (first touch / cudaHostAlloc)
    lands on that NUMA node.  Returns (previous affinity, node) or None if nothing was done."""
    import os

    info = gpu_local_cpus(device)
    if not info:
        return None
    cpus, node = info
    try:
        prev = os.sched_getaffinity(0)
        allowed = prev & set(cpus)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return prev, node
    except (AttributeError, OSError):
        return None


def measure(device: int, total_bytes: int = 2 << 30, chunk_bytes: int = 32 << 20, pinned: bool = True, repeats: int = 3,
            host_src=None, host_dst=None):
    """-> dict(h2d, d2h, both_h2d, both_d2h, both) in GB/s (best of `repeats`).  `both` = bytes moved in the two
    directions together / wall time, i.e. directly comparable with bench.py's e2e value (input + output bytes / time).
    host_src / host_dst: optional existing host tensors (uint8 views are taken) so that no new memory is pinned."""
    import torch

    torch.cuda.set_device(device)
    dev = torch.device("cuda", device)
    nchunks = max(1, total_bytes // chunk_bytes)
    total = nchunks * chunk_bytes
    if host_src is None:
        host_src = torch.empty(total, dtype=torch.uint8, pin_memory=pinned)
        host_src[::4096] = 1                                   # touch every page
    else:
        host_src = host_src.view(torch.uint8).reshape(-1)[:total]
    if host_dst is None:
        host_dst = torch.empty(total, dtype=torch.uint8, pin_memory=pinned)
        host_dst[::4096] = 1
    else:
        host_dst = host_dst.view(torch.uint8).reshape(-1)[:total]
    slots = 3
    d_in = [torch.empty(chunk_bytes, dtype=torch.uint8, device=dev) for _ in range(slots)]
    d_out = [torch.ones(chunk_bytes, dtype=torch.uint8, device=dev) for _ in range(slots)]
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def run(h2d: bool, d2h: bool) -> float:
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for i in range(nchunks):
            lo = i * chunk_bytes
            if h2d:
                with torch.cuda.stream(s_in):
                    d_in[i % slots].copy_(host_src[lo:lo + chunk_bytes], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s_out):
                    host_dst[lo:lo + chunk_bytes].copy_(d_out[i % slots], non_blocking=True)
        torch.cuda.synchronize(dev)
        return time.perf_counter() - t0

    run(True, True)                                            # warm-up (first-touch, stream creation)
    best = {}
    for name, (a, b) in {"h2d": (True, False), "d2h": (False, True), "both": (True, True)}.items():
        best[name] = min(run(a, b) for _ in range(repeats))
    gb = total / 1e9
    return {"h2d": round(gb / best["h2d"], 2), "d2h": round(gb / best["d2h"], 2), "both": round(2 * gb / best["both"], 2),
            "bytes_each_way": total, "chunk_bytes": chunk_bytes, "pinned": bool(pinned)}


def main():
    import torch

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--mb", type=int, default=2048)
    ap.add_argument("--chunk-mb", type=int, default=32)
    ap.add_argument("--pageable", action="store_true")
    ap.add_argument("--numa", action="store_true", help="bind each worker to its GPU's NUMA node before it allocates host memory")
    args = ap.parse_args()
    n = min(args.gpus, torch.cuda.device_count())
    out = {"gpus": n, "pinned": not args.pageable, "numa_bound": bool(args.numa),
           "gpu_numa_nodes": [(gpu_local_cpus(i) or (None, None))[1] for i in range(n)]}
    for concurrent in sorted({1, n}):
        results = [None] * concurrent
        gate = threading.Barrier(concurrent)

        def work(i):
            torch.cuda.set_device(i)
            if args.numa:
                bind_to_gpu_node(i)
            gate.wait()
            results[i] = measure(i, args.mb << 20, args.chunk_mb << 20, not args.pageable)

        th = [threading.Thread(target=work, args=(i,)) for i in range(concurrent)]
        [t.start() for t in th]
        [t.join() for t in th]
        out[f"concurrent_{concurrent}"] = {
            "per_gpu": results,
            "sum_both": round(sum(r["both"] for r in results), 2),
            "sum_h2d": round(sum(r["h2d"] for r in results), 2),
            "sum_d2h": round(sum(r["d2h"] for r in results), 2),
        }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
