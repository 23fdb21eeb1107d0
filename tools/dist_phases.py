#!/usr/bin/env python3
"""Per-phase device time of the distributed six-step (rank 0's view), under torchrun."""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ckfft_b200.distributed import DistributedFFT, split_n

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 28
n = 1 << lg
d = DistributedFFT(n)
be = d.backend
x = torch.view_as_complex(torch.empty((n // world, 2), dtype=torch.float32, device=dev).uniform_(-1, 1))
n1, n2 = split_n(n, world)
names, evs = [], []
def mark(name):
    e = torch.cuda.Event(enable_timing=True); e.record(); names.append(name); evs.append(e)
lib = be.lib
def exchange(a, rows, cols, tag):
    w = cols // world
    send = torch.empty_like(a)
    lib.CkFftB200PackColumnsAsync(a.data_ptr(), send.data_ptr(), rows, world, w, be._stream(a)); mark(tag + " pack")
    recv = torch.empty_like(a)
    dist.all_to_all_single(torch.view_as_real(recv), torch.view_as_real(send)); mark(tag + " all_to_all")
    out = torch.empty_like(a)
    lib.CkFftB200UnpackTransposeAsync(recv.data_ptr(), out.data_ptr(), world, rows, w, be._stream(a)); mark(tag + " unpack")
    return out
for it in range(3):
    names, evs = [], []
    dist.barrier(device_ids=[local]); torch.cuda.synchronize()
    mark("start")
    a = exchange(x, n1 // world, n2, "x1")
    a = be.local_fft(a, n2 // world, n1, False); mark("fft n1")
    be.twiddle(a, n, n2 // world, n1, rank * (n2 // world), False); mark("twiddle")
    a = exchange(a, n2 // world, n1, "x2")
    a = be.local_fft(a, n1 // world, n2, False); mark("fft n2")
    a = exchange(a, n1 // world, n2, "x3")
    torch.cuda.synchronize()
if rank == 0:
    tot = evs[0].elapsed_time(evs[-1])
    print(f"N=2^{lg} P={world} total {tot:.3f} ms")
    for i in range(1, len(evs)):
        print(f"  {names[i]:16s} {evs[i-1].elapsed_time(evs[i]):8.3f} ms")
dist.destroy_process_group()
