#!/usr/bin/env python3
"""Isolate failures of the fused real-inverse kernel: one case per subprocess (a faulting kernel kills the context).
    python tools/twist_probe.py            runs all cases
    python tools/twist_probe.py <log2n> <batch> <offset>   one case"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1:
    import numpy as np
    import torch

    import ckfft_b200 as ck

    lg, batch, off = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    n = 1 << lg
    rng = np.random.default_rng(1)
    spec = (rng.uniform(-1, 1, (batch + off, n // 2 + 1)) + 1j * rng.uniform(-1, 1, (batch + off, n // 2 + 1))).astype(np.complex64)
    sd = torch.from_numpy(spec).cuda()
    with ck.Context(n, ck.BOTH) as ctx:
        x = ctx.real_inverse(sd[off:], n)
        torch.cuda.synchronize()
        os.environ["CKFFT_B200_PIPE_REAL"] = "0"
        x0 = ctx.real_inverse(sd[off:], n)
        torch.cuda.synchronize()
        bad = (x != x0).nonzero()
        err = float(torch.linalg.vector_norm((x - x0).double()) / torch.linalg.vector_norm(x0.double()))
        print(f"n=2^{lg} batch={batch} offset={off}: rel err vs separate twist pass {err:.2e} equal={bool(torch.equal(x, x0))} mismatches={bad.shape[0]}"
              + (f" first={bad[0].tolist()} cols mod: {sorted(set((bad[:64, 1] // 2 % (1 << (lg - 1 - (lg - 1) // 2))).tolist()))[:20]}" if bad.shape[0] else ""))
    sys.exit(0)

for lg in (16, 17, 20):
    for batch, off in ((1, 0), (1, 1), (2, 0), (5, 0), (5, 1)):
        r = subprocess.run([sys.executable, __file__, str(lg), str(batch), str(off)], capture_output=True, text=True,
                           env=dict(os.environ, CUDA_LAUNCH_BLOCKING="1"))
        out = (r.stdout.strip().splitlines() or ["(no output)"])[-1]
        err = [l for l in r.stderr.strip().splitlines() if "Error" in l or "error" in l][-1:] if r.returncode else []
        print(out if r.returncode == 0 else f"n=2^{lg} batch={batch} offset={off}: FAILED rc={r.returncode} {err}", flush=True)
