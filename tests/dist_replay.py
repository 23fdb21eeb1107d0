"""CPU stand-ins for the distributed transforms -- TEST INFRASTRUCTURE, not product code.

  * `NumpyBackend`: the backend interface of ckfft_b200.distributed.six_step with numpy slab arithmetic and a caller-
    supplied local transform (the tests pass the oracle), exchanges over gloo;
  * `replay_fused`: replays the exchange + pass descriptors of the fused distributed transform exactly as the library
    hands them to its kernels (CkFftB200DistDescribe) in numpy: a host-logic check of the routing algebra.
"""
from __future__ import annotations

import numpy as np

from ckfft_b200.distributed import fused_passes


def replay_fused(slices, layout, inverse: bool = False, deliver=None, ranks=None):
    """Replay the fused transform's exchange + pass descriptors in numpy (host-logic check of the routing algebra;
    the arithmetic of a pass is numpy's FFT).  `slices[r]` is rank r's natural-order slice.  With `deliver=None` all
    ranks are simulated in this process and the list of output slices is returned.  Otherwise only `ranks` are
    simulated and every store goes through deliver(buffer_id, {dest_rank: (flat_indices, values)}, bufs), which moves
    the data between processes (tests/test_distributed_cpu.py does that with gloo)."""
    world = layout.world
    n1, n2 = 1 << layout.log2n1, 1 << layout.log2n2
    h, w = n1 // world, n2 // world
    per = (n1 * n2) // world
    ranks = list(range(world)) if ranks is None else list(ranks)
    bufs = {r: [np.zeros(per, np.complex64) for _ in range(3)] for r in ranks}

    def default_deliver(buf_id, outgoing, _bufs):
        for q, (idx, val) in outgoing.items():
            bufs[q][buf_id][idx] = val

    deliver_fn = deliver or default_deliver
    sign = 2.0 if inverse else -2.0

    def fft(a, axis):
        a = a.astype(np.complex128)
        return (np.fft.ifft(a, axis=axis) * a.shape[axis] if inverse else np.fft.fft(a, axis=axis))

    # exchange: rank s pushes the column blocks of its rows (exchange_push_kernel); none in a pull layout
    for r in ([] if layout.pull else ranks):
        x = np.asarray(slices[r], np.complex64).reshape(h, n2)
        rows = (r * h + np.arange(h))[:, None]
        outgoing = {}
        for q in range(world):
            idx = rows * w + np.arange(w)[None, :]
            outgoing[q] = (idx.reshape(-1), x[:, q * w:(q + 1) * w].reshape(-1))
        deliver_fn(0, outgoing, bufs[r])
    npass = len(fused_passes(layout, 0))
    for i in range(npass):
        staged = []
        for r in ranks:
            d = fused_passes(layout, r)[i]
            src = bufs[r][d.src] if d.src < 3 else None
            L, npr, nc = d.L, int(d.nproblems), d.ncols
            k = np.arange(L, dtype=np.int64)
            c = np.arange(nc, dtype=np.int64)
            prob = np.arange(npr, dtype=np.int64)
            kk = prob[:, None] * d.kProbMul + k[None, :] * d.kMul                     # [prob][k]
            if d.kind == 0:
                if d.pull:
                    # row a of the problem = row a % pullRows of rank a // pullRows's INPUT slice (TMA boxes from peer memory)
                    assert npr == 1
                    col = (c // d.pullW) * d.pullN2 + d.pullCol0 + c % d.pullW
                    rows = [np.asarray(slices[a // d.pullRows], np.complex64)[(a % d.pullRows) * d.pullRowLen + col] for a in range(L)]
                    y = fft(np.stack(rows)[None, :, :], 1)
                else:
                    y = fft(src[:npr * L * nc].reshape(npr, L, nc), 1)
                kt = kk if d.routed else np.broadcast_to(k[None, :], (npr, L))
                cc = (d.twColBase + c) >> d.twColShift
                tn = 1 << d.twLog2
                e = (kt[:, :, None] * cc[None, None, :]) % tn
                y = y * np.exp(1j * sign * np.pi * e / tn)
            else:
                assert d.routed
                start = c[None, :] * d.inColStride + prob[:, None] * d.inProbStride  # [prob][c]
                x = src[start[:, :, None] + k[None, None, :]]                        # [prob][c][k]
                y = fft(x, 2).transpose(0, 2, 1)                                     # [prob][k][c]
            y = y.astype(np.complex64)
            if d.routed:
                dest = kk >> d.rankShift
                row = kk & ((1 << d.rankShift) - 1)
                idx = row[:, :, None] * d.outRowStride + d.outColBase + c[None, None, :]
                outgoing = {}
                for q in range(world):
                    m = np.broadcast_to((dest == q)[:, :, None], idx.shape)
                    outgoing[q] = (idx[m], y[m])
                staged.append((r, d.dst, outgoing))
            else:
                staged.append((r, d.dst, {r: (np.arange(npr * L * nc), y.reshape(-1))}))
        for r, dst, outgoing in staged:      # all ranks finish the pass before anything lands (the flag barrier)
            if list(outgoing.keys()) == [r]:
                bufs[r][dst][outgoing[r][0]] = outgoing[r][1]
            else:
                deliver_fn(dst, outgoing, bufs[r])
    return [bufs[r][2] for r in ranks]


class NumpyBackend:
    """CPU stand-in used by the gloo tests: same slab arithmetic, numpy for the local steps.
    `fft_rows(a2d, inverse)` supplies the local transform (the tests pass the oracle)."""

    def __init__(self, fft_rows, dist=None, group=None):
        self.fft_rows, self.dist, self.group = fft_rows, dist, group

    def exchange_transpose(self, a, rows, cols, world):
        import torch

        w = cols // world
        send = np.ascontiguousarray(a.reshape(rows, world, w).transpose(1, 0, 2))       # [P][rows][w]
        if world > 1:
            t_send = torch.from_numpy(send.view(np.float32).reshape(-1).copy())
            t_recv = torch.empty_like(t_send)
            self.dist.all_to_all_single(t_recv, t_send, group=self.group)
            recv = t_recv.numpy().view(np.complex64).reshape(world, rows, w)
        else:
            recv = send
        return np.ascontiguousarray(recv.transpose(2, 0, 1)).reshape(-1)                 # [w][P*rows]

    def local_fft(self, a, rows, length, inverse):
        return self.fft_rows(a.reshape(rows, length), inverse).reshape(-1)

    def twiddle(self, a, n, rows, cols, first_row, inverse):
        i = (first_row + np.arange(rows, dtype=np.int64))[:, None]
        k = np.arange(cols, dtype=np.int64)[None, :]
        ang = (2.0 if inverse else -2.0) * np.pi * ((i * k) % n).astype(np.float64) / n
        v = a.reshape(rows, cols)
        v *= np.exp(1j * ang).astype(np.complex64)
