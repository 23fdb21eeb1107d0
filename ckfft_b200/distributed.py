"""Distributed six-step transform of ONE very large 1-D complex FFT over P GPUs (BASELINE.json config 5,
SURVEY.md 8e).  One process per GPU; `torch.distributed` (NCCL over NVLink / NVSwitch) does the three
all-to-all exchanges, every local step is a kernel of libckfft_b200.so called through the C ABI.

Index algebra (the classic six-step; the reference ships the same scheme, unbuilt, in
ext/fftw-3.3.2/mpi/dft-rank1.c:58-79).  N = N1*N2, x viewed as [N1][N2] (n = N2*n1 + n2), rank r owns the
rows n1 in block r -- which is exactly its natural-order slice x[r*N/P : (r+1)*N/P]:

    1. exchange+transpose   [N1/P][N2] -> [N2/P][N1]      rank r now owns the columns n2 in block r
    2. N1-point FFTs along each of its N2/P rows          A[n2][k1]
    3. twiddle               A[n2][k1] *= W_N^(n2*k1)
    4. exchange+transpose   [N2/P][N1] -> [N1/P][N2]      rank r owns k1 in block r
    5. N2-point FFTs along the rows                       B[k1][k2] = X[k1 + N1*k2]
    6. exchange+transpose   [N1/P][N2] -> [N2/P][N1]      [k2][k1]: natural order, rank r owns X[r*N/P : ...]

Each exchange is: pack the slab for every peer (CkFftB200PackColumnsAsync), all_to_all_single, tiled
transpose of what arrived (CkFftB200UnpackTransposeAsync).  Each all-to-all moves (P-1)/P * 8N/P bytes per GPU.

The algorithm is written against a small backend interface so that the host logic (slab arithmetic, exchange
order) is tested on CPU with gloo (tests/test_distributed_cpu.py, where the oracle stands in for the kernels).
"""
from __future__ import annotations

import numpy as np


def split_n(n: int, world: int) -> tuple[int, int]:
    """N = N1 * N2 with N1 <= N2, both powers of two and multiples of `world`."""
    if n <= 0 or n & (n - 1):
        raise ValueError("n must be a power of two")
    lg = n.bit_length() - 1
    n1 = 1 << (lg // 2)
    n2 = n // n1
    if n1 % world or n2 % world:
        raise ValueError(f"n={n} is too small to spread over {world} ranks")
    return n1, n2


def six_step(x_local, n: int, rank: int, world: int, be, inverse: bool = False):
    """Run the six steps on this rank's natural-order slice (length n/world).  `be` is the backend."""
    n1, n2 = split_n(n, world)
    a = be.exchange_transpose(x_local, n1 // world, n2, world)        # -> [n2/P][n1]
    a = be.local_fft(a, n2 // world, n1, inverse)
    be.twiddle(a, n, n2 // world, n1, rank * (n2 // world), inverse)
    a = be.exchange_transpose(a, n2 // world, n1, world)              # -> [n1/P][n2]
    a = be.local_fft(a, n1 // world, n2, inverse)
    return be.exchange_transpose(a, n1 // world, n2, world)           # -> [n2/P][n1] = natural order


class CudaBackend:
    """Local steps on the GPU through the C ABI; exchanges through torch.distributed (NCCL)."""

    def __init__(self, ctx, group=None):
        import torch
        import torch.distributed as dist

        from . import _lib

        self.torch, self.dist, self.lib, self.ctx, self.group = torch, dist, _lib.load(), ctx, group

    def _stream(self, t):
        return self.torch.cuda.current_stream(t.device).cuda_stream

    def exchange_transpose(self, a, rows, cols, world):
        """a: rank-local [rows][cols] (flat complex64) -> [cols/world][rows*world]"""
        torch = self.torch
        w = cols // world
        send = torch.empty_like(a)
        if not self.lib.CkFftB200PackColumnsAsync(a.data_ptr(), send.data_ptr(), rows, world, w, self._stream(a)):
            raise RuntimeError("pack failed")
        if world > 1:
            recv = torch.empty_like(a)
            self.dist.all_to_all_single(torch.view_as_real(recv), torch.view_as_real(send), group=self.group)
        else:
            recv = send
        out = torch.empty_like(a)
        if not self.lib.CkFftB200UnpackTransposeAsync(recv.data_ptr(), out.data_ptr(), world, rows, w, self._stream(a)):
            raise RuntimeError("unpack failed")
        return out

    def local_fft(self, a, rows, length, inverse):
        out = self.torch.empty_like(a)
        x, y = a.view(rows, length), out.view(rows, length)
        (self.ctx.complex_inverse if inverse else self.ctx.complex_forward)(x, y)
        return out

    def twiddle(self, a, n, rows, cols, first_row, inverse):
        if not self.lib.CkFftB200TwiddleRowsAsync(self.ctx.handle, n, a.data_ptr(), rows, cols, first_row, int(inverse),
                                                  self._stream(a)):
            from .api import last_error

            raise RuntimeError("twiddle failed: " + last_error())


class DistributedFFT:
    """One N-point complex transform spread over the ranks of a process group (N up to 2^30)."""

    def __init__(self, n: int, group=None):
        import torch.distributed as dist

        from .api import BOTH, Context

        self.n = n
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        split_n(n, self.world)
        self.ctx = Context(n, BOTH)          # nMax = n: two-level twiddles of W_n + tables for the local transforms
        self.backend = CudaBackend(self.ctx, group)

    def forward(self, x_local):
        """x_local: this rank's natural-order slice, complex64 CUDA tensor of n/world elements."""
        return six_step(x_local.reshape(-1), self.n, self.rank, self.world, self.backend, inverse=False)

    def inverse(self, x_local):
        return six_step(x_local.reshape(-1), self.n, self.rank, self.world, self.backend, inverse=True)

    def bytes_per_exchange(self) -> int:
        """bytes each GPU sends in one all-to-all"""
        return (self.world - 1) * 8 * self.n // (self.world * self.world)

    def close(self):
        self.ctx.close()


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
