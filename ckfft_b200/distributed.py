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
order) is tested on CPU with gloo (tests/test_distributed_cpu.py; the CPU stand-in backend and the numpy replay of the
fused pass descriptors live with the tests, tests/dist_replay.py -- nothing in this package computes on the CPU).
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
        self.ctx = Context(max(n, 1 << 15), BOTH)   # a multi-pass context: two-level twiddles + tables for the local transforms
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


# ---------------------------------------------------------------------------------------------------------------
# Fused variant: no collective library on the data path.  The transform passes store their results directly into
# the memory of the GPU that needs them next (csrc/dist_fused.cu, include/ckfft/ckfft_b200.h "Fused distributed
# transform"); torch.distributed only ships the 64-byte memory handles once, at set-up.
# ---------------------------------------------------------------------------------------------------------------
def fused_layout(n: int, world: int, prefer_passes: int = 0, pull: bool = False):
    """CkFftB200DistGetLayout: how n = n1*n2 is factored into passes (pure host arithmetic, no GPU needed).
    pull: no exchange kernel, the first pass reads the ranks' input arrays directly."""
    import ctypes as C

    from . import _lib

    lay = _lib.DistLayout()
    if not _lib.load().CkFftB200DistGetLayout(int(n), int(world), int(prefer_passes), C.byref(lay)):
        raise ValueError(f"n={n} cannot be spread over {world} ranks by the fused distributed transform")
    lay.pull = int(bool(pull))
    return lay


def fused_passes(layout, rank: int):
    """CkFftB200DistDescribe: the pass descriptors of `rank` exactly as the library hands them to its kernels."""
    from . import _lib

    arr = (_lib.DistPass * 4)()
    cnt = _lib.load().CkFftB200DistDescribe(layout, int(rank), arr)
    return [arr[i] for i in range(cnt)]


class _DeviceArray:
    """Raw device memory as a CUDA-array-interface object (so torch can view a buffer the library allocated)."""

    def __init__(self, ptr: int, nfloats: int):
        self.__cuda_array_interface__ = {"shape": (nfloats,), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class FusedDistributedFFT:
    """One N-point complex transform spread over the ranks of a process group, NVLink stores instead of collectives.

    forward(x_local) / inverse(x_local): x_local is this rank's natural-order slice (complex64 CUDA tensor of
    n/world elements); the result is a view of the plan's own output buffer, valid until the next call."""

    def __init__(self, n: int, group=None, prefer_passes: int = 0, pull=None):
        """pull=True: no exchange kernel -- the first pass fetches its tiles from the peers' input arrays with TMA (put
        the input into `self.input` to avoid a local copy).  pull=False: a push exchange kernel runs first.
        Default (None): pull for four-pass layouts (n >= 2^28), where it hides a local pass behind the NVLink reads
        (2^30 on 8 GPUs: 5.11 -> 4.83 ms); push below, where the first pass would both pull and push and gains nothing."""
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import _lib
        from .api import BOTH, Context, CkFftError, last_error

        self.torch, self.n, self.group = torch, n, group
        self.lib = lib = _lib.load()
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.layout = fused_layout(n, self.world, prefer_passes)
        if pull is None:
            pull = self.layout.la > 1
        self.ctx = Context(max(n, 1 << 15), BOTH)      # a multi-pass context: it carries the two-level twiddles of W_nMax
        self.device = torch.device("cuda", torch.cuda.current_device())
        per_bytes = 8 * n // self.world
        sizes = [per_bytes, per_bytes, per_bytes, 256] + ([per_bytes] if pull else [])
        self.layout.pull = int(bool(pull))
        self._own = []
        for b in sizes:
            ptr = lib.CkFftB200PeerAlloc(b)
            if not ptr:
                raise CkFftError("CkFftB200PeerAlloc: " + last_error())
            self._own.append(ptr)
        self._opened = []
        table = [[None] * self.world for _ in sizes]      # [buffer][rank] -> device pointer valid in this process
        if self.world > 1:
            mine = []
            for ptr in self._own:
                h = C.create_string_buffer(64)
                if not lib.CkFftB200PeerExport(ptr, h):
                    raise CkFftError("CkFftB200PeerExport: " + last_error())
                mine.append(h.raw)
            everyone = [None] * self.world
            dist.all_gather_object(everyone, mine, group=group)
            for q in range(self.world):
                for b in range(len(sizes)):
                    if q == self.rank:
                        table[b][q] = self._own[b]
                    else:
                        ptr = lib.CkFftB200PeerOpen(everyone[q][b])
                        if not ptr:
                            raise CkFftError(f"CkFftB200PeerOpen(rank {q}): " + last_error())
                        self._opened.append(ptr)
                        table[b][q] = ptr
        else:
            for b in range(len(sizes)):
                table[b][0] = self._own[b]
        arrs = [(C.c_void_p * self.world)(*table[b]) for b in range(len(sizes))]
        if not pull:
            arrs.append(None)
        self._plan = lib.CkFftB200DistPlanCreate(self.ctx.handle, n, self.rank, self.world, prefer_passes, *arrs)
        if not self._plan:
            raise CkFftError("CkFftB200DistPlanCreate: " + last_error())
        flat = torch.as_tensor(_DeviceArray(self._own[2], 2 * n // self.world), device=self.device)
        self.out = torch.view_as_complex(flat.view(-1, 2))
        self.input = None
        if pull:
            flat_in = torch.as_tensor(_DeviceArray(self._own[4], 2 * n // self.world), device=self.device)
            self.input = torch.view_as_complex(flat_in.view(-1, 2))
        if self.world > 1:
            dist.barrier(group=group)        # every rank has mapped everything before anyone starts storing

    def _run(self, x_local, inverse):
        from .api import CkFftError, last_error

        x = x_local.reshape(-1)
        assert x.is_cuda and x.dtype == self.torch.complex64 and x.numel() == self.n // self.world and x.is_contiguous()
        stream = self.torch.cuda.current_stream(self.device).cuda_stream
        if not self.lib.CkFftB200DistExecAsync(self._plan, x.data_ptr(), int(inverse), stream):
            raise CkFftError("CkFftB200DistExecAsync: " + last_error())
        return self.out

    def forward(self, x_local):
        return self._run(x_local, False)

    def inverse(self, x_local):
        return self._run(x_local, True)

    def check(self):
        """Synchronise and raise if a barrier of this plan ever timed out."""
        from .api import CkFftError, last_error

        if not self.lib.CkFftB200DistPlanStatus(self._plan):
            raise CkFftError(last_error())

    def profile(self, x_local, inverse: bool = False):
        """One profiled execution: list of (phase name, device ms) as seen by this rank."""
        import ctypes as C

        self.lib.CkFftB200DistPlanSetProfiling(self._plan, 1)
        self._run(x_local, inverse)
        ms = (C.c_float * 12)()
        names = C.create_string_buffer(512)
        cnt = self.lib.CkFftB200DistPlanPhases(self._plan, ms, names, 512)
        self.lib.CkFftB200DistPlanSetProfiling(self._plan, 0)
        return list(zip(names.value.decode().split(","), [float(ms[i]) for i in range(cnt)]))

    def bytes_per_exchange(self) -> int:
        return (self.world - 1) * 8 * self.n // (self.world * self.world)

    def close(self):
        if getattr(self, "_plan", None):
            self.torch.cuda.synchronize(self.device)
            self.lib.CkFftB200DistPlanDestroy(self._plan)
            self._plan = None
            self.out = None
            self.input = None
            for p in self._opened:
                self.lib.CkFftB200PeerClose(p)
            for p in self._own:
                self.lib.CkFftB200PeerFree(p)
            self._opened, self._own = [], []
            self.ctx.close()
