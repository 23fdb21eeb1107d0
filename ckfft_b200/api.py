"""Python mirror of the ckfft interface on top of the C ABI (libckfft_b200.so).

Names and argument meaning follow the reference's public API (inc/ckfft/ckfft.h:59-158):
`Context(n_max, direction)` = CkFftInit, `.complex_forward` = CkFftComplexForward, ... with one
difference of convenience: the leading dimensions of an array are the batch, so one call maps to the
batched entry point of the same name.  Error behaviour is the reference's: the C call returns 0 and
the wrapper raises `CkFftError` with the library's reason.

Arrays may be numpy arrays (host path: staged through the GPU by the library) or CUDA torch tensors
(device path: enqueued on torch's current stream, no synchronisation).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

FORWARD, INVERSE, BOTH = 1, 2, 3   # CkFftDirection, inc/ckfft/ckfft.h:21-27


class CkFftError(RuntimeError):
    pass


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def last_error() -> str:
    return _lib.load().CkFftB200LastError().decode()


def kernel_launches() -> int:
    return int(_lib.load().CkFftB200KernelLaunches())


def get_plan(n: int, real: bool = False):
    """Host-side planner (no GPU needed): dict describing how a transform of n points will run."""
    p = _lib.Plan()
    if not _lib.load().CkFftB200GetPlan(n, int(real), C.byref(p)):
        return None
    return {"n": p.n, "real": bool(p.isReal), "complex_points": p.complexPoints, "passes": p.passes,
            "radix": [[r for r in row if r] for row in p.radix][: p.passes],
            "threads_per_transform": p.threadsPerTransform, "elems_per_thread": p.elemsPerThread,
            "transforms_per_cta": p.transformsPerCta, "shared_bytes": p.sharedBytes}


class Context:
    """CkFftContext.  Immutable after creation; safe to share between threads (inc/ckfft/ckfft.h:39-41)."""

    def __init__(self, n_max: int, direction: int = BOTH):
        self._lib = _lib.load()
        self._ctx = self._lib.CkFftInit(int(n_max), int(direction), None, None)
        if not self._ctx:
            raise CkFftError(f"CkFftInit({n_max}, {direction}) returned NULL: {last_error()}")
        self.n_max = n_max
        self.direction = direction

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.CkFftShutdown(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def handle(self):
        return self._ctx

    @property
    def device(self) -> int:
        return int(self._lib.CkFftB200ContextDevice(self._ctx))

    # -- helpers -------------------------------------------------------------------------------
    def _fail(self, what):
        raise CkFftError(f"{what} returned 0: {last_error()}")

    def _run(self, name, n, x, out, batch):
        if _is_torch(x):
            import torch  # plumbing only: device memory and the current stream

            stream = torch.cuda.current_stream(x.device).cuda_stream
            fn = getattr(self._lib, name + "BatchAsync")
            ok = fn(self._ctx, n, x.data_ptr(), out.data_ptr(), batch, 0, 0, stream)
        else:
            fn = getattr(self._lib, name + "Batch")
            if name == "CkFftRealInverse":
                ok = fn(self._ctx, n, x.ctypes.data, out.ctypes.data, None, batch)
            else:
                ok = fn(self._ctx, n, x.ctypes.data, out.ctypes.data, batch)
        if not ok:
            self._fail(name)
        return out

    @staticmethod
    def _prep(x, np_dtype, torch_name):
        if _is_torch(x):
            import torch

            dt = getattr(torch, torch_name)
            if x.dtype != dt or not x.is_cuda:
                raise CkFftError(f"expected a CUDA tensor of dtype {torch_name}")
            return x.contiguous()
        return np.ascontiguousarray(x, dtype=np_dtype)

    @staticmethod
    def _check_out(out, like, shape, np_dtype, torch_name, what="out"):
        """A caller-supplied output goes to the C ABI as a bare pointer with dense strides: it must be exactly the
        array the call would have allocated (dtype, shape, contiguous, same side / device as the input)."""
        shape = tuple(int(d) for d in shape)
        if _is_torch(like):
            import torch

            if not _is_torch(out) or not out.is_cuda or out.device != like.device:
                raise CkFftError(f"{what} must be a CUDA tensor on {like.device}")
            if out.dtype != getattr(torch, torch_name) or tuple(out.shape) != shape or not out.is_contiguous():
                raise CkFftError(f"{what} must be a contiguous {torch_name} tensor of shape {shape}")
        else:
            if not isinstance(out, np.ndarray) or out.dtype != np.dtype(np_dtype) or tuple(out.shape) != shape \
                    or not out.flags.c_contiguous or not out.flags.writeable:
                raise CkFftError(f"{what} must be a writeable C-contiguous numpy {np.dtype(np_dtype).name} array of shape {shape}")
        return out

    @staticmethod
    def _empty_like(x, shape, np_dtype, torch_name):
        if _is_torch(x):
            import torch

            return torch.empty(shape, dtype=getattr(torch, torch_name), device=x.device)
        return np.empty(shape, dtype=np_dtype)

    # -- the four transforms -------------------------------------------------------------------
    def complex_forward(self, x, out=None):
        """CkFftComplexForward over the last axis; x complex64[..., n]."""
        return self._complex(x, out, "CkFftComplexForward")

    def complex_inverse(self, x, out=None):
        """CkFftComplexInverse over the last axis (not divided by n)."""
        return self._complex(x, out, "CkFftComplexInverse")

    def _complex(self, x, out, name):
        x = self._prep(x, np.complex64, "complex64")
        n = x.shape[-1]
        batch = int(np.prod(x.shape[:-1], dtype=np.int64)) if x.ndim > 1 else 1
        if out is None:
            out = self._empty_like(x, tuple(x.shape), np.complex64, "complex64")
        else:
            self._check_out(out, x, tuple(x.shape), np.complex64, "complex64")
        return self._run(name, n, x, out, batch)

    def complex_planar(self, re, im, inverse: bool = False, out=None):
        """Split-complex transform (CkFftB200Complex{Forward,Inverse}PlanarBatchAsync): re, im float32[..., n] CUDA
        tensors -> (re, im) of the spectrum.  `out=(re, im)` transforms in place.  Stream-ordered on torch's stream."""
        import torch

        re = self._prep(re, np.float32, "float32")
        im = self._prep(im, np.float32, "float32")
        if not _is_torch(re) or not _is_torch(im) or re.shape != im.shape:
            raise CkFftError("complex_planar works on two CUDA tensors of the same shape")
        n = re.shape[-1]
        batch = int(np.prod(re.shape[:-1], dtype=np.int64)) if re.ndim > 1 else 1
        if out is None:
            out = (torch.empty_like(re), torch.empty_like(im))
        else:
            if not isinstance(out, (tuple, list)) or len(out) != 2:
                raise CkFftError("out must be a pair (re, im)")
            for o, name in zip(out, ("out[0]", "out[1]")):
                self._check_out(o, re, tuple(re.shape), np.float32, "float32", name)
        fn = self._lib.CkFftB200ComplexInversePlanarBatchAsync if inverse else self._lib.CkFftB200ComplexForwardPlanarBatchAsync
        stream = torch.cuda.current_stream(re.device).cuda_stream
        if not fn(self._ctx, n, re.data_ptr(), im.data_ptr(), out[0].data_ptr(), out[1].data_ptr(), batch, 0, 0, stream):
            self._fail("CkFftB200ComplexPlanarBatchAsync")
        return out

    def real_forward(self, x, out=None):
        """CkFftRealForward: float32[..., n] -> complex64[..., n/2+1] (= 2 * rfft)."""
        x = self._prep(x, np.float32, "float32")
        n = x.shape[-1]
        batch = int(np.prod(x.shape[:-1], dtype=np.int64)) if x.ndim > 1 else 1
        if out is None:
            out = self._empty_like(x, tuple(x.shape[:-1]) + (n // 2 + 1,), np.complex64, "complex64")
        else:
            self._check_out(out, x, tuple(x.shape[:-1]) + (n // 2 + 1,), np.complex64, "complex64")
        return self._run("CkFftRealForward", n, x, out, batch)

    def real_forward_power(self, x, window=None, out=None):
        """Fused audio front end: |CkFftRealForward(window * x)|^2, float32[..., n] -> float32[..., n/2+1].
        CUDA tensors only (stream-ordered on torch's current stream)."""
        import torch

        x = self._prep(x, np.float32, "float32")
        if not _is_torch(x):
            raise CkFftError("real_forward_power works on CUDA tensors")
        n = x.shape[-1]
        batch = int(np.prod(x.shape[:-1], dtype=np.int64)) if x.ndim > 1 else 1
        if window is not None:
            window = self._prep(window, np.float32, "float32")
            if tuple(window.shape) != (n,):
                raise CkFftError(f"window must have {n} samples")
        if out is None:
            out = torch.empty(tuple(x.shape[:-1]) + (n // 2 + 1,), dtype=torch.float32, device=x.device)
        else:
            self._check_out(out, x, tuple(x.shape[:-1]) + (n // 2 + 1,), np.float32, "float32")
        stream = torch.cuda.current_stream(x.device).cuda_stream
        ok = self._lib.CkFftB200RealForwardPowerBatchAsync(self._ctx, n, x.data_ptr(), window.data_ptr() if window is not None else None,
                                                           out.data_ptr(), batch, 0, 0, stream)
        if not ok:
            self._fail("CkFftB200RealForwardPowerBatchAsync")
        return out

    def real_inverse(self, y, n: int, out=None):
        """CkFftRealInverse: complex64[..., n/2+1] -> float32[..., n]."""
        y = self._prep(y, np.complex64, "complex64")
        if y.shape[-1] != n // 2 + 1:
            raise CkFftError(f"expected {n // 2 + 1} bins for n={n}, got {y.shape[-1]}")
        batch = int(np.prod(y.shape[:-1], dtype=np.int64)) if y.ndim > 1 else 1
        if out is None:
            out = self._empty_like(y, tuple(y.shape[:-1]) + (n,), np.float32, "float32")
        else:
            self._check_out(out, y, tuple(y.shape[:-1]) + (n,), np.float32, "float32")
        return self._run("CkFftRealInverse", n, y, out, batch)


class MultiContext:
    """CkFftB200Multi: the batched scheduler behind `CkFft*BatchMulti` -- one call spreads a batch of independent
    transforms held in HOST arrays over several GPUs (contiguous shards, one context replica + one host thread per
    device, no collective; include/ckfft/ckfft_b200.h).  `devices=None`: every visible device."""

    def __init__(self, n_max: int, direction: int = BOTH, devices=None):
        self._lib = _lib.load()
        if devices is None:
            arr, cnt = None, 0
        else:
            devices = [int(d) for d in devices]
            arr, cnt = (C.c_int * len(devices))(*devices), len(devices)
        self._m = self._lib.CkFftB200MultiInit(int(n_max), int(direction), arr, cnt)
        if not self._m:
            raise CkFftError(f"CkFftB200MultiInit({n_max}, {direction}, {devices}) returned NULL: {last_error()}")
        self.n_max = n_max

    @property
    def devices(self) -> list[int]:
        return [int(self._lib.CkFftB200MultiDevice(self._m, i)) for i in range(self._lib.CkFftB200MultiDeviceCount(self._m))]

    def close(self):
        if getattr(self, "_m", None):
            self._lib.CkFftB200MultiShutdown(self._m)
            self._m = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _run(self, name, n, x, out_shape, out_dtype, out, batch):
        if out is None:
            out = np.empty(out_shape, dtype=out_dtype)
        else:
            Context._check_out(out, x, out_shape, out_dtype, "", "out")
        if not getattr(self._lib, name)(self._m, n, x.ctypes.data, out.ctypes.data, batch):
            raise CkFftError(f"{name} returned 0: {last_error()}")
        return out

    @staticmethod
    def _batch(x):
        return int(np.prod(x.shape[:-1], dtype=np.int64)) if x.ndim > 1 else 1

    def complex_forward(self, x, out=None):
        x = np.ascontiguousarray(x, dtype=np.complex64)
        return self._run("CkFftComplexForwardBatchMulti", x.shape[-1], x, tuple(x.shape), np.complex64, out, self._batch(x))

    def complex_inverse(self, x, out=None):
        x = np.ascontiguousarray(x, dtype=np.complex64)
        return self._run("CkFftComplexInverseBatchMulti", x.shape[-1], x, tuple(x.shape), np.complex64, out, self._batch(x))

    def real_forward(self, x, out=None):
        x = np.ascontiguousarray(x, dtype=np.float32)
        n = x.shape[-1]
        return self._run("CkFftRealForwardBatchMulti", n, x, tuple(x.shape[:-1]) + (n // 2 + 1,), np.complex64, out, self._batch(x))

    def real_inverse(self, y, n: int, out=None):
        y = np.ascontiguousarray(y, dtype=np.complex64)
        if y.shape[-1] != n // 2 + 1:
            raise CkFftError(f"expected {n // 2 + 1} bins for n={n}, got {y.shape[-1]}")
        return self._run("CkFftRealInverseBatchMulti", n, y, tuple(y.shape[:-1]) + (n,), np.float32, out, self._batch(y))
