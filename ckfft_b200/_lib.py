"""ctypes binding of libckfft_b200.so -- the C ABI declared in include/ckfft/ckfft.h and
include/ckfft/ckfft_b200.h.  Fails loudly when the library is missing: there is no Python or CPU
fallback for any transform."""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CKFFT_B200_LIB") or os.path.join(PKG, "lib", "libckfft_b200.so")   # override: A/B builds (build.py)

# every symbol the two public headers declare (tests check the .so exports exactly these)
CLASSIC_SYMBOLS = ["CkFftInit", "CkFftRealForward", "CkFftRealInverse", "CkFftComplexForward",
                   "CkFftComplexInverse", "CkFftShutdown"]
B200_SYMBOLS = ["CkFftComplexForwardBatch", "CkFftComplexInverseBatch", "CkFftRealForwardBatch",
                "CkFftRealInverseBatch", "CkFftComplexForwardBatchAsync", "CkFftComplexInverseBatchAsync",
                "CkFftRealForwardBatchAsync", "CkFftRealInverseBatchAsync", "CkFftB200GetPlan",
                "CkFftB200LastError", "CkFftB200KernelLaunches", "CkFftB200HostAlloc", "CkFftB200HostFree",
                "CkFftB200ContextDevice", "CkFftB200PackColumnsAsync", "CkFftB200UnpackTransposeAsync",
                "CkFftB200TwiddleRowsAsync", "CkFftB200RealForwardPowerBatchAsync",
                "CkFftB200ComplexForwardPlanarBatchAsync", "CkFftB200ComplexInversePlanarBatchAsync",
                "CkFftB200DistGetLayout", "CkFftB200DistDescribe", "CkFftB200PeerAlloc", "CkFftB200PeerFree",
                "CkFftB200PeerExport", "CkFftB200PeerOpen", "CkFftB200PeerClose", "CkFftB200DistPlanCreate",
                "CkFftB200DistExecAsync", "CkFftB200DistPlanStatus", "CkFftB200DistPlanDestroy",
                "CkFftB200DistPlanSetProfiling", "CkFftB200DistPlanPhases",
                "CkFftB200MultiInit", "CkFftB200MultiShutdown", "CkFftB200MultiDeviceCount", "CkFftB200MultiDevice",
                "CkFftB200MultiContext", "CkFftB200ShardRange", "CkFftComplexForwardBatchMulti",
                "CkFftComplexInverseBatchMulti", "CkFftRealForwardBatchMulti", "CkFftRealInverseBatchMulti"]


class Plan(C.Structure):
    """CkFftB200Plan (include/ckfft/ckfft_b200.h)."""
    _fields_ = [("n", C.c_int), ("isReal", C.c_int), ("complexPoints", C.c_int), ("passes", C.c_int),
                ("radix", (C.c_int * 3) * 2), ("threadsPerTransform", C.c_int), ("elemsPerThread", C.c_int),
                ("transformsPerCta", C.c_int), ("sharedBytes", C.c_int)]


class DistLayout(C.Structure):
    """CkFftB200DistLayout"""
    _fields_ = [("log2n", C.c_int), ("world", C.c_int), ("log2n1", C.c_int), ("log2n2", C.c_int), ("la", C.c_int),
                ("lb", C.c_int), ("lc", C.c_int), ("ld", C.c_int), ("passes", C.c_int), ("pull", C.c_int)]


class DistPass(C.Structure):
    """CkFftB200DistPass"""
    _fields_ = [("kind", C.c_int), ("routed", C.c_int), ("L", C.c_int), ("nproblems", C.c_longlong), ("ncols", C.c_int),
                ("twLog2", C.c_int), ("twColBase", C.c_int), ("twColShift", C.c_int), ("kProbMul", C.c_int),
                ("kMul", C.c_int), ("rankShift", C.c_int), ("outRowStride", C.c_longlong), ("outColBase", C.c_longlong),
                ("inColStride", C.c_longlong), ("inProbStride", C.c_longlong), ("src", C.c_int), ("dst", C.c_int),
                ("pull", C.c_int), ("pullRows", C.c_int), ("pullRowLen", C.c_longlong), ("pullN2", C.c_longlong),
                ("pullCol0", C.c_longlong), ("pullW", C.c_int)]


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m ckfft_b200.build` (needs nvcc). "
            "ckfft_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, sz, i = C.c_void_p, C.c_size_t, C.c_int
    lib.CkFftInit.restype = vp
    lib.CkFftInit.argtypes = [i, i, vp, C.POINTER(sz)]
    lib.CkFftShutdown.restype = None
    lib.CkFftShutdown.argtypes = [vp]
    for name in ("CkFftComplexForward", "CkFftComplexInverse", "CkFftRealForward"):
        getattr(lib, name).argtypes = [vp, i, vp, vp]
    lib.CkFftRealInverse.argtypes = [vp, i, vp, vp, vp]
    for name in ("CkFftComplexForwardBatch", "CkFftComplexInverseBatch", "CkFftRealForwardBatch"):
        getattr(lib, name).argtypes = [vp, i, vp, vp, sz]
    lib.CkFftRealInverseBatch.argtypes = [vp, i, vp, vp, vp, sz]
    for name in ("CkFftComplexForwardBatchAsync", "CkFftComplexInverseBatchAsync", "CkFftRealForwardBatchAsync",
                 "CkFftRealInverseBatchAsync"):
        getattr(lib, name).argtypes = [vp, i, vp, vp, sz, sz, sz, vp]
    lib.CkFftB200GetPlan.argtypes = [i, i, C.POINTER(Plan)]
    lib.CkFftB200LastError.restype = C.c_char_p
    lib.CkFftB200KernelLaunches.restype = C.c_ulonglong
    lib.CkFftB200HostAlloc.restype = vp
    lib.CkFftB200HostAlloc.argtypes = [sz]
    lib.CkFftB200HostFree.restype = None
    lib.CkFftB200HostFree.argtypes = [vp]
    lib.CkFftB200ContextDevice.argtypes = [vp]
    lib.CkFftB200PackColumnsAsync.argtypes = [vp, vp, sz, i, sz, vp]
    lib.CkFftB200UnpackTransposeAsync.argtypes = [vp, vp, i, sz, sz, vp]
    lib.CkFftB200TwiddleRowsAsync.argtypes = [vp, i, vp, sz, sz, sz, i, vp]
    lib.CkFftB200RealForwardPowerBatchAsync.argtypes = [vp, i, vp, vp, vp, sz, sz, sz, vp]
    for name in ("CkFftB200ComplexForwardPlanarBatchAsync", "CkFftB200ComplexInversePlanarBatchAsync"):
        getattr(lib, name).argtypes = [vp, i, vp, vp, vp, vp, sz, sz, sz, vp]
    lib.CkFftB200DistGetLayout.argtypes = [C.c_longlong, i, i, C.POINTER(DistLayout)]
    lib.CkFftB200DistDescribe.argtypes = [C.POINTER(DistLayout), i, C.POINTER(DistPass)]
    lib.CkFftB200PeerAlloc.restype = vp
    lib.CkFftB200PeerAlloc.argtypes = [sz]
    lib.CkFftB200PeerFree.restype = None
    lib.CkFftB200PeerFree.argtypes = [vp]
    lib.CkFftB200PeerExport.argtypes = [vp, C.c_char_p]
    lib.CkFftB200PeerOpen.restype = vp
    lib.CkFftB200PeerOpen.argtypes = [C.c_char_p]
    lib.CkFftB200PeerClose.restype = None
    lib.CkFftB200PeerClose.argtypes = [vp]
    lib.CkFftB200DistPlanCreate.restype = vp
    lib.CkFftB200DistPlanCreate.argtypes = [vp, C.c_longlong, i, i, i, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp),
                                            C.POINTER(vp)]
    lib.CkFftB200DistExecAsync.argtypes = [vp, vp, i, vp]
    lib.CkFftB200DistPlanStatus.argtypes = [vp]
    lib.CkFftB200DistPlanSetProfiling.argtypes = [vp, i]
    lib.CkFftB200DistPlanPhases.argtypes = [vp, C.POINTER(C.c_float), C.c_char_p, sz]
    lib.CkFftB200DistPlanDestroy.restype = None
    lib.CkFftB200DistPlanDestroy.argtypes = [vp]
    lib.CkFftB200MultiInit.restype = vp
    lib.CkFftB200MultiInit.argtypes = [i, i, C.POINTER(C.c_int), i]
    lib.CkFftB200MultiShutdown.restype = None
    lib.CkFftB200MultiShutdown.argtypes = [vp]
    lib.CkFftB200MultiDeviceCount.argtypes = [vp]
    lib.CkFftB200MultiDevice.argtypes = [vp, i]
    lib.CkFftB200MultiContext.restype = vp
    lib.CkFftB200MultiContext.argtypes = [vp, i]
    lib.CkFftB200ShardRange.argtypes = [sz, i, i, C.POINTER(sz), C.POINTER(sz)]
    for name in ("CkFftComplexForwardBatchMulti", "CkFftComplexInverseBatchMulti", "CkFftRealForwardBatchMulti",
                 "CkFftRealInverseBatchMulti"):
        getattr(lib, name).argtypes = [vp, i, vp, vp, sz]
    _lib = lib
    return lib
