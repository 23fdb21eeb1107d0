#!/usr/bin/env python3
"""In-tree build of libckfft_b200.so (nvcc, sm_100a only).

    python -m ckfft_b200.build            # incremental
    python -m ckfft_b200.build --force

The library is a plain C-ABI shared object (include/ckfft/*.h); nothing in it depends on torch or
Python.  Objects are compiled in parallel (the single-pass kernel family is instantiated once per
variant, see csrc/fft_variants.cu) and linked with the static CUDA runtime, so the .so is
self-contained and travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
# CKFFT_B200_BUILD_TAG=<tag> (development): objects in build_<tag>/, library lib/libckfft_b200_<tag>.so, so that an A/B
# build (CKFFT_B200_NVCC_FLAGS="-D...") sits next to the product library; CKFFT_B200_LIB=<path> makes _lib.py load it.
_TAG = os.environ.get("CKFFT_B200_BUILD_TAG", "")
OBJ = os.path.join(PKG, "build" + ("_" + _TAG if _TAG else ""))
LIB = os.path.join(PKG, "lib", "libckfft_b200" + ("_" + _TAG if _TAG else "") + ".so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", *ARCH, "-Xcompiler", "-fPIC,-fvisibility=default",
          f"-I{os.path.join(ROOT, 'include')}", f"-I{CSRC}", *os.environ.get("CKFFT_B200_NVCC_FLAGS", "").split()]

# (object name, source, extra flags)
UNITS = [
    ("api.o", "api.cu", []),
    ("multi.o", "multi.cu", []),
    ("host_copy.o", "host_copy.cpp", []),
    ("tiny.o", "tiny.cu", []),
    ("small.o", "small.cu", []),
    ("four_step.o", "four_step.cu", []),
    ("dist_glue.o", "dist_glue.cu", []),
    ("dist_fused.o", "dist_fused.cu", []),
    ("c2c_fwd.o", "fft_variants.cu", ["-DCKB_VARIANT=0"]),
    ("c2c_inv.o", "fft_variants.cu", ["-DCKB_VARIANT=1"]),
    ("r2c.o", "fft_variants.cu", ["-DCKB_VARIANT=2"]),
    ("c2r.o", "fft_variants.cu", ["-DCKB_VARIANT=3"]),
    ("r2c_audio.o", "fft_variants.cu", ["-DCKB_VARIANT=4"]),
    ("c2c_fwd_planar.o", "fft_variants.cu", ["-DCKB_VARIANT=5"]),
    ("c2c_inv_planar.o", "fft_variants.cu", ["-DCKB_VARIANT=6"]),
]


# CKFFT_B200_UNIT_FLAGS="four_step.o:-DX=1 -DY=2;api.o:-DZ" adds flags to single translation units (development A/B builds)
_UNIT_FLAGS = {}
for _item in os.environ.get("CKFFT_B200_UNIT_FLAGS", "").split(";"):
    if ":" in _item:
        _obj, _fl = _item.split(":", 1)
        _UNIT_FLAGS[_obj.strip()] = _fl.split()


def _source_stamp() -> str:
    h = hashlib.sha256()
    for d in (CSRC, os.path.join(ROOT, "include", "ckfft")):
        for name in sorted(os.listdir(d)):
            with open(os.path.join(d, name), "rb") as f:
                h.update(name.encode())
                h.update(f.read())
    h.update(" ".join(COMMON).encode())
    h.update(repr(sorted(_UNIT_FLAGS.items())).encode())
    h.update(repr(UNITS).encode())
    return h.hexdigest()


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: " + " ".join(cmd) + "\n" + r.stdout)
    return r.stdout


def build(force: bool = False, verbose: bool = False) -> str:
    stamp_file = os.path.join(OBJ, "stamp")
    stamp = _source_stamp()
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return LIB
    if not os.path.exists(NVCC):
        if os.path.exists(LIB):
            # GPU box without a toolkit: use the library that travelled with the snapshot
            return LIB
        raise RuntimeError(f"nvcc not found at {NVCC} and no prebuilt {LIB}")
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(os.path.dirname(LIB), exist_ok=True)

    def compile_unit(unit):
        obj, src, extra = unit
        cmd = [NVCC, *COMMON, *extra, *_UNIT_FLAGS.get(obj, []), "-c", os.path.join(CSRC, src), "-o", os.path.join(OBJ, obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        out = _run(cmd)
        if verbose:
            print(out)
        return os.path.join(OBJ, obj)

    with ThreadPoolExecutor(max_workers=min(len(UNITS), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_unit, UNITS))
    _run([NVCC, *ARCH, "-shared", "-o", LIB, *objs])
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
