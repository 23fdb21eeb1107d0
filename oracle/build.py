#!/usr/bin/env python3
"""Build recipe for the parity oracle (TEST INFRASTRUCTURE, never product code).

Targets
  oracle/libckfft_oracle.so        our C restatement (oracle/ckfft_oracle.c); always built.
  oracle/_ref/libckfft_ref.so      the UNMODIFIED reference, compiled from the sources where they
                                   lie under /root/reference (nine files of src/ckfft + KISS FFT
                                   1.3.0 from ext/kiss_fft130), plus our driver ref_driver.cpp.
  oracle/_ref/ckfft_test           the reference's own regression+timing harness (src/test/test.cpp).
  oracle/_ref/libfftw3_ref.so      vendored FFTW 3.3.2 (ext/fftw-3.3.2), double precision: the fp64
                                   truth oracle named by BASELINE.json's north_star.

The `_ref` targets are built only when /root/reference exists (this container).  Outputs go to
oracle/_ref/ only; that directory is git-ignored but NOT gpurun-ignored, so the binaries travel
to the GPU box, where /root/reference does not exist.  No reference source is copied into the repo.

Flags: -O2 -DNDEBUG -ffp-contract=off, so the scalar arithmetic is identical on every x86-64 host
and bit-comparable with oracle/ckfft_oracle.c.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CKFFT_REFERENCE", "/root/reference")
OUT_REF = os.path.join(HERE, "_ref")
CFLAGS = ["-O2", "-DNDEBUG", "-ffp-contract=off", "-fPIC", "-fopenmp"]

REF_LIB_SOURCES = [
    "src/ckfft/ckfft.cpp", "src/ckfft/context.cpp", "src/ckfft/debug.cpp", "src/ckfft/fft.cpp",
    "src/ckfft/fft_default.cpp", "src/ckfft/fft_neon.cpp", "src/ckfft/fft_real.cpp",
    "src/ckfft/fft_real_default.cpp", "src/ckfft/fft_real_neon.cpp",
]
KISS_SOURCES = ["ext/kiss_fft130/kiss_fft.c", "ext/kiss_fft130/tools/kiss_fftr.c"]
HARNESS_SOURCES = [
    "src/test/test.cpp", "src/test/stats.cpp", "src/test/timer.cpp", "src/test/timer_android.cpp",
    "src/test/macos/main.cpp", "src/test/macos/platform.cpp",
    "ext/tinyxml/tinyxml.cpp", "ext/tinyxml/tinystr.cpp", "ext/tinyxml/tinyxmlerror.cpp",
    "ext/tinyxml/tinyxmlparser.cpp",
]


def run(cmd, **kw):
    if os.environ.get("CKFFT_BUILD_VERBOSE"):
        print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd, **kw)


def newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources if os.path.exists(s))


def build_restatement():
    src = os.path.join(HERE, "ckfft_oracle.c")
    out = os.path.join(HERE, "libckfft_oracle.so")
    if newer(out, [src]):
        return out
    run(["gcc", "-std=c99", *CFLAGS, "-shared", "-o", out, src, "-lm"])
    return out


def ref_includes():
    shim = os.path.join(HERE, "shim")
    return [f"-I{shim}", f"-I{REF}/inc", f"-I{REF}/src", f"-I{REF}/ext", f"-I{REF}/ext/kiss_fft130",
            f"-I{REF}/ext/kiss_fft130/tools", f"-I{REF}/src/test"]


def build_reference_lib():
    os.makedirs(OUT_REF, exist_ok=True)
    out = os.path.join(OUT_REF, "libckfft_ref.so")
    driver = os.path.join(HERE, "ref_driver.cpp")
    srcs = [os.path.join(REF, s) for s in REF_LIB_SOURCES + KISS_SOURCES] + [driver]
    if newer(out, srcs):
        return out
    with tempfile.TemporaryDirectory() as tmp:
        objs = []
        for s in srcs:
            o = os.path.join(tmp, os.path.basename(s) + ".o")
            cc = "gcc" if s.endswith(".c") else "g++"
            # only our driver uses OpenMP; KISS FFT has its own (unwanted) _OPENMP code path
            flags = CFLAGS if s == driver else [f for f in CFLAGS if f != "-fopenmp"]
            run([cc, *flags, *ref_includes(), "-c", s, "-o", o])
            objs.append(o)
        # -Bsymbolic: the library's own CkFft* calls bind inside it even if the product
        # library (same symbol names) is loaded in the same process.
        run(["g++", "-shared", "-fopenmp", "-Wl,-Bsymbolic", "-o", out, *objs, "-lm"])
    return out


def build_reference_harness():
    os.makedirs(OUT_REF, exist_ok=True)
    out = os.path.join(OUT_REF, "ckfft_test")
    srcs = [os.path.join(REF, s) for s in REF_LIB_SOURCES + KISS_SOURCES + HARNESS_SOURCES]
    if newer(out, srcs):
        return out
    with tempfile.TemporaryDirectory() as tmp:
        objs = []
        for s in srcs:
            o = os.path.join(tmp, s.replace("/", "_") + ".o")
            cc = "gcc" if s.endswith(".c") else "g++"
            run([cc, "-O2", "-DNDEBUG", "-ffp-contract=off", "-w", *ref_includes(), "-c", s, "-o", o])
            objs.append(o)
        run(["g++", "-o", out, *objs, "-lm", "-lrt"])
    return out


def build_b200_harness():
    """The reference's own harness (regression vs KISS FFT + timing tables), linked against the PRODUCT library:
    proves the drop-in claim with the reference's own caller.  Needs ckfft_b200/lib/libckfft_b200.so."""
    os.makedirs(OUT_REF, exist_ok=True)
    out = os.path.join(OUT_REF, "ckfft_test_b200")
    root = os.path.dirname(HERE)
    lib_dir = os.path.join(root, "ckfft_b200", "lib")
    lib = os.path.join(lib_dir, "libckfft_b200.so")
    if not os.path.exists(lib):
        raise FileNotFoundError(lib)
    compat = os.path.join(HERE, "harness_compat.cpp")
    srcs = [os.path.join(REF, s) for s in ["src/ckfft/debug.cpp"] + KISS_SOURCES + HARNESS_SOURCES] + [compat]
    if newer(out, srcs + [lib]):
        return out
    with tempfile.TemporaryDirectory() as tmp:
        objs = []
        for s in srcs:
            o = os.path.join(tmp, s.replace("/", "_") + ".o")
            cc = "gcc" if s.endswith(".c") else "g++"
            run([cc, "-O2", "-DNDEBUG", "-w", *ref_includes(), "-c", s, "-o", o])
            objs.append(o)
        # $ORIGIN-relative rpath: the binary finds the library inside the repo snapshot on the GPU box too
        run(["g++", "-o", out, *objs, f"-L{lib_dir}", "-lckfft_b200", "-Wl,-rpath,$ORIGIN/../../ckfft_b200/lib", "-lm", "-lrt"])
    return out


def build_fftw():
    os.makedirs(OUT_REF, exist_ok=True)
    out = os.path.join(OUT_REF, "libfftw3_ref.so")
    if os.path.exists(out):
        return out
    src = os.path.join(REF, "ext/fftw-3.3.2")
    tmp = tempfile.mkdtemp(prefix="fftw_build_")
    try:
        log = open(os.path.join(tmp, "build.log"), "w")
        run(["sh", os.path.join(src, "configure"), "--disable-fortran", "--enable-shared=no", "--with-pic",
             "CFLAGS=-O2 -fPIC"], cwd=tmp, stdout=log, stderr=subprocess.STDOUT)
        run(["make", f"-j{os.cpu_count() or 4}"], cwd=tmp, stdout=log, stderr=subprocess.STDOUT)
        run(["gcc", "-shared", "-o", out, "-Wl,--whole-archive", os.path.join(tmp, ".libs/libfftw3.a"),
             "-Wl,--no-whole-archive", "-lm"])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return out


def build_all(with_fftw=True, with_harness=True, verbose=True):
    built = {"oracle": build_restatement()}
    if os.path.isdir(os.path.join(REF, "src/ckfft")):
        built["ref"] = build_reference_lib()
        if with_harness:
            try:
                built["harness"] = build_reference_harness()
            except subprocess.CalledProcessError as e:  # harness is optional
                print("reference harness build failed:", e, file=sys.stderr)
        try:
            built["harness_b200"] = build_b200_harness()
        except (subprocess.CalledProcessError, OSError) as e:
            print("reference harness (B200 link) build failed:", e, file=sys.stderr)
        if with_fftw:
            try:
                built["fftw"] = build_fftw()
            except (subprocess.CalledProcessError, OSError) as e:  # fp64 FFTW is optional (numpy fp64 is the fallback truth)
                print("FFTW build failed:", e, file=sys.stderr)
    elif verbose:
        print(f"{REF} not present: using prebuilt oracle/_ref binaries if any", file=sys.stderr)
    return built


if __name__ == "__main__":
    print(build_all(with_fftw="--no-fftw" not in sys.argv, with_harness="--no-harness" not in sys.argv))
