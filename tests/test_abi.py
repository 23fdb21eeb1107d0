"""CPU tests of the drop-in boundary: the C-ABI library loads, exports exactly what include/*.h declare,
the headers are valid C, and the host-side logic that needs no GPU (argument checks, the size-query
protocol of CkFftInit, the launch planner) behaves like the reference (src/ckfft/ckfft.cpp:14-34,
src/ckfft/context.cpp:27-51).  No compute calls here."""
import ctypes as C
import os
import re
import subprocess

import pytest

import ckfft_b200 as ck
from ckfft_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")


def declared_functions(header):
    text = open(os.path.join(INC, "ckfft", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(CkFft\w+)\s*\(", text))


def exported_symbols():
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    return {line.split()[-1] for line in out.splitlines() if " T " in line}


def test_library_exports_every_declared_symbol():
    declared = declared_functions("ckfft.h") | declared_functions("ckfft_b200.h")
    assert declared == set(_lib.CLASSIC_SYMBOLS) | set(_lib.B200_SYMBOLS)
    exported = exported_symbols()
    missing = declared - exported
    assert not missing, f"declared but not exported: {missing}"
    extra = {s for s in exported if s.startswith("CkFft")} - declared
    assert not extra, f"exported but not declared: {extra}"


def test_integration_doc_names_every_entry_point():
    """INTEGRATION.md is the binding guide of the drop-in boundary: every function the headers declare appears in it
    (brace groups such as CkFft{ComplexForward,RealForward}Batch are expanded)."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    named = set(re.findall(r"\bCkFft\w+", text))
    for m in re.finditer(r"(CkFft\w*)\{([^}]*)\}(\w*)", text):
        named |= {m.group(1) + alt.strip() + m.group(3) for alt in m.group(2).split(",")}
    declared = declared_functions("ckfft.h") | declared_functions("ckfft_b200.h")
    assert not declared - named, f"not documented in INTEGRATION.md: {sorted(declared - named)}"


def test_classic_header_is_the_reference_surface():
    """the six functions of inc/ckfft/ckfft.h:59-158, nothing more"""
    assert declared_functions("ckfft.h") == {"CkFftInit", "CkFftRealForward", "CkFftRealInverse",
                                             "CkFftComplexForward", "CkFftComplexInverse", "CkFftShutdown"}


@pytest.mark.parametrize("header", ["ckfft.h", "ckfft_b200.h"])
def test_headers_are_valid_c(header, tmp_path):
    """the reference keeps its header C-clean with src/test/test.c; same check here"""
    src = tmp_path / "t.c"
    src.write_text(f'#include "ckfft/{header}"\nint main(void) {{ CkFftComplex c; c.real = 0; c.imag = 0; return (int) c.real; }}\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", f"-I{INC}", str(src)])


def test_python_binding_loads():
    lib = _lib.load()
    for name in _lib.CLASSIC_SYMBOLS + _lib.B200_SYMBOLS:
        assert hasattr(lib, name)
    assert isinstance(ck.kernel_launches(), int)


def test_init_argument_checks():
    """src/ckfft/ckfft.cpp:16-31: every invalid argument gives NULL before anything else happens."""
    lib = _lib.load()
    for nmax in (0, -4, 3, 1000, -(2 ** 31)):
        assert not lib.CkFftInit(nmax, ck.BOTH, None, None)
    for direction in (0, 4, 7, -1):
        assert not lib.CkFftInit(1024, direction, None, None)
    buf = C.create_string_buffer(1 << 20)
    assert not lib.CkFftInit(1024, ck.BOTH, buf, None)      # buf without bufSize
    assert ck.last_error() != ""


def test_size_query_protocol():
    """src/ckfft/context.cpp:47-51 and the usage in inc/ckfft/ckfft.h:51-55: a NULL or too-small buffer
    makes CkFftInit write the required size and return NULL."""
    lib = _lib.load()
    sizes = {}
    for direction in (ck.FORWARD, ck.INVERSE, ck.BOTH):
        sz = C.c_size_t(0)
        assert not lib.CkFftInit(1024, direction, None, C.byref(sz))
        assert sz.value > 0
        sizes[direction] = sz.value
        small = C.create_string_buffer(16)
        sz2 = C.c_size_t(16)
        assert not lib.CkFftInit(1024, direction, small, C.byref(sz2))
        assert sz2.value == sz.value
    # one table per requested direction, 8 bytes per entry (context.cpp:34-45)
    assert sizes[ck.FORWARD] == sizes[ck.INVERSE]
    assert sizes[ck.BOTH] - sizes[ck.FORWARD] == 1024 * 8


def test_transform_calls_reject_null_context():
    lib = _lib.load()
    buf = C.create_string_buffer(64)
    assert lib.CkFftComplexForward(None, 4, buf, buf) == 0
    assert lib.CkFftRealInverse(None, 4, buf, buf, None) == 0
    assert lib.CkFftComplexForwardBatch(None, 4, buf, buf, 1) == 0
    lib.CkFftShutdown(None)   # NULL-safe like the reference (context.cpp:116-122)


def test_planner_covers_every_power_of_two():
    for lg in range(0, 15):
        n = 1 << lg
        p = ck.get_plan(n, real=False)
        assert p is not None and p["complex_points"] == n
        prod = 1
        for r in p["radix"][0]:
            prod *= r
        assert prod == n, p
        assert p["passes"] == 1
        assert p["threads_per_transform"] * p["elems_per_thread"] == n
        assert p["shared_bytes"] <= 227 * 1024
        q = ck.get_plan(2 * n, real=True)
        assert q is not None and q["complex_points"] == n
    for lg in range(15, 31):
        p = ck.get_plan(1 << lg)
        assert p is not None and p["passes"] == (2 if lg <= 20 else 3), (lg, p)
        if lg <= 20:
            prod = 1
            for row in p["radix"]:
                for r in row:
                    prod *= r
            assert prod == 1 << lg
        q = ck.get_plan(1 << lg, real=True)
        assert q is not None and q["complex_points"] == 1 << (lg - 1)
    assert ck.get_plan(24) is None and ck.get_plan(0) is None
    # the headline shapes: one warp per 1024-point transform, 32x32 radix split, 4 transforms per CTA
    p = ck.get_plan(1024)
    assert p["radix"] == [[32, 32]] and p["threads_per_transform"] == 32 and p["elems_per_thread"] == 32
    q = ck.get_plan(4096, real=True)
    assert q["complex_points"] == 2048
    # short rows (small_kernel.cuh): one register network per thread, 128 rows per CTA; 64 complex points: two threads per row
    for n in (8, 16, 32):
        p = ck.get_plan(n)
        assert p["radix"][0][0] == n and p["threads_per_transform"] == 1 and p["transforms_per_cta"] == 128, p
    p = ck.get_plan(64)
    assert p["radix"][0][:2] == [32, 2] and p["threads_per_transform"] == 2 and p["transforms_per_cta"] == 64, p
    q = ck.get_plan(64, real=True)
    assert q["complex_points"] == 32 and q["threads_per_transform"] == 1
    q = ck.get_plan(128, real=True)                      # real n = 128 stays on the cooperative kernel (8 x 8 radix split)
    assert q["complex_points"] == 64 and q["threads_per_transform"] == 8 and q["radix"][0][:2] == [8, 8], q


def test_no_gpu_fails_loudly():
    """Without a usable GPU the product must not compute anything: CkFftInit -> NULL with a reason."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(ck.CkFftError) as e:
        ck.Context(1024)
    assert "no CPU path" in str(e.value)


def test_shard_range_and_multi_init_checks_need_no_gpu():
    """CkFftB200ShardRange is pure host arithmetic; CkFftB200MultiInit validates its arguments like CkFftInit
    (src/ckfft/ckfft.cpp:16-31) before it touches a device."""
    lib = _lib.load()
    first, count = C.c_size_t(0), C.c_size_t(0)
    covered = 0
    for part in range(5):
        assert lib.CkFftB200ShardRange(13, part, 5, C.byref(first), C.byref(count)) == 1
        assert first.value == covered and count.value in (2, 3)
        covered += count.value
    assert covered == 13
    assert lib.CkFftB200ShardRange(13, 5, 5, C.byref(first), C.byref(count)) == 0
    assert lib.CkFftB200ShardRange(13, 0, 0, C.byref(first), C.byref(count)) == 0
    assert lib.CkFftB200ShardRange(13, 0, 2, None, C.byref(count)) == 0
    assert not lib.CkFftB200MultiInit(1000, ck.BOTH, None, 0)
    assert "power of two" in ck.last_error()
    assert not lib.CkFftB200MultiInit(1024, 5, None, 0)
    assert lib.CkFftB200MultiDeviceCount(None) == 0
    lib.CkFftB200MultiShutdown(None)          # NULL-safe like CkFftShutdown (src/ckfft/ckfft.cpp:116-119)
