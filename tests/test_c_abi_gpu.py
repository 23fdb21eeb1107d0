"""The drop-in boundary used from plain C: tests/c/abi_smoke.c is compiled with gcc against include/ckfft and linked
to libckfft_b200.so, then run on the GPU (the reference's example program flow, src/example/main.cpp:44-84)."""
import os
import subprocess

import pytest

from ckfft_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


def test_c_program_links_and_runs(tmp_path):
    exe = tmp_path / "abi_smoke"
    lib_dir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-O1", f"-I{os.path.join(ROOT, 'include')}",
                           os.path.join(ROOT, "tests", "c", "abi_smoke.c"), "-o", str(exe),
                           f"-L{lib_dir}", "-lckfft_b200", f"-Wl,-rpath,{lib_dir}", "-lm"])
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "error: 0.000000" in r.stdout and r.stdout.strip().endswith("ok")
