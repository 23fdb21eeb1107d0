"""CPU test of the host-side staging copy of the pageable path (ckfft_b200/csrc/host_copy.cpp: non-temporal stores, chosen at
run time among SSE2 / AVX2 / AVX-512): byte-exact against memcpy for every level, awkward sizes and alignments, and nothing
outside the destination range is touched.  Plain host C++ -- no GPU, no CUDA runtime."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = tmp_path / "host_copy_check"
    subprocess.check_call(["g++", "-O2", "-Wall", "-Werror", "-o", str(exe),
                           os.path.join(ROOT, "tests", "c", "host_copy_check.cpp"),
                           os.path.join(ROOT, "ckfft_b200", "csrc", "host_copy.cpp")])
    return exe


def test_stream_copy_matches_memcpy(tmp_path):
    out = subprocess.run([str(_build(tmp_path))], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.startswith("ok default=1"), out.stdout
    assert "0: memcpy" in out.stdout and "non-temporal" in out.stdout


def test_stream_copy_level_follows_the_environment(tmp_path):
    exe = _build(tmp_path)
    for value, want in (("0", 0), ("2", 2), ("", 1), ("9", 1)):
        out = subprocess.run([str(exe), "level"], capture_output=True, text=True, timeout=60, env={**os.environ, "CKFFT_B200_NT_COPY": value})
        assert out.returncode == 0 and out.stdout.startswith(f"ok default={want}"), out.stdout
