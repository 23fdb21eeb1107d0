import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)          # test-only helpers (dist_replay.py)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # The tests that pin the oracle to the UNMODIFIED reference must not skip where the reference can be compiled: in the
    # build container (/root/reference present) the library is built on demand (nine source files, seconds; the recipe is
    # oracle/build.py).  On the GPU box the prebuilt oracle/_ref binaries travel with the snapshot.
    try:
        from oracle import build as oracle_build

        ref_lib = os.path.join(ROOT, "oracle", "_ref", "libckfft_ref.so")
        if not os.path.exists(ref_lib) and os.path.isdir(os.path.join(oracle_build.REF, "src", "ckfft")):
            oracle_build.build_reference_lib()
    except Exception as e:  # noqa: BLE001 -- the affected tests then report themselves as skipped
        print(f"conftest: could not build oracle/_ref/libckfft_ref.so: {e}", file=sys.stderr)


def rel_rms(a, b):
    """relative RMS error ||a-b|| / ||b|| in float64 (the metric of BASELINE.json's north_star)."""
    a = np.asarray(a).astype(np.complex128 if np.iscomplexobj(a) or np.iscomplexobj(b) else np.float64)
    b = np.asarray(b).astype(a.dtype)
    den = np.linalg.norm(b.ravel())
    return float(np.linalg.norm((a - b).ravel()) / (den if den > 0 else 1.0))


def tolerance(n):
    """north_star: relative RMS error <= 1e-6 * log2(N) (floor of one bit for n = 1, 2)."""
    return 1e-6 * max(1.0, np.log2(max(n, 2)))


def uniform_complex(rng, shape):
    return (rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)).astype(np.complex64)


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "ckfft_golden.npz"))
