#!/usr/bin/env python3
"""Generate tests/golden/ckfft_golden.npz from the UNMODIFIED reference compiled here.

Run in the build container only (needs /root/reference and oracle/_ref/libckfft_ref.so, see
oracle/build.py).  The committed .npz is what travels: nothing at test time reads /root/reference.

Contents (all produced by the reference's scalar path, -O2 -ffp-contract=off, x86-64):
  input              complex64[4096]   the reference's own fixture src/test/input.txt
                                       (read exactly like src/test/test.cpp:90-107)
  cfwd_<n>, cinv_<n> complex64[n]      CkFftComplexForward / CkFftComplexInverse of input[:n], n = 1..4096
                                       (the cases of regressionTestComplex, src/test/test.cpp:754-786,879-894)
  kfwd_<n>, kinv_<n> complex64[n]      KISS FFT 1.3.0 on the same data (the harness's own oracle, RMS <= 1e-3)
  rfwd_<n>           complex64[n/2+1]  CkFftRealForward of input[:n].real (src/test/test.cpp:788-864,897-918)
  rinv_<n>           float32[n]        CkFftRealInverse of the complex-forward spectrum of the real input
                                       (the harness feeds refOutput, test.cpp:846-848)
  example_in         float32[1024]     the hard-coded signal of src/example/main.cpp:7-43
  example_fwd / example_rt             its CkFftRealForward and the CkFftRealInverse of that (= 2048 * input)
Context variants maxCount == n and maxCount == 8192 give bit-identical results (asserted here), as in
the harness's two passes (test.cpp:886,891).
"""
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

REF = "/root/reference"


def main():
    vals = np.array(open(os.path.join(REF, "src/test/input.txt")).read().split(), dtype=np.float32)
    inp = vals.view(np.complex64).copy()
    assert inp.shape == (4096,)
    out = {"input": inp}
    n = 4096
    while n >= 1:
        x = inp[:n]
        res = []
        for nmax in (n, 8192):
            R = oracle.Reference(nmax, 3)
            cf, ci = R.complex(x, False), R.complex(x, True)
            xr = np.ascontiguousarray(x.real)
            rf = R.real_forward(xr)
            spec = R.complex(xr.astype(np.complex64), False)[: n // 2 + 1]
            ri = R.real_inverse(spec, n)
            res.append((cf, ci, rf, ri))
            R.close()
        for a, b in zip(res[0], res[1]):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), n
        R = oracle.Reference(n, 3)
        out[f"cfwd_{n}"], out[f"cinv_{n}"], out[f"rfwd_{n}"], out[f"rinv_{n}"] = res[0]
        out[f"kfwd_{n}"], out[f"kinv_{n}"] = R.kiss(x, False), R.kiss(x, True)
        R.close()
        n //= 2
    src = open(os.path.join(REF, "src/example/main.cpp")).read()
    body = src[src.index("{", src.index("float input[]")) + 1: src.index("};")]
    ex = np.array([float(t.rstrip("f")) for t in re.findall(r"-?\d[\d.eE+-]*f", body)], np.float32)
    assert ex.shape == (1024,), ex.shape
    R = oracle.Reference(1024, 3)
    out["example_in"] = ex
    out["example_fwd"] = R.real_forward(ex)
    out["example_rt"] = R.real_inverse(out["example_fwd"], 1024)
    R.close()
    path = os.path.join(ROOT, "tests/golden/ckfft_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
