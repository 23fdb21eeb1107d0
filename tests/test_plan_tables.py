"""Host-side check of the single-pass plan tables (csrc/plans.h) against the hardware limits the kernels rely on.

Replays the shared-memory arithmetic of `Cfg` (csrc/fft_kernel.cuh) for every row of every plan list and checks that
  * the radices multiply to the length and fit a thread's registers (R <= E),
  * a CTA has at most 1024 threads and, where groups span several warps, at most 15 named barriers,
  * MINB CTAs of that plan fit the 228 KiB of shared memory of an SM (227 KiB per CTA, 1 KiB reserved per CTA),
  * MINB CTAs leave every thread at least 64 registers (65536 per SM),
so that a plan edited on a machine without a GPU cannot silently drop to one CTA per SM or fail to launch.
"""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TEXT = open(os.path.join(ROOT, "ckfft_b200", "csrc", "plans.h")).read()

PF_NONE, PF_DOUBLE, PF_INPLACE, PF_SPLIT = 0, 1, 2, 3
LISTS = {
    "CKB_SINGLE_PASS_PLANS": PF_NONE,
    "CKB_PREFETCH_PLANS": PF_DOUBLE,
    "CKB_INPLACE_PREFETCH_PLANS": PF_INPLACE,
    "CKB_INPLACE_PREFETCH_PLANS_PLANAR": PF_INPLACE,
    "CKB_INPLACE_PREFETCH_PLANS_C2R": PF_INPLACE,
    "CKB_INPLACE_PREFETCH_PLANS_AUDIO": PF_INPLACE,
    "CKB_SPLIT_PREFETCH_PLANS_C2C": PF_SPLIT,
    "CKB_SPLIT_PREFETCH_PLANS_R2C": PF_SPLIT,
    "CKB_SPLIT_PREFETCH_PLANS_AUDIO": PF_SPLIT,
    "CKB_SPLIT_PREFETCH_PLANS_PLANAR": PF_SPLIT,
}


def _rows(name):
    start = TEXT.index("#define " + name + "(X)")
    body = []
    for line in TEXT[start:].split("\n"):
        body.append(line)
        if not line.rstrip().endswith("\\"):
            break
    rows = re.findall(r"X\(\s*(\d+),\s*(\d+),\s*(\d+),\s*(\d+),\s*(\d+),\s*(\d+),\s*(\d+),\s*(\d+)\)", "\n".join(body))
    assert rows, name
    return [tuple(int(v) for v in r) for r in rows]


CASES = [(name, pf, row) for name, pf in LISTS.items() for row in _rows(name)]


@pytest.mark.parametrize("name,pf,row", CASES, ids=[f"{n[4:].lower()}-{r[0]}" for n, _, r in CASES])
def test_plan_fits_the_sm(name, pf, row):
    M, E, R0, R1, R2, G, MINB, TWR = row
    assert R0 * R1 * R2 == M and max(R0, R1, R2) <= E and M % E == 0
    T = M // E
    threads = G * T
    assert threads <= 1024
    assert T <= 32 or G <= 15, "named barriers 1..15"
    if TWR:
        assert R1 == E, "register stage twiddles need one stage-1 butterfly per thread"
    if pf == PF_SPLIT:
        assert R2 > 1 and E == R0, "split prefetch: three stages, one stage-0 butterfly per thread"
    real = "R2C" in name or "C2R" in name or "AUDIO" in name
    xraw = M + M // R0 + 2
    xbuf = xraw
    if real and T == 8:
        xbuf = xraw + (8 + 16 - xraw % 16) % 16
    if real and T == 4:
        xbuf = xraw + (12 + 16 - xraw % 16) % 16
    lut1 = 0 if TWR else (R1 - 1) * R0
    lut2 = (R2 - 1) * R0 * R1 if (R2 > 1 and (R2 - 1) * R0 * R1 <= 4096) else 0
    group = xbuf + (M if pf == PF_DOUBLE else M // 2 if pf == PF_SPLIT else 0)
    smem = 8 * (lut1 + lut2 + G * group) + (16 * G if pf != PF_NONE else 0)
    assert smem <= 227 * 1024, (name, row, smem)
    assert MINB * (smem + 1024) <= 228 * 1024, (name, row, smem, "MINB CTAs do not fit the SM's shared memory")
    assert MINB * threads * 64 <= 65536, (name, row, "fewer than 64 registers per thread at MINB CTAs per SM")
