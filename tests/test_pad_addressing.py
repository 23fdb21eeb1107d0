"""Host-side mirror of the padded exchange-buffer addressing of csrc/fft_kernel.cuh (stage_gather / stage_scatter).

The kernels address a butterfly's R values as one run-time base plus compile-time offsets, using
    padidx(a + c) == padidx(a) + padoff(c)      for c a multiple of the padding period 2^LOGPAD.
This test replays that arithmetic for every row of the single-pass plan table (csrc/plans.h) and checks it against the
defining formula p -> p + (p >> LOGPAD), including the mirror-paired map of the real-forward last stage.
"""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _plans():
    text = open(os.path.join(ROOT, "ckfft_b200", "csrc", "plans.h")).read()
    block = text[text.index("#define CKB_SINGLE_PASS_PLANS"):text.index("// Bulk-prefetch variants")]
    rows = re.findall(r"X\((\d+),\s*(\d+),\s*(\d+),\s*(\d+),\s*(\d+),", block)
    assert len(rows) == 11
    return [tuple(int(v) for v in r) for r in rows]


def _ilog2(x):
    return x.bit_length() - 1


def _padidx(p, lg):
    return p + (p >> lg)


def _padoff(c, lg):
    return c + (c >> lg)


def _bfly_index(j, q, T, STR, paired):
    if not paired:
        return j + q * T
    p = j + (q >> 1) * T
    if q & 1 == 0:
        return p
    return STR // 2 if p == 0 else STR - p


@pytest.mark.parametrize("plan", _plans(), ids=lambda p: f"M{p[0]}")
def test_linear_addressing_matches_padidx(plan):
    M, E, R0, R1, R2 = plan
    T, lg = M // E, _ilog2(R0)
    padw = 1 << lg
    stages = [(R0, 1), (R1, R0)] + ([(R2, R0 * R1)] if R2 > 1 else [])
    for si, (R, NS) in enumerate(stages):
        B, STR = E // R, M // R
        last = si == len(stages) - 1
        for paired in ((False, True) if (last and B % 2 == 0 and si > 0) else (False,)):
            for j in range(T):
                for q in range(B):
                    jq = _bfly_index(j, q, T, STR, paired)
                    if si > 0 and STR % padw == 0:           # stage_gather, SRC_XBUF
                        if T % padw != 0:
                            base = _padidx(jq, lg)
                        elif not paired or q % 2 == 0:
                            base = _padidx(j, lg) + _padoff((q >> 1 if paired else q) * T, lg)
                        elif q == 1:
                            base = _padidx(jq, lg)
                        else:
                            base = _padidx(STR - j, lg) - _padoff((q >> 1) * T, lg)
                        for t in range(R):
                            assert base + _padoff(t * STR, lg) == _padidx(jq + t * STR, lg), (M, si, paired, j, q, t)
                    if paired:
                        continue
                    jq = j + q * T
                    if not last:                              # stage_scatter, DST_XCHG
                        for u in range(R):
                            want = _padidx((jq // NS) * (NS * R) + (jq & (NS - 1)) + u * NS, lg)
                            if NS == 1 and R == padw:
                                got = j * (R + 1) + q * T * (R + 1) + u
                            elif NS > 1 and NS % padw == 0:
                                if T % NS == 0:
                                    got = _padidx((j // NS) * (NS * R) + (j & (NS - 1)), lg) + _padoff(q * T * R, lg) + _padoff(u * NS, lg)
                                else:
                                    got = _padidx((jq // NS) * (NS * R) + (jq & (NS - 1)), lg) + _padoff(u * NS, lg)
                            else:
                                got = want
                            assert got == want, (M, si, j, q, u)
                    if STR % padw == 0:                       # stage_scatter, DST_XNAT
                        base = _padidx(j, lg) + _padoff(q * T, lg) if T % padw == 0 else _padidx(jq, lg)
                        for u in range(R):
                            assert base + _padoff(u * STR, lg) == _padidx(jq + u * STR, lg)
