#!/usr/bin/env python3
"""Bank model of the single-pass cooperative kernels' exchange buffer (fft_kernel.cuh), per plan of plans.h: average number
of conflicting lanes per 8-byte warp request for the current group pitch and the smallest conflict-free pitch.  No GPU needed.
Finding (round 1): plans with several groups per half-warp (M = 16 .. 128, T = 4 or 8) collide when the group pitch is the
raw M + M/R0 + 2 (6 or 10 mod 16); a pitch of 8 mod 16 (T = 8) / 12 mod 16 (T = 4) is free.  Applied in round 2 (Cfg::XBUF)."""
import os, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(ROOT, "ckfft_b200", "csrc", "plans.h")).read()
m = re.search(r"#define CKB_SINGLE_PASS_PLANS\(X\)(.*?)\n\n", src, re.S)
plans = [tuple(map(int, x)) for x in re.findall(r"X\((\d+),\s*(\d+),\s*(\d+),\s*(\d+),\s*(\d+),\s*(\d+),\s*(\d+),\s*(\d+)\)", m.group(1))]


def conflicts(addrs):
    return sum(len(h) - len({a % 16 for a in h}) for h in (addrs[:16], addrs[16:]))


def model(M, E, R0, R1, G, GS):
    T = M // E
    pad = lambda p: p + p // R0
    tot = cnt = 0
    for warp in range(max(1, G * T // 32)):
        lanes = [t for t in range(32 * warp, 32 * warp + 32) if t < G * T]
        if len(lanes) < 32:
            continue
        for q in range(E // R0):
            for u in range(R0):          # stage-0 scatter
                tot += conflicts([(t // T) * GS + pad(((t % T) + q * T) * R0 + u) for t in lanes]); cnt += 1
        for q in range(E // R1):
            for tt in range(R1):         # stage-1 gather
                tot += conflicts([(t // T) * GS + pad((t % T) + q * T + tt * (M // R1)) for t in lanes]); cnt += 1
    return tot / max(cnt, 1)


print("M      T  pitch  conflicting lanes/request   smallest free pitch (mod 16)   [raw pitch: conflicts]")
for (M, E, R0, R1, R2, G, MINB, TWR) in plans:
    T, raw = M // E, M + M // R0 + 2
    X = raw + (8 + 16 - raw % 16) % 16 if T == 8 else raw + (12 + 16 - raw % 16) % 16 if T == 4 else raw      # Cfg::XBUF (fft_kernel.cuh)
    best = min(range(X, X + 17), key=lambda gs: (model(M, E, R0, R1, G, gs), gs))
    print(f"{M:<6d} {T:<2d} {X:<6d} {model(M, E, R0, R1, G, X):<27.2f} {best} ({best % 16})   [{raw}: {model(M, E, R0, R1, G, raw):.2f}]")
