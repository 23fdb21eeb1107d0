"""CPU model of the shared-memory bank mapping of the exchange buffers (8-byte accesses: a warp request is served as two
128-byte wavefronts of 16 lanes; it is conflict-free when the 16 word addresses of each half are distinct mod 16).
Restates the index arithmetic of four_step.cuh / pipe_kernel.cuh (tile plans, along-the-columns thread map) and of
small_kernel.cuh (row pitch) -- the check that would have caught the 2-way conflict of the first 8-column pitch
(DESIGN.md 3.3)."""
import re
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TILE_SRC = open(os.path.join(ROOT, "ckfft_b200", "csrc", "tile_launch.h")).read()
PIPE_SRC = open(os.path.join(ROOT, "ckfft_b200", "csrc", "four_step.cu")).read()


def tile_plans():
    plans = set()
    for m in re.finditer(r"X\((\d+),\s*(\d+),\s*(\d+),\s*(\d+),\s*(\d+),\s*(\d+)\)", TILE_SRC):
        L, E, R0, R1, C, _ = map(int, m.groups())
        plans.add((L, E, R0, R1, C))
    for m in re.finditer(r"TileCfg<(\d+), (\d+), (\d+), (\d+), (\d+), INV, KIND", PIPE_SRC):
        plans.add(tuple(map(int, m.groups())))
    assert len(plans) >= 4
    return sorted(plans)


def xbuf(L, R0, C):
    """TileCfg::XBUF (four_step.cuh)"""
    xraw = L + L // R0 + 1
    return (xraw | 1) if C >= 16 else xraw + ((6 - xraw % 4) % 4)


def conflict_free(addresses):
    for half in (addresses[:16], addresses[16:]):
        if len({a % 16 for a in half}) != len(half):
            return False
    return True


def test_tile_exchange_is_conflict_free_for_every_plan():
    for L, E, R0, R1, C in tile_plans():
        T, X = L // E, xbuf(L, R0, C)
        pad = lambda p: p + p // R0
        threads = C * T
        for warp in range(threads // 32):
            lanes = range(32 * warp, 32 * warp + 32)
            for q in range(E // R0):              # stage-0 scatter (column pass): p = jq*R0 + u
                for u in (0, 1, R0 - 1):
                    a = [(t % C) * X + pad(((t // C) + q * T) * R0 + u) for t in lanes]
                    assert conflict_free(a), ("scatter", L, C, warp, q, u)
            for q in range(E // R1):              # stage-1 gather: jq + t*STR1
                for tt in (0, 1, R1 - 1):
                    a = [(t % C) * X + pad((t // C) + q * T + tt * (L // R1)) for t in lanes]
                    assert conflict_free(a), ("gather", L, C, warp, q, tt)


def test_the_first_pitch_of_the_8_column_tiles_was_a_two_way_conflict():
    L, R0, C, T = 1024, 32, 8, 32
    X = (L + L // R0 + 1) | 1                      # the odd pitch used before
    a = [(t % C) * X + (((t // C) * R0) + ((t // C) * R0) // R0) for t in range(32)]
    assert not conflict_free(a)


def test_small_kernel_row_pitches():
    for M, tpr in ((8, 1), (16, 1), (32, 1), (64, 2)):
        pitch = ((M + 1) | 1) if tpr == 1 else M + 2          # SmallCfg::PITCH
        for c in (0, 1, M - 1):
            if tpr == 1:
                a = [r * pitch + c for r in range(32)]
            else:
                a = [(t >> 1) * pitch + 2 * (c // 2) + (t & 1) for t in range(32)]
            assert conflict_free(a), (M, c)


def test_cooperative_plans_group_pitch_is_conflict_free():
    """fft_kernel.cuh Cfg::XBUF: plans whose half-warps hold several groups (T = 4, 8) need a group pitch of 12 resp. 8 mod 16;
    the raw pitch M + M/R0 + 2 collided on 4-8 lanes per request (tools/bank_model.py restates the index arithmetic)."""
    import importlib.util
    import io
    import contextlib

    spec = importlib.util.spec_from_file_location("bank_model", os.path.join(ROOT, "tools", "bank_model.py"))
    mod = importlib.util.module_from_spec(spec)
    with contextlib.redirect_stdout(io.StringIO()):
        spec.loader.exec_module(mod)
    kernel_src = open(os.path.join(ROOT, "ckfft_b200", "csrc", "fft_kernel.cuh")).read()
    assert "XRAW + (8 + 16 - XRAW % 16) % 16" in kernel_src and "XRAW + (12 + 16 - XRAW % 16) % 16" in kernel_src      # real-mode kernels
    for (M, E, R0, R1, R2, G, MINB, TWR) in mod.plans:
        T, raw = M // E, M + M // R0 + 2
        X = raw + (8 + 16 - raw % 16) % 16 if T == 8 else raw + (12 + 16 - raw % 16) % 16 if T == 4 else raw
        assert mod.model(M, E, R0, R1, G, X) == 0.0, (M, X)
        assert X % 2 == 0                      # group buffers stay 16-byte aligned (bulk-copy destinations)
    assert mod.model(64, 8, 8, 8, 16, 74) > 0  # the raw pitch did collide
