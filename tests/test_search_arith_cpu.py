"""The arithmetic of the encode-P search levels (pretty_fast_video_b200/csrc/pfv_kernels.cu: search_level, search_level_fine),
restated with Python integers and held against the plain definition of the reference's error (src/common.rs:125-139: the sum of
squared differences over the 16x16 block) - no GPU needed.  What is checked:

  * sum (a - b)^2 = sum a^2 + sum b^2 - 2 sum a b with sum b^2 of the three vertical candidates taken as differences of ONE
    running prefix down the column (the snapshots t_s, t_2s, t_16, t_16s of the kernel);
  * the fine levels' byte extraction: three aligned words per row, brought to the left-most candidate's alignment by two funnel
    shifts (X0, X1), the candidate word of horizontal offset mx by one more, clamped, funnel shift of (mx + 1) * STEP bytes -
    for every alignment of the column and both steps.

The -m gpu parity tests (tests/test_gpu_parity.py -k encode_p) prove the kernel itself, bit for bit, against the oracle.
"""
import random

import pytest

WIN_W, WIN_H = 160, 46          # EP2_WIN_W, WIN_H (pfv_internal.h)
M32 = 0xFFFFFFFF


def word_at(win, off):
    """aligned little-endian 32-bit load (off % 4 == 0); bytes past the window read as the kernel's stage padding would: anything"""
    return sum((win[off + k] if off + k < len(win) else 0xA5) << (8 * k) for k in range(4))


def funnel_r(lo, hi, shift):     # __funnelshift_r: shift & 31
    return (((hi << 32) | lo) >> (shift & 31)) & M32


def funnel_rc(lo, hi, shift):    # __funnelshift_rc: shift clamped to 32
    return (((hi << 32) | lo) >> min(shift, 32)) & M32


def dp4a(a, b, c):
    return (c + sum(((a >> (8 * k)) & 255) * ((b >> (8 * k)) & 255) for k in range(4))) & M32


def strip_ssd(win, S, off):
    """the plain definition: squared differences of the 4-pixel strip whose top-left byte is at window offset `off`"""
    e = 0
    for r in range(16):
        for k in range(4):
            d = ((S[r] >> (8 * k)) & 255) - win[off + r * WIN_W + k]
            e += d * d
    return e


def level_like_kernel(win, S, base, step, fine):
    """the body of search_level / search_level_fine for one lane: {(mx, my): this strip's share of the candidate's error}"""
    A = 0
    for r in range(16):
        A = dp4a(S[r], S[r], A)
    rows = 16 + 2 * step
    out = {}
    if fine:
        col0 = base - step - step * WIN_W
        sh = (col0 & 3) * 8
        X0, X1 = [], []
        for rho in range(rows):
            a = (col0 & ~3) + rho * WIN_W
            w0, w1, w2 = word_at(win, a), word_at(win, a + 4), word_at(win, a + 8)
            X0.append(funnel_r(w0, w1, sh))
            X1.append(funnel_r(w1, w2, sh))
    for mx in (-1, 0, 1):
        cu = cm = cd = 0
        t = t_s = t_2s = t_16 = t_16s = 0
        col = base + mx * step - step * WIN_W
        for rho in range(rows):
            if fine:
                w = funnel_rc(X0[rho], X1[rho], (mx + 1) * step * 8)
            else:
                assert (col + rho * WIN_W) % 4 == 0
                w = word_at(win, col + rho * WIN_W)
            if rho == step:
                t_s = t
            if rho == 2 * step:
                t_2s = t
            if rho == 16:
                t_16 = t
            if rho == 16 + step:
                t_16s = t
            t = dp4a(w, w, t)
            if rho < 16:
                cu = dp4a(S[rho], w, cu)
            if step <= rho < 16 + step:
                cm = dp4a(S[rho - step], w, cm)
            if rho >= 2 * step:
                cd = dp4a(S[rho - 2 * step], w, cd)
        out[(mx, -1)] = (A + t_16 - 2 * cu) & M32
        out[(mx, 0)] = (A + (t_16s - t_s) - 2 * cm) & M32
        out[(mx, 1)] = (A + (t - t_2s) - 2 * cd) & M32
    return out


@pytest.mark.parametrize("step,fine", [(8, False), (4, False), (2, True), (1, True)])
def test_search_level_arithmetic_equals_the_plain_sum_of_squared_differences(step, fine):
    rng = random.Random(1000 + step)
    for trial in range(60):
        extreme = trial % 5 == 0                                    # all-0 / all-255 bytes: the largest sums
        win = [rng.choice((0, 255)) if extreme else rng.randrange(256) for _ in range(WIN_W * WIN_H)]
        S = [sum((rng.choice((0, 255)) if extreme else rng.randrange(256)) << (8 * k) for k in range(4)) for _ in range(16)]
        m8, q = rng.randrange(8), rng.randrange(4)
        # a centre the earlier levels can have left: a multiple of 2 * step no further out than those levels reach (8, 12, 14)
        lim = 16 - 2 * step
        grid = [v for v in range(-lim, lim + 1) if v % (2 * step) == 0]
        cx, cy = rng.choice(grid), rng.choice(grid)
        base = (15 + cy) * WIN_W + 16 + m8 * 16 + q * 4 + cx         # o0 + cy * EP2_WIN_W + cx
        got = level_like_kernel(win, S, base, step, fine)
        for (mx, my), v in got.items():
            want = strip_ssd(win, S, base + mx * step + my * step * WIN_W)
            assert v == want, (step, trial, mx, my, cx, cy, m8, q)


def test_every_alignment_of_the_fine_levels():
    rng = random.Random(7)
    win = [rng.randrange(256) for _ in range(WIN_W * WIN_H)]
    S = [rng.getrandbits(32) for _ in range(16)]
    for step in (1, 2):
        for cx in range(-13, 14):                                    # every byte alignment of the column, both parities
            base = 15 * WIN_W + 16 + 3 * 16 + 2 * 4 + cx
            got = level_like_kernel(win, S, base, step, True)
            for (mx, my), v in got.items():
                assert v == strip_ssd(win, S, base + mx * step + my * step * WIN_W), (step, cx, mx, my)
