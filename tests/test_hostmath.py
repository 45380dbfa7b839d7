"""The CUDA kernels' register-level arithmetic, compiled for the CPU from the kernels' own header
(pretty_fast_video_b200/csrc/pfv_dct.cuh -> tests/hostmath/libpfv_hostmath.so), against the oracle.

No GPU needed: transforms (src/dct.rs:176-293), the quantiser's reciprocal multiply (src/dct.rs:88-99), the sub-block
drivers (src/common.rs:287-325) and the run-length bookkeeping (src/rle.rs:9-39) are the same source the kernels
inline; the -m gpu tests are left with the memory side."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import pfvo

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hm():
    d = os.path.join(HERE, "hostmath")
    subprocess.check_call(["make", "-s", "-C", d])
    lib = C.CDLL(os.path.join(d, "libpfv_hostmath.so"))
    lib.pfv_hm_quant_check.restype = C.c_long
    lib.pfv_hm_quant_f32_check.restype = C.c_long
    lib.pfv_hm_quant_f32_floor_check.restype = C.c_long
    lib.pfv_hm_escapes_check.restype = C.c_long
    lib.pfv_hm_mb_entry_count.restype = C.c_uint32
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_quantiser_reciprocal_is_exact(hm):
    # every |n| the transforms can produce (|v * scale| >> 16 <= ~1100 needs |v| < 2^21.3; go well beyond), every divisor
    # the reference's tables can hold at any quality, then a sparse sweep up to the limit the engine accepts
    assert hm.pfv_hm_quant_check(-(1 << 22), 1 << 22, 37, 1, 512) == 0
    assert hm.pfv_hm_quant_check(-(1 << 22), 1 << 22, 4099, 513, 65535) == 0
    assert hm.pfv_hm_quant_check(-70000, 70000, 1, 1, 64) == 0          # dense around zero, incl. n = -1, 0, exact multiples
    assert hm.pfv_hm_escapes_check() == 0


def test_fp32_quantiser_is_exact(hm):
    """quant_one_f32 (the encode-I kernel's quantiser: floor by a round-down FMA, division by a float reciprocal a little
    above 1/q): every |n| <= 4096 (the transforms cannot produce more than ~1 700) against every divisor 1..65535, and the
    floor step over every transform output 256 V in +-2^22."""
    assert hm.pfv_hm_quant_f32_check(4096, 1, 65535) == 0
    assert hm.pfv_hm_quant_f32_floor_check(-(1 << 22), 1 << 22, 1) == 0


@pytest.mark.parametrize("quality", [0, 1, 2, 5, 8, 10])
def test_encode_subblock_matches_oracle(hm, quality):
    rng = np.random.default_rng(100 + quality)
    qt, _ = pfvo.make_qtables(quality)
    for trial in range(400):
        q = qt[trial % 2]                                   # intra_l, intra_c
        kind = trial % 4
        if kind == 0:
            px = rng.integers(0, 256, 64)
        elif kind == 1:
            px = np.clip(128 + rng.normal(0, 12, 64), 0, 255)
        elif kind == 2:
            px = np.where(rng.random(64) < 0.5, 0, 255)        # extremes: the largest transform outputs
        else:
            px = np.full(64, rng.integers(0, 256))
        px = px.astype(np.uint8)
        out = np.zeros(64, np.int16)
        hm.pfv_hm_encode_sb(_p(px), 0, _p(np.ascontiguousarray(q)), _p(out))
        assert np.array_equal(out, pfvo.encode_subblock(px, q)), (quality, trial)
        outf = np.zeros(64, np.int16)
        hm.pfv_hm_encode_sb_f32(_p(px), 0, _p(np.ascontiguousarray(q)), _p(outf))      # the fp32 formulation: same bits
        assert np.array_equal(outf, out), (quality, trial, "f32")
        hm.pfv_hm_encode_sb_f32_generic(_p(px), _p(np.ascontiguousarray(q)), _p(outf))
        assert np.array_equal(outf, out), (quality, trial, "f32 generic")


@pytest.mark.parametrize("quality", [0, 3, 5, 10])
def test_encode_subblock_delta_matches_oracle(hm, quality):
    rng = np.random.default_rng(200 + quality)
    qt, _ = pfvo.make_qtables(quality)
    for trial in range(400):
        q = qt[2 + trial % 2]                               # inter_l, inter_c
        d = rng.integers(-255, 256, 64) if trial % 3 else np.where(rng.random(64) < 0.5, -255, 255)
        if trial % 3 == 1:
            d = np.clip(rng.normal(0, 6, 64), -255, 255)
        d = d.astype(np.int16)
        out = np.zeros(64, np.int16)
        hm.pfv_hm_encode_sb(_p(d), 1, _p(np.ascontiguousarray(q)), _p(out))
        assert np.array_equal(out, pfvo.encode_subblock_delta(d, q)), (quality, trial)
        outf = np.zeros(64, np.int16)
        hm.pfv_hm_encode_sb_f32(_p(d), 1, _p(np.ascontiguousarray(q)), _p(outf))
        assert np.array_equal(outf, out), (quality, trial, "f32")


def test_encode_with_arbitrary_divisors(hm):
    """Divisors beyond what Encoder::new derives (1 .. 65535: what a .pfv header's u16 fields can carry)."""
    rng = np.random.default_rng(7)
    for trial in range(300):
        q = rng.integers(1, [2, 16, 300, 65536][trial % 4], 64).astype(np.int32)
        px = rng.integers(0, 256, 64).astype(np.uint8)
        out = np.zeros(64, np.int16)
        hm.pfv_hm_encode_sb(_p(px), 0, _p(q), _p(out))
        assert np.array_equal(out, pfvo.encode_subblock(px, q))
        outf = np.zeros(64, np.int16)
        hm.pfv_hm_encode_sb_f32(_p(px), 0, _p(q), _p(outf))
        assert np.array_equal(outf, out)


def test_decode_subblock_matches_oracle(hm):
    rng = np.random.default_rng(3)
    qt, _ = pfvo.make_qtables(5)
    for trial in range(600):
        q = qt[trial % 4]
        mode = trial % 3
        if mode == 0:
            c = rng.integers(-32768, 32768, 64)                # wrapping i32 arithmetic
        elif mode == 1:
            c = rng.integers(-300, 301, 64) * (rng.random(64) < 0.2)
        else:
            c = np.zeros(64, np.int64); c[0] = rng.integers(-2048, 2048)
        c = c.astype(np.int16)
        out = np.zeros(64, np.uint8)
        hm.pfv_hm_decode_sb(_p(c), _p(np.ascontiguousarray(q)), _p(out))
        assert np.array_equal(out, pfvo.decode_subblock(c, q)), trial
        out2 = np.zeros(64, np.uint8)
        hm.pfv_hm_decode_sb_rolled(_p(c), _p(np.ascontiguousarray(q)), _p(out2))
        assert np.array_equal(out2, out), (trial, "rolled")


def test_macroblock_entry_count_matches_rle_encode(hm):
    """sb_runs + mb_entry_count (what the encode kernels leave in mb_cnt for the sparse seam) == len(rle_encode(mb))."""
    rng = np.random.default_rng(11)
    cases = [np.zeros(256, np.int16)]
    for pos in (0, 1, 15, 16, 17, 31, 32, 63, 64, 79, 80, 127, 128, 240, 241, 254, 255):
        c = np.zeros(256, np.int16); c[pos] = 5
        cases.append(c)
    for gap in range(1, 70):                                   # every inner gap around the 15-zero escape thresholds
        for start in (0, 3, 60, 64, 100, 190):
            if start + gap + 1 < 256:
                c = np.zeros(256, np.int16); c[start] = 1; c[start + gap + 1] = -1
                cases.append(c)
    for dens in (0.002, 0.01, 0.03, 0.1, 0.3, 0.7, 1.0):
        for _ in range(150):
            c = (rng.integers(-40, 41, 256) * (rng.random(256) < dens)).astype(np.int16)
            cases.append(c)
    for c in cases:
        tok, _, _ = pfvo.rle_frame(c.reshape(1, 256))
        assert hm.pfv_hm_mb_entry_count(_p(np.ascontiguousarray(c))) == tok.size, np.flatnonzero(c)
