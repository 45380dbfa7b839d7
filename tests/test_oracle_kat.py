"""CPU: pins the C oracle (oracle/pfv_oracle.c) against (i) the derived KATs of SURVEY.md Appendix C,
(ii) the independent Python restatement oracle/pfv_ref.py, (iii) the committed golden fixtures under
tests/golden/ and (iv) the reference's own asserting tests (src/lib.rs:96-158 entropy round trip).
The reference publishes no golden outputs for this path ("parity unpinned")."""
import importlib.util
import os

import numpy as np
import pytest

import pfvo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("pfv_ref", os.path.join(ROOT, "oracle", "pfv_ref.py"))
ref = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(ref)


def test_tables_match_reference_listing():
    assert list(pfvo.table("pfvo_dct_scale_factor", np.int32)) == ref.DCT_SCALE_FACTOR
    assert list(pfvo.table("pfvo_q_table_intra", np.int32)) == ref.Q_TABLE_INTRA
    assert list(pfvo.table("pfvo_q_table_inter", np.int32)) == ref.Q_TABLE_INTER
    zz, izz = list(pfvo.table("pfvo_zigzag", np.uint8)), list(pfvo.table("pfvo_inv_zigzag", np.uint8))
    assert zz == ref.ZIGZAG_TABLE and izz == ref.INV_ZIGZAG_TABLE
    assert all(izz[zz[i]] == i for i in range(64))        # exact inverses


def test_kat_b_fdct_idct():
    v = [x << 8 for x in (0, 10, 20, 30, 40, 50, 60, 70)]  # src/lib.rs:38
    f = pfvo.fdct8(v)
    assert list(f) == [71680, -40000, 0, -6400, 0, -1280, 0, -320]
    assert list(pfvo.idct8(f)) == [13700, 36060, 54260, 66220, 77140, 89100, 107300, 129660]


def test_kat_c_negative_truncation():
    v = [-256, 768, -19712, 32512, -32768, 5, -5, 1]
    assert list(pfvo.fdct8(v)) == [-19455, 3232, 10236, -28537, 18433, 86133, -25587, -97833]
    assert list(pfvo.idct8(v)) == [-10013, 13262, -11972, -38665, 21901, 96696, 32062, -105319]


def test_kat_qtables():
    q0, _ = pfvo.make_qtables(0)
    assert (q0 == 1).all()
    q2, px2 = pfvo.make_qtables(2)
    assert list(q2[0][:8]) == [2, 4, 4, 5, 6, 6, 7, 8] and list(q2[1][:8]) == [4, 8, 9, 11, 13, 13, 14, 17]
    assert (q2[2] == 4).all() and (q2[3] == 8).all() and px2 == 3.0
    q5, px5 = pfvo.make_qtables(5)
    assert list(q5[0][:8]) == [5, 10, 11, 13, 16, 16, 18, 21] and list(q5[1][:8]) == [10, 20, 23, 27, 32, 33, 36, 42]
    assert (q5[2] == 10).all() and (q5[3] == 20).all() and px5 == 7.5
    q10, _ = pfvo.make_qtables(10)
    assert list(q10[0]) == list(q5[1]) and list(q10[1][:8]) == [20, 40, 47, 55, 65, 67, 72, 85]
    for q in range(11):
        assert pfvo.make_qtables(q)[0].tolist() == ref.make_qtables(q)


def test_kat_a_intra_block():
    """src/lib.rs:57-94 test_dct_encode's block with the intra-luma table at quality 5."""
    q = pfvo.make_qtables(5)[0][0]
    golden = np.load(os.path.join(ROOT, "tests", "golden", "kat.npz"))
    px = golden["kat_a_px"]
    c = pfvo.encode_subblock(px, q)
    assert list(c) == list(golden["kat_a_coeff"])
    assert list(c[:12]) == [-150, 2, 5, 0, 0, -2, 0, 0, 0, 0, 0, -1] and not c[12:].any()
    d = pfvo.decode_subblock(c, q)
    assert list(d) == [39, 41, 44, 47, 46, 44, 41, 38, 41, 43, 45, 46, 44, 40, 36, 33, 42, 43, 45, 44, 41, 36, 31, 28,
                       37, 39, 40, 41, 38, 34, 29, 26, 29, 31, 34, 36, 36, 33, 29, 26, 24, 27, 30, 33, 33, 30, 27, 25,
                       25, 27, 30, 31, 30, 26, 22, 20, 28, 29, 30, 31, 28, 23, 18, 14]


def test_kat_d_delta_block():
    q = pfvo.make_qtables(5)[0][2]
    src = np.array([(7 * i + 3) % 256 for i in range(64)], np.int16)
    prev = np.array([(5 * i + 40) % 256 for i in range(64)], np.int16)
    delta = np.clip(src - prev, -255, 255).astype(np.int16)
    c = pfvo.encode_subblock_delta(delta, q)
    assert list(c) == [0, -2, -7, 9, 3, -2, 1, 0, -4, -14, 1, -4, 2, -1, 1, -1, 0, 0, -1, 10, 6, -6, -4, -2, 2, -2, 1, 0,
                       2, 0, 1, 1, -2, 2, -10, 0, 13, 1, 0, 1, -2, 1, -1, -2, 0, 0, -1, 2, -2, -3, 0, 1, 0, 2, 0, 1, -1,
                       2, 0, 0, -1, 0, -1, 0]
    d = pfvo.decode_subblock(c, q)
    recon = ref.apply_residuals([int(x) for x in d], [int(x) for x in prev])
    assert recon == [0, 0, 0, 0, 18, 27, 18, 47, 78, 81, 86, 81, 96, 97, 104, 81, 158, 165, 162, 189, 176, 177, 196, 215,
                     214, 203, 224, 223, 222, 243, 238, 237, 180, 199, 204, 209, 224, 35, 24, 53, 48, 45, 34, 55, 2, 23,
                     8, 13, 80, 71, 86, 79, 84, 105, 116, 131, 144, 161, 168, 171, 176, 191, 190, 195]


@pytest.mark.parametrize("seed", range(6))
def test_c_oracle_equals_python_restatement(seed):
    rng = np.random.default_rng(seed)
    for quality in (0, 2, 5, 10):
        qt = pfvo.make_qtables(quality)[0]
        for t in range(4):
            q = qt[t]
            v = rng.integers(-2 ** 31, 2 ** 31, 8)
            assert list(pfvo.fdct8(v)) == ref.fdct([int(x) for x in v])
            assert list(pfvo.idct8(v)) == ref.idct([int(x) for x in v])
            px = rng.integers(0, 256, 64).astype(np.uint8)
            assert list(pfvo.encode_subblock(px, q)) == ref.encode_subblock([int(x) for x in px], list(q))
            d = rng.integers(-255, 256, 64).astype(np.int16)
            assert list(pfvo.encode_subblock_delta(d, q)) == ref.encode_subblock_delta([int(x) for x in d], list(q))
            hi = (64, 2048, 32768)[seed % 3]
            c = rng.integers(-hi, hi, 64).astype(np.int16)
            assert list(pfvo.decode_subblock(c, q)) == ref.decode_subblock([int(x) for x in c], [int(x) for x in q])


def test_block_search_equals_python_restatement():
    import ctypes as C
    rng = np.random.default_rng(11)
    rw, rh = 64, 48
    base = rng.integers(0, 256, (rh + 32, rw + 32)).astype(np.uint8)
    for trial in range(12):
        refp = np.ascontiguousarray(base[16:16 + rh, 16:16 + rw])
        bx, by = int(rng.integers(0, rw // 16)) * 16, int(rng.integers(0, rh // 16)) * 16
        ox, oy = int(rng.integers(-9, 10)), int(rng.integers(-9, 10))
        src = np.ascontiguousarray(base[16 + by + oy:32 + by + oy, 16 + bx + ox:32 + bx + ox]).copy()
        if trial % 3 == 0:
            src = np.clip(src.astype(int) + rng.integers(-20, 21, src.shape), 0, 255).astype(np.uint8)
        if trial % 4 == 3:
            refp[:] = 77; src[:] = 77                         # all ties: the centre must win
        dx, dy = C.c_int(), C.c_int()
        best = np.zeros(256, np.uint8)
        err = pfvo.lib().pfvo_block_search(pfvo._p(src), pfvo._p(refp), C.c_int(rw), C.c_int(rh), C.c_int(bx), C.c_int(by),
                                           C.c_int(8), C.byref(dx), C.byref(dy), pfvo._p(best))
        want = ref.block_search([int(x) for x in src.ravel()], [int(x) for x in refp.ravel()], rw, rh, bx, by, 8)
        assert (dx.value, dy.value, float(err)) == want
        assert np.array_equal(best.reshape(16, 16), refp[by + dy.value:by + dy.value + 16, bx + dx.value:bx + dx.value + 16])


def test_reference_entropy_roundtrip():
    """src/lib.rs:96-158 test_entropy: the reference's only asserting test, same literal input."""
    data = [10, 0, 0, -5, 2, 0, 0, 0, 0, 0] if True else None
    rc, out = pfvo.entropy_roundtrip(np.array(data, np.int16))
    assert rc > 0 and list(out) == data
    rng = np.random.default_rng(3)
    big = rng.integers(-300, 301, 589824 // 2).astype(np.int16)      # src/lib.rs:160-239 uses a 589 824-byte frame
    big[rng.random(big.size) < 0.9] = 0
    rc, out = pfvo.entropy_roundtrip(big)
    assert rc > 0 and np.array_equal(out, big)


def test_oracle_stream_matches_golden_fixture():
    """Golden stream committed with its generator (tests/golden/make_golden.py): the oracle encoder must
    reproduce the bytes, and the oracle decoder the per-frame checksums."""
    import hashlib
    from pretty_fast_video_b200.synth import SynthVideo
    g = np.load(os.path.join(ROOT, "tests", "golden", "stream_96x64_q3.npz"))
    sv = SynthVideo(96, 64, 77)
    enc = pfvo.Encoder(96, 64, 24, 3, nthreads=2)
    for t in range(int(g["nframes"])):
        y, u, v = sv.frame(t)
        if t % 4 == 0:
            enc.encode_iframe(y, u, v)
        elif t == 5:
            enc.encode_dropframe()
        else:
            enc.encode_pframe(y, u, v)
    enc.finish()
    data = enc.bytes()
    assert data == g["stream"].tobytes()
    dec = pfvo.Decoder(data)
    assert (dec.width, dec.height, dec.framerate) == (96, 64, 24)
    sums = []
    while True:
        more, fr = dec.advance_frame()
        if fr is not None:
            sums.append(hashlib.sha256(b"".join(p.tobytes() for p in fr)).hexdigest())
        if not more:
            break
    assert sums == [s for s in g["sha256"]]


def test_decoder_header_errors():
    enc = pfvo.Encoder(32, 32, 30, 5)
    enc.finish()
    good = enc.bytes()
    assert len(good) == 532 + 5
    with pytest.raises(ValueError, match="FormatError"):
        pfvo.Decoder(b"NOTAPFV\0" + good[8:])
    with pytest.raises(ValueError, match="VersionError"):
        pfvo.Decoder(good[:8] + (210).to_bytes(4, "little") + good[12:])
    with pytest.raises(ValueError, match="IOError"):
        pfvo.Decoder(good[:100])


def test_rle_restatements_agree():
    """rle_encode + update_table (src/rle.rs:9-47): the C oracle and the Python restatement give the same sequence and symbol
    counts on macroblocks that exercise every rule - runs of 15/16/30/31, a leading run of 255, all zeros, all non-zero,
    the largest representable magnitudes."""
    rng = np.random.default_rng(20)
    blocks = []
    for density in (0.0, 0.01, 0.05, 0.3, 1.0):
        for _ in range(6):
            c = rng.integers(-300, 300, 256).astype(np.int16)
            c[rng.random(256) >= density] = 0
            blocks.append(c)
    for gap in (14, 15, 16, 29, 30, 31, 45, 46):
        c = np.zeros(256, np.int16); c[gap] = 5; c[min(255, 2 * gap + 1)] = -7
        blocks.append(c)
    c = np.zeros(256, np.int16); c[255] = 1; blocks.append(c)
    c = np.zeros(256, np.int16); c[0] = 16383; c[1] = -16383; c[2] = 1; c[3] = -1; blocks.append(c)
    coeff = np.stack(blocks)
    tok, table, mb_off = pfvo.rle_frame(coeff)
    want_table = [0] * 16
    for m, c in enumerate(coeff):
        seq = ref.rle_encode(c)
        ref.update_table(want_table, seq)
        got = tok[mb_off[m]:mb_off[m + 1]]
        assert [(int(t) & 15, (int(t) >> 4) & 15, int(np.int16(np.uint16(int(t) >> 16)))) for t in got] == seq
    assert list(table) == want_table
