"""GPU parity: the CUDA macroblock path, called through the C ABI, against the CPU oracle (bit-exact).

Mirrors what the reference's own tests exercise (src/lib.rs:241-335: encode I+P frames, decode them back)
with the assertions the reference lacks.  Run with `-m gpu` on a B200.
"""
import numpy as np
import pytest

import pfvo
from pretty_fast_video_b200 import (PFV_FRAME_I, PFV_FRAME_P, Engine, PfvError, make_qtables)
from pretty_fast_video_b200.engine import DecodeJob, EncodeJob
from pretty_fast_video_b200.synth import SynthVideo

pytestmark = pytest.mark.gpu

SIZES = [(16, 16), (64, 48), (50, 38), (512, 384), (1918, 1078 + 2)]


def rand_coeffs(rng, nb, mode):
    if mode == "small":
        c = rng.integers(-64, 65, nb * 256)
        c[rng.random(nb * 256) < 0.7] = 0
    elif mode == "mid":
        c = rng.integers(-1024, 1025, nb * 256)
    elif mode in ("dc_only", "mixed"):
        # stream-like sparsity: most sub-blocks carry only a DC term (or nothing), incl. DC extremes that
        # saturate or wrap; "mixed" sprinkles sub-blocks with a few AC terms among them so the kernel's
        # DC-only path, its queue of general sub-blocks and the carry between tiles are all exercised
        sbk = np.zeros((nb * 4, 64), np.int64)
        sbk[:, 0] = rng.integers(-300, 301, nb * 4)
        sbk[rng.random(nb * 4) < 0.1, 0] = 0
        ext = rng.random(nb * 4) < 0.05
        sbk[ext, 0] = rng.choice([-32768, 32767, -2048, 2047], int(ext.sum()))
        if mode == "mixed":
            for frac, npos in ((0.15, 3), (0.05, 40)):
                pick = np.flatnonzero(rng.random(nb * 4) < frac)
                for i in pick:
                    pos = rng.integers(1, 64, npos)
                    sbk[i, pos] = rng.integers(-200, 201, npos)
            sbk[np.flatnonzero(rng.random(nb * 4) < 0.02), 63] = 1     # lone last coefficient, zero DC
        c = sbk.reshape(-1)
    else:  # full i16 range: exercises the wrapping i32 arithmetic
        c = rng.integers(-32768, 32768, nb * 256)
    return c.astype(np.int16)


def rand_headers(rng, g, p_coded=0.5):
    hdr = np.zeros((g.nb, 4), np.uint8)
    mv = hdr[:, :2].view(np.int8)
    idx = 0
    for (pw, ph) in ((g.pw, g.ph), (g.cpw, g.cph), (g.cpw, g.cph)):
        bw, bh = pw // 16, ph // 16
        for by in range(bh):
            for bx in range(bw):
                lo_x, hi_x = max(-15, -bx * 16), min(15, pw - 16 - bx * 16)
                lo_y, hi_y = max(-15, -by * 16), min(15, ph - 16 - by * 16)
                if rng.random() < 0.8:
                    mv[idx, 0] = rng.integers(lo_x, hi_x + 1)
                    mv[idx, 1] = rng.integers(lo_y, hi_y + 1)
                hdr[idx, 2] = rng.random() < p_coded
                idx += 1
    return hdr


@pytest.mark.parametrize("size", SIZES)
@pytest.mark.parametrize("mode", ["small", "mid", "full", "dc_only", "mixed"])
def test_decode_iframe_matches_oracle(size, mode):
    w, h = size
    rng = np.random.default_rng(w * 7919 + h * 31 + len(mode))
    qt, _ = make_qtables(5)
    og = pfvo.geometry_for(w, h)
    coeff = rand_coeffs(rng, og.nb, mode)
    want = pfvo.frame_init(og)
    pfvo.decode_iframe_coeffs(og, qt, (0, 1, 1), coeff, want)
    with Engine(w, h, qt, nslots=2, max_jobs=1) as e:
        assert e.geometry.nb == og.nb and e.geometry.frame_bytes == want.size
        e.decode_submit([DecodeJob(PFV_FRAME_I, 1, coeff, (0, 1, 1))])
        e.sync()
        got = e.slot_read(1)
        # PFV_JOB_DENSE is a hint: it picks the kernel that transforms every sub-block, never the result
        e.decode_submit([DecodeJob(PFV_FRAME_I, 0, coeff, (0, 1, 1), dense_hint=True)])
        e.sync()
        got_dense = e.slot_read(0)
    assert np.array_equal(got, want)
    assert np.array_equal(got_dense, want)


@pytest.mark.parametrize("size", SIZES)
@pytest.mark.parametrize("mode", ["small", "full", "dc_only", "mixed"])
def test_decode_pframe_matches_oracle(size, mode):
    w, h = size
    rng = np.random.default_rng(w * 7919 + h * 31 + len(mode) + 5)
    qt, _ = make_qtables(3)
    og = pfvo.geometry_for(w, h)
    ref = rng.integers(0, 256, pfvo.frame_init(og).size).astype(np.uint8)
    hdr = rand_headers(rng, og)
    coeff = rand_coeffs(rng, og.nb, mode)
    coeff.reshape(-1, 256)[hdr[:, 2] == 0] = 0            # dec.rs:376: skipped blocks carry zeros
    want = ref.copy()
    pfvo.decode_pframe_coeffs(og, qt, (2, 3, 3), hdr, coeff, want)
    with Engine(w, h, qt, nslots=2, max_jobs=1) as e:
        e.slot_write(0, ref)
        e.decode_submit([DecodeJob(PFV_FRAME_P, 1, coeff, (2, 3, 3), ref_slot=0, hdr=hdr)])
        e.sync()
        got = e.slot_read(1)
        assert np.array_equal(e.slot_read(0), ref)        # the reference plane is only read
    assert np.array_equal(got, want)


@pytest.mark.parametrize("ivar,pvar", [("stream", "fused"), ("stream", "win"), ("sb", "warp"), ("warp", "win")])
def test_decode_kernel_variants_agree(ivar, pvar, monkeypatch):
    """One default kernel per path plus independently written second implementations (PFV_DECODE_*_VARIANT): all agree
    with the oracle."""
    monkeypatch.setenv("PFV_DECODE_I_VARIANT", ivar)
    monkeypatch.setenv("PFV_DECODE_P_VARIANT", pvar)
    w, h = 208, 112
    rng = np.random.default_rng(99)
    qt, _ = make_qtables(4)
    og = pfvo.geometry_for(w, h)
    ci = rand_coeffs(rng, og.nb, "mixed")
    cp = rand_coeffs(rng, og.nb, "mixed")
    hdr = rand_headers(rng, og)
    cp.reshape(-1, 256)[hdr[:, 2] == 0] = 0
    want0 = pfvo.frame_init(og)
    pfvo.decode_iframe_coeffs(og, qt, (0, 1, 1), ci, want0)
    want1 = want0.copy()
    pfvo.decode_pframe_coeffs(og, qt, (2, 3, 3), hdr, cp, want1)
    with Engine(w, h, qt, nslots=2, max_jobs=1) as e:
        e.decode_submit([DecodeJob(PFV_FRAME_I, 0, ci, (0, 1, 1))])
        e.decode_submit([DecodeJob(PFV_FRAME_P, 1, cp, (2, 3, 3), ref_slot=0, hdr=hdr)])
        e.sync()
        assert np.array_equal(e.slot_read(0), want0)
        assert np.array_equal(e.slot_read(1), want1)


@pytest.mark.parametrize("ivar", ["stream", "sb"])
@pytest.mark.parametrize("qidx", [(0, 1, 1), (0, 1, 3), (2, 2, 2)])
def test_decode_iframe_batch_with_three_tables(ivar, qidx, monkeypatch):
    """A batch of key frames in one submit with U and V on different q-tables, on the same one, and all three planes on one;
    "mixed" coefficients keep the streaming kernel's ring busy."""
    monkeypatch.setenv("PFV_DECODE_I_VARIANT", ivar)
    w, h, n = 336, 208, 7
    rng = np.random.default_rng(4242 + sum(qidx))
    qt, _ = make_qtables(6)
    og = pfvo.geometry_for(w, h)
    coeffs = [rand_coeffs(rng, og.nb, "mixed" if i % 3 else "small") for i in range(n)]
    with Engine(w, h, qt, nslots=n, max_jobs=n) as e:
        e.decode_submit([DecodeJob(PFV_FRAME_I, i, coeffs[i], qidx) for i in range(n)])
        e.sync()
        for i in range(n):
            want = pfvo.frame_init(og)
            pfvo.decode_iframe_coeffs(og, qt, qidx, coeffs[i], want)
            assert np.array_equal(e.slot_read(i), want), f"job {i}"


@pytest.mark.parametrize("pvar", ["fused", "win", "warp"])
def test_decode_pframe_long_motion_vectors(pvar, monkeypatch):
    """The stream format carries 7-bit vectors (src/dec.rs:367-368) and the reference decoder follows any vector that
    stays inside the padded plane (src/common.rs:255-261), also ones its own encoder (|mv| <= 15) never emits."""
    monkeypatch.setenv("PFV_DECODE_P_VARIANT", pvar)
    w, h = 400, 240
    rng = np.random.default_rng(17)
    qt, _ = make_qtables(5)
    og = pfvo.geometry_for(w, h)
    hdr = np.zeros((og.nb, 4), np.uint8)
    mv = hdr[:, :2].view(np.int8)
    idx = 0
    for (pw, ph) in ((og.pw, og.ph), (og.cpw, og.cph), (og.cpw, og.cph)):
        for by in range(ph // 16):
            for bx in range(pw // 16):
                mv[idx, 0] = rng.integers(max(-64, -bx * 16), min(63, pw - 16 - bx * 16) + 1)
                mv[idx, 1] = rng.integers(max(-64, -by * 16), min(63, ph - 16 - by * 16) + 1)
                hdr[idx, 2] = rng.random() < 0.4
                idx += 1
    coeff = rand_coeffs(rng, og.nb, "mixed")
    coeff.reshape(-1, 256)[hdr[:, 2] == 0] = 0
    ref = rng.integers(0, 256, pfvo.frame_init(og).size).astype(np.uint8)
    want = ref.copy()
    pfvo.decode_pframe_coeffs(og, qt, (2, 3, 3), hdr, coeff, want)
    with Engine(w, h, qt, nslots=2, max_jobs=1) as e:
        e.slot_write(0, ref)
        e.decode_submit([DecodeJob(PFV_FRAME_P, 1, coeff, (2, 3, 3), ref_slot=0, hdr=hdr)])
        e.sync()
        assert np.array_equal(e.slot_read(1), want)


@pytest.mark.parametrize("compact", ["0", "1"])
@pytest.mark.parametrize("mode", ["small", "mixed", "full"])
def test_host_compaction_of_dense_buffers_is_transparent(compact, mode, monkeypatch):
    """pfv_decode_submit with dense HOST buffers: the engine may compact them to tokens on the host before the copy
    (PFV_HOST_COMPACT=1, off by default; too-dense frames go as they are) - same pictures either way."""
    monkeypatch.setenv("PFV_HOST_COMPACT", compact)
    w, h = 336, 208
    rng = np.random.default_rng(5)
    qt, _ = make_qtables(6)
    og = pfvo.geometry_for(w, h)
    jobs_i = [rand_coeffs(rng, og.nb, mode) for _ in range(3)]
    hdrs = [rand_headers(rng, og) for _ in range(3)]
    jobs_p = []
    for hd in hdrs:
        c = rand_coeffs(rng, og.nb, mode)
        # garbage in the coefficients of skipped macroblocks must never be read (src/dec.rs:381)
        c.reshape(-1, 256)[hd[:, 2] == 0] = 12345
        jobs_p.append(c)
    with Engine(w, h, qt, nslots=6, max_jobs=3) as e:
        e.decode_submit([DecodeJob(PFV_FRAME_I, 2 * i, jobs_i[i], (0, 1, 1)) for i in range(3)])
        e.decode_submit([DecodeJob(PFV_FRAME_P, 2 * i + 1, jobs_p[i], (2, 3, 3), ref_slot=2 * i, hdr=hdrs[i]) for i in range(3)])
        e.sync()
        for i in range(3):
            want = pfvo.frame_init(og)
            pfvo.decode_iframe_coeffs(og, qt, (0, 1, 1), jobs_i[i], want)
            assert np.array_equal(e.slot_read(2 * i), want)
            cz = jobs_p[i].copy()
            cz.reshape(-1, 256)[hdrs[i][:, 2] == 0] = 0
            pfvo.decode_pframe_coeffs(og, qt, (2, 3, 3), hdrs[i], cz, want)
            assert np.array_equal(e.slot_read(2 * i + 1), want)


@pytest.mark.parametrize("pvar", ["fused", "win"])
def test_chained_batches_of_pframes_match_oracle(pvar, monkeypatch):
    """A chain of batched P submits (5 lanes x 4 frames, different coded fractions per lane: the fused kernel's ring
    mixes macroblocks of different frames in one transform pass) gives the oracle's pictures."""
    monkeypatch.setenv("PFV_DECODE_P_VARIANT", pvar)
    w, h = 400, 240
    rng = np.random.default_rng(77)
    qt, _ = make_qtables(5)
    og = pfvo.geometry_for(w, h)
    L, T = 5, 4
    state = [rng.integers(0, 256, pfvo.frame_init(og).size).astype(np.uint8) for _ in range(L)]
    with Engine(w, h, qt, nslots=2 * L, max_jobs=L) as e:
        for l in range(L):
            e.slot_write(2 * l, state[l])
        cur = [2 * l for l in range(L)]
        for t in range(T):
            jobs = []
            for l in range(L):
                hdr = rand_headers(rng, og, p_coded=0.3 + 0.1 * l)
                coeff = rand_coeffs(rng, og.nb, "mixed")
                coeff.reshape(-1, 256)[hdr[:, 2] == 0] = 0
                pfvo.decode_pframe_coeffs(og, qt, (2, 3, 3), hdr, coeff, state[l])
                jobs.append(DecodeJob(PFV_FRAME_P, cur[l] ^ 1, coeff, (2, 3, 3), ref_slot=cur[l], hdr=hdr))
                cur[l] ^= 1
            e.decode_submit(jobs)
        e.sync()
        for l in range(L):
            assert np.array_equal(e.slot_read(cur[l]), state[l])


def test_decode_pframe_all_skipped_is_a_copy():
    w, h = 320, 240
    qt, _ = make_qtables(5)
    og = pfvo.geometry_for(w, h)
    rng = np.random.default_rng(7)
    ref = rng.integers(0, 256, pfvo.frame_init(og).size).astype(np.uint8)
    hdr = np.zeros((og.nb, 4), np.uint8)
    coeff = np.zeros(og.nb * 256, np.int16)
    with Engine(w, h, qt) as e:
        e.slot_write(0, ref)
        e.decode_submit([DecodeJob(PFV_FRAME_P, 1, coeff, (2, 3, 3), ref_slot=0, hdr=hdr)])
        e.sync()
        assert np.array_equal(e.slot_read(1), ref)


def test_bad_motion_vector_is_reported_not_followed():
    w, h = 64, 64
    qt, _ = make_qtables(5)
    og = pfvo.geometry_for(w, h)
    hdr = np.zeros((og.nb, 4), np.uint8)
    hdr[:, :2].view(np.int8)[0] = (-5, 0)                 # leaves the plane on the left (common.rs:258)
    coeff = np.zeros(og.nb * 256, np.int16)
    with Engine(w, h, qt) as e:
        e.decode_submit([DecodeJob(PFV_FRAME_P, 1, coeff, (2, 3, 3), ref_slot=0, hdr=hdr)])
        with pytest.raises(PfvError) as ei:
            e.sync()
        assert ei.value.code == -4
        e.sync()                                           # error is cleared once reported


def test_argument_errors():
    qt, _ = make_qtables(5)
    with pytest.raises(PfvError):
        Engine(63, 64, qt)                                 # odd width (frame.rs:13)
    with Engine(64, 64, qt) as e:
        c = np.zeros(e.geometry.nb * 256, np.int16)
        with pytest.raises(PfvError):
            e.decode_submit([DecodeJob(PFV_FRAME_P, 1, c, (2, 3, 3), ref_slot=1, hdr=np.zeros((e.geometry.nb, 4), np.uint8))])
        with pytest.raises(PfvError):
            e.decode_submit([DecodeJob(PFV_FRAME_I, 5, c, (0, 1, 1))])
        with pytest.raises(PfvError):
            e.decode_submit([DecodeJob(PFV_FRAME_I, 1, c, (0, 9, 1))])
        e.sync()


@pytest.mark.parametrize("size", SIZES)
@pytest.mark.parametrize("quality", [0, 2, 5, 10])
def test_encode_iframe_matches_oracle(size, quality):
    w, h = size
    if quality not in (2, 5) and w > 600:
        pytest.skip("large sizes at two qualities only")
    qt, _ = make_qtables(quality)
    oqt, _ = pfvo.make_qtables(quality)
    assert np.array_equal(qt, oqt)
    og = pfvo.geometry_for(w, h)
    kind = "random" if quality == 10 else "moving"
    y, u, v = SynthVideo(w, h, 99, kind).frame(3)
    prev = pfvo.frame_init(og)
    want_c = pfvo.encode_iframe_coeffs(og, qt, y, u, v, prev)
    got_c = np.zeros(og.nb * 256, np.int16)
    with Engine(w, h, qt) as e:
        e.encode_submit([EncodeJob(PFV_FRAME_I, 1, (y, u, v), got_c)])
        e.sync()
        got_recon = e.slot_read(1)
    assert np.array_equal(got_c, want_c)                   # bar is +-1 LSB; integer math gives 0
    assert np.array_equal(got_recon, prev)


@pytest.mark.parametrize("size", SIZES)
@pytest.mark.parametrize("quality,kind", [(0, "moving"), (2, "moving"), (5, "moving"), (5, "random"), (5, "static"), (10, "moving")])
def test_encode_pframe_matches_oracle(size, quality, kind):
    w, h = size
    qt, px_err = make_qtables(quality)
    _, opx = pfvo.make_qtables(quality)
    assert px_err == opx
    og = pfvo.geometry_for(w, h)
    sv = SynthVideo(w, h, 1000 + quality, kind)
    prev = pfvo.frame_init(og)
    y0, u0, v0 = sv.frame(0)
    pfvo.encode_iframe_coeffs(og, qt, y0, u0, v0, prev)   # prev = reconstructed key frame
    ref = prev.copy()
    y1, u1, v1 = sv.frame(1)
    want_h, want_c = pfvo.encode_pframe_coeffs(og, qt, px_err, y1, u1, v1, prev)
    got_c = np.zeros(og.nb * 256, np.int16)
    got_h = np.zeros((og.nb, 4), np.uint8)
    with Engine(w, h, qt) as e:
        e.slot_write(0, ref)
        e.encode_submit([EncodeJob(PFV_FRAME_P, 1, (y1, u1, v1), got_c, ref_slot=0, px_err=px_err, hdr_out=got_h)])
        e.sync()
        got_recon = e.slot_read(1)
    assert np.array_equal(got_h, want_h), "motion vectors / skip flags differ"
    coded = want_h[:, 2] != 0
    assert np.array_equal(got_c.reshape(-1, 256)[coded], want_c.reshape(-1, 256)[coded])
    assert np.array_equal(got_recon, prev)


@pytest.mark.parametrize("ivar", ["persist", "warp"])
def test_encode_i_kernel_variants_agree(ivar, monkeypatch):
    """PFV_ENCODE_I_VARIANT: the persistent thread-per-sub-block kernel (default) and the first-generation warp-per-macroblock
    kernel give the oracle's coefficients and reconstruction; a ragged size (padding with the clear colour, planes
    whose width is not a multiple of 8) and a batch of jobs in one launch."""
    monkeypatch.setenv("PFV_ENCODE_I_VARIANT", ivar)
    test_encode_iframe_matches_oracle((50, 38), 2)
    test_encode_iframe_matches_oracle((512, 384), 5)
    w, h, n = 336, 208, 5
    qt, _ = make_qtables(4)
    og = pfvo.geometry_for(w, h)
    sv = SynthVideo(w, h, 5, "moving")
    outs = [np.zeros(og.nb * 256, np.int16) for _ in range(n)]
    with Engine(w, h, qt, nslots=n, max_jobs=n) as e:
        e.encode_submit([EncodeJob(PFV_FRAME_I, i, sv.frame(i), outs[i]) for i in range(n)])
        e.sync()
        for i in range(n):
            prev = pfvo.frame_init(og)
            want = pfvo.encode_iframe_coeffs(og, qt, *sv.frame(i), prev)
            assert np.array_equal(outs[i], want)
            assert np.array_equal(e.slot_read(i), prev)


@pytest.mark.parametrize("pvar", ["strip", "v1"])
def test_encode_p_kernel_variants_agree(pvar, monkeypatch):
    """PFV_ENCODE_P_VARIANT: the warp-per-tile column-strip search (default) and the first-generation warp-per-macroblock
    kernel give the oracle's motion vectors, skip decisions, coefficients and reconstruction - ragged sizes, every content
    kind, and a batch of dependent submits."""
    monkeypatch.setenv("PFV_ENCODE_P_VARIANT", pvar)
    test_encode_pframe_matches_oracle((50, 38), 2, "moving")
    test_encode_pframe_matches_oracle((512, 384), 5, "moving")
    test_encode_pframe_matches_oracle((512, 384), 5, "random")
    test_encode_pframe_matches_oracle((64, 48), 5, "static")
    test_encode_pframe_matches_oracle((1918, 1080), 0, "moving")


def test_encode_rejects_zero_divisors():
    """The reference's quantiser divides by the table entry (src/dct.rs:95: a zero would panic); Encoder::new clamps its
    tables to >= 1 (src/enc.rs:48-51).  A context with a zero divisor in tables 0..3 decodes, but refuses to encode."""
    qt, _ = make_qtables(5)
    bad = qt.copy()
    bad[2, 17] = 0
    y = np.zeros((48, 64), np.uint8); u = np.zeros((24, 32), np.uint8)
    with Engine(64, 48, bad) as e:
        with pytest.raises(PfvError):
            e.encode_submit([EncodeJob(PFV_FRAME_I, 0, (y, u, u), np.zeros(e.geometry.nb * 256, np.int16))])
        e.decode_submit([DecodeJob(PFV_FRAME_I, 0, np.zeros(e.geometry.nb * 256, np.int16), (0, 1, 1))])
        e.sync()


def _oracle_stream(w, h, nframes, quality, key_every, seed, kind="moving"):
    """Encode with the oracle; return per-frame seam data and decoded visible planes."""
    sv = SynthVideo(w, h, seed, kind)
    enc = pfvo.Encoder(w, h, 30, quality, nthreads=8)
    frames = []
    for t in range(nframes):
        y, u, v = sv.frame(t)
        if t % key_every == 0:
            enc.encode_iframe(y, u, v)
            frames.append((PFV_FRAME_I, None, enc.last_coeffs(), (0, 1, 1)))
        else:
            enc.encode_pframe(y, u, v)
            frames.append((PFV_FRAME_P, enc.last_headers(), enc.last_coeffs(), (2, 3, 3)))
    enc.finish()
    data = enc.bytes()
    dec = pfvo.Decoder(data, nthreads=8)
    planes = []
    while True:
        more, fr = dec.advance_frame()
        if fr is not None:
            planes.append(fr)
        if not more:
            break
    return frames, planes, data


def test_config1_stream_512x384_bit_exact():
    """BASELINE.json configs[0] restated: 512x384, 161 frames, quality 2, key frame every 60
    (the parameters of src/lib.rs:271-292), decoded frame by frame on the GPU."""
    w, h, n = 512, 384, 161
    frames, planes, _ = _oracle_stream(w, h, n, 2, 60, 0x50465600)
    assert len(planes) == n
    qt, _ = make_qtables(2)
    with Engine(w, h, qt, nslots=2, max_jobs=1) as e:
        g = e.geometry
        cur = 0
        for t, (kind, hdr, coeff, qidx) in enumerate(frames):
            y = np.empty((g.height, g.width), np.uint8)
            u = np.empty((g.cheight, g.cwidth), np.uint8)
            v = np.empty((g.cheight, g.cwidth), np.uint8)
            dst = cur ^ 1
            e.decode_submit([DecodeJob(kind, dst, coeff, qidx, ref_slot=cur, hdr=hdr, out=(y, u, v))])
            e.sync()
            cur = dst
            for got, want, name in zip((y, u, v), planes[t], "yuv"):
                assert np.array_equal(got, want), f"frame {t} plane {name}"


def test_batched_gops_and_pipelining():
    """Frame k of several independent GOPs per submit, many submits in flight, outputs copied to host."""
    w, h, gop, ngops = 320, 208, 6, 5
    qt, _ = make_qtables(5)
    streams = [_oracle_stream(w, h, gop, 5, gop, 500 + i) for i in range(ngops)]
    with Engine(w, h, qt, nslots=2 * ngops, max_jobs=ngops) as e:
        g = e.geometry
        outs = [[(np.empty((g.height, g.width), np.uint8), np.empty((g.cheight, g.cwidth), np.uint8),
                  np.empty((g.cheight, g.cwidth), np.uint8)) for _ in range(gop)] for _ in range(ngops)]
        cur = [2 * i for i in range(ngops)]
        for k in range(gop):
            jobs = []
            for i in range(ngops):
                kind, hdr, coeff, qidx = streams[i][0][k]
                dst = cur[i] ^ 1
                jobs.append(DecodeJob(kind, dst, coeff, qidx, ref_slot=cur[i], hdr=hdr, out=outs[i][k]))
                cur[i] = dst
            e.decode_submit(jobs)                           # no sync between dependent submits
        e.sync()
    for i in range(ngops):
        for k in range(gop):
            for got, want in zip(outs[i][k], streams[i][1][k]):
                assert np.array_equal(got, want), (i, k)


def test_encoder_chain_matches_oracle_stream():
    """I + P chain encoded on the GPU frame by frame (state kept on the device) vs the oracle encoder."""
    w, h, n = 352, 288, 8
    qt, px_err = make_qtables(5)
    sv = SynthVideo(w, h, 4242)
    enc = pfvo.Encoder(w, h, 30, 5, nthreads=8)
    with Engine(w, h, qt) as e:
        g = e.geometry
        cur = 0
        for t in range(n):
            y, u, v = sv.frame(t)
            c = np.zeros(g.nb * 256, np.int16)
            hd = np.zeros((g.nb, 4), np.uint8)
            dst = cur ^ 1
            if t == 0:
                enc.encode_iframe(y, u, v)
                e.encode_submit([EncodeJob(PFV_FRAME_I, dst, (y, u, v), c)])
            else:
                enc.encode_pframe(y, u, v)
                e.encode_submit([EncodeJob(PFV_FRAME_P, dst, (y, u, v), c, ref_slot=cur, px_err=px_err, hdr_out=hd)])
            e.sync()
            cur = dst
            wc, wh = enc.last_coeffs().reshape(-1, 256), enc.last_headers()
            if t == 0:
                assert np.array_equal(c.reshape(-1, 256), wc)
            else:
                assert np.array_equal(hd, wh), f"frame {t} headers"
                coded = wh[:, 2] != 0
                assert np.array_equal(c.reshape(-1, 256)[coded], wc[coded]), f"frame {t} coefficients"
            assert np.array_equal(e.slot_read(cur), enc.prev_frame()), f"frame {t} reconstruction"


def test_full_size_1080p_properties():
    """BASELINE full size: (i) one 1080p I and P frame against the oracle; (ii) closed loop: what the GPU
    encoder reconstructed is exactly what the GPU decoder produces from the encoder's output."""
    w, h = 1920, 1080
    qt, px_err = make_qtables(5)
    og = pfvo.geometry_for(w, h)
    sv = SynthVideo(w, h, 0x50465602)
    with Engine(w, h, qt, nslots=4, max_jobs=1) as e:
        g = e.geometry
        assert (g.nb, g.frame_bytes) == (12240, 3133440)
        c0 = np.zeros(g.nb * 256, np.int16)
        c1 = np.zeros(g.nb * 256, np.int16)
        h1 = np.zeros((g.nb, 4), np.uint8)
        y0, u0, v0 = sv.frame(0)
        y1, u1, v1 = sv.frame(1)
        e.encode_submit([EncodeJob(PFV_FRAME_I, 0, (y0, u0, v0), c0)])
        e.encode_submit([EncodeJob(PFV_FRAME_P, 1, (y1, u1, v1), c1, ref_slot=0, px_err=px_err, hdr_out=h1)])
        e.sync()
        prev = pfvo.frame_init(og)
        wc0 = pfvo.encode_iframe_coeffs(og, qt, y0, u0, v0, prev, nthreads=8)
        assert np.array_equal(c0, wc0)
        assert np.array_equal(e.slot_read(0), prev)
        wh1, wc1 = pfvo.encode_pframe_coeffs(og, qt, px_err, y1, u1, v1, prev, nthreads=8)
        assert np.array_equal(h1, wh1)
        coded = wh1[:, 2] != 0
        assert np.array_equal(c1.reshape(-1, 256)[coded], wc1.reshape(-1, 256)[coded])
        assert np.array_equal(e.slot_read(1), prev)
        # closed loop through the decoder kernels
        c1z = c1.copy()
        c1z.reshape(-1, 256)[~coded] = 0
        e.decode_submit([DecodeJob(PFV_FRAME_I, 2, c0, (0, 1, 1))])
        e.decode_submit([DecodeJob(PFV_FRAME_P, 3, c1z, (2, 3, 3), ref_slot=2, hdr=h1)])
        e.sync()
        assert np.array_equal(e.slot_read(2), e.slot_read(0))
        assert np.array_equal(e.slot_read(3), e.slot_read(1))


def test_contexts_on_two_devices_in_one_process():
    """One process may own contexts on several GPUs (the sharded path normally uses one process per GPU): per-device
    kernel attributes, pools and streams must not leak between them."""
    from pretty_fast_video_b200 import _native as N
    if N.lib().pfv_device_count() < 2:
        pytest.skip("needs two GPUs")
    w, h = 336, 208
    rng = np.random.default_rng(41)
    qt, _ = make_qtables(5)
    og = pfvo.geometry_for(w, h)
    ci = rand_coeffs(rng, og.nb, "mixed")
    hdr = rand_headers(rng, og)
    cp = rand_coeffs(rng, og.nb, "mixed")
    cp.reshape(-1, 256)[hdr[:, 2] == 0] = 0
    want0 = pfvo.frame_init(og)
    pfvo.decode_iframe_coeffs(og, qt, (0, 1, 1), ci, want0)
    want1 = want0.copy()
    pfvo.decode_pframe_coeffs(og, qt, (2, 3, 3), hdr, cp, want1)
    engines = [Engine(w, h, qt, nslots=2, max_jobs=1, device=d) for d in (1, 0, 1)]
    try:
        for e in engines:
            e.decode_submit([DecodeJob(PFV_FRAME_I, 0, ci, (0, 1, 1))])
        for e in engines:
            e.decode_submit([DecodeJob(PFV_FRAME_P, 1, cp, (2, 3, 3), ref_slot=0, hdr=hdr)])
        for e in engines:
            e.sync()
            assert np.array_equal(e.slot_read(0), want0) and np.array_equal(e.slot_read(1), want1)
    finally:
        for e in engines:
            e.close()
