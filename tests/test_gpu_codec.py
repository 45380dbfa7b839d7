"""GPU: the Decoder / Encoder objects (csrc/pfv_codec.cpp) and the sparse coefficient transport, against the oracle's
restatement of pfv_rs::dec::Decoder / pfv_rs::enc::Encoder on the same streams.  Bit-exact everywhere."""
import numpy as np
import pytest

import pfvo
from pretty_fast_video_b200 import PFV_FRAME_I, PFV_FRAME_P, Engine, PfvError, codec, geometry_for, make_qtables
from pretty_fast_video_b200 import _native as N
from pretty_fast_video_b200.engine import DecodeJob, EncodeJob, PinnedArena, SparseDecodeJob, SparseEncodeJob
from pretty_fast_video_b200.synth import SynthVideo
from test_codec_cpu import oracle_stream

pytestmark = pytest.mark.gpu


def oracle_decode_all(data):
    """[(frame planes or None for a drop frame)], final framebuffer"""
    dec = pfvo.Decoder(data)
    out = []
    while True:
        more, fr = dec.advance_frame()
        if not more:
            break
        out.append(None if fr is None else tuple(p.copy() for p in fr))
    return out, dec.framebuffer().copy()


def gpu_decode_all(data, **kw):
    out = []
    with codec.Decoder(data, **kw) as dec:
        while True:
            got = []
            more = dec.advance_frame(lambda fr: got.append(tuple(p.copy() for p in fr)))
            if not more:
                assert not got
                break
            out.append(got[0] if got else None)
        fb = dec.framebuffer()
        assert dec.advance_frame(lambda fr: None) is False           # eof stays eof (src/dec.rs:171-173)
    return out, fb


def same_frames(a, b):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert (x is None) == (y is None)
        if x is not None:
            for p, q in zip(x, y):
                assert np.array_equal(p, q)


@pytest.mark.parametrize("mode", ["small", "mixed", "mid", "full"])
@pytest.mark.parametrize("size", [(64, 48), (176, 144), (1920, 1080)])
def test_sparse_submit_equals_dense_submit(size, mode):
    from test_gpu_parity import rand_coeffs, rand_headers
    w, h = size
    rng = np.random.default_rng(hash((size, mode)) & 0xFFFF)
    qt, _ = make_qtables(5)
    g = geometry_for(w, h)
    ci = rand_coeffs(rng, g.nb, mode)
    cp = rand_coeffs(rng, g.nb, mode)
    hdr = rand_headers(rng, g)
    cp.reshape(g.nb, 256)[hdr[:, 2] == 0] = 0
    with Engine(w, h, qt, nslots=4, max_jobs=2) as e:
        e.decode_submit([DecodeJob(PFV_FRAME_I, 0, ci, (0, 1, 1))])
        e.decode_submit([DecodeJob(PFV_FRAME_P, 1, cp, (2, 3, 3), ref_slot=0, hdr=hdr)])
        mo_i, tk_i = codec.dense_to_tokens(ci, g.nb)
        mo_p, tk_p = codec.dense_to_tokens(cp, g.nb)
        e.decode_submit_sparse([SparseDecodeJob(PFV_FRAME_I, 2, mo_i, tk_i, (0, 1, 1))])
        e.decode_submit_sparse([SparseDecodeJob(PFV_FRAME_P, 3, mo_p, tk_p, (2, 3, 3), ref_slot=2, hdr=hdr)])
        e.sync()
        assert np.array_equal(e.slot_read(0), e.slot_read(2))
        assert np.array_equal(e.slot_read(1), e.slot_read(3))
        # and against the oracle
        of = pfvo.frame_init(pfvo.geometry_for(w, h))
        pfvo.decode_iframe_coeffs(pfvo.geometry_for(w, h), qt, (0, 1, 1), ci, of)
        assert np.array_equal(e.slot_read(2), of)
        pfvo.decode_pframe_coeffs(pfvo.geometry_for(w, h), qt, (2, 3, 3), hdr, cp, of)
        assert np.array_equal(e.slot_read(3), of)


def test_sparse_submit_argument_errors():
    qt, _ = make_qtables(5)
    g = geometry_for(64, 48)
    with Engine(64, 48, qt, nslots=2, max_jobs=1) as e:
        mo = np.zeros(g.nb + 1, np.uint32)
        tk = np.zeros(4, np.uint32)
        mo[1:] = 4
        with pytest.raises(PfvError):                                # ntok does not match mb_off[nb]
            e.decode_submit_sparse([SparseDecodeJob(PFV_FRAME_I, 0, mo, tk[:3], (0, 1, 1))])
        mo2 = mo.copy(); mo2[1] = 7                                  # not monotonic
        with pytest.raises(PfvError):
            e.decode_submit_sparse([SparseDecodeJob(PFV_FRAME_I, 0, mo2, tk, (0, 1, 1))])
        e.decode_submit_sparse([SparseDecodeJob(PFV_FRAME_I, 0, mo, tk, (0, 1, 1))])
        e.sync()


@pytest.mark.parametrize("size,quality,kind,n,key", [((96, 64), 3, "moving", 12, 4), ((130, 70), 5, "moving", 9, 3),
                                                     ((64, 48), 10, "random", 6, 3), ((176, 144), 0, "moving", 6, 6),
                                                     ((64, 64), 5, "static", 5, 5)])
def test_decoder_matches_oracle_decoder(size, quality, kind, n, key):
    w, h = size
    data, _ = oracle_stream(w, h, n, quality, key, 4321, kind=kind, drop_at=(5,) if n > 6 else ())
    want, want_fb = oracle_decode_all(data)
    got, fb = gpu_decode_all(data, num_threads=3)
    same_frames(got, want)
    assert np.array_equal(fb, want_fb)


def test_decoder_accessors_reset_and_delta():
    data, _ = oracle_stream(96, 64, 8, 3, 4, 7)
    want, _ = oracle_decode_all(data)
    with codec.Decoder(data, num_threads=2) as dec:
        assert (dec.width(), dec.height(), dec.framerate()) == (96, 64, 30)
        first = []
        for _ in range(3):
            assert dec.advance_frame(lambda fr: first.append(tuple(p.copy() for p in fr)))
        dec.reset()                                                  # src/dec.rs:148: rewind only
        again = []
        while dec.advance_frame(lambda fr: again.append(tuple(p.copy() for p in fr))):
            pass
        same_frames(first, want[:3])
        same_frames(again, want)                                     # the stream starts with a key frame
        dec.reset()
        # advance_delta: 2.5 frame times -> two frames now, the third with the next half
        frames = []
        assert dec.advance_delta(2.5 / 30.0, lambda fr: frames.append(1))
        assert len(frames) == 2
        assert dec.advance_delta(0.5 / 30.0 + 1e-9, lambda fr: frames.append(1))
        assert len(frames) == 3
        assert dec.advance_delta(100.0, lambda fr: frames.append(1)) is False    # runs into EOF
        assert len(frames) == len(want)


@pytest.mark.parametrize("threads,ahead", [(1, 1), (4, 6), (8, 3)])
def test_decoder_result_does_not_depend_on_read_ahead(threads, ahead):
    data, _ = oracle_stream(176, 144, 14, 4, 5, 11)
    want, want_fb = oracle_decode_all(data)
    got, fb = gpu_decode_all(data, num_threads=threads, read_ahead=ahead)
    same_frames(got, want)
    assert np.array_equal(fb, want_fb)


def test_decoder_errors():
    data, _ = oracle_stream(96, 64, 6, 3, 3, 5)
    with pytest.raises(codec.DecodeError) as e:
        codec.Decoder(b"XXXXXXXX" + data[8:])
    assert e.value.kind == "FormatError"
    with pytest.raises(codec.DecodeError) as e:
        codec.Decoder(data[:200])
    assert e.value.kind == "IOError"
    # cut inside the 4th frame's payload: three frames decode, then an IOError (read_exact fails, src/dec.rs:192/205)
    info, _ = codec.parse_header(data)
    pk, _ = codec.index_packets(data, info.first_packet)
    cut = pk[3][2] + pk[3][1] // 2
    want, _ = oracle_decode_all(data)
    with codec.Decoder(data[:cut], num_threads=2) as dec:
        got = []
        for _ in range(3):
            assert dec.advance_frame(lambda fr: got.append(tuple(p.copy() for p in fr)))
        same_frames(got, want[:3])
        with pytest.raises(codec.DecodeError) as e:
            dec.advance_frame(lambda fr: None)
        assert e.value.kind == "IOError"
    # stream without its EOF packet: all frames, then the read of the next packet header fails
    with codec.Decoder(data[:-5], num_threads=2) as dec:
        n = 0
        for _ in range(len(want)):
            assert dec.advance_frame(lambda fr: None)
            n += 1
        with pytest.raises(codec.DecodeError):
            dec.advance_frame(lambda fr: None)
    # corrupt payload of frame 2 (flip bytes in the token stream): an error or a different picture, never a crash;
    # later key frames still decode
    bad = bytearray(data)
    p2 = pk[1][2]
    for i in range(30, 60):
        bad[p2 + i] ^= 0xFF
    with codec.Decoder(bytes(bad), num_threads=2) as dec:
        for _ in range(len(want)):
            try:
                if not dec.advance_frame(lambda fr: None):
                    break
            except PfvError:
                pass


# ---- sparse encode seam: the run-length pass on the device ---------------------------------------------------------------
def _sparse_outputs(arena, nb, cap=None):
    cap = nb * 256 if cap is None else cap
    return (arena.take((max(cap, 1),), np.uint32), arena.take((N.PFV_TOKSTATS_WORDS,), np.uint32), arena.take((nb + 1,), np.uint32),
            arena.take((nb, 4), np.uint8))


@pytest.mark.parametrize("size", [(16, 16), (64, 48), (50, 38), (512, 384), (1918, 1080)])
@pytest.mark.parametrize("quality,kind", [(0, "moving"), (5, "moving"), (10, "random"), (5, "static")])
def test_sparse_encode_matches_oracle_rle(size, quality, kind):
    """pfv_encode_submit_sparse: the device's RLE sequence, symbol statistics and offsets are rle_encode + update_table
    (src/rle.rs:9-47) of the oracle's coefficients, for a key frame and the P frame that follows it; the packets written from
    them are the oracle encoder's packets."""
    w, h = size
    if w > 600 and (quality, kind) not in ((5, "moving"), (10, "random")):
        pytest.skip("large sizes at two settings only")
    qt, px_err = make_qtables(quality)
    og = pfvo.geometry_for(w, h)
    geo = geometry_for(w, h)
    sv = SynthVideo(w, h, 4242 + quality, kind)
    f0, f1 = sv.frame(0), sv.frame(1)
    prev = pfvo.frame_init(og)
    want_c0 = pfvo.encode_iframe_coeffs(og, qt, *f0, prev)
    want_h1, want_c1 = pfvo.encode_pframe_coeffs(og, qt, px_err, *f1, prev)
    arena = PinnedArena(2 * (og.nb * 256 * 4 + og.nb * 8 + 4096) + 65536)
    tok0, st0, off0, _ = _sparse_outputs(arena, og.nb)
    tok1, st1, off1, h1 = _sparse_outputs(arena, og.nb)
    with Engine(w, h, qt, nslots=3) as e:
        e.encode_submit_sparse([SparseEncodeJob(PFV_FRAME_I, 1, f0, tok0, st0, mb_off_out=off0)])
        e.encode_submit_sparse([SparseEncodeJob(PFV_FRAME_P, 2, f1, tok1, st1, ref_slot=1, px_err=px_err, hdr_out=h1, mb_off_out=off1)])
        e.sync()
        recon = e.slot_read(2)
    assert np.array_equal(recon, prev)
    assert np.array_equal(h1, want_h1)
    for fk, tok, st, off, want_c, hdr in ((PFV_FRAME_I, tok0, st0, off0, want_c0, None), (PFV_FRAME_P, tok1, st1, off1, want_c1, want_h1)):
        o_tok, o_table, o_off = pfvo.rle_frame(want_c, None if hdr is None else hdr[:, 2])
        n = int(st[N.PFV_TOKSTATS_NTOK])
        assert n == o_tok.size and int(st[N.PFV_TOKSTATS_FLAGS]) == 0
        assert np.array_equal(tok[:n], o_tok)
        assert np.array_equal(off, o_off)
        assert np.array_equal(st[:16], np.bincount(o_tok & 15, minlength=16))
        assert np.array_equal(st[16:32], np.bincount((o_tok >> 4) & 15, minlength=16))
        assert np.array_equal(st[:16].astype(np.int64) + st[16:32], o_table)
        assert codec.encode_packet_tokens(geo, fk, tok[:n], st, hdr) == codec.encode_packet(geo, fk, want_c, hdr)
    arena.close()


def test_sparse_encode_batches_overflow_and_device_outputs():
    """Several independent frames in one submit (I and P mixed, dense and sparse seams agree); a token buffer that is too small
    is reported; outputs in device memory work like pinned ones."""
    import torch
    w, h, quality = 176, 144, 4
    qt, px_err = make_qtables(quality)
    geo = geometry_for(w, h)
    nb = geo.nb
    sv = SynthVideo(w, h, 77, "moving")
    frames = [sv.frame(t) for t in range(4)]
    arena = PinnedArena(8 * (nb * 256 * 4 + nb * 8 + 4096) + 65536)
    dense_c = [np.zeros(nb * 256, np.int16) for _ in range(4)]
    dense_h = [np.zeros((nb, 4), np.uint8) for _ in range(4)]
    outs = [_sparse_outputs(arena, nb) for _ in range(4)]
    with Engine(w, h, qt, nslots=10, max_jobs=4) as e:
        # slot 0 = key frame of frame 0 (reference of every P job below)
        e.encode_submit([EncodeJob(PFV_FRAME_I, 0, frames[0], dense_c[0])])
        e.sync()
        kinds = [PFV_FRAME_I, PFV_FRAME_P, PFV_FRAME_P, PFV_FRAME_I]
        e.encode_submit([EncodeJob(k, 1 + i, frames[i], dense_c[i], ref_slot=0, px_err=px_err, hdr_out=dense_h[i])
                         for i, k in enumerate(kinds)])
        e.sync()
        want_recon = [e.slot_read(1 + i) for i in range(4)]
        e.encode_submit_sparse([SparseEncodeJob(k, 5 + i, frames[i], outs[i][0], outs[i][1], ref_slot=0, px_err=px_err,
                                                hdr_out=outs[i][3], mb_off_out=outs[i][2]) for i, k in enumerate(kinds)])
        e.sync()
        for i, k in enumerate(kinds):
            tok, st, off, hdr = outs[i]
            assert np.array_equal(e.slot_read(5 + i), want_recon[i])
            hd = dense_h[i] if k == PFV_FRAME_P else None
            if hd is not None:
                assert np.array_equal(hdr, hd)
            t_tok, t_st, t_off = codec.tokenize(geo, k, dense_c[i], hd)
            n = int(st[N.PFV_TOKSTATS_NTOK])
            assert np.array_equal(tok[:n], t_tok) and np.array_equal(st, t_st) and np.array_equal(off, t_off)
        # too small a buffer: flagged, the first tok_cap entries are still the right ones
        n_full = int(outs[0][1][N.PFV_TOKSTATS_NTOK])
        small_tok, small_st, _, _ = _sparse_outputs(arena, nb, cap=n_full // 2)
        small_tok[:] = 0xdeadbeef
        guard = arena.take((64,), np.uint32)
        guard[:] = 0x5a5a5a5a
        e.encode_submit_sparse([SparseEncodeJob(PFV_FRAME_I, 9, frames[0], small_tok, small_st, tok_cap=n_full // 2)])
        e.sync()
        assert int(small_st[N.PFV_TOKSTATS_FLAGS]) & N.PFV_TOKFLAG_OVERFLOW and int(small_st[N.PFV_TOKSTATS_NTOK]) == n_full
        assert np.array_equal(small_tok[:n_full // 2], outs[0][0][:n_full // 2])
        assert (guard == 0x5a5a5a5a).all()
        with pytest.raises(PfvError, match="more than the token buffer"):
            codec.encode_packet_tokens(geo, PFV_FRAME_I, small_tok, small_st)
        # pageable host memory is refused: the device could not write it
        with pytest.raises(PfvError, match="pinned"):
            e.encode_submit_sparse([SparseEncodeJob(PFV_FRAME_I, 9, frames[0], np.zeros(nb * 256, np.uint32),
                                                    np.zeros(N.PFV_TOKSTATS_WORDS, np.uint32))])
        # device-resident outputs
        d_tok = torch.zeros(nb * 256, dtype=torch.int32, device="cuda")
        d_st = torch.zeros(N.PFV_TOKSTATS_WORDS, dtype=torch.int32, device="cuda")
        e.encode_submit_sparse([SparseEncodeJob(PFV_FRAME_I, 9, frames[0], d_tok.data_ptr(), d_st.data_ptr(), tok_cap=nb * 256)])
        e.sync()
        assert np.array_equal(d_st.cpu().numpy().view(np.uint32), outs[0][1])
        assert np.array_equal(d_tok.cpu().numpy().view(np.uint32)[:n_full], outs[0][0][:n_full])
    arena.close()


def test_interleaved_encode_and_decode_submits_share_the_staging_rings():
    """One context, 24 submits alternating between the sparse encode seam and the decode seam without a sync in between: every
    staging stage is reused several times by both kinds (the store kernel of an encode submit and the copies of a decode submit must
    not overtake each other)."""
    w, h, quality = 176, 144, 6
    qt, px_err = make_qtables(quality)
    geo = geometry_for(w, h)
    nb = geo.nb
    sv = SynthVideo(w, h, 1234, "moving")
    frames = [sv.frame(t) for t in range(12)]
    arena = PinnedArena(12 * (nb * 256 * 4 + 8192) + 12 * (w * h * 2) + 65536)
    outs = [_sparse_outputs(arena, nb) for _ in range(12)]
    pics = [(arena.take((h, w), np.uint8), arena.take((h // 2, w // 2), np.uint8), arena.take((h // 2, w // 2), np.uint8)) for _ in range(12)]
    dense = []
    with Engine(w, h, qt, nslots=4) as e:
        for t in range(12):                                          # the expected coefficients, one frame at a time
            c = np.zeros(nb * 256, np.int16)
            e.encode_submit([EncodeJob(PFV_FRAME_I, 0, frames[t], c)])
            e.sync()
            dense.append(c)
        for t in range(12):
            e.encode_submit_sparse([SparseEncodeJob(PFV_FRAME_I, 1, frames[t], outs[t][0], outs[t][1], mb_off_out=outs[t][2])])
            e.decode_submit([DecodeJob(PFV_FRAME_I, 2 + (t & 1), dense[t], (0, 1, 1), out=pics[t])])
        e.sync()
        for t in range(12):
            tok, st, off = codec.tokenize(geo, PFV_FRAME_I, dense[t])
            n = int(outs[t][1][N.PFV_TOKSTATS_NTOK])
            assert np.array_equal(outs[t][0][:n], tok) and np.array_equal(outs[t][1], st) and np.array_equal(outs[t][2], off)
        og = pfvo.geometry_for(w, h)
        for t in (0, 5, 11):
            fr = pfvo.frame_init(og)
            pfvo.decode_iframe_coeffs(og, qt, (0, 1, 1), dense[t], fr)
            y, u, v = pfvo.crop_frame(og, fr)
            assert np.array_equal(pics[t][0], y) and np.array_equal(pics[t][1], u) and np.array_equal(pics[t][2], v)
    arena.close()


def test_sparse_encode_of_an_all_zero_frame():
    """Every macroblock all zeros: 17 escapes + one run entry each (src/rle.rs:31-38), the longest escape chains the tokenizer
    can meet.  (An unrepresentable coefficient cannot come out of the transform, |c| < 2^14; the RANGE flag is exercised on the
    host restatement in tests/test_codec_cpu.py.)"""
    w, h = 64, 48
    qt, _ = make_qtables(5)
    geo = geometry_for(w, h)
    y = np.full((h, w), 128, np.uint8)                               # (px - 128) = 0 everywhere
    u = np.full((h // 2, w // 2), 128, np.uint8)
    arena = PinnedArena(geo.nb * 256 * 4 + 65536)
    tok, st, off, _ = _sparse_outputs(arena, geo.nb)
    with Engine(w, h, qt) as e:
        e.encode_submit_sparse([SparseEncodeJob(PFV_FRAME_I, 1, (y, u, u), tok, st, mb_off_out=off)])
        e.sync()
    want = pfvo.encode_iframe_coeffs(pfvo.geometry_for(w, h), qt, y, u, u, pfvo.frame_init(pfvo.geometry_for(w, h)))
    o_tok, o_table, o_off = pfvo.rle_frame(want)
    n = int(st[N.PFV_TOKSTATS_NTOK])
    assert np.array_equal(tok[:n], o_tok) and np.array_equal(off, o_off)
    assert np.array_equal(st[:16].astype(np.int64) + st[16:32], o_table)
    arena.close()


@pytest.mark.parametrize("dense", ["0", "1"])
@pytest.mark.parametrize("size,quality,kind,n,key", [((96, 64), 3, "moving", 10, 4), ((130, 70), 5, "moving", 7, 3),
                                                     ((64, 48), 10, "random", 5, 2), ((176, 144), 0, "moving", 5, 5),
                                                     ((64, 64), 5, "static", 4, 4), ((320, 240), 2, "moving", 8, 60)])
def test_encoder_stream_is_byte_identical_to_oracle_encoder(size, quality, kind, n, key, dense, monkeypatch):
    monkeypatch.setenv("PFV_ENCODER_DENSE", dense)                   # both seams: device run-length pass (default) and host
    w, h = size
    drop = (5,) if n > 6 else ()
    want, _ = oracle_stream(w, h, n, quality, key, 2468, kind=kind, drop_at=drop)
    sv = SynthVideo(w, h, 2468, kind=kind)
    with codec.Encoder(w, h, 30, quality, num_threads=3) as enc:
        for t in range(n):
            if t in drop:
                enc.encode_dropframe()
            elif t % key == 0:
                enc.encode_iframe(sv.frame(t))
            else:
                enc.encode_pframe(sv.frame(t))
        enc.finish()
        mine = enc.bytes()
        prev = enc.prev_frame()
        with pytest.raises(PfvError):                                # assert!(!self.finished)
            enc.encode_iframe(sv.frame(0))
        with pytest.raises(PfvError):
            enc.finish()
    assert mine == want
    _, fb = oracle_decode_all(want)
    assert np.array_equal(prev, fb)                                  # closed loop: encoder recon == decoder picture


def test_encoder_stream_buffer_grows_in_place():
    """The Encoder's own stream buffer starts as an 8 MB mapping and grows by mremap (pfv_codec.cpp, GrowBuf): a stream of noise
    at quality 10 passes that size and must still equal the oracle encoder's, byte for byte."""
    w, h, n, key, quality = 640, 480, 48, 4, 10
    want, _ = oracle_stream(w, h, n, quality, key, 1357, kind="random")
    assert len(want) > (8 << 20), len(want)                          # (10.7 MB)
    sv = SynthVideo(w, h, 1357, kind="random")
    with codec.Encoder(w, h, 30, quality, num_threads=4) as enc:
        for t in range(n):
            (enc.encode_iframe if t % key == 0 else enc.encode_pframe)(sv.frame(t))
        enc.finish()
        mine = enc.bytes()
    assert mine == want


def test_encoder_writer_and_decoder_reader_stream_through_callbacks():
    """Encoder<W: Write> / Decoder<R: Read + Seek> (src/enc.rs:12-26, src/dec.rs:15-28): the stream leaves through a writer
    as packets finish (header first, same bytes as the in-memory encoder) and comes back in through a reader that hands out
    odd-sized chunks; a failing writer surfaces as an I/O error."""
    import io
    w, h, n, key = 176, 144, 9, 4
    want, _ = oracle_stream(w, h, n, 4, key, 97)
    sv = SynthVideo(w, h, 97)
    sink = io.BytesIO()
    sizes = []

    class W:
        def write(self, b):
            sizes.append(len(b))
            sink.write(b)

    with codec.Encoder(w, h, 30, 4, num_threads=2, writer=W()) as enc:
        assert sink.getvalue()[:8] == b"PFVIDEO\0"                   # the header goes out at once (write_header, src/enc.rs:190-219)
        for t in range(n):
            (enc.encode_iframe if t % key == 0 else enc.encode_pframe)(sv.frame(t))
        enc.finish()
        assert enc.bytes() == b""
    assert sink.getvalue() == want
    assert len(sizes) >= n + 2                                       # header, one write per packet, eof

    class R:
        def __init__(self, data):
            self.b, self.i = io.BytesIO(data), 0

        def read(self, nbytes):
            self.i += 1
            return self.b.read(min(nbytes, 1 + (self.i * 7919) % 5000))

    planes_a, planes_b = [], []
    with codec.Decoder(R(want), num_threads=2) as dec:
        while dec.advance_frame(lambda fr: planes_a.append(tuple(p.copy() for p in fr))):
            pass
    with codec.Decoder(want, num_threads=2) as dec:
        while dec.advance_frame(lambda fr: planes_b.append(tuple(p.copy() for p in fr))):
            pass
    assert len(planes_a) == len(planes_b) == n
    for a, b in zip(planes_a, planes_b):
        assert all(np.array_equal(x, y) for x, y in zip(a, b))

    class Broken:
        def write(self, b):
            raise OSError("disk full")

    with pytest.raises(PfvError):
        codec.Encoder(w, h, 30, 4, writer=Broken())


def test_encoder_argument_errors():
    with pytest.raises(PfvError):
        codec.Encoder(64, 48, 30, 11)                                # assert!(quality >= 0 && quality <= 10)
    with pytest.raises(PfvError):
        codec.Encoder(63, 48, 30, 5)                                 # odd size (src/frame.rs:13)
    with codec.Encoder(64, 48, 30, 5) as enc:
        with pytest.raises(PfvError):
            enc.encode_iframe((np.zeros((48, 32), np.uint8), np.zeros((24, 32), np.uint8), np.zeros((24, 32), np.uint8)))
        assert enc.bytes()[:8] == b"PFVIDEO\0"
    # close() without finish() still terminates the stream (Drop, src/enc.rs:28-34) - nothing to observe but no hang


def test_round_trip_1080p_encoder_to_decoder():
    """Full-size property: our Encoder's stream decodes (our Decoder) to exactly the encoder's own reconstruction,
    frame after frame (closed loop, src/enc.rs:85/135), and the oracle agrees on a sample frame."""
    w, h = 1920, 1080
    sv = SynthVideo(w, h, 31337)
    recon = []
    with codec.Encoder(w, h, 30, 5, num_threads=4) as enc:
        for t in range(5):
            (enc.encode_iframe if t == 0 else enc.encode_pframe)(sv.frame(t))
            recon.append(enc.prev_frame())
        enc.finish()
        data = enc.bytes()
    got, fb = gpu_decode_all(data, num_threads=4)
    assert len(got) == 5 and np.array_equal(fb, recon[-1])
    og = pfvo.geometry_for(w, h)
    for fr, rec in zip(got, recon):
        y, u, v = pfvo.crop_frame(og, rec)
        assert np.array_equal(fr[0], y) and np.array_equal(fr[1], u) and np.array_equal(fr[2], v)
    want, want_fb = oracle_decode_all(data)
    assert np.array_equal(want_fb, fb)


def test_round_trip_4k_p_stream_sharded_like_config5():
    """BASELINE config 5 at full size (3840x2160 P-frame stream, GOPs sharded): two GOPs encoded on the GPU, each decoded from its
    own sub-stream (the per-rank unit of shard.py), equal to the whole-stream decode, to the encoder's reconstruction and - for the
    last frame of the stream - to the oracle decoder."""
    from pretty_fast_video_b200 import shard
    w, h, gop = 3840, 2160, 3
    sv = SynthVideo(w, h, 4711)
    with codec.Encoder(w, h, 30, 5, num_threads=4) as enc:
        for t in range(2 * gop):
            (enc.encode_iframe if t % gop == 0 else enc.encode_pframe)(sv.frame(t))
        last_recon = enc.prev_frame()
        enc.finish()
        data = enc.bytes()
    whole, fb = gpu_decode_all(data, num_threads=4)
    _, want_fb = oracle_decode_all(data)
    assert len(whole) == 2 * gop
    # (each side against the oracle's decode of the stream, so that a failure says which one is off)
    assert np.array_equal(last_recon, want_fb), f"encoder reconstruction: {int((last_recon != want_fb).sum())} bytes differ from the oracle's decode of its stream"
    assert np.array_equal(fb, want_fb), f"decoder framebuffer: {int((fb != want_fb).sum())} bytes differ from the oracle's decode"
    merged = []
    for rank in range(2):
        info, gops, mine = shard.plan(data, rank, 2)
        assert [g.nframes for g in gops] == [gop, gop] and len(mine) == 1
        for g in mine:
            part, _ = gpu_decode_all(shard.substream(data, info.first_packet, g), num_threads=2)
            merged += [(g.first_frame + i, fr) for i, fr in enumerate(part)]
    same_frames(shard.gather_ordered(merged), whole)


def test_gop_sharded_gpu_decode_equals_whole_stream():
    """The multi-GPU partitioning (shard.py) with the GPU Decoder as the per-rank worker: every GOP decoded from its
    own sub-stream in its own context, merged by display index, equals the whole-stream decode (no collective)."""
    from pretty_fast_video_b200 import shard
    data, _ = oracle_stream(176, 144, 17, 3, 4, 31, drop_at=(6,))
    whole, _ = gpu_decode_all(data, num_threads=2)
    world = 3
    merged = []
    for rank in range(world):
        info, gops, mine = shard.plan(data, rank, world)
        for g in mine:
            part, _ = gpu_decode_all(shard.substream(data, info.first_packet, g), num_threads=2)
            assert len(part) == g.nframes
            merged += [(g.first_frame + i, fr) for i, fr in enumerate(part)]
    same_frames(shard.gather_ordered(merged), whole)


def test_decoder_reproduces_the_committed_golden_stream():
    """tests/golden/stream_96x64_q3.npz (made by tests/golden/make_golden.py, frozen): the GPU Decoder yields the
    per-frame SHA-256 digests stored with the stream, and the GPU Encoder re-creates the stored bytes."""
    import hashlib
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    g = np.load(os.path.join(root, "tests", "golden", "stream_96x64_q3.npz"))
    data = g["stream"].tobytes()
    got, _ = gpu_decode_all(data, num_threads=2)
    sums = [hashlib.sha256(b"".join(p.tobytes() for p in fr)).hexdigest() for fr in got if fr is not None]
    assert sums == [s for s in g["sha256"]]
    assert [fr is None for fr in got].count(True) == 1              # the drop frame at t = 5
    sv = SynthVideo(96, 64, 77)
    with codec.Encoder(96, 64, 24, 3, num_threads=2) as enc:
        for t in range(int(g["nframes"])):
            if t % 4 == 0:
                enc.encode_iframe(sv.frame(t))
            elif t == 5:
                enc.encode_dropframe()
            else:
                enc.encode_pframe(sv.frame(t))
        enc.finish()
        assert enc.bytes() == data


def test_unknown_packets_are_skipped_and_p_first_streams_decode():
    """src/dec.rs:216-219: a packet of an unknown type is skipped over by its length.  And a stream whose first picture
    is a P frame decodes against the blank initial framebuffer (Y = 0, U = V = 128, src/frame.rs:38-43)."""
    data, _ = oracle_stream(96, 64, 6, 3, 3, 9)
    info, _ = codec.parse_header(data)
    pk, _ = codec.index_packets(data, info.first_packet)
    junk = bytes([7]) + (11).to_bytes(4, "little") + bytes(range(11))
    cut = pk[2][2] - 5                                               # in front of the third frame packet
    spliced = data[:cut] + junk + data[cut:]
    want, want_fb = oracle_decode_all(spliced)
    got, fb = gpu_decode_all(spliced, num_threads=2)
    same_frames(got, want)
    assert np.array_equal(fb, want_fb)
    base, _ = oracle_decode_all(data)
    same_frames(got, base)                                           # the junk packet changes nothing
    # drop the leading key frame: the stream now starts with two P frames
    headless = data[:info.first_packet] + data[pk[1][2] - 5:]
    want, want_fb = oracle_decode_all(headless)
    got, fb = gpu_decode_all(headless, num_threads=2)
    same_frames(got, want)
    assert np.array_equal(fb, want_fb)


def test_cpp_decode_speed_example_on_the_test2_geometry(tmp_path):
    """examples/decode_speed.cpp = src/lib.rs:310-335 (test_decode_speed_2) on the C ABI: 512x384, 161 frames, fps 30,
    quality 2, a key frame every 60 (the parameters of test_encode_2, src/lib.rs:271-292).  The stream it writes is
    also decoded by the oracle and by the Python mirror: same pictures."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["make", "-C", os.path.join(root, "examples")], stdout=subprocess.DEVNULL)
    exe = os.path.join(root, "examples", "decode_speed")
    path = str(tmp_path / "test2_like.pfv")
    subprocess.check_call([exe, "--make", path, "512", "384", "161"], stdout=subprocess.DEVNULL)
    out = subprocess.run([exe, path, "3", "6"], capture_output=True, text=True, check=True).stdout.splitlines()
    assert [l.split()[1] for l in out if l.startswith("Decoded")] == ["161"] * 3
    data = open(path, "rb").read()
    want, want_fb = oracle_decode_all(data)
    got, fb = gpu_decode_all(data, num_threads=4)
    assert len(want) == 161
    same_frames(got, want)
    assert np.array_equal(fb, want_fb)
