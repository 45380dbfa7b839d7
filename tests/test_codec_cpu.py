"""CPU: the host side of the codec (csrc/pfv_codec.cpp: container, Huffman/RLE entropy layer) against the oracle's
restatement of src/enc.rs:190-481, src/dec.rs:38-448, src/rle.rs, src/huffman.rs and the committed golden stream.
No compute kernels are called here (no GPU in this tier)."""
import os

import numpy as np
import pytest

import pfvo
from pretty_fast_video_b200 import PFV_FRAME_I, PFV_FRAME_P, geometry_for
from pretty_fast_video_b200 import codec
from pretty_fast_video_b200.synth import SynthVideo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def oracle_stream(w, h, n, quality, key_every, seed, kind="moving", drop_at=()):
    """-> (stream bytes, [(kind, hdr, coeff)] per coded frame) from the oracle encoder."""
    sv = SynthVideo(w, h, seed, kind=kind)
    enc = pfvo.Encoder(w, h, 30, quality, nthreads=2)
    seam = []
    for t in range(n):
        y, u, v = sv.frame(t)
        if t in drop_at:
            enc.encode_dropframe()
        elif t % key_every == 0:
            enc.encode_iframe(y, u, v)
            seam.append((PFV_FRAME_I, None, enc.last_coeffs().copy()))
        else:
            enc.encode_pframe(y, u, v)
            seam.append((PFV_FRAME_P, enc.last_headers().copy(), enc.last_coeffs().copy()))
    enc.finish()
    return enc.bytes(), seam


def frame_packets(data):
    info, qt = codec.parse_header(data)
    pk, trunc = codec.index_packets(data, info.first_packet)
    return info, qt, pk, trunc


def test_header_and_index_of_golden_stream():
    g = np.load(os.path.join(ROOT, "tests", "golden", "stream_96x64_q3.npz"))
    data = g["stream"].tobytes()
    info, qt, pk, trunc = frame_packets(data)
    assert (info.version, info.width, info.height, info.framerate, info.num_qtables) == (211, 96, 64, 24, 4)
    assert info.first_packet == 532
    assert np.array_equal(qt, pfvo.make_qtables(3)[0])
    assert not trunc
    types = [t for t, _, _ in pk]
    # 10 frames: I at 0,4,8; drop frame (type 1, len 0) at 5; EOF last
    assert types == [1, 2, 2, 2, 1, 1, 2, 2, 1, 2, 0]
    assert pk[5][1] == 0 and all(l > 0 for t, l, _ in pk[:5])
    off = 532
    for t, l, p in pk:
        assert p == off + 5
        off = p + (l if t != 0 else 0)


def test_header_errors_match_the_reference_classes():
    enc = pfvo.Encoder(32, 32, 30, 5)
    enc.finish()
    good = enc.bytes()
    with pytest.raises(codec.DecodeError) as e:
        codec.parse_header(b"NOTAPFV\0" + good[8:])
    assert e.value.kind == "FormatError"                             # src/dec.rs:48-52
    with pytest.raises(codec.DecodeError) as e:
        codec.parse_header(good[:8] + (210).to_bytes(4, "little") + good[12:])
    assert e.value.kind == "VersionError"                            # src/dec.rs:55-60
    for cut in (3, 10, 19, 100, 531):
        with pytest.raises(codec.DecodeError) as e:
            codec.parse_header(good[:cut])
        assert e.value.kind == "IOError"
    info, _ = codec.parse_header(good)
    pk, trunc = codec.index_packets(good, info.first_packet)
    assert pk == [(0, 0, 537)] and not trunc
    pk, trunc = codec.index_packets(good[:534], info.first_packet)   # cut inside the EOF packet
    assert pk == [] and trunc
    pk, trunc = codec.index_packets(good[:532], info.first_packet)   # no EOF packet at all
    assert pk == [] and trunc


@pytest.mark.parametrize("size,quality,kind", [((96, 64), 3, "moving"), ((176, 144), 0, "moving"), ((64, 48), 10, "random"),
                                               ((130, 70), 5, "moving"), ((64, 64), 5, "static")])
def test_entropy_decode_matches_oracle_seam(size, quality, kind):
    """pfv_packet_decode: tokens -> dense must equal the dense coefficients / headers the oracle decoder produced."""
    w, h = size
    data, seam = oracle_stream(w, h, 6, quality, 3, 1234, kind=kind)
    info, qt, pk, _ = frame_packets(data)
    geo = geometry_for(w, h)
    dec = pfvo.Decoder(data)
    k = 0
    for t, l, p in pk:
        if t == 0 or l == 0:
            continue
        more, fr = dec.advance_frame()
        fk = PFV_FRAME_I if t == 1 else PFV_FRAME_P
        qidx, hdr, mb_off, tok = codec.decode_packet(geo, fk, data[p:p + l])
        assert tuple(qidx) == tuple(dec.last_qidx())
        dense = codec.tokens_to_dense(geo.nb, mb_off, tok)
        assert np.array_equal(dense, dec.last_coeffs())
        assert np.array_equal(dense, seam[k][2] if fk == PFV_FRAME_I else np.where(np.repeat(seam[k][1][:, 2] != 0, 256), seam[k][2], 0))
        if fk == PFV_FRAME_P:
            assert np.array_equal(hdr, dec.last_headers())
            assert np.array_equal(hdr, seam[k][1])
            # a macroblock without coefficients owns no tokens
            cnt = np.diff(mb_off.astype(np.int64))
            assert (cnt[hdr[:, 2] == 0] == 0).all()
        # tokens are emitted in increasing position order inside a macroblock
        pos = (tok >> 16).astype(np.int64) + 256 * np.repeat(np.arange(geo.nb), np.diff(mb_off.astype(np.int64)))
        assert (np.diff(pos) > 0).all()
        k += 1
    assert k == len(seam)


@pytest.mark.parametrize("size,quality,kind", [((96, 64), 3, "moving"), ((176, 144), 0, "moving"), ((64, 48), 10, "random"),
                                               ((64, 64), 5, "static")])
def test_entropy_encode_is_byte_identical_to_oracle(size, quality, kind):
    """pfv_packet_encode on the oracle's seam data reproduces the oracle's packets byte for byte (RLE split rules,
    weight normalisation, tree shape and tie-breaks, LSB-first packing, zero padding)."""
    w, h = size
    data, seam = oracle_stream(w, h, 6, quality, 3, 99, kind=kind)
    info, qt, pk, _ = frame_packets(data)
    geo = geometry_for(w, h)
    frames = [(t, l, p) for t, l, p in pk if t != 0 and l > 0]
    assert len(frames) == len(seam)
    for (t, l, p), (fk, hdr, coeff) in zip(frames, seam):
        mine = codec.encode_packet(geo, fk, coeff, hdr)
        assert mine == data[p - 5:p + l]


def test_golden_stream_round_trips_through_the_entropy_layer():
    g = np.load(os.path.join(ROOT, "tests", "golden", "stream_96x64_q3.npz"))
    data = g["stream"].tobytes()
    info, qt, pk, _ = frame_packets(data)
    geo = geometry_for(96, 64)
    for t, l, p in pk:
        if t == 0 or l == 0:
            continue
        fk = PFV_FRAME_I if t == 1 else PFV_FRAME_P
        qidx, hdr, mb_off, tok = codec.decode_packet(geo, fk, data[p:p + l])
        assert codec.encode_packet(geo, fk, codec.tokens_to_dense(geo.nb, mb_off, tok), hdr) == data[p - 5:p + l]


def test_reference_entropy_test_literals():
    """src/lib.rs:96-158 (test_entropy): the ten literals, padded into one macroblock, survive encode -> decode."""
    lits = np.array([10, 0, 0, 5, 0, -3, 0, 0, 0, 1], np.int16)      # src/lib.rs:98
    geo = geometry_for(16, 16)                                        # nb = 1 + 2 (chroma planes pad to 16x16)
    coeff = np.zeros(geo.nb * 256, np.int16)
    coeff[:10] = lits
    coeff[256 + 255] = -1                                             # a run of 255 zeros split into (15,0) tokens
    coeff[512] = 16383                                                # largest size symbol: 14 bits + sign = 15
    coeff[513] = -16383
    pkt = codec.encode_packet(geo, PFV_FRAME_I, coeff)
    qidx, _, mb_off, tok = codec.decode_packet(geo, PFV_FRAME_I, pkt[5:])
    assert tuple(qidx) == (0, 1, 1)
    assert np.array_equal(codec.tokens_to_dense(geo.nb, mb_off, tok), coeff)
    # |v| >= 16384 needs a size symbol >= 16 (debug_assert in src/rle.rs:43): rejected, never written corrupt
    from pretty_fast_video_b200 import PfvError
    coeff[514] = 16384
    with pytest.raises(PfvError):
        codec.encode_packet(geo, PFV_FRAME_I, coeff)


def test_degenerate_trees():
    geo = geometry_for(16, 16)
    # all-zero frame: every token is (15, 0) or (1, 0): symbols {0, 1, 15}
    z = np.zeros(geo.nb * 256, np.int16)
    pkt = codec.encode_packet(geo, PFV_FRAME_I, z)
    _, _, mb_off, tok = codec.decode_packet(geo, PFV_FRAME_I, pkt[5:])
    assert tok.size == 0 and (mb_off == 0).all()
    # P frame with nothing coded: empty histogram -> all-zero weight table, no tokens
    hdr = np.zeros((geo.nb, 4), np.uint8)
    pkt = codec.encode_packet(geo, PFV_FRAME_P, z, hdr)
    assert pkt[5:21] == bytes(16)
    q, h2, mb_off, tok = codec.decode_packet(geo, PFV_FRAME_P, pkt[5:])
    assert tuple(q) == (2, 3, 3) and np.array_equal(h2, hdr) and tok.size == 0


def test_malformed_payloads_fail_loudly():
    geo = geometry_for(32, 32)
    coeff = (np.arange(geo.nb * 256) % 7 - 3).astype(np.int16)
    pkt = codec.encode_packet(geo, PFV_FRAME_I, coeff)
    with pytest.raises(codec.DecodeError) as e:                      # truncated: bit stream ends (io::Error in the reference)
        codec.decode_packet(geo, PFV_FRAME_I, pkt[5:5 + 40])
    assert e.value.kind == "IOError"
    with pytest.raises(codec.DecodeError):
        codec.decode_packet(geo, PFV_FRAME_I, pkt[5:15])
    # motion vector leaving the plane is rejected on the host (src/common.rs:258-259 is only a debug_assert)
    hdr = np.zeros((geo.nb, 4), np.uint8)
    hdr[0, 0] = np.uint8(-3 & 0xFF)
    bad = codec.encode_packet(geo, PFV_FRAME_P, np.zeros(geo.nb * 256, np.int16), hdr)
    from pretty_fast_video_b200 import PfvError
    with pytest.raises(PfvError) as e2:
        codec.decode_packet(geo, PFV_FRAME_P, bad[5:])
    assert e2.value.code == -4


def test_dense_token_helpers_are_inverse():
    rng = np.random.default_rng(5)
    nb = 7
    c = rng.integers(-300, 300, nb * 256).astype(np.int16)
    c[rng.random(nb * 256) < 0.9] = 0
    mb_off, tok = codec.dense_to_tokens(c, nb)
    assert np.array_equal(codec.tokens_to_dense(nb, mb_off, tok), c)


def test_entropy_decoder_survives_corrupted_payloads():
    """Mutated and truncated payloads: pfv_packet_decode returns an error or a (different) token list, never crashes,
    never loops forever, never writes outside its buffers (the reference would panic / return io::Error)."""
    from pretty_fast_video_b200 import PfvError
    rng = np.random.default_rng(2024)
    w, h = 96, 64
    data, seam = oracle_stream(w, h, 4, 3, 2, 55)
    info, qt, pk, _ = frame_packets(data)
    geo = geometry_for(w, h)
    frames = [(t, l, p) for t, l, p in pk if t != 0 and l > 0]
    outcomes = {"ok": 0, "err": 0}
    for t, l, p in frames:
        fk = PFV_FRAME_I if t == 1 else PFV_FRAME_P
        good = bytearray(data[p:p + l])
        for trial in range(150):
            bad = bytearray(good)
            mode = trial % 3
            if mode == 0:                                            # flip a few bytes anywhere (incl. the weight table)
                for _ in range(int(rng.integers(1, 6))):
                    bad[int(rng.integers(0, len(bad)))] ^= int(rng.integers(1, 256))
            elif mode == 1:                                          # truncate
                bad = bad[:int(rng.integers(0, len(bad)))]
            else:                                                    # random weight table, payload kept
                bad[:16] = bytes(rng.integers(0, 256, 16, dtype=np.uint8))
            try:
                qidx, hdr, mb_off, tok = codec.decode_packet(geo, fk, bytes(bad))
                assert mb_off[0] == 0 and mb_off[-1] == tok.size and (np.diff(mb_off.astype(np.int64)) >= 0).all()
                assert ((tok >> 16) < 256).all()
                outcomes["ok"] += 1
            except PfvError:
                outcomes["err"] += 1
    assert outcomes["err"] > 0 and outcomes["ok"] + outcomes["err"] == 150 * len(frames)


def test_single_symbol_tree_is_decodable():
    """A frame whose tokens are all (run 1, size 1): the weight table has one entry, both codes have length zero
    (src/huffman.rs:125-131), every token costs one value bit.  Legal, and far more tokens per byte than usual."""
    geo = geometry_for(16, 16)
    nb = geo.nb
    # every second coefficient is -1 or 0 (size 1 holds only 0 / -1): position pattern z v z v ...
    rng = np.random.default_rng(3)
    vals = -rng.integers(0, 2, nb * 128).astype(np.int16)
    coeff = np.zeros(nb * 256, np.int16)
    coeff[1::2] = vals
    # payload by hand: table with weight only for symbol 1, qidx, then nb*128 value bits LSB-first
    bits = (vals != 0).astype(np.uint8)
    body = np.packbits(bits, bitorder="little").tobytes()
    table = bytes([0, 200] + [0] * 14)
    payload = table + bytes([0, 1, 1]) + body
    qidx, _, mb_off, tok = codec.decode_packet(geo, PFV_FRAME_I, payload)
    assert tok.size == nb * 128                                      # zero-valued tokens are kept: they occupy a position
    assert np.array_equal(codec.tokens_to_dense(nb, mb_off, tok), coeff)


# ---- sparse encode seam: host restatement of the device tokenizer + the writer that takes its output -------------------
@pytest.mark.parametrize("size,quality,kind", [((96, 64), 3, "moving"), ((176, 144), 0, "moving"), ((64, 48), 10, "random")])
def test_tokenize_matches_oracle_rle_and_packets_are_byte_identical(size, quality, kind):
    """pfv_packet_tokenize == rle_encode per macroblock + update_table of the oracle (src/rle.rs:9-47); the packet written from
    (tok, stats) is the oracle's packet byte for byte."""
    from pretty_fast_video_b200 import _native as N
    w, h = size
    data, seam = oracle_stream(w, h, 6, quality, 3, 7, kind=kind)
    info, qt, pk, _ = frame_packets(data)
    geo = geometry_for(w, h)
    frames = [(t, l, p) for t, l, p in pk if t != 0 and l > 0]
    for (t, l, p), (fk, hdr, coeff) in zip(frames, seam):
        coded = None if fk == PFV_FRAME_I else np.asarray(hdr).reshape(-1, 4)[:, 2]
        o_tok, o_table, o_off = pfvo.rle_frame(coeff, coded)
        tok, stats, mb_off = codec.tokenize(geo, fk, coeff, hdr)
        assert np.array_equal(tok, o_tok)
        assert np.array_equal(mb_off, o_off)
        assert np.array_equal(stats[:16].astype(np.int64) + stats[16:32], o_table)
        assert int(stats[N.PFV_TOKSTATS_NTOK]) == o_tok.size and int(stats[N.PFV_TOKSTATS_FLAGS]) == 0
        # per-symbol split: [0..15] = num_zeroes, [16..31] = coeff_size
        assert np.array_equal(stats[:16], np.bincount(o_tok & 15, minlength=16))
        assert np.array_equal(stats[16:32], np.bincount((o_tok >> 4) & 15, minlength=16))
        assert codec.encode_packet_tokens(geo, fk, tok, stats, hdr) == data[p - 5:p + l]


def test_tokenize_edge_cases():
    from pretty_fast_video_b200 import _native as N
    geo = geometry_for(16, 16)                                        # 3 macroblocks
    coeff = np.zeros(geo.nb * 256, np.int16)
    coeff[255] = 7                  # 255 zeros first: 16 escapes + (15, size)
    coeff[256 + 16] = -1            # run of exactly 16: one escape + (1, size); tail 239 zeros = 15 escapes + (14, 0)
    # third macroblock all zero: 17 escapes + (1, 0)
    tok, stats, mb_off = codec.tokenize(geo, PFV_FRAME_I, coeff)
    o_tok, o_table, o_off = pfvo.rle_frame(coeff)
    assert np.array_equal(tok, o_tok) and np.array_equal(mb_off, o_off)
    assert list(np.diff(mb_off)) == [17, 18, 18]
    assert np.array_equal(stats[:16].astype(np.int64) + stats[16:32], o_table)
    pkt = codec.encode_packet_tokens(geo, PFV_FRAME_I, tok, stats)
    assert pkt == codec.encode_packet(geo, PFV_FRAME_I, coeff)
    _, _, off2, tok2 = codec.decode_packet(geo, PFV_FRAME_I, pkt[5:])
    assert np.array_equal(codec.tokens_to_dense(geo.nb, off2, tok2), coeff)
    # a token buffer that is too small and an unrepresentable coefficient are reported, never written
    t_small, st_small, _ = codec.tokenize(geo, PFV_FRAME_I, coeff, tok_cap=10)
    assert int(st_small[N.PFV_TOKSTATS_FLAGS]) & N.PFV_TOKFLAG_OVERFLOW and int(st_small[N.PFV_TOKSTATS_NTOK]) == 53
    with pytest.raises(Exception, match="more than the token buffer"):
        codec.encode_packet_tokens(geo, PFV_FRAME_I, t_small, st_small)
    coeff[3] = -16384
    t_bad, st_bad, _ = codec.tokenize(geo, PFV_FRAME_I, coeff)
    assert int(st_bad[N.PFV_TOKSTATS_FLAGS]) & N.PFV_TOKFLAG_RANGE
    with pytest.raises(Exception, match="not representable"):
        codec.encode_packet_tokens(geo, PFV_FRAME_I, t_bad, st_bad)
    # statistics that do not describe the sequence are caught by the exact-size check
    st_wrong = stats.copy()
    st_wrong[16 + 4] += 1
    with pytest.raises(Exception, match="disagree"):
        codec.encode_packet_tokens(geo, PFV_FRAME_I, tok, st_wrong)
