"""Host-side mirror of the reference's public objects: ``pfv_rs::dec::Decoder`` (src/dec.rs:15-224) and
``pfv_rs::enc::Encoder`` (src/enc.rs:12-188), plus the container / entropy helpers under them.

Everything forwards to libpfv_b200.so (csrc/pfv_codec.cpp on top of the hot-path C ABI); nothing is computed in
Python.  Names, argument meaning and error classes follow the reference:

    Decoder(data, num_threads)            Decoder::new(reader, num_threads)      -> FormatError / VersionError / IOError
      .width() .height() .framerate()     src/dec.rs:136-146
      .reset()                            src/dec.rs:148
      .advance_frame(onvideo) -> bool     src/dec.rs:169
      .advance_delta(delta, onvideo)      src/dec.rs:154
    Encoder(width, height, framerate, quality, num_threads)   Encoder::new (the writer is an in-memory buffer: .bytes())
      .encode_iframe(frame) .encode_pframe(frame) .encode_dropframe() .finish()
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional, Tuple

import numpy as np

from . import _native as N
from ._native import PfvError


class DecodeError(PfvError):
    """DecodeError of src/dec.rs:31-35; ``kind`` is 'FormatError', 'VersionError' or 'IOError'."""

    def __init__(self, code: int, msg: str):
        super().__init__(code, msg)
        self.kind = {N.PFV_ERR_BAD_STREAM: "FormatError", N.PFV_ERR_BAD_VERSION: "VersionError",
                     N.PFV_ERR_IO: "IOError"}.get(code, "Other")


def _check_dec(rc: int):
    if rc < 0:
        msg = N.lib().pfv_last_error().decode("utf-8", "replace")
        if rc in (N.PFV_ERR_BAD_STREAM, N.PFV_ERR_BAD_VERSION, N.PFV_ERR_IO):
            raise DecodeError(rc, msg)
        raise PfvError(rc, msg)


# ---- container / entropy helpers (host only: usable without a GPU) ---------------------------------
def parse_header(data: bytes):
    """-> (StreamInfo, qtables int32[nq,64]); raises DecodeError like Decoder::new (src/dec.rs:38-118)."""
    buf = np.frombuffer(data, np.uint8)
    info = N.StreamInfo()
    _check_dec(N.lib().pfv_stream_parse_header(buf.ctypes.data, buf.size, C.byref(info), None, 0))
    qt = np.zeros((info.num_qtables, 64), np.int32)
    _check_dec(N.lib().pfv_stream_parse_header(buf.ctypes.data, buf.size, C.byref(info), qt.ctypes.data, info.num_qtables))
    return info, qt


def index_packets(data: bytes, offset: int):
    """Packet scan (u8 type + u32 len per packet, src/dec.rs:179-180) -> ([(type, len, payload_offset)], truncated)."""
    buf = np.frombuffer(data, np.uint8)
    cap = 1024
    while True:
        arr = (N.Packet * cap)()
        n, tr = C.c_uint32(), C.c_int()
        N.check(N.lib().pfv_stream_index(buf.ctypes.data, buf.size, offset, arr, cap, C.byref(n), C.byref(tr)))
        if n.value < cap:
            return [(arr[i].type, arr[i].len, arr[i].payload) for i in range(n.value)], bool(tr.value)
        cap *= 4


def decode_packet(geo: N.Geometry, kind: int, payload: bytes):
    """Entropy-decodes one frame payload -> (qidx[3], hdr uint8[nb,4] or None, mb_off uint32[nb+1], tok uint32[ntok])."""
    buf = np.frombuffer(payload, np.uint8)
    nb = geo.nb
    qidx = np.zeros(3, np.uint8)
    hdr = np.zeros((nb, 4), np.uint8)
    mb_off = np.zeros(nb + 1, np.uint32)
    cap = int(N.lib().pfv_packet_token_bound(C.byref(geo), buf.ctypes.data, buf.size))
    tok = np.zeros(max(cap, 1), np.uint32)
    ntok = C.c_uint32()
    _check_dec(N.lib().pfv_packet_decode(C.byref(geo), kind, buf.ctypes.data, buf.size, qidx.ctypes.data, hdr.ctypes.data,
                                         mb_off.ctypes.data, tok.ctypes.data, cap, C.byref(ntok)))
    return qidx, (hdr if kind == N.PFV_FRAME_P else None), mb_off, tok[:ntok.value].copy()


def tokens_to_dense(nb: int, mb_off: np.ndarray, tok: np.ndarray) -> np.ndarray:
    """The scatter of src/dec.rs:288 / :410: sparse tokens -> dense int16[nb*256] (test aid)."""
    dense = np.zeros(nb * 256, np.int16)
    mb = np.repeat(np.arange(nb, dtype=np.int64), np.diff(mb_off.astype(np.int64)))
    dense[mb * 256 + (tok >> 16).astype(np.int64)] = (tok & 0xFFFF).astype(np.uint16).view(np.int16)
    return dense


def dense_to_tokens(coeff: np.ndarray, nb: int):
    """dense int16[nb*256] -> (mb_off uint32[nb+1], tok uint32[ntok]) in the order the entropy decoder emits them."""
    c = np.ascontiguousarray(coeff, np.int16).reshape(nb, 256)
    mbi, pos = np.nonzero(c)
    tok = (pos.astype(np.uint32) << 16) | c[mbi, pos].view(np.uint16).astype(np.uint32)
    mb_off = np.zeros(nb + 1, np.uint32)
    mb_off[1:] = np.cumsum(np.bincount(mbi, minlength=nb))
    return mb_off, np.ascontiguousarray(tok, np.uint32)


def encode_packet(geo: N.Geometry, kind: int, coeff: np.ndarray, hdr: Optional[np.ndarray] = None) -> bytes:
    """Entropy-codes one frame from the dense seam -> packet bytes (5-byte packet header + payload)."""
    c = np.ascontiguousarray(coeff, np.int16)
    h = np.ascontiguousarray(hdr, np.uint8) if hdr is not None else None
    cap = N.lib().pfv_packet_encode_bound(C.byref(geo))
    out = np.empty(cap, np.uint8)
    n = C.c_size_t()
    N.check(N.lib().pfv_packet_encode(C.byref(geo), kind, h.ctypes.data if h is not None else None, c.ctypes.data,
                                      out.ctypes.data, cap, C.byref(n)))
    return out[:n.value].tobytes()


def tokenize(geo: N.Geometry, kind: int, coeff: np.ndarray, hdr: Optional[np.ndarray] = None, tok_cap: Optional[int] = None):
    """Host run-length pass (rle_encode, src/rle.rs:9-39, per macroblock): dense coefficients -> (tok, stats, mb_off) in the
    format of the sparse encode seam.  tok[i] = run | size << 4 | uint16(value) << 16."""
    c = np.ascontiguousarray(coeff, np.int16)
    h = np.ascontiguousarray(hdr, np.uint8) if hdr is not None else None
    cap = geo.nb * 256 if tok_cap is None else int(tok_cap)
    tok = np.empty(max(cap, 1), np.uint32)
    stats = np.zeros(N.PFV_TOKSTATS_WORDS, np.uint32)
    mb_off = np.zeros(geo.nb + 1, np.uint32)
    N.check(N.lib().pfv_packet_tokenize(C.byref(geo), kind, h.ctypes.data if h is not None else None, c.ctypes.data,
                                        tok.ctypes.data, cap, mb_off.ctypes.data, stats.ctypes.data))
    return tok[:min(int(stats[N.PFV_TOKSTATS_NTOK]), cap)], stats, mb_off


def encode_packet_tokens(geo: N.Geometry, kind: int, tok: np.ndarray, stats: np.ndarray, hdr: Optional[np.ndarray] = None) -> bytes:
    """Entropy-codes one frame from the sparse encode seam (RLE sequence + symbol statistics) -> packet bytes."""
    t = np.ascontiguousarray(tok, np.uint32)
    st = np.ascontiguousarray(stats, np.uint32)
    h = np.ascontiguousarray(hdr, np.uint8) if hdr is not None else None
    cap = N.lib().pfv_packet_encode_bound(C.byref(geo))
    out = np.empty(cap, np.uint8)
    n = C.c_size_t()
    N.check(N.lib().pfv_packet_encode_tokens(C.byref(geo), kind, h.ctypes.data if h is not None else None, t.ctypes.data,
                                             st.ctypes.data, out.ctypes.data, cap, C.byref(n)))
    return out[:n.value].tobytes()


# ---- Decoder -------------------------------------------------------------------------------------------
Frame = Tuple[np.ndarray, np.ndarray, np.ndarray]


class Decoder:
    """pfv_rs::dec::Decoder<R> (src/dec.rs:15-224).  `data` is the reader R: bytes-like, or any object with a
    read(n) method (a file, a socket wrapper, io.BytesIO ...), which is then drained through pfv_decoder_open_reader."""

    def __init__(self, data, num_threads: int = 1, device: int = 0, read_ahead: int = 0):
        self._d = C.c_void_p()
        if hasattr(data, "read"):
            def _read(user, buf, cap):
                try:
                    chunk = data.read(min(int(cap), 1 << 20))
                except Exception:
                    return -1
                if not chunk:
                    return 0
                C.memmove(buf, chunk, len(chunk))
                return len(chunk)
            cb = N.READ_FN(_read)
            _check_dec(N.lib().pfv_decoder_open_reader(cb, None, device, num_threads, read_ahead, C.byref(self._d)))
        else:
            self._buf = np.frombuffer(data, np.uint8)      # the reader; must outlive the native decoder
            _check_dec(N.lib().pfv_decoder_open(self._buf.ctypes.data, self._buf.size, device, num_threads, read_ahead,
                                                C.byref(self._d)))
        self._w, self._h = self.width(), self.height()
        self._cache = {}
        self._got, self._y, self._u, self._v = C.c_int(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._adv = N.lib().pfv_decoder_advance_frame
        self._args = (self._d, C.byref(self._got), C.byref(self._y), C.byref(self._u), C.byref(self._v))

    def width(self) -> int:
        return int(N.lib().pfv_decoder_width(self._d))

    def height(self) -> int:
        return int(N.lib().pfv_decoder_height(self._d))

    def framerate(self) -> int:
        return int(N.lib().pfv_decoder_framerate(self._d))

    def reset(self) -> None:
        _check_dec(N.lib().pfv_decoder_reset(self._d))

    def _views(self, y, u, v) -> Frame:
        # the decoder hands out pointers into a small ring of pinned pictures: build the numpy views once per buffer
        fr = self._cache.get(y)
        if fr is None:
            w, h = self._w, self._h
            mk = lambda p, n, shape: np.ctypeslib.as_array((C.c_uint8 * n).from_address(p)).reshape(shape)
            fr = (mk(y, w * h, (h, w)), mk(u, (w // 2) * (h // 2), (h // 2, w // 2)), mk(v, (w // 2) * (h // 2), (h // 2, w // 2)))
            self._cache[y] = fr
        return fr

    def advance_frame(self, onvideo: Callable[[Frame], None]) -> bool:
        """Ok(true) / Ok(false) of src/dec.rs:169-224; onvideo gets views valid until the next call."""
        rc = self._adv(*self._args)
        if rc < 0:
            _check_dec(rc)
        if self._got.value:
            onvideo(self._views(self._y.value, self._u.value, self._v.value))
        return rc == 1

    def advance_delta(self, delta: float, onvideo: Callable[[Frame], None]) -> bool:
        cb = N.ONVIDEO(lambda user, y, u, v: onvideo(self._views(y, u, v)))
        rc = N.lib().pfv_decoder_advance_delta(self._d, float(delta), cb, None)
        _check_dec(rc)
        return rc == 1

    def framebuffer(self) -> np.ndarray:
        """Decoder.framebuffer (padded Y|U|V) as it stands after the last returned picture (test aid)."""
        ctx = N.lib().pfv_decoder_ctx(self._d)
        g = N.Geometry()
        N.check(N.lib().pfv_ctx_geometry(ctx, C.byref(g)))
        out = np.empty(g.frame_bytes, np.uint8)
        N.check(N.lib().pfv_slot_read(ctx, N.lib().pfv_decoder_framebuffer_slot(self._d), out.ctypes.data))
        return out

    def close(self):
        if self._d:
            N.lib().pfv_decoder_close(self._d)
            self._d = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- Encoder -------------------------------------------------------------------------------------------
class Encoder:
    """pfv_rs::enc::Encoder<W> (src/enc.rs:12-188).  `writer` is W: any object with a write(bytes) method; packets reach it in
    stream order as soon as they are finished.  Without one the stream is kept in memory (bytes())."""

    def __init__(self, width: int, height: int, framerate: int, quality: int, num_threads: int = 1, device: int = 0,
                 writer=None):
        self._e = C.c_void_p()
        N.check(N.lib().pfv_encoder_open(width, height, framerate, quality, num_threads, device, C.byref(self._e)))
        self._w, self._h = width, height
        self._wcb = None
        if writer is not None:
            def _write(user, data, n):
                try:
                    writer.write(C.string_at(data, n))
                    return 0
                except Exception:
                    return 1
            self._wcb = N.WRITE_FN(_write)                  # must outlive the native encoder
            N.check(N.lib().pfv_encoder_set_writer(self._e, self._wcb, None))

    def _planes(self, frame):
        y, u, v = (np.ascontiguousarray(p, np.uint8) for p in frame)
        w, h = self._w, self._h
        # assert!(frame.width == self.width ...) src/enc.rs:76-79
        if y.shape != (h, w) or u.shape != (h // 2, w // 2) or v.shape != (h // 2, w // 2):
            raise PfvError(N.PFV_ERR_BAD_ARG, "frame planes do not match the encoder's size (src/enc.rs:76-79)")
        return y, u, v

    def encode_iframe(self, frame) -> None:
        y, u, v = self._planes(frame)
        N.check(N.lib().pfv_encoder_encode_iframe(self._e, y.ctypes.data, u.ctypes.data, v.ctypes.data))

    def encode_pframe(self, frame) -> None:
        y, u, v = self._planes(frame)
        N.check(N.lib().pfv_encoder_encode_pframe(self._e, y.ctypes.data, u.ctypes.data, v.ctypes.data))

    def encode_dropframe(self) -> None:
        N.check(N.lib().pfv_encoder_encode_dropframe(self._e))

    def finish(self) -> None:
        N.check(N.lib().pfv_encoder_finish(self._e))

    def bytes(self) -> bytes:
        p, n = C.c_void_p(), C.c_size_t()
        N.check(N.lib().pfv_encoder_bytes(self._e, C.byref(p), C.byref(n)))
        return C.string_at(p.value, n.value)

    def prev_frame(self) -> np.ndarray:
        """Encoder.prev_frame (padded Y|U|V) after everything submitted so far (test aid)."""
        self.bytes()
        ctx = N.lib().pfv_encoder_ctx(self._e)
        g = N.Geometry()
        N.check(N.lib().pfv_ctx_geometry(ctx, C.byref(g)))
        out = np.empty(g.frame_bytes, np.uint8)
        N.check(N.lib().pfv_slot_read(ctx, N.lib().pfv_encoder_prev_frame_slot(self._e), out.ctypes.data))
        return out

    def close(self):
        if self._e:
            N.lib().pfv_encoder_close(self._e)
            self._e = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
