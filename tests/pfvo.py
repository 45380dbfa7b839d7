"""ctypes binding of the CPU oracle (oracle/libpfv_oracle.so).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "libpfv_oracle.so")


class Geometry(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "width", "height", "cwidth", "cheight", "pw", "ph", "cpw", "cph", "nb_y", "nb_c", "nb")]


_lib = None
_QT = C.POINTER(C.c_int32 * 64)
_I16 = C.POINTER(C.c_int16)
_U8 = C.POINTER(C.c_uint8)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            subprocess.check_call(["make", "-C", ORACLE_DIR])
        l = C.CDLL(LIB_PATH)
        l.pfvo_px_err.restype = C.c_float
        l.pfvo_frame_bytes.restype = C.c_size_t
        l.pfvo_block_search.restype = C.c_float
        l.pfvo_encoder_new.restype = C.c_void_p
        l.pfvo_encoder_bytes.restype = C.c_void_p
        l.pfvo_encoder_last_coeffs.restype = C.c_void_p
        l.pfvo_encoder_last_headers.restype = C.c_void_p
        l.pfvo_encoder_prev_frame.restype = C.c_void_p
        l.pfvo_decoder_new.restype = C.c_void_p
        l.pfvo_decoder_last_coeffs.restype = C.c_void_p
        l.pfvo_decoder_last_headers.restype = C.c_void_p
        l.pfvo_decoder_last_qidx.restype = C.c_void_p
        l.pfvo_decoder_framebuffer.restype = C.c_void_p
        l.pfvo_decoder_qtables.restype = C.c_void_p
        l.pfvo_entropy_roundtrip.restype = C.c_long
        l.pfvo_rle_encode.restype = C.c_size_t
        for n in ("pfvo_dct_scale_factor", "pfvo_q_table_intra", "pfvo_q_table_inter", "pfvo_zigzag", "pfvo_inv_zigzag"):
            getattr(l, n).restype = C.c_void_p
        _lib = l
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def geometry_for(w, h) -> Geometry:
    g = Geometry()
    lib().pfvo_geometry_for(C.c_uint32(w), C.c_uint32(h), C.byref(g))
    return g


def make_qtables(quality):
    out = np.zeros((4, 64), np.int32)
    lib().pfvo_make_qtables(C.c_int(quality), _p(out))
    return out, float(lib().pfvo_px_err(C.c_int(quality)))


def table(name, dtype, n=64):
    p = getattr(lib(), name)()
    return np.ctypeslib.as_array((np.ctypeslib.as_ctypes_type(dtype) * n).from_address(p)).copy()


def fdct8(v):
    a = np.array(v, np.int32)
    lib().pfvo_fdct8(_p(a))
    return a


def idct8(v):
    a = np.array(v, np.int32)
    lib().pfvo_idct8(_p(a))
    return a


def encode_subblock(px, q):
    px = np.ascontiguousarray(px, np.uint8); q = np.ascontiguousarray(q, np.int32)
    out = np.zeros(64, np.int16)
    lib().pfvo_encode_subblock(_p(px), _p(q), _p(out))
    return out


def encode_subblock_delta(d, q):
    d = np.ascontiguousarray(d, np.int16); q = np.ascontiguousarray(q, np.int32)
    out = np.zeros(64, np.int16)
    lib().pfvo_encode_subblock_delta(_p(d), _p(q), _p(out))
    return out


def decode_subblock(c, q):
    c = np.ascontiguousarray(c, np.int16); q = np.ascontiguousarray(q, np.int32)
    out = np.zeros(64, np.uint8)
    lib().pfvo_decode_subblock(_p(c), _p(q), _p(out))
    return out


def frame_init(g):
    n = lib().pfvo_frame_bytes(C.byref(g))
    f = np.zeros(n, np.uint8)
    lib().pfvo_frame_init(C.byref(g), _p(f))
    return f


def decode_iframe_coeffs(g, qtables, qidx, coeff, frame, nthreads=1):
    qt = np.ascontiguousarray(qtables, np.int32); qi = np.array(qidx, np.uint8)
    lib().pfvo_decode_iframe_coeffs(C.byref(g), _p(qt), _p(qi), _p(coeff), _p(frame), C.c_int(nthreads))


def decode_pframe_coeffs(g, qtables, qidx, hdr, coeff, frame, nthreads=1):
    qt = np.ascontiguousarray(qtables, np.int32); qi = np.array(qidx, np.uint8)
    lib().pfvo_decode_pframe_coeffs(C.byref(g), _p(qt), _p(qi), _p(hdr), _p(coeff), _p(frame), C.c_int(nthreads))


def encode_iframe_coeffs(g, qtables, y, u, v, prev_frame, nthreads=1):
    qt = np.ascontiguousarray(qtables, np.int32)
    coeff = np.zeros(g.nb * 256, np.int16)
    lib().pfvo_encode_iframe_coeffs(C.byref(g), _p(qt), _p(np.ascontiguousarray(y)), _p(np.ascontiguousarray(u)),
                                    _p(np.ascontiguousarray(v)), _p(coeff), _p(prev_frame), C.c_int(nthreads))
    return coeff


def encode_pframe_coeffs(g, qtables, px_err, y, u, v, prev_frame, nthreads=1):
    qt = np.ascontiguousarray(qtables, np.int32)
    coeff = np.zeros(g.nb * 256, np.int16)
    hdr = np.zeros((g.nb, 4), np.uint8)
    lib().pfvo_encode_pframe_coeffs(C.byref(g), _p(qt), C.c_float(px_err), _p(np.ascontiguousarray(y)),
                                    _p(np.ascontiguousarray(u)), _p(np.ascontiguousarray(v)), _p(hdr), _p(coeff),
                                    _p(prev_frame), C.c_int(nthreads))
    return hdr, coeff


def crop_frame(g, frame):
    y = np.zeros((g.height, g.width), np.uint8)
    u = np.zeros((g.cheight, g.cwidth), np.uint8)
    v = np.zeros((g.cheight, g.cwidth), np.uint8)
    lib().pfvo_crop_frame(C.byref(g), _p(frame), _p(y), _p(u), _p(v))
    return y, u, v


def rgb_to_yuv420(rgb):
    h, w, _ = rgb.shape
    rgb = np.ascontiguousarray(rgb, np.uint8)
    y = np.zeros((h, w), np.uint8); u = np.zeros((h // 2, w // 2), np.uint8); v = np.zeros((h // 2, w // 2), np.uint8)
    lib().pfvo_rgb_to_yuv420(_p(rgb), w, h, _p(y), _p(u), _p(v))
    return y, u, v


def yuv420_to_rgb(y, u, v):
    h, w = y.shape
    rgb = np.zeros((h, w, 3), np.uint8)
    lib().pfvo_yuv420_to_rgb(_p(np.ascontiguousarray(y)), _p(np.ascontiguousarray(u)), _p(np.ascontiguousarray(v)), w, h, _p(rgb))
    return rgb


class Encoder:
    """pfv_rs::enc::Encoder restated (oracle)."""

    def __init__(self, width, height, framerate, quality, nthreads=1):
        self.g = geometry_for(width, height)
        self.h = lib().pfvo_encoder_new(C.c_int(width), C.c_int(height), C.c_int(framerate), C.c_int(quality), C.c_int(nthreads))
        if not self.h:
            raise ValueError("bad encoder arguments")

    def _planes(self, y, u, v):
        return _p(np.ascontiguousarray(y, np.uint8)), _p(np.ascontiguousarray(u, np.uint8)), _p(np.ascontiguousarray(v, np.uint8))

    def encode_iframe(self, y, u, v):
        assert lib().pfvo_encoder_encode_iframe(C.c_void_p(self.h), *self._planes(y, u, v)) == 0

    def encode_pframe(self, y, u, v):
        assert lib().pfvo_encoder_encode_pframe(C.c_void_p(self.h), *self._planes(y, u, v)) == 0

    def encode_dropframe(self):
        assert lib().pfvo_encoder_encode_dropframe(C.c_void_p(self.h)) == 0

    def finish(self):
        assert lib().pfvo_encoder_finish(C.c_void_p(self.h)) == 0

    def bytes(self) -> bytes:
        n = C.c_size_t()
        p = lib().pfvo_encoder_bytes(C.c_void_p(self.h), C.byref(n))
        return C.string_at(p, n.value)

    def last_coeffs(self):
        p = lib().pfvo_encoder_last_coeffs(C.c_void_p(self.h))
        return np.ctypeslib.as_array((C.c_int16 * (self.g.nb * 256)).from_address(p)).copy()

    def last_headers(self):
        p = lib().pfvo_encoder_last_headers(C.c_void_p(self.h))
        return np.ctypeslib.as_array((C.c_uint8 * (self.g.nb * 4)).from_address(p)).reshape(-1, 4).copy()

    def prev_frame(self):
        n = lib().pfvo_frame_bytes(C.byref(self.g))
        p = lib().pfvo_encoder_prev_frame(C.c_void_p(self.h))
        return np.ctypeslib.as_array((C.c_uint8 * n).from_address(p)).copy()

    def close(self):
        if self.h:
            lib().pfvo_encoder_free(C.c_void_p(self.h))
            self.h = None

    def __del__(self):
        self.close()


class Decoder:
    """pfv_rs::dec::Decoder restated (oracle) over an in-memory stream."""

    def __init__(self, data: bytes, nthreads=1):
        self._data = np.frombuffer(data, np.uint8).copy()
        err = C.c_int()
        self.h = lib().pfvo_decoder_new(_p(self._data), C.c_size_t(self._data.size), C.c_int(nthreads), C.byref(err))
        self.err = err.value
        if not self.h:
            raise ValueError({1: "FormatError", 2: "VersionError", 3: "IOError"}.get(err.value, "error"))
        self.width = lib().pfvo_decoder_width(C.c_void_p(self.h))
        self.height = lib().pfvo_decoder_height(C.c_void_p(self.h))
        self.framerate = lib().pfvo_decoder_framerate(C.c_void_p(self.h))
        self.g = geometry_for(self.width, self.height)

    def reset(self):
        lib().pfvo_decoder_reset(C.c_void_p(self.h))

    def advance_frame(self):
        """-> (more, frame or None); frame = (y, u, v)"""
        g = self.g
        y = np.zeros((g.height, g.width), np.uint8)
        u = np.zeros((g.cheight, g.cwidth), np.uint8)
        v = np.zeros((g.cheight, g.cwidth), np.uint8)
        got = C.c_int()
        rc = lib().pfvo_decoder_advance_frame(C.c_void_p(self.h), _p(y), _p(u), _p(v), C.byref(got))
        if rc < 0:
            raise IOError("truncated stream")
        return rc == 1, ((y, u, v) if got.value else None)

    def last_kind(self):
        return lib().pfvo_decoder_last_kind(C.c_void_p(self.h))

    def last_coeffs(self):
        p = lib().pfvo_decoder_last_coeffs(C.c_void_p(self.h))
        return np.ctypeslib.as_array((C.c_int16 * (self.g.nb * 256)).from_address(p)).copy()

    def last_headers(self):
        p = lib().pfvo_decoder_last_headers(C.c_void_p(self.h))
        return np.ctypeslib.as_array((C.c_uint8 * (self.g.nb * 4)).from_address(p)).reshape(-1, 4).copy()

    def last_qidx(self):
        p = lib().pfvo_decoder_last_qidx(C.c_void_p(self.h))
        return np.ctypeslib.as_array((C.c_uint8 * 3).from_address(p)).copy()

    def qtables(self):
        n = C.c_int()
        p = lib().pfvo_decoder_qtables(C.c_void_p(self.h), C.byref(n))
        return np.ctypeslib.as_array((C.c_int32 * (n.value * 64)).from_address(p)).reshape(-1, 64).copy()

    def framebuffer(self):
        n = lib().pfvo_frame_bytes(C.byref(self.g))
        p = lib().pfvo_decoder_framebuffer(C.c_void_p(self.h))
        return np.ctypeslib.as_array((C.c_uint8 * n).from_address(p)).copy()

    def close(self):
        if self.h:
            lib().pfvo_decoder_free(C.c_void_p(self.h))
            self.h = None

    def __del__(self):
        self.close()


def entropy_roundtrip(data):
    d = np.ascontiguousarray(data, np.int16)
    out = np.zeros_like(d)
    rc = lib().pfvo_entropy_roundtrip(_p(d), C.c_size_t(d.size), _p(out))
    return rc, out


def rle_frame(coeff, coded=None):
    """rle_encode per macroblock (enc.rs:256-262 / :370-378) over a frame's dense coefficients [nb, 256]; `coded` = per-macroblock
    has_coeff (P frames, enc.rs:357-358).  Returns (tok, table, mb_off) with tok = run | size << 4 | uint16(value) << 16."""
    c = np.ascontiguousarray(coeff, np.int16).reshape(-1, 256)
    nb = c.shape[0]
    z = np.zeros(256, np.uint8); sz = np.zeros(256, np.uint8); v = np.zeros(256, np.int16)
    table = np.zeros(16, np.int32)
    toks, mb_off = [], np.zeros(nb + 1, np.uint32)
    for m in range(nb):
        if coded is None or coded[m]:
            n = lib().pfvo_rle_encode(_p(c[m]), C.c_size_t(256), _p(z), _p(sz), _p(v), C.c_size_t(256), _p(table))
            assert n <= 256
            toks.append(z[:n].astype(np.uint32) | (sz[:n].astype(np.uint32) << 4) | (v[:n].view(np.uint16).astype(np.uint32) << 16))
            mb_off[m + 1] = mb_off[m] + n
        else:
            mb_off[m + 1] = mb_off[m]
    tok = np.concatenate(toks) if toks else np.zeros(0, np.uint32)
    return tok, table, mb_off
