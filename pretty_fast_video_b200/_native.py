"""ctypes binding of the engine's C ABI (include/pfv_b200.h).

The shared library is built in-tree by ``pretty_fast_video_b200/csrc/Makefile`` (see
``__graft_entry__.build``).  There is no Python or CPU fallback: if the library is missing,
or no sm_100 device is present, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpfv_b200.so")

PFV_OK = 0
PFV_ERR_BAD_ARG = -1
PFV_ERR_CUDA = -2
PFV_ERR_NO_DEVICE = -3
PFV_ERR_BAD_MV = -4
PFV_ERR_NOMEM = -5
PFV_ERR_BAD_STREAM = -6
PFV_ERR_BAD_VERSION = -7
PFV_ERR_IO = -8
PFV_ERR_STATE = -9

PFV_FRAME_I = 1
PFV_FRAME_P = 2
PFV_JOB_DEVICE_PTRS = 1
PFV_JOB_SRC_RGB = 2
PFV_JOB_DENSE = 4


class PfvError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"pfv status {code}: {msg}")
        self.code = code


class Geometry(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "width", "height", "cwidth", "cheight", "pw", "ph", "cpw", "cph", "nb_y", "nb_c", "nb", "frame_bytes")]


class MbHdr(C.Structure):
    _fields_ = [("mx", C.c_int8), ("my", C.c_int8), ("has_coeff", C.c_uint8), ("reserved", C.c_uint8)]


class DecodeJob(C.Structure):
    _fields_ = [
        ("kind", C.c_uint32), ("flags", C.c_uint32), ("dst_slot", C.c_uint32), ("ref_slot", C.c_uint32),
        ("qidx", C.c_uint8 * 3), ("reserved", C.c_uint8),
        ("hdr", C.c_void_p), ("coeff", C.c_void_p),
        ("out_y", C.c_void_p), ("out_u", C.c_void_p), ("out_v", C.c_void_p),
    ]


class EncodeJob(C.Structure):
    _fields_ = [
        ("kind", C.c_uint32), ("flags", C.c_uint32), ("dst_slot", C.c_uint32), ("ref_slot", C.c_uint32),
        ("px_err", C.c_float), ("reserved", C.c_uint32),
        ("src_y", C.c_void_p), ("src_u", C.c_void_p), ("src_v", C.c_void_p),
        ("hdr_out", C.c_void_p), ("coeff_out", C.c_void_p),
    ]


class DecodeJobSparse(C.Structure):
    _fields_ = [
        ("kind", C.c_uint32), ("flags", C.c_uint32), ("dst_slot", C.c_uint32), ("ref_slot", C.c_uint32),
        ("qidx", C.c_uint8 * 3), ("reserved", C.c_uint8),
        ("hdr", C.c_void_p), ("mb_off", C.c_void_p), ("tok", C.c_void_p),
        ("ntok", C.c_uint32), ("reserved2", C.c_uint32),
        ("out_y", C.c_void_p), ("out_u", C.c_void_p), ("out_v", C.c_void_p),
    ]


class EncodeJobSparse(C.Structure):
    _fields_ = [
        ("kind", C.c_uint32), ("flags", C.c_uint32), ("dst_slot", C.c_uint32), ("ref_slot", C.c_uint32),
        ("px_err", C.c_float), ("tok_cap", C.c_uint32),
        ("src_y", C.c_void_p), ("src_u", C.c_void_p), ("src_v", C.c_void_p),
        ("hdr_out", C.c_void_p), ("mb_off_out", C.c_void_p), ("tok_out", C.c_void_p), ("stats_out", C.c_void_p),
    ]


PFV_TOKSTATS_WORDS, PFV_TOKSTATS_NTOK, PFV_TOKSTATS_FLAGS = 36, 32, 33
PFV_TOKFLAG_OVERFLOW, PFV_TOKFLAG_RANGE = 1, 2


class StreamInfo(C.Structure):
    _fields_ = [("version", C.c_uint32), ("width", C.c_uint32), ("height", C.c_uint32), ("framerate", C.c_uint32),
                ("num_qtables", C.c_uint32), ("first_packet", C.c_uint64)]


class Packet(C.Structure):
    _fields_ = [("type", C.c_uint8), ("reserved", C.c_uint8 * 3), ("len", C.c_uint32), ("payload", C.c_uint64)]


ONVIDEO = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p)
READ_FN = C.CFUNCTYPE(C.c_longlong, C.c_void_p, C.c_void_p, C.c_size_t)      # pfv_read_fn
WRITE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t)          # pfv_write_fn

# every symbol include/pfv_b200.h declares: name -> (restype, argtypes)
_QT = C.POINTER(C.c_int32 * 64)
SYMBOLS = {
    "pfv_abi_version": (C.c_int, []),
    "pfv_last_error": (C.c_char_p, []),
    "pfv_device_count": (C.c_int, []),
    "pfv_geometry_for": (None, [C.c_uint32, C.c_uint32, C.POINTER(Geometry)]),
    "pfv_make_qtables": (C.c_int, [C.c_int, _QT, C.POINTER(C.c_float)]),
    "pfv_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "pfv_host_free": (None, [C.c_void_p]),
    "pfv_ctx_create": (C.c_int, [C.c_int, C.c_uint32, C.c_uint32, _QT, C.c_uint32, C.c_uint32, C.c_uint32,
                                 C.c_void_p, C.POINTER(C.c_void_p)]),
    "pfv_ctx_destroy": (None, [C.c_void_p]),
    "pfv_ctx_geometry": (C.c_int, [C.c_void_p, C.POINTER(Geometry)]),
    "pfv_sync": (C.c_int, [C.c_void_p]),
    "pfv_slot_reset": (C.c_int, [C.c_void_p, C.c_uint32]),
    "pfv_slot_read": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p]),
    "pfv_slot_write": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p]),
    "pfv_slot_read_visible": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pfv_slot_device_ptr": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p)]),
    "pfv_slot_read_rgb": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p]),
    "pfv_slot_convert_rgb": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p]),
    "pfv_slots_convert_rgb": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t]),
    "pfv_decode_submit": (C.c_int, [C.c_void_p, C.POINTER(DecodeJob), C.c_uint32]),
    "pfv_encode_submit": (C.c_int, [C.c_void_p, C.POINTER(EncodeJob), C.c_uint32]),
    "pfv_decode_submit_sparse": (C.c_int, [C.c_void_p, C.POINTER(DecodeJobSparse), C.c_uint32]),
    "pfv_encode_submit_sparse": (C.c_int, [C.c_void_p, C.POINTER(EncodeJobSparse), C.c_uint32]),
    "pfv_ctx_last_submit_id": (C.c_uint64, [C.c_void_p]),
    "pfv_ctx_wait_submit": (C.c_int, [C.c_void_p, C.c_uint64]),
    "pfv_ctx_launch_count": (C.c_uint64, [C.c_void_p]),
    "pfv_ctx_last_kernel_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    # host codec layer
    "pfv_stream_parse_header": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(StreamInfo), C.c_void_p, C.c_uint32]),
    "pfv_stream_index": (C.c_int, [C.c_void_p, C.c_size_t, C.c_uint64, C.POINTER(Packet), C.c_uint32,
                                   C.POINTER(C.c_uint32), C.POINTER(C.c_int)]),
    "pfv_packet_decode": (C.c_int, [C.POINTER(Geometry), C.c_uint32, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]),
    "pfv_packet_encode": (C.c_int, [C.POINTER(Geometry), C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                    C.POINTER(C.c_size_t)]),
    "pfv_packet_encode_bound": (C.c_size_t, [C.POINTER(Geometry)]),
    "pfv_packet_encode_tokens": (C.c_int, [C.POINTER(Geometry), C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_size_t, C.POINTER(C.c_size_t)]),
    "pfv_packet_tokenize": (C.c_int, [C.POINTER(Geometry), C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                      C.c_void_p, C.c_void_p]),
    "pfv_packet_token_bound": (C.c_uint32, [C.POINTER(Geometry), C.c_void_p, C.c_size_t]),
    "pfv_decoder_open": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p)]),
    "pfv_decoder_open_reader": (C.c_int, [READ_FN, C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p)]),
    "pfv_decoder_close": (None, [C.c_void_p]),
    "pfv_decoder_width": (C.c_uint32, [C.c_void_p]),
    "pfv_decoder_height": (C.c_uint32, [C.c_void_p]),
    "pfv_decoder_framerate": (C.c_uint32, [C.c_void_p]),
    "pfv_decoder_reset": (C.c_int, [C.c_void_p]),
    "pfv_decoder_advance_frame": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                            C.POINTER(C.c_void_p)]),
    "pfv_decoder_advance_delta": (C.c_int, [C.c_void_p, C.c_double, ONVIDEO, C.c_void_p]),
    "pfv_decoder_ctx": (C.c_void_p, [C.c_void_p]),
    "pfv_decoder_framebuffer_slot": (C.c_uint32, [C.c_void_p]),
    "pfv_encoder_open": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.c_int, C.POINTER(C.c_void_p)]),
    "pfv_encoder_close": (None, [C.c_void_p]),
    "pfv_encoder_set_writer": (C.c_int, [C.c_void_p, WRITE_FN, C.c_void_p]),
    "pfv_encoder_encode_iframe": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pfv_encoder_encode_pframe": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pfv_encoder_encode_dropframe": (C.c_int, [C.c_void_p]),
    "pfv_encoder_finish": (C.c_int, [C.c_void_p]),
    "pfv_encoder_bytes": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "pfv_encoder_ctx": (C.c_void_p, [C.c_void_p]),
    "pfv_encoder_prev_frame_slot": (C.c_uint32, [C.c_void_p]),
}

_lib = None


def lib() -> C.CDLL:
    """Loads libpfv_b200.so (once).  Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing - build it with `make -C pretty_fast_video_b200/csrc` "
                "(or __graft_entry__.build()); the engine has no CPU fallback")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int) -> None:
    if rc != PFV_OK:
        raise PfvError(rc, lib().pfv_last_error().decode("utf-8", "replace"))
