"""pretty_fast_video_b200 - B200 (sm_100a) macroblock engine for the Pretty Fast Video codec.

Only the per-macroblock hot path of pfv-rs 0.2.2 lives here: CUDA kernels behind the C ABI of
``include/pfv_b200.h`` plus the host-side mirror of the reference's Encoder/Decoder interface.
"""
from ._native import (  # noqa: F401
    PFV_FRAME_I, PFV_FRAME_P, PFV_JOB_DEVICE_PTRS, PFV_JOB_SRC_RGB, Geometry, MbHdr, PfvError, lib,
)
from .engine import Engine, PinnedArena, geometry_for, make_qtables  # noqa: F401
from . import codec  # noqa: F401
from .codec import Decoder, Encoder, DecodeError  # noqa: F401

__all__ = ["Decoder", "Encoder", "DecodeError", "codec", "Engine", "PinnedArena", "geometry_for", "make_qtables", "Geometry", "MbHdr", "PfvError",
           "PFV_FRAME_I", "PFV_FRAME_P", "PFV_JOB_DEVICE_PTRS", "PFV_JOB_SRC_RGB", "lib"]
