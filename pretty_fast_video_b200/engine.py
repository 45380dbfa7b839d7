"""Thin numpy-facing wrapper over the C ABI (include/pfv_b200.h).

`Engine` is the reference's macroblock seam (the plane loops of src/common.rs:351-521 as called from
src/enc.rs:84-97,134-147 and src/dec.rs:303-310,425-432) on one B200.  Everything here forwards to
libpfv_b200.so; nothing is computed in Python.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence, Union

import numpy as np

from . import _native as N

Buf = Union[np.ndarray, int, None]


def geometry_for(width: int, height: int) -> N.Geometry:
    """VideoFrame::new_padded geometry (src/frame.rs:28-49)."""
    g = N.Geometry()
    N.lib().pfv_geometry_for(width, height, C.byref(g))
    return g


def make_qtables(quality: int):
    """Encoder::new q-tables and px_err (src/enc.rs:40-51) -> (int32[4,64], float)."""
    out = ((C.c_int32 * 64) * 4)()
    px = C.c_float()
    N.check(N.lib().pfv_make_qtables(quality, out, C.byref(px)))
    return np.ctypeslib.as_array(out).reshape(4, 64).copy(), float(px.value)


def _addr(b: Buf) -> Optional[int]:
    if b is None:
        return None
    if isinstance(b, np.ndarray):
        if not b.flags["C_CONTIGUOUS"]:
            raise ValueError("buffers handed to the engine must be C-contiguous")
        return b.ctypes.data
    return int(b)


class PinnedArena:
    """Page-locked host memory from pfv_host_alloc, carved into numpy views."""

    def __init__(self, nbytes: int):
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        N.check(N.lib().pfv_host_alloc(C.byref(p), self.nbytes))
        self._ptr = p.value
        self._raw = (C.c_uint8 * self.nbytes).from_address(self._ptr)
        self.bytes = np.ctypeslib.as_array(self._raw)
        self._off = 0

    def take(self, shape, dtype) -> np.ndarray:
        dt = np.dtype(dtype)
        n = int(np.prod(shape)) * dt.itemsize
        start = (self._off + 255) & ~255
        if start + n > self.nbytes:
            raise MemoryError("pinned arena exhausted")
        self._off = start + n
        return self.bytes[start:start + n].view(dt).reshape(shape)

    def close(self):
        if self._ptr:
            self.bytes = None
            self._raw = None
            N.lib().pfv_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


@dataclass
class DecodeJob:
    kind: int                      # PFV_FRAME_I / PFV_FRAME_P
    dst_slot: int
    coeff: Buf                     # int16[nb*256] (host) or device address with device_ptrs
    qidx: Sequence[int] = (0, 1, 1)
    ref_slot: int = 0
    hdr: Buf = None                # MbHdr[nb] as uint8[nb,4] / structured, P only
    out: Optional[Sequence[Buf]] = None   # (y, u, v) tight host planes
    device_ptrs: bool = False
    dense_hint: bool = False       # PFV_JOB_DENSE: most sub-blocks carry AC terms (key frames; a hint, never changes results)


@dataclass
class SparseDecodeJob:
    kind: int
    dst_slot: int
    mb_off: np.ndarray             # uint32[nb+1]
    tok: np.ndarray                # uint32[ntok]: (position << 16) | uint16(value)
    qidx: Sequence[int] = (0, 1, 1)
    ref_slot: int = 0
    hdr: Buf = None
    out: Optional[Sequence[Buf]] = None


@dataclass
class EncodeJob:
    kind: int
    dst_slot: int
    src: Optional[Sequence[Buf]]   # (y, u, v) tight planes, or None when `rgb` is given
    coeff_out: Buf
    ref_slot: int = 0
    px_err: float = 0.0
    hdr_out: Buf = None
    device_ptrs: bool = False
    rgb: Buf = None                # packed RGB8 [h, w, 3] source: converted on the device (PFV_JOB_SRC_RGB)


@dataclass
class SparseEncodeJob:
    """pfv_encode_job_sparse: the frame's RLE sequence comes back instead of dense coefficients.  tok_out / stats_out /
    mb_off_out must be pinned (PinnedArena) or device memory: the device stores them itself."""
    kind: int
    dst_slot: int
    src: Optional[Sequence[Buf]]
    tok_out: Buf                   # uint32[tok_cap]: run | size << 4 | uint16(value) << 16
    stats_out: Buf                 # uint32[PFV_TOKSTATS_WORDS]
    tok_cap: int = 0               # entries; 0 = tok_out.size
    ref_slot: int = 0
    px_err: float = 0.0
    hdr_out: Buf = None
    mb_off_out: Buf = None
    device_ptrs: bool = False
    rgb: Buf = None


class Engine:
    def __init__(self, width: int, height: int, qtables: np.ndarray, nslots: int = 2, max_jobs: int = 1,
                 device: int = 0, stream: Optional[int] = None):
        qt = np.ascontiguousarray(qtables, dtype=np.int32).reshape(-1, 64)
        self._qt = qt
        self._ctx = C.c_void_p()
        self._keep = []
        l = N.lib()
        N.check(l.pfv_ctx_create(device, width, height, qt.ctypes.data_as(N._QT), qt.shape[0], nslots, max_jobs,
                                 C.c_void_p(stream) if stream else None, C.byref(self._ctx)))
        self.geometry = N.Geometry()
        N.check(l.pfv_ctx_geometry(self._ctx, C.byref(self.geometry)))
        self.nslots, self.max_jobs, self.device = nslots, max_jobs, device

    # -- lifetime ------------------------------------------------------------------------------
    def close(self):
        if self._ctx:
            N.lib().pfv_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()
            self._keep.clear()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- hot path ------------------------------------------------------------------------------
    def build_decode_jobs(self, jobs: Sequence[DecodeJob]):
        arr = (N.DecodeJob * len(jobs))()
        for a, j in zip(arr, jobs):
            a.kind, a.dst_slot, a.ref_slot = j.kind, j.dst_slot, j.ref_slot
            a.flags = (N.PFV_JOB_DEVICE_PTRS if j.device_ptrs else 0) | (N.PFV_JOB_DENSE if j.dense_hint else 0)
            for p in range(3):
                a.qidx[p] = int(j.qidx[p])
            a.hdr, a.coeff = _addr(j.hdr), _addr(j.coeff)
            if j.out is not None:
                a.out_y, a.out_u, a.out_v = (_addr(b) for b in j.out)
        return arr

    def build_encode_jobs(self, jobs: Sequence[EncodeJob]):
        arr = (N.EncodeJob * len(jobs))()
        for a, j in zip(arr, jobs):
            a.kind, a.dst_slot, a.ref_slot = j.kind, j.dst_slot, j.ref_slot
            a.flags = N.PFV_JOB_DEVICE_PTRS if j.device_ptrs else 0
            a.px_err = j.px_err
            if j.rgb is not None:
                a.flags |= N.PFV_JOB_SRC_RGB
                a.src_y = _addr(j.rgb)
            else:
                a.src_y, a.src_u, a.src_v = (_addr(b) for b in j.src)
            a.hdr_out, a.coeff_out = _addr(j.hdr_out), _addr(j.coeff_out)
        return arr

    def build_sparse_encode_jobs(self, jobs: Sequence[SparseEncodeJob]):
        arr = (N.EncodeJobSparse * len(jobs))()
        for a, j in zip(arr, jobs):
            a.kind, a.dst_slot, a.ref_slot = j.kind, j.dst_slot, j.ref_slot
            a.flags = N.PFV_JOB_DEVICE_PTRS if j.device_ptrs else 0
            a.px_err = j.px_err
            if j.rgb is not None:
                a.flags |= N.PFV_JOB_SRC_RGB
                a.src_y = _addr(j.rgb)
            else:
                a.src_y, a.src_u, a.src_v = (_addr(b) for b in j.src)
            a.hdr_out, a.mb_off_out, a.tok_out, a.stats_out = (_addr(b) for b in (j.hdr_out, j.mb_off_out, j.tok_out, j.stats_out))
            a.tok_cap = int(j.tok_cap) if j.tok_cap else int(j.tok_out.size)
        return arr

    def encode_submit_sparse(self, jobs, prebuilt=None):
        """Sparse encode transport (pfv_encode_submit_sparse): the run-length pass runs on the GPU, only its output crosses PCIe."""
        arr = prebuilt if prebuilt is not None else self.build_sparse_encode_jobs(jobs)
        self._keep.append((jobs, arr))
        N.check(N.lib().pfv_encode_submit_sparse(self._ctx, arr, len(arr)))

    def build_sparse_decode_jobs(self, jobs: Sequence[SparseDecodeJob]):
        arr = (N.DecodeJobSparse * len(jobs))()
        for a, j in zip(arr, jobs):
            a.kind, a.dst_slot, a.ref_slot, a.flags = j.kind, j.dst_slot, j.ref_slot, 0
            for p in range(3):
                a.qidx[p] = int(j.qidx[p])
            a.hdr, a.mb_off, a.tok = _addr(j.hdr), _addr(j.mb_off), _addr(j.tok)
            a.ntok = int(j.tok.size) if isinstance(j.tok, np.ndarray) else 0
            if j.out is not None:
                a.out_y, a.out_u, a.out_v = (_addr(b) for b in j.out)
        return arr

    def decode_submit_sparse(self, jobs, prebuilt=None):
        """Sparse coefficient transport (pfv_decode_submit_sparse): tokens cross PCIe, the dense layout is rebuilt on the GPU."""
        arr = prebuilt if prebuilt is not None else self.build_sparse_decode_jobs(jobs)
        self._keep.append((jobs, arr))
        N.check(N.lib().pfv_decode_submit_sparse(self._ctx, arr, len(arr)))

    def decode_submit(self, jobs, prebuilt=None):
        """Asynchronous; buffers must stay alive until sync() (a reference is kept for you)."""
        arr = prebuilt if prebuilt is not None else self.build_decode_jobs(jobs)
        self._keep.append((jobs, arr))
        N.check(N.lib().pfv_decode_submit(self._ctx, arr, len(arr)))

    def encode_submit(self, jobs, prebuilt=None):
        arr = prebuilt if prebuilt is not None else self.build_encode_jobs(jobs)
        self._keep.append((jobs, arr))
        N.check(N.lib().pfv_encode_submit(self._ctx, arr, len(arr)))

    def sync(self):
        try:
            N.check(N.lib().pfv_sync(self._ctx))
        finally:
            self._keep.clear()

    # -- slots ---------------------------------------------------------------------------------
    def slot_reset(self, slot: int):
        N.check(N.lib().pfv_slot_reset(self._ctx, slot))

    def slot_read(self, slot: int) -> np.ndarray:
        out = np.empty(self.geometry.frame_bytes, np.uint8)
        N.check(N.lib().pfv_slot_read(self._ctx, slot, out.ctypes.data))
        return out

    def slot_write(self, slot: int, frame: np.ndarray):
        f = np.ascontiguousarray(frame, np.uint8)
        assert f.size == self.geometry.frame_bytes
        N.check(N.lib().pfv_slot_write(self._ctx, slot, f.ctypes.data))

    def slot_read_visible(self, slot: int):
        g = self.geometry
        y = np.empty((g.height, g.width), np.uint8)
        u = np.empty((g.cheight, g.cwidth), np.uint8)
        v = np.empty((g.cheight, g.cwidth), np.uint8)
        N.check(N.lib().pfv_slot_read_visible(self._ctx, slot, y.ctypes.data, u.ctypes.data, v.ctypes.data))
        self.sync()
        return y, u, v

    def slot_read_rgb(self, slot: int) -> np.ndarray:
        """Visible crop of a slot as packed RGB8 [h, w, 3], converted on the device (save_frame, src/lib.rs:365-395)."""
        g = self.geometry
        rgb = np.empty((g.height, g.width, 3), np.uint8)
        N.check(N.lib().pfv_slot_read_rgb(self._ctx, slot, rgb.ctypes.data))
        self.sync()
        return rgb

    def slot_convert_rgb(self, slot: int, device_ptr: int):
        N.check(N.lib().pfv_slot_convert_rgb(self._ctx, slot, device_ptr))

    def slots_convert_rgb(self, slots, device_ptr: int, stride: int):
        """Batched pfv_slot_convert_rgb: picture i of `slots` goes to device_ptr + i * stride, one launch per 64 pictures."""
        a = np.ascontiguousarray(slots, np.uint32)
        N.check(N.lib().pfv_slots_convert_rgb(self._ctx, a.ctypes.data, a.size, device_ptr, stride))

    def slot_device_ptr(self, slot: int) -> int:
        p = C.c_void_p()
        N.check(N.lib().pfv_slot_device_ptr(self._ctx, slot, C.byref(p)))
        return p.value

    # -- instrumentation -------------------------------------------------------------------------
    @property
    def launch_count(self) -> int:
        return int(N.lib().pfv_ctx_launch_count(self._ctx))

    def enable_kernel_timing(self) -> None:
        """pfv_ctx_last_kernel_ms is off until asked for once (it costs two driver calls per submit): switch it on."""
        ms = C.c_float()
        N.lib().pfv_ctx_last_kernel_ms(self._ctx, C.byref(ms))

    def last_kernel_ms(self) -> float:
        ms = C.c_float()
        N.check(N.lib().pfv_ctx_last_kernel_ms(self._ctx, C.byref(ms)))
        return float(ms.value)
