#!/usr/bin/env python
"""Encoder.encode_* over 480 frames of the bench's 1080p sequence, PFV_TRACE=1: where the calling thread's time goes.
   python tools/exp/enc_trace.py [threads]"""
import os, sys, time
os.environ["PFV_TRACE"] = "1"
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from pretty_fast_video_b200 import codec
from pretty_fast_video_b200.synth import SynthVideo
nthreads = int(sys.argv[1]) if len(sys.argv) > 1 else (os.cpu_count() or 1)
w, h, gop = 1920, 1080, 15
sv = SynthVideo(w, h, 0x50465602)
src = [sv.frame(t) for t in range(gop + 3 * 3)]
for n in (gop, 32 * gop, 32 * gop, 32 * gop, 32 * gop, 32 * gop, 32 * gop):
    with codec.Encoder(w, h, 30, 5, num_threads=nthreads, device=0) as enc:
        t0 = time.perf_counter()
        for t in range(n):
            (enc.encode_iframe if t % gop == 0 else enc.encode_pframe)(src[t % gop + ((t // gop) % 4) * 3])
        t1 = time.perf_counter()
        enc.finish()
        t2 = time.perf_counter()
        nbytes = len(enc.bytes())
        t3 = time.perf_counter()
    print(f"{n} frames, {nthreads} threads: {n / (t2 - t0):.0f} frames/s (calls {1e3 * (t1 - t0):.1f} ms, finish {1e3 * (t2 - t1):.1f} ms; "
          f"bytes() of {nbytes / 1e6:.1f} MB {1e3 * (t3 - t2):.1f} ms)", flush=True)
