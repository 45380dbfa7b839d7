#!/bin/bash
# run-to-run spread of the Encoder object (60-frame passes like bench.py's sub-leg), with the calling thread's time split
timeout 600 python - <<'PY' 2>&1 | tail -24
import time, os, numpy as np
os.environ["PFV_TRACE"] = "1"
from pretty_fast_video_b200 import codec
from pretty_fast_video_b200.synth import SynthVideo
w, h, gop = 1920, 1080, 15
sv = SynthVideo(w, h, 0x50465602)
src = [sv.frame(t) for t in range(gop + 9)]
def enc_pass(n, threads):
    with codec.Encoder(w, h, 30, 5, num_threads=threads) as enc:
        t0 = time.perf_counter()
        for t in range(n):
            (enc.encode_iframe if t % gop == 0 else enc.encode_pframe)(src[t % gop + ((t // gop) % 4) * 3])
        enc.finish()
        dt = time.perf_counter() - t0
    return n / dt
for threads in (16,):
    for spin in ("0",):
        os.environ["PFV_EVENT_SPIN"] = spin
        r = [enc_pass(60, threads) for _ in range(6)]
        print("threads %2d spin=%s: " % (threads, spin) + " ".join("%.0f" % x for x in r), flush=True)
PY
