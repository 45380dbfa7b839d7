#!/bin/bash
# Encoder object throughput with the calling thread's time split: tools/gpu_encobj.sh
timeout 600 python - <<'PY' 2>&1 | tail -12
import time, os, numpy as np
os.environ["PFV_TRACE"] = "1"
from pretty_fast_video_b200 import codec
from pretty_fast_video_b200.synth import SynthVideo
w, h, gop = 1920, 1080, 15
sv = SynthVideo(w, h, 0x50465602)
src = [sv.frame(t) for t in range(gop + 9)]
def enc_pass(n, threads, dense):
    os.environ["PFV_ENCODER_DENSE"] = dense
    with codec.Encoder(w, h, 30, 5, num_threads=threads) as enc:
        t0 = time.perf_counter()
        for t in range(n):
            (enc.encode_iframe if t % gop == 0 else enc.encode_pframe)(src[t % gop + ((t // gop) % 4) * 3])
        enc.finish()
        dt = time.perf_counter() - t0
        return n / dt, len(enc.bytes())
for threads in (4, 16):
    for dense in ("1", "0"):
        enc_pass(60, threads, dense)
        r = [enc_pass(240, threads, dense) for _ in range(2)]
        print("Encoder threads %2d dense=%s: %.0f frames/s (bytes %d)" % (threads, dense, max(x[0] for x in r), r[0][1]), flush=True)
PY
# decoder side of the same event change: the C++ example over a 1080p stream, sleeping and spinning waits
make -C examples >/dev/null 2>&1
LD_LIBRARY_PATH=pretty_fast_video_b200 examples/decode_speed --make /tmp/x.pfv 1920 1080 240 | tail -1
for spin in 0 1; do
  for th in 6 16; do
    echo "decode_speed spin=$spin threads=$th: $(PFV_EVENT_SPIN=$spin LD_LIBRARY_PATH=pretty_fast_video_b200 examples/decode_speed /tmp/x.pfv 6 $th | grep Decoded | sort -k5 -n | head -1)"
  done
done
