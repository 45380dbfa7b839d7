"""PFV_TRACE=1 python tools/exp/dec_trace.py : where the time of a long single-stream decode goes"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from pretty_fast_video_b200 import codec
from pretty_fast_video_b200.synth import SynthVideo
w, h, gop, ngop = 1920, 1080, 15, 16
sv = SynthVideo(w, h, 0x50465602)
frames = [sv.frame(t) for t in range(gop)]
with codec.Encoder(w, h, 30, 5, num_threads=16) as enc:
    t0 = time.perf_counter()
    for t in range(gop * ngop):
        (enc.encode_iframe if t % gop == 0 else enc.encode_pframe)(frames[t % gop])
    enc.finish()
    data = enc.bytes()
    print("encoder: %.0f fps" % (gop * ngop / (time.perf_counter() - t0)))
cfgs = ((16, 0), (16, 48), (16, 24), (12, 48))
if len(sys.argv) > 1:                                   # python tools/exp/dec_trace.py 10 12 14 16 : a sweep over pool sizes
    cfgs = tuple((int(a), 0) for a in sys.argv[1:])
for threads, ahead in cfgs:
    for rep in range(3 if len(sys.argv) > 1 else 2):
        dec = codec.Decoder(data, num_threads=threads, read_ahead=ahead)
        t1 = time.perf_counter()
        n = 0
        while dec.advance_frame(lambda fr: None):
            n += 1
        dt = time.perf_counter() - t1
        dec.close()
        print("threads %d ahead %d: %d frames in %.1f ms = %.0f fps" % (threads, ahead, n, dt * 1e3, n / dt))
