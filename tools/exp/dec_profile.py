"""per-frame latency of codec.Decoder on a 1080p stream (diagnostic)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from pretty_fast_video_b200 import codec
from pretty_fast_video_b200.synth import SynthVideo
w, h, gop, ngop = 1920, 1080, 15, 4
sv = SynthVideo(w, h, 0x50465602)
frames = [sv.frame(t) for t in range(gop)]
t0 = time.perf_counter()
with codec.Encoder(w, h, 30, 5, num_threads=8) as enc:
    for t in range(gop * ngop):
        (enc.encode_iframe if t % gop == 0 else enc.encode_pframe)(frames[t % gop])
    enc.finish()
    data = enc.bytes()
print("encode: %.1f fps, stream %d bytes" % (gop * ngop / (time.perf_counter() - t0), len(data)))
for threads, ahead in ((16, 0), (16, 6), (8, 0), (4, 0), (1, 1)):
    for rep in range(2):
        t_open = time.perf_counter()
        dec = codec.Decoder(data, num_threads=threads, read_ahead=ahead)
        t1 = time.perf_counter()
        stamps = []
        while dec.advance_frame(lambda fr: None):
            stamps.append(time.perf_counter())
        dt = np.diff([t1] + stamps) * 1e3
        dec.close()
        print("threads %2d ahead %d rep %d: open %.1f ms, total %.1f ms (%.0f fps); first 8 frame gaps ms: %s ; median gap %.2f" % (
            threads, ahead, rep, (t1 - t_open) * 1e3, (stamps[-1] - t1) * 1e3, len(stamps) / (stamps[-1] - t1),
            np.round(dt[:8], 2), np.median(dt)))

# raw ctypes loop (no numpy views)
import ctypes as C
from pretty_fast_video_b200 import _native as N
buf = np.frombuffer(data, np.uint8)
for rep in range(2):
    d = C.c_void_p()
    N.check(N.lib().pfv_decoder_open(buf.ctypes.data, buf.size, 0, 16, 6, C.byref(d)))
    got = C.c_int(); y = C.c_void_p(); u = C.c_void_p(); v = C.c_void_p()
    t1 = time.perf_counter(); stamps = []
    while N.lib().pfv_decoder_advance_frame(d, C.byref(got), C.byref(y), C.byref(u), C.byref(v)) == 1:
        stamps.append(time.perf_counter())
    dt = np.diff([t1] + stamps) * 1e3
    print("raw ctypes: total %.1f ms (%.0f fps) first gaps %s median %.3f" % ((stamps[-1] - t1) * 1e3, len(stamps) / (stamps[-1] - t1), np.round(dt[:8], 2), np.median(dt)))
    N.lib().pfv_decoder_close(d)
# view creation cost
t0 = time.perf_counter()
for _ in range(100):
    a = np.ctypeslib.as_array((C.c_uint8 * (1920 * 1080)).from_address(buf.ctypes.data))
print("as_array per call ms", (time.perf_counter() - t0) * 10)
# encoder raw timing
e = C.c_void_p()
N.check(N.lib().pfv_encoder_open(w, h, 30, 5, 8, 0, C.byref(e)))
ys = [tuple(np.ascontiguousarray(p) for p in f) for f in frames]
t0 = time.perf_counter()
for t in range(30):
    y_, u_, v_ = ys[t % gop]
    fn = N.lib().pfv_encoder_encode_iframe if t % gop == 0 else N.lib().pfv_encoder_encode_pframe
    N.check(fn(e, y_.ctypes.data, u_.ctypes.data, v_.ctypes.data))
t1 = time.perf_counter()
N.check(N.lib().pfv_encoder_finish(e))
t2 = time.perf_counter()
print("encoder raw: submit loop %.1f ms, finish %.1f ms -> %.0f fps" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, 30 / (t2 - t0)))
N.lib().pfv_encoder_close(e)
