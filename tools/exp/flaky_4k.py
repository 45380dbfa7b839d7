"""Repeats the 4K round trip of tests/test_gpu_codec.py::test_round_trip_4k_p_stream_sharded_like_config5 and says which side is
   off when it fails: the encoder's reconstruction, the stream it wrote, or the decoder.   python tools/exp/flaky_4k.py [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import pfvo
from pretty_fast_video_b200 import codec
from pretty_fast_video_b200.synth import SynthVideo
from test_gpu_codec import gpu_decode_all, oracle_decode_all
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
nthreads = int(sys.argv[2]) if len(sys.argv) > 2 else 4
w, h, gop = 3840, 2160, 3
sv = SynthVideo(w, h, 4711)
frames = [sv.frame(t) for t in range(2 * gop)]
ref_data = None
for rep in range(reps):
    with codec.Encoder(w, h, 30, 5, num_threads=nthreads) as enc:
        for t in range(2 * gop):
            (enc.encode_iframe if t % gop == 0 else enc.encode_pframe)(frames[t])
        recon = enc.prev_frame()
        enc.finish()
        data = enc.bytes()
    if ref_data is None:
        ref_data = data
        _, want_fb = oracle_decode_all(data)
    same_stream = data == ref_data
    whole, fb = gpu_decode_all(data, num_threads=nthreads)
    if same_stream:
        ofb = want_fb
    else:
        _, ofb = oracle_decode_all(data)
    bad_recon = int((recon != ofb).sum())
    bad_dec = int((fb != ofb).sum())
    where = ""
    if bad_recon:
        idx = np.flatnonzero(recon != ofb)
        where += f" recon first/last bad byte {idx[0]} {idx[-1]}"
    if bad_dec:
        idx = np.flatnonzero(fb != ofb)
        where += f" decoder first/last bad byte {idx[0]} {idx[-1]}"
    if bad_recon or bad_dec or not same_stream or rep == reps - 1:
        print(f"threads {nthreads} rep {rep}: stream {'same' if same_stream else 'DIFFERENT'} ({len(data)} bytes); encoder recon vs oracle decode: {bad_recon} bytes off; "
              f"GPU decoder vs oracle decode: {bad_dec} bytes off{where}", flush=True)
